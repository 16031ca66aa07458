import torch


class MessagePassing(torch.nn.Module):
    """propagate(edge_index, x=..., W=...) = scatter-add of message(x[edge_index[0]], W)
    at edge_index[1] (PyG default flow source_to_target)."""

    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        x = kwargs.pop("x")
        x_j = x[edge_index[0]]
        msg = self.message(x_j, **kwargs)
        out = torch.zeros_like(x[:, : msg.shape[1]]) if msg.shape[1] != x.shape[1] else torch.zeros_like(x)
        return out.index_add(0, edge_index[1], msg)
