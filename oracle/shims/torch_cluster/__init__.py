"""Stand-in for the absent `torch_cluster` package (test infrastructure only).

Restates the documented behaviour of torch_cluster's CUDA `radius` kernel
(torch-cluster is un-vendored and unpinned in the reference's pyproject.toml:19):
one query point at a time, candidates of the same example visited in ascending
index order, strict `dist2 < r*r` in the dtype of `x`, at most
`max_num_neighbors` hits kept per query.  Parity with the real wheel is
UNPINNED (it cannot be installed here).
"""
import torch


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(y.shape[0], dtype=torch.long, device=y.device)
    diff = y[:, None, :] - x[None, :, :]
    d2 = (diff * diff).sum(-1)
    r2 = torch.as_tensor(r, dtype=x.dtype) ** 2
    hit = (d2 < r2) & (batch_y[:, None] == batch_x[None, :])
    # keep the first max_num_neighbors hits of every query row
    rank = torch.cumsum(hit.to(torch.long), dim=1)
    hit = hit & (rank <= max_num_neighbors)
    row, col = torch.nonzero(hit, as_tuple=True)  # row-major: query asc, candidate asc
    return torch.stack([row, col], dim=0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target",
                 num_workers=1, batch_size=None):
    assert flow in ("source_to_target", "target_to_source")
    edge_index = radius(x, x, r, batch, batch,
                        max_num_neighbors if loop else max_num_neighbors + 1, num_workers, batch_size)
    if flow == "source_to_target":
        row, col = edge_index[1], edge_index[0]
    else:
        row, col = edge_index[0], edge_index[1]
    if not loop:
        mask = row != col
        row, col = row[mask], col[mask]
    return torch.stack([row, col], dim=0)
