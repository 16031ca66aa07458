import torch


def scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    assert reduce in ("sum", "add")
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.index_add(dim, index, src)
