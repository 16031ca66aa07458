"""Stand-in for the absent `torch_cluster` package (test infrastructure only).

Restates the documented behaviour of torch_cluster's CUDA `radius` kernel
(torch-cluster is un-vendored and unpinned in the reference's pyproject.toml:19):
one query point at a time, candidates of the same example visited in ascending
index order, strict `dist2 < r*r` in the dtype of `x`, at most
`max_num_neighbors` hits kept per query.  Parity with the real wheel is
UNPINNED (it cannot be installed here).
"""
import torch


def _radius_dense(x, y, r, batch_x, batch_y, max_num_neighbors):
    diff = y[:, None, :] - x[None, :, :]
    d2 = (diff * diff).sum(-1)
    r2 = torch.as_tensor(r, dtype=x.dtype) ** 2
    hit = (d2 < r2) & (batch_y[:, None] == batch_x[None, :])
    # keep the first max_num_neighbors hits of every query row
    rank = torch.cumsum(hit.to(torch.long), dim=1)
    hit = hit & (rank <= max_num_neighbors)
    row, col = torch.nonzero(hit, as_tuple=True)  # row-major: query asc, candidate asc
    return torch.stack([row, col], dim=0)


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1, batch_size=None):
    if batch_x is None:
        batch_x = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(y.shape[0], dtype=torch.long, device=y.device)
    # Sorted batches (the only case the reference produces): one dense block per example instead of an [N, N] table over
    # the whole batch - same pairs in the same order, and the timed CPU baseline is not charged for O(N^2) over molecules
    # that cannot interact.
    sorted_x = bool((batch_x[1:] >= batch_x[:-1]).all()) if batch_x.numel() > 1 else True
    sorted_y = bool((batch_y[1:] >= batch_y[:-1]).all()) if batch_y.numel() > 1 else True
    if not (sorted_x and sorted_y) or batch_x.numel() == 0 or batch_y.numel() == 0:
        return _radius_dense(x, y, r, batch_x, batch_y, max_num_neighbors)
    nb = int(max(batch_x[-1], batch_y[-1])) + 1
    px = torch.searchsorted(batch_x, torch.arange(nb + 1, device=x.device))
    py = torch.searchsorted(batch_y, torch.arange(nb + 1, device=y.device))
    out = []
    for b in range(nb):
        x0, x1, y0, y1 = int(px[b]), int(px[b + 1]), int(py[b]), int(py[b + 1])
        if x1 == x0 or y1 == y0:
            continue
        e = _radius_dense(x[x0:x1], y[y0:y1], r, batch_x[x0:x1], batch_y[y0:y1], max_num_neighbors)
        out.append(torch.stack([e[0] + y0, e[1] + x0]))
    return torch.cat(out, dim=1) if out else torch.zeros((2, 0), dtype=torch.long, device=x.device)


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
