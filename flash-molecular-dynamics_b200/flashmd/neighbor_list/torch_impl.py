"""Per-step radius graph on the GPU (replaces reference neighbor_list/torch_impl.py:175-226, which
delegates to the un-vendored torch_cluster.radius_graph CUDA kernel)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import _lib as L


def radius_graph_csr(pos: torch.Tensor, mol_ptr: torch.Tensor, rcut: float, max_num_neighbors: int = 1000,
                     idx_dtype=torch.int64, capacity: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Radius graph of a batch of molecules (nodes of a molecule contiguous, `mol_ptr` [B+1]).

    Returns edge_index [2,E] in torch_cluster's order (centre ascending, neighbour ascending;
    strict fp32 d^2 < r^2; no self loops; flow="target_to_source"), `src_ptr` [N+1] (CSR over the
    centre), `rev` [E] (index of the reverse edge, == csr_perm of the dst-major CSR when the list
    is symmetric) and `dist` [E].  One host sync (edge count) unless `capacity` is given.
    """
    if not pos.is_cuda:
        raise RuntimeError("radius_graph_csr needs CUDA tensors (no CPU fallback)")
    pos = pos.contiguous().float()
    N = pos.shape[0]
    B = mol_ptr.numel() - 1
    mp = mol_ptr.to(torch.int32).contiguous()
    dev = pos.device
    st = L.stream_ptr()
    sizes = mp[1:] - mp[:-1]
    max_mol = int(sizes.max().item()) if B > 0 else 0
    deg = torch.empty(N, dtype=torch.int32, device=dev)
    seg = torch.zeros(N + 1, dtype=torch.int32, device=dev)
    ws = torch.empty(N // 1024 + 4, dtype=torch.int32, device=dev)
    L.call("fmd_nl_count", L.ptr(pos), L.ptr(mp), B, N, max_mol, float(rcut), int(max_num_neighbors), L.ptr(deg), st)
    L.call("fmd_exclusive_scan_i32", L.ptr(deg), L.ptr(seg), N, L.ptr(ws), st)
    E = int(seg[N].item()) if capacity is None else int(capacity)
    src = torch.empty(E, dtype=idx_dtype, device=dev)
    dst = torch.empty(E, dtype=idx_dtype, device=dev)
    rev = torch.empty(E, dtype=idx_dtype, device=dev)
    dist = torch.empty(E, dtype=torch.float32, device=dev)
    if E > 0 and N > 0:
        ib = L.idx_bytes(src)
        L.call("fmd_nl_fill", L.ptr(pos), L.ptr(mp), B, N, max_mol, float(rcut), int(max_num_neighbors), L.ptr(seg), E,
               L.ptr(src), L.ptr(dst), ib, L.ptr(dist), st)
        L.call("fmd_nl_reverse", L.ptr(seg), L.ptr(src), L.ptr(dst), ib, N, E, L.ptr(rev), st)
    return {"edge_index": torch.stack([src, dst]), "src_ptr": seg.to(idx_dtype), "rev": rev, "dist": dist}


def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target", num_workers: int = 1) -> torch.Tensor:
    """Drop-in for torch_cluster.radius_graph on CUDA tensors (sorted `batch`)."""
    assert flow in ("source_to_target", "target_to_source")
    if loop:
        raise NotImplementedError("loop=True is not supported (the reference always passes loop=False)")
    N = x.shape[0]
    if batch is None:
        ptr = torch.tensor([0, N], device=x.device)
    else:
        counts = torch.bincount(batch, minlength=int(batch[-1].item()) + 1 if N > 0 else 0)
        ptr = torch.zeros(counts.numel() + 1, dtype=torch.long, device=x.device)
        ptr[1:] = torch.cumsum(counts, 0)
    ei = radius_graph_csr(x, ptr, r, max_num_neighbors)["edge_index"]
    return ei if flow == "target_to_source" else ei.flip(0)


def _radius_graph_torch(pos: torch.Tensor, ptr: torch.Tensor, rcut: float, max_num_neighbors: int) -> torch.Tensor:
    """Plain-PyTorch radius graph with the same semantics / order (the CPU `--disable_optim` path)."""
    out = []
    r2 = torch.tensor(float(rcut), dtype=torch.float32) ** 2
    for b in range(ptr.numel() - 1):
        lo, hi = int(ptr[b]), int(ptr[b + 1])
        x = pos[lo:hi].float()
        d = x[:, None, :] - x[None, :, :]
        d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
        hit = d2 < r2
        if max_num_neighbors + 1 < hi - lo:
            keep = torch.cumsum(hit.to(torch.int32), dim=1) <= max_num_neighbors + 1
            hit = hit & keep
        hit.fill_diagonal_(False)
        ij = hit.nonzero().t()
        out.append(ij + lo)
    return torch.cat(out, dim=1) if out else torch.zeros((2, 0), dtype=torch.long)


def torch_neighbor_list(data, rcut: float, self_interaction: bool = False, num_workers: int = 1,
                        max_num_neighbors: int = 1000) -> torch.Tensor:
    """edge_index [2,E] of a collated batch (reference torch_impl.py:175-226, no-PBC branch; `num_workers` is accepted
    for signature compatibility and ignored)."""
    if self_interaction:
        raise NotImplementedError("self_interaction=True is not supported (the reference never passes it)")
    pos = data.pos
    ptr = data["ptr"] if "ptr" in data else torch.tensor([0, pos.shape[0]], device=pos.device)
    if pos.is_cuda:
        return radius_graph_csr(pos, ptr, rcut, max_num_neighbors)["edge_index"]
    return _radius_graph_torch(pos, ptr, rcut, max_num_neighbors)


def torch_neighbor_list_no_pbc(data, rcut: float, self_interaction: bool = False, num_workers: int = 1,
                               max_num_neighbors: int = 1000) -> torch.Tensor:
    """Reference torch_impl.py:175-226 by its own name."""
    return torch_neighbor_list(data, rcut, self_interaction, num_workers, max_num_neighbors)


def torch_neighbor_list_pbc(data, rcut: float, self_interaction: bool = False, num_workers: int = 1,
                            max_num_neighbors: int = 1000):
    raise NotImplementedError("periodic neighbour lists are out of scope (coarse-grained proteins are simulated without a "
                              "cell; reference torch_impl.py:252-329)")
