"""Multi-GPU plumbing: one process per GPU (torchrun), replicas sharded contiguously over the ranks.

Plain Langevin batches need no communication at all (independent molecules, edges never cross molecules).
Parallel tempering needs one exchange step every `exchange_interval` steps (reference
simulation/parallel_tempering.py:368-481 does it on one device with a host round trip):

  1. all_gather of the per-sim potential energies (n_sims * 4 bytes);
  2. every rank evaluates the SAME Metropolis decisions from the same uniforms (a seeded generator keyed by
     the exchange index: no broadcast needed);
  3. accepted pairs swap positions and sqrt(beta_old/beta_new)-rescaled velocities: rank-local pairs in
     place, cross-rank pairs with ONE packed send + recv per peer (NCCL over NVLink on GPUs, gloo on CPU).

Works on CPU tensors (gloo) and CUDA tensors (NCCL) alike; the data moved is 24 bytes per bead per accepted
pair, i.e. latency-bound."""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def dist_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` for `rank` (requires divisibility: equal work per GPU)."""
    if n_items % world != 0:
        raise ValueError(f"{n_items} simulations cannot be split evenly over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per


def shard_configurations(configurations: List, rank: Optional[int] = None, world: Optional[int] = None) -> List:
    r, w = dist_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    lo, hi = shard_range(len(configurations), rank, world)
    return configurations[lo:hi]


def exchange_uniforms(seed: int, exchange_index: int, n_pairs: int) -> torch.Tensor:
    """Identical on every rank without communication."""
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(exchange_index)) & 0x7FFFFFFFFFFFFFFF)
    return torch.rand(n_pairs, generator=g)


class ShardedExchange:
    """Replica exchange over sims sharded contiguously across ranks."""

    def __init__(self, betas_all: torch.Tensor, n_atoms: int, rank: int, world: int, group=None):
        self.betas_all = betas_all.detach().float().cpu()
        self.n_total = int(betas_all.numel())
        self.n_atoms = int(n_atoms)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = shard_range(self.n_total, rank, world)
        self.per = self.hi - self.lo

    def gather_energies(self, energy_local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return energy_local
        out = torch.empty(self.n_total, dtype=energy_local.dtype, device=energy_local.device)
        dist.all_gather_into_tensor(out, energy_local.contiguous(), group=self.group)
        return out

    def decide(self, energies_all: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor,
               uniforms: torch.Tensor) -> torch.Tensor:
        e = energies_all.detach().float().cpu()
        p = torch.exp((e[pair_a] - e[pair_b]) * (self.betas_all[pair_a] - self.betas_all[pair_b]))
        return uniforms < p

    def swap(self, pos: torch.Tensor, vel: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor,
             accepted: torch.Tensor) -> None:
        """In-place swap of the accepted pairs' rows in the LOCAL pos/vel [per * n_atoms, 3]."""
        n, dev = self.n_atoms, pos.device
        x = pos.view(self.per, n, 3)
        v = vel.view(self.per, n, 3)
        a_all, b_all = pair_a[accepted].tolist(), pair_b[accepted].tolist()
        # (local sim, partner sim, partner rank) for every accepted pair that touches this rank
        mine = []
        for a, b in zip(a_all, b_all):
            if self.lo <= a < self.hi:
                mine.append((a, b, b // self.per))
            if self.lo <= b < self.hi:
                mine.append((b, a, a // self.per))
        if not mine:
            return
        scale = lambda new, old: float(torch.sqrt(self.betas_all[new] / self.betas_all[old]))  # noqa: E731
        # ---- rank-local pairs: both ends here (each pair appears twice in `mine`; handle it once)
        loc = [(s, p) for s, p, r in mine if r == self.rank and s < p]
        if loc:
            ia = torch.tensor([s - self.lo for s, _ in loc], device=dev)
            ib = torch.tensor([p - self.lo for _, p in loc], device=dev)
            sab = torch.tensor([scale(s, p) for s, p in loc], device=dev, dtype=vel.dtype)[:, None, None]
            xa, xb, va, vb = x[ia].clone(), x[ib].clone(), v[ia].clone(), v[ib].clone()
            x[ia], x[ib] = xb, xa
            v[ia], v[ib] = vb * sab, va / sab
        # ---- cross-rank pairs: one packed send + recv per peer, rows ordered by the (global) pair order
        peers = sorted({r for _, _, r in mine if r != self.rank})
        if not peers:
            return
        ops, recv_bufs, meta = [], {}, {}
        for peer in peers:
            rows = [(s, p) for s, p, r in mine if r == peer]
            rows.sort(key=lambda sp: (min(sp), max(sp)))
            idx = torch.tensor([s - self.lo for s, _ in rows], device=dev)
            send = torch.stack([x[idx], v[idx]], dim=1).contiguous()           # [k, 2, n, 3]
            recv = torch.empty_like(send)
            recv_bufs[peer], meta[peer] = recv, (idx, rows)
            ops.append(dist.P2POp(dist.isend, send, peer, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv, peer, group=self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for peer in peers:
            idx, rows = meta[peer]
            recv = recv_bufs[peer]
            sc = torch.tensor([scale(s, p) for s, p in rows], device=dev, dtype=vel.dtype)[:, None, None]
            x[idx] = recv[:, 0]
            v[idx] = recv[:, 1] * sc

    def exchange(self, pos, vel, energy_local, pair_a, pair_b, uniforms) -> torch.Tensor:
        accepted = self.decide(self.gather_energies(energy_local), pair_a, pair_b, uniforms)
        self.swap(pos, vel, pair_a, pair_b, accepted)
        return accepted
