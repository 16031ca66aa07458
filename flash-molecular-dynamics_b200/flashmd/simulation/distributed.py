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
    """Replica exchange over sims sharded contiguously across ranks.

    CUDA tensors (NCCL): everything stays on the device and on the stream - no `.cpu()`, no Python loop over accepted
    pairs, no data-dependent host control flow:
      1. all_gather of the per-sim energies into a device buffer;
      2. `fmd_pt_decide` on every rank (same energies, same uniforms - a seeded host generator copied asynchronously, or
         counter-based Philox keyed by (seed, exchange index, pair) inside the kernel) -> accept[n_pairs] on the device;
      3. rank-local pairs: `fmd_pt_swap`;
      4. cross-rank pairs: the communication pattern is STATIC (every proposed cross-rank pair exchanges its rows, one
         packed send + recv per peer in one grouped NCCL call: <= 24 bytes x beads x local sims, microseconds over
         NVLink); the received rows are committed with a device-side select on accept[] and the
         sqrt(beta_old / beta_new) scale, so the decision never has to reach the host.
    CPU tensors (gloo; module path and tests): the same protocol with torch ops for the decision."""

    def __init__(self, betas_all: torch.Tensor, n_atoms: int, rank: int, world: int, group=None):
        self.betas_all = betas_all.detach().float().cpu()
        self.n_total = int(betas_all.numel())
        self.n_atoms = int(n_atoms)
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = shard_range(self.n_total, rank, world)
        self.per = self.hi - self.lo
        self._plans = {}
        self._betas_dev = None

    # ------------------------------------------------------------------ static plan per proposed pair set
    def _plan(self, pair_a: torch.Tensor, pair_b: torch.Tensor, device):
        key = (int(pair_a[0]), int(pair_b[0]), int(pair_a.numel()), str(device))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        a_all, b_all = pair_a.tolist(), pair_b.tolist()
        beta = self.betas_all
        loc_idx, loc_a, loc_b = [], [], []
        peers = {}
        for k, (a, b) in enumerate(zip(a_all, b_all)):
            ra, rb = a // self.per, b // self.per
            if ra == self.rank and rb == self.rank:
                loc_idx.append(k); loc_a.append(a - self.lo); loc_b.append(b - self.lo)
                continue
            s_ab = float(torch.sqrt(beta[a] / beta[b]))
            if ra == self.rank:      # my sim is `a`: it receives b's rows, v scaled by s_ab (reference :465-477)
                peers.setdefault(rb, []).append((k, a - self.lo, s_ab))
            elif rb == self.rank:    # my sim is `b`: it receives a's rows, v scaled by 1 / s_ab
                peers.setdefault(ra, []).append((k, b - self.lo, 1.0 / s_ab))
        i32, i64 = torch.int32, torch.int64
        plan = {
            "pa": pair_a.to(device=device, dtype=i32).contiguous(), "pb": pair_b.to(device=device, dtype=i32).contiguous(),
            "loc_pairs": torch.tensor(loc_idx, dtype=i64, device=device),
            "loc_a": torch.tensor(loc_a, dtype=i32, device=device), "loc_b": torch.tensor(loc_b, dtype=i32, device=device),
            "peers": [],
        }
        for peer in sorted(peers):
            rows = sorted(peers[peer])                      # pair order: both sides enumerate the same pairs
            plan["peers"].append({
                "peer": peer,
                "pair": torch.tensor([r[0] for r in rows], dtype=i64, device=device),
                "sim": torch.tensor([r[1] for r in rows], dtype=i64, device=device),
                "scale": torch.tensor([r[2] for r in rows], dtype=torch.float32, device=device)[:, None, None],
            })
        self._plans[key] = plan
        return plan

    def gather_energies(self, energy_local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return energy_local
        out = torch.empty(self.n_total, dtype=energy_local.dtype, device=energy_local.device)
        dist.all_gather_into_tensor(out, energy_local.contiguous(), group=self.group)
        return out

    def decide(self, energies_all: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor,
               uniforms: Optional[torch.Tensor], seed: int = 0, exchange_index: int = 0) -> torch.Tensor:
        """accept[n_pairs] (bool on CPU tensors, int32 on CUDA tensors), identical on every rank."""
        if energies_all.is_cuda:
            from .. import _lib as L
            dev = energies_all.device
            plan = self._plan(pair_a, pair_b, dev)
            if self._betas_dev is None or self._betas_dev.device != dev:
                self._betas_dev = self.betas_all.to(dev).contiguous()
            acc = torch.empty(pair_a.numel(), dtype=torch.int32, device=dev)
            uni = None if uniforms is None else uniforms.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            L.call("fmd_pt_decide", L.ptr(energies_all.float().contiguous()), L.ptr(self._betas_dev), L.ptr(plan["pa"]),
                   L.ptr(plan["pb"]), pair_a.numel(), L.ptr(uni), int(seed) & ((1 << 64) - 1), int(exchange_index),
                   L.ptr(acc), L.stream_ptr())
            return acc
        e = energies_all.detach().float()
        p = torch.exp((e[pair_a] - e[pair_b]) * (self.betas_all[pair_a] - self.betas_all[pair_b]))
        return uniforms < p

    def swap(self, pos: torch.Tensor, vel: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor,
             accepted: torch.Tensor) -> None:
        """In-place swap of the accepted pairs' rows in the LOCAL pos/vel [per * n_atoms, 3]; `accepted` stays on the
        device of pos (no host read)."""
        n, dev = self.n_atoms, pos.device
        plan = self._plan(pair_a, pair_b, dev)
        acc = accepted.to(device=dev)
        x = pos.view(self.per, n, 3)
        v = vel.view(self.per, n, 3)
        # ---- cross-rank pairs first (they read the pre-swap rows): static pattern, one grouped send/recv
        ops, recvs = [], []
        for pp in plan["peers"]:
            send = torch.stack([x.index_select(0, pp["sim"]), v.index_select(0, pp["sim"])], dim=1).contiguous()  # [k,2,n,3]
            recv = torch.empty_like(send)
            recvs.append(recv)
            ops.append(dist.P2POp(dist.isend, send, pp["peer"], group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv, pp["peer"], group=self.group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        # ---- rank-local pairs
        if plan["loc_pairs"].numel() > 0:
            acc_loc = acc.index_select(0, plan["loc_pairs"])
            if pos.is_cuda:
                from .. import _lib as L
                if self._betas_dev is None or self._betas_dev.device != dev:
                    self._betas_dev = self.betas_all.to(dev).contiguous()
                beta_loc = self._betas_dev[self.lo:self.hi].contiguous()
                L.call("fmd_pt_swap", L.ptr(pos), L.ptr(vel), L.ptr(beta_loc), L.ptr(plan["loc_a"]), L.ptr(plan["loc_b"]),
                       L.ptr(acc_loc.to(torch.int32).contiguous()), int(plan["loc_pairs"].numel()), n, L.stream_ptr())
            else:
                ia, ib = plan["loc_a"].long(), plan["loc_b"].long()
                m = acc_loc.bool()[:, None, None]
                sab = torch.sqrt(self.betas_all[self.lo:self.hi][ia] / self.betas_all[self.lo:self.hi][ib]).to(vel.dtype)[:, None, None]
                xa, xb, va, vb = x[ia].clone(), x[ib].clone(), v[ia].clone(), v[ib].clone()
                x[ia], x[ib] = torch.where(m, xb, xa), torch.where(m, xa, xb)
                v[ia], v[ib] = torch.where(m, vb * sab, va), torch.where(m, va / sab, vb)
        for req in reqs:
            req.wait()
        # ---- commit the received rows where the pair was accepted (device-side select)
        for pp, recv in zip(plan["peers"], recvs):
            m = acc.index_select(0, pp["pair"]).bool()[:, None, None]
            x.index_copy_(0, pp["sim"], torch.where(m, recv[:, 0], x.index_select(0, pp["sim"])))
            v.index_copy_(0, pp["sim"], torch.where(m, recv[:, 1] * pp["scale"].to(vel.dtype), v.index_select(0, pp["sim"])))

    def exchange(self, pos, vel, energy_local, pair_a, pair_b, uniforms, seed: int = 0, exchange_index: int = 0) -> torch.Tensor:
        accepted = self.decide(self.gather_energies(energy_local), pair_a, pair_b, uniforms, seed, exchange_index)
        self.swap(pos, vel, pair_a, pair_b, accepted)
        return accepted
