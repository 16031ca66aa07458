"""Parallel-tempering Langevin dynamics (reference simulation/parallel_tempering.py:74-517).

Replica layout as in the reference: sim index = beta_index * n_indep + r; every `exchange_interval` steps
adjacent-beta pairs (even / odd sets alternating) are proposed, accepted with min(1, exp((U_a-U_b)(b_a-b_b)))
and accepted pairs swap positions and sqrt(beta_old/beta_new)-rescaled velocities.

Execution: module path (CPU) = the reference's torch code path; fused path (1 GPU) = fmd_pt_decide +
fmd_pt_swap on the engine's buffers, no host round trip except the tiny acceptance bookkeeping;
multi-GPU (torch.distributed initialised, sims sharded contiguously over ranks) = all_gather of the
energies over NCCL, identical decisions on every rank, peer exchange of the accepted rows
(simulation/distributed.py)."""
from copy import deepcopy
from typing import Any, Dict, List

import numpy as np
import torch

from ..data._keys import ENERGY_KEY, MASS_KEY, POSITIONS_KEY, VELOCITY_KEY
from .langevin import LangevinSimulation


def adjacent_pairs(n_replicas: int, n_indep: int):
    """(even_a, even_b), (odd_a, odd_b) index tensors (reference :256-284)."""
    def build(starts):
        a = [torch.arange(n_indep) + s * n_indep for s in starts]
        b = [torch.arange(n_indep) + (s + 1) * n_indep for s in starts]
        return [torch.cat(a), torch.cat(b)]
    even = list(range(0, n_replicas - 1, 2))
    odd = list(range(1, n_replicas - 1, 2)) or even
    return build(even), build(odd)


class PTSimulation(LangevinSimulation):
    def __init__(self, friction: float = 1e-3, exchange_interval: int = 100, exchange_rng: str = "torch", **kwargs: Any):
        if kwargs.get("sim_subroutine") is not None:
            raise ValueError("PTSimulation installs its own sim_subroutine (the replica exchange)")
        kwargs["sim_subroutine"] = self.detect_and_exchange_replicas
        kwargs["sim_subroutine_interval"] = exchange_interval
        kwargs.setdefault("save_subroutine", self.save_exchanges)
        super().__init__(friction=friction, **kwargs)
        self.exchange_interval = exchange_interval
        self.exchange_rng = exchange_rng          # "torch": CPU global RNG like the reference; "philox": on device
        self._replica_exchange_attempts = 0
        self._replica_exchange_approved = 0
        self._exchange_index = 0
        self._pending = []

    def attach_model_and_configurations(self, model, configurations, betas: List[float], **kw):
        super().attach_model_and_configurations(model, configurations, betas)

    def _attach_configurations(self, configurations: List, betas: List[float]):
        if not isinstance(betas, list):
            raise ValueError(f"Parallel tempering requires multiple temperatures, but only {betas} was supplied.")
        if not all(b >= 0 and np.isfinite(b) for b in betas):
            raise ValueError(f"All betas must be positive, but {betas} contains an illegal value.")
        if list(betas) != sorted(betas, reverse=True):
            raise ValueError("Betas must be in order of increasing temperature.")
        self.n_indep_sims, self.n_replicas = len(configurations), len(betas)
        extended = [deepcopy(c) for _ in betas for c in configurations]
        ext_betas = [b for b in betas for _ in configurations]
        # multi-GPU: the replicas are sharded contiguously over the ranks (simulation/distributed.py)
        from .distributed import ShardedExchange, dist_info, shard_range
        rank, world = dist_info()
        self._sharded = None
        if world > 1:
            lo, hi = shard_range(len(extended), rank, world)
            self._sharded = ShardedExchange(torch.tensor(ext_betas), len(configurations[0].atom_types), rank, world)
            extended, ext_betas = extended[lo:hi], ext_betas[lo:hi]
            self._node_offset = lo * len(configurations[0].atom_types)     # Philox noise keyed by the global bead index
            if self.filename is not None and not self.filename.endswith(f"_rank{rank}"):
                self.filename = f"{self.filename}_rank{rank}"
        super()._attach_configurations(extended, ext_betas)
        self._propose_even_pairs = True
        self._even_pairs, self._odd_pairs = adjacent_pairs(self.n_replicas, self.n_indep_sims)
        self.pair_to_beta_idx = torch.arange(self.n_replicas).repeat_interleave(self.n_indep_sims)
        self.acceptance_matrix = torch.zeros(self.n_replicas, self.n_replicas)
        return extended

    # ------------------------------------------------------------------ proposal bookkeeping
    def _get_proposed_pairs(self):
        pairs = self._even_pairs if self._propose_even_pairs else self._odd_pairs
        self._propose_even_pairs = not self._propose_even_pairs
        return pairs[0], pairs[1]

    def _record(self, pair_a, pair_b, approved: torch.Tensor):
        if approved.is_cuda:
            # bookkeeping only: resolved at the next save / summary so that an exchange never waits for the GPU
            self._pending.append((pair_a, pair_b, approved))
            return
        self._record_now(pair_a, pair_b, approved)

    def _flush_records(self):
        for pa, pb, acc in self._pending:
            self._record_now(pa, pb, acc)
        self._pending = []

    def _record_now(self, pair_a, pair_b, approved: torch.Tensor):
        approved = approved.cpu().bool()
        self._replica_exchange_approved += int(approved.sum())
        self._replica_exchange_attempts += len(pair_a)
        ia, ib = self.pair_to_beta_idx[pair_a], self.pair_to_beta_idx[pair_b]
        beta_pairs = torch.unique(torch.stack((ia, ib)), dim=1)
        per_beta = approved.reshape(beta_pairs.shape[1], self.n_indep_sims).sum(dim=1).to(self.acceptance_matrix.dtype)
        self.acceptance_matrix[beta_pairs[0], beta_pairs[1]] += per_beta
        self.acceptance_matrix[beta_pairs[1], beta_pairs[0]] += self.n_indep_sims - per_beta

    def _uniforms(self, n: int) -> torch.Tensor:
        return torch.rand(n)   # CPU global RNG, as the reference (:394)

    # ------------------------------------------------------------------ module path
    def _detect_exchange(self, data) -> Dict:
        pair_a, pair_b = self._get_proposed_pairs()
        if self._sharded is not None:
            # sharded module path (CPU / gloo, fp64, ...): pair indices are GLOBAL, the data is local -> the same
            # all-gather + identical-decision + peer-swap protocol as the fused path
            from .distributed import exchange_uniforms
            uni = exchange_uniforms(self.random_seed or 0, self._exchange_index, len(pair_a))
            x = data[POSITIONS_KEY].detach().clone().contiguous()
            v = data[VELOCITY_KEY].detach().clone().contiguous()
            acc = self._sharded.exchange(x, v, data.out[ENERGY_KEY].detach(), pair_a, pair_b, uni,
                                         self.random_seed or 0, self._exchange_index)
            self._exchange_index += 1
            self._record(pair_a, pair_b, acc)
            data[POSITIONS_KEY], data[VELOCITY_KEY] = x, v
            return {"a": pair_a[:0], "b": pair_b[:0]}
        u = data.out[ENERGY_KEY]
        beta = self.beta
        p = torch.exp((u[pair_a] - u[pair_b]) * (beta[pair_a] - beta[pair_b])).cpu()
        approved = self._uniforms(len(p)) < p
        self._record(pair_a, pair_b, approved)
        return {"a": pair_a[approved], "b": pair_b[approved]}

    def _perform_exchange(self, data, pairs: Dict):
        a, b = pairs["a"], pairs["b"]
        if len(a) == 0:
            return data
        n = self.n_atoms
        x = data[POSITIONS_KEY].detach().clone().view(self.n_sims, n, -1)
        v = data[VELOCITY_KEY].detach().clone().view(self.n_sims, n, -1)
        xa, xb, va, vb = x[a].clone(), x[b].clone(), v[a].clone(), v[b].clone()
        s_ab = torch.sqrt(self.beta[a] / self.beta[b])[:, None, None]
        x[a], x[b] = xb, xa
        v[a], v[b] = vb * s_ab, va / s_ab
        data[POSITIONS_KEY], data[VELOCITY_KEY] = x.view(-1, x.shape[-1]), v.view(-1, v.shape[-1])
        return data

    def detect_and_exchange_replicas(self, data):
        return self._perform_exchange(data, self._detect_exchange(data))

    # ------------------------------------------------------------------ fused path
    def _engine_subroutine(self, eng):
        from .. import _lib as L
        pair_a, pair_b = self._get_proposed_pairs()
        if self._sharded is not None:
            from .distributed import exchange_uniforms
            uni = None if self.exchange_rng == "philox" else exchange_uniforms(self.random_seed or 0, self._exchange_index, len(pair_a))
            acc = self._sharded.exchange(eng.pos, eng.vel, eng.ff.energy, pair_a, pair_b, uni, self.random_seed or 0,
                                         self._exchange_index)
            self._exchange_index += 1
            self._record(pair_a, pair_b, acc)
            return
        dev = eng.pos.device
        pa = pair_a.to(dev, torch.int32).contiguous()
        pb = pair_b.to(dev, torch.int32).contiguous()
        acc = torch.zeros(len(pa), dtype=torch.int32, device=dev)
        st = L.stream_ptr()
        uni = None if self.exchange_rng == "philox" else self._uniforms(len(pa)).to(dev)
        seed = (self.random_seed or 0) & ((1 << 64) - 1)
        L.call("fmd_pt_decide", L.ptr(eng.ff.energy), L.ptr(eng.beta), L.ptr(pa), L.ptr(pb), len(pa), L.ptr(uni), seed,
               self._exchange_index, L.ptr(acc), st)
        L.call("fmd_pt_swap", L.ptr(eng.pos), L.ptr(eng.vel), L.ptr(eng.beta), L.ptr(pa), L.ptr(pb), L.ptr(acc), len(pa),
               self.n_atoms, st)
        self._exchange_index += 1
        self._record(pair_a, pair_b, acc)
        # like the reference, the forces carried into the next half-kick are the pre-swap forces

    # ------------------------------------------------------------------ output
    def save_exchanges(self, data, save_step: int) -> None:
        self._flush_records()
        if self.filename is None:
            return
        np.save(f"{self.filename}_acceptance_{self._get_numpy_count()}.npy", self.acceptance_matrix.cpu().numpy())
        self.acceptance_matrix = torch.zeros(self.n_replicas, self.n_replicas)

    def get_replica_info(self, replica_num: int = 0) -> Dict:
        return {"beta": float(self.beta[replica_num * self.n_indep_sims]),
                "indices_in_the_output": list(range(replica_num * self.n_indep_sims, (replica_num + 1) * self.n_indep_sims))}

    def summary(self):
        self._flush_records()
        att = max(self._replica_exchange_attempts, 1)
        self.exchange_summary = {"attempted": self._replica_exchange_attempts,
                                 "approved": self._replica_exchange_approved,
                                 "ratio": self._replica_exchange_approved / att}
