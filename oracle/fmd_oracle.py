"""CPU oracle for the CGSchNet force-field + Langevin step.  TEST INFRASTRUCTURE ONLY.

This file is a plain restatement (numpy / torch-CPU, autograd for the forces) of the reference
algorithm; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import it.  The product path (`flash-molecular-dynamics_b200/`) never does.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (/root/reference/src/flashmd,
behind the import stand-ins in `oracle/shims/`) on CPU and stores inputs+outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against them.
Parts that cannot be pinned that way are marked "parity unpinned":
  * `radius_graph` — the reference delegates to the un-vendored `torch_cluster` CUDA kernel
    (neighbor_list/torch_impl.py:216-224; pyproject.toml:19, no version pin); restated from its
    published algorithm.
  * the W16A16 rounding model (`precision="w16a16"`) — the reference's FP16 Triton kernels
    (kernels/cfconv_kernels.py:644-720, 896-952) only run on a GPU.

All paths cited are relative to /root/reference/src/flashmd/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# neighbour list + CSR  (integer work: bit-exact bar)
# --------------------------------------------------------------------------------------------


def pair_dist2_f32(pos_c: np.ndarray, pos_n: np.ndarray) -> np.ndarray:
    """fp32 squared distance with the accumulation order of torch_cluster's CUDA radius kernel:
    dist = 0; for d in x,y,z: dist += (x_n[d]-x_c[d])^2, which nvcc contracts to
    fma(dz,dz, fma(dy,dy, dx*dx)).  The FMA is emulated through float64 (products of two fp32
    are exact in fp64)."""
    d = pos_n.astype(np.float32) - pos_c.astype(np.float32)
    dx, dy, dz = (d[..., k].astype(np.float64) for k in range(3))
    acc = (dx * dx).astype(np.float32)
    acc = (dy * dy + acc.astype(np.float64)).astype(np.float32)
    acc = (dz * dz + acc.astype(np.float64)).astype(np.float32)
    return acc


def radius_graph(pos: np.ndarray, ptr: np.ndarray, rc: float, max_num_neighbors: int = 1000) -> np.ndarray:
    """edge_index [2,E] int64: row 0 = centre ("src"), ascending; row 1 = neighbour ("dst"),
    ascending inside a centre; strict fp32 d^2 < rc^2; same molecule only; at most
    max_num_neighbors+1 hits (self included) are kept per centre before the self pair is
    dropped.  Follows neighbor_list/torch_impl.py:175-226 (flow="target_to_source", loop=False)
    -> torch_cluster.radius_graph / radius CUDA kernel.  parity unpinned (see module docstring)."""
    pos = np.asarray(pos, dtype=np.float32)
    r2 = np.float32(np.float32(rc) * np.float32(rc))
    src_all, dst_all = [], []
    for b in range(len(ptr) - 1):
        lo, hi = int(ptr[b]), int(ptr[b + 1])
        p = pos[lo:hi]
        d2 = pair_dist2_f32(p[:, None, :], p[None, :, :])
        hit = d2 < r2
        rank = np.cumsum(hit, axis=1)
        hit &= rank <= (max_num_neighbors + 1)
        c, n = np.nonzero(hit)
        keep = c != n
        src_all.append(c[keep] + lo)
        dst_all.append(n[keep] + lo)
    if not src_all:
        return np.zeros((2, 0), dtype=np.int64)
    return np.stack([np.concatenate(src_all), np.concatenate(dst_all)]).astype(np.int64)


def build_csr(keys: np.ndarray, num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """(ptr [N+1], perm [E]) of a counting sort by `keys`.  kernels/csr_kernels.py:88-169 and
    :229-294 produce `ptr` deterministically and `perm` in a run-dependent order inside each
    segment (atomic cursors, :80-85); the canonical form used for parity is the STABLE sort."""
    keys = np.asarray(keys, dtype=np.int64)
    counts = np.bincount(keys, minlength=num_nodes)[:num_nodes]
    ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    perm = np.argsort(keys, kind="stable").astype(np.int64)
    return ptr, perm


def reverse_edge_index(edge_index: np.ndarray, num_nodes: int) -> np.ndarray:
    """rev[e] = index of the edge (dst_e -> src_e); -1 when it does not exist."""
    src, dst = edge_index
    key = src * num_nodes + dst
    rkey = dst * num_nodes + src
    order = np.argsort(key, kind="stable")
    pos = np.searchsorted(key[order], rkey)
    pos = np.clip(pos, 0, max(len(key) - 1, 0))
    rev = order[pos] if len(key) else np.zeros(0, dtype=np.int64)
    ok = key[rev] == rkey if len(key) else np.zeros(0, dtype=bool)
    return np.where(ok, rev, -1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# SchNet pieces (floating point: tolerance bar)
# --------------------------------------------------------------------------------------------


def cosine_cutoff(d: torch.Tensor, rc: float) -> torch.Tensor:
    """models/cutoff.py:137-145 (cutoff_lower == 0 branch)."""
    return 0.5 * (torch.cos(d * math.pi / rc) + 1.0) * (d < rc).to(d.dtype)


def gaussian_rbf(d: torch.Tensor, centers: torch.Tensor, gamma: float, rc: float) -> torch.Tensor:
    """models/radial_basis/gaussian.py:83-102 with CosineCutoff(0, rc) as the basis cutoff;
    identical to the fused kernel kernels/cfconv_kernels.py:1545-1575."""
    return torch.exp(gamma * (d[:, None] - centers[None, :]) ** 2) * cosine_cutoff(d, rc)[:, None]


def rbf_params(rc: float, num_rbf: int, dtype=torch.float32, lower: float = 0.0):
    """models/radial_basis/gaussian.py:64-75: centres = linspace(lo, hi, R) (fp32), gamma = -0.5/(c1-c0)^2."""
    centers = torch.linspace(lower, rc, num_rbf)
    gamma = float(-0.5 / (centers[1] - centers[0]) ** 2)
    return centers.to(dtype), gamma


def edge_distances(pos: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """geometry/internal_coordinates.py:96-101: ||pos[dst] - pos[src]||_2."""
    return (pos[edge_index[1]] - pos[edge_index[0]]).norm(p=2, dim=1)


def _tanh_w16(x: torch.Tensor) -> torch.Tensor:
    """kernels/cfconv_kernels.py:449-454 (_triton_tanh): clamp to +-10, (e^{2x}-1)/(e^{2x}+1)."""
    xc = x.clamp(-10.0, 10.0)
    e = torch.exp(2.0 * xc)
    return (e - 1.0) / (e + 1.0)


def _h(x: torch.Tensor) -> torch.Tensor:
    """round to fp16 and come back (value-preserving cast used to model fp16 storage)."""
    return x.to(torch.float16).to(x.dtype)


class _RoundF16(torch.autograd.Function):
    """fp16 rounding with a straight-through gradient that is itself rounded to fp16 when
    `round_grad` (models the fp16 gradient tensors of kernels/cfconv_kernels.py:1023,1208,252)."""

    @staticmethod
    def forward(ctx, x, round_grad):
        ctx.round_grad = round_grad
        return _h(x)

    @staticmethod
    def backward(ctx, g):
        return (_h(g) if ctx.round_grad else g), None


def rh(x, round_grad=False):
    return _RoundF16.apply(x, round_grad)


class SchNetParams:
    """Flat weight container (all torch tensors, `nn.Linear` layout [out, in]).

    keys: embedding [T,F]; per block l: lin1_w [F,F]; f0_w [F,R], f0_b [F]; f1_w [F,F];
    lin2_w [F,F], lin2_b [F]; lin_w [F,F], lin_b [F]; output MLP: out_w[i], out_b[i] (last bias None).
    Mirrors StandardSchNet (models/schnet.py:778-841).
    """

    def __init__(self, tensors: Dict[str, torch.Tensor], num_blocks: int, num_out_layers: int,
                 cutoff: float, num_rbf: int):
        self.t = tensors
        self.num_blocks = num_blocks
        self.num_out_layers = num_out_layers
        self.cutoff = float(cutoff)
        self.num_rbf = int(num_rbf)

    def to(self, dtype):
        return SchNetParams({k: (v.to(dtype) if v is not None else None) for k, v in self.t.items()},
                            self.num_blocks, self.num_out_layers, self.cutoff, self.num_rbf)


def random_schnet_params(seed: int, cutoff: float, num_rbf: int = 50, hidden: int = 128, filters: int = 128,
                         num_blocks: int = 3, out_widths=(128, 64), embedding_size: int = 25,
                         bias_scale: float = 0.0) -> SchNetParams:
    """Xavier-uniform weights, zero (or small random) biases — models/_module_init.py:4-28.
    Independent generator (numpy) so that tests do not depend on torch's RNG stream."""
    rng = np.random.default_rng(seed)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return torch.from_numpy(rng.uniform(-a, a, size=(o, i)).astype(np.float32))

    def bias(o):
        return torch.from_numpy((bias_scale * rng.standard_normal(o)).astype(np.float32))

    t = {"embedding": torch.from_numpy(rng.standard_normal((embedding_size, hidden)).astype(np.float32))}
    for l in range(num_blocks):
        t[f"b{l}.lin1_w"] = xavier(filters, hidden)
        t[f"b{l}.f0_w"] = xavier(filters, num_rbf)
        t[f"b{l}.f0_b"] = bias(filters)
        t[f"b{l}.f1_w"] = xavier(filters, filters)
        t[f"b{l}.lin2_w"] = xavier(hidden, filters)
        t[f"b{l}.lin2_b"] = bias(hidden)
        t[f"b{l}.lin_w"] = xavier(hidden, hidden)
        t[f"b{l}.lin_b"] = bias(hidden)
    widths = [hidden] + list(out_widths) + [1]
    for i in range(len(widths) - 1):
        t[f"out{i}_w"] = xavier(widths[i + 1], widths[i])
        t[f"out{i}_b"] = bias(widths[i + 1]) if i < len(widths) - 2 else None
    return SchNetParams(t, num_blocks, len(widths) - 1, cutoff, num_rbf)


def schnet_energy(P: SchNetParams, pos: torch.Tensor, atom_types: torch.Tensor, batch: torch.Tensor,
                  n_mols: int, edge_index: torch.Tensor, precision: str = "fp32",
                  drop_cutoff_grad: bool = False, return_intermediates: bool = False):
    """Energy per molecule [B].  Restates SchNet.forward (models/schnet.py:177-369),
    InteractionBlock.forward (:497-548), CFConv.forward PyTorch path (:706-719), MLP (mlp.py:41-57).

    precision:
      "fp32"/"fp64": everything in pos.dtype (the reference's `--disable_optim` path).
      "w16a16": fp16 rounding points of the default GPU path — filter network
         (models/gptq.py:92-130: x->fp16, fp16 weights/bias, fp32 accumulate, _triton_tanh, t and W
         stored fp16), output network (gptq.py:266-306: two fused linear+tanh layers stored fp16,
         last layer fp16 in -> fp32 out); CFConv math, lin1/lin2/lin in fp32 (csr_kernels.py:705-717).
    drop_cutoff_grad: reproduce the Triton paths' missing d(cutoff)/d(distance) term in CFConv
      (kernels/csr_kernels.py:912 returns None for edge_weight) by detaching C(d) there.
    """
    T = P.t
    w16 = precision == "w16a16"
    src, dst = edge_index[0], edge_index[1]
    d = edge_distances(pos, edge_index)
    centers, gamma = rbf_params(P.cutoff, P.num_rbf, dtype=pos.dtype)
    rbf = gaussian_rbf(d, centers, gamma, P.cutoff)
    C = cosine_cutoff(d, P.cutoff)
    if drop_cutoff_grad:
        C = C.detach()
    h = T["embedding"][atom_types]
    inter = {"d": d, "rbf": rbf}
    for l in range(P.num_blocks):
        a = h @ T[f"b{l}.lin1_w"].t()
        if w16:
            w0, b0, w1 = _h(T[f"b{l}.f0_w"]), _h(T[f"b{l}.f0_b"]), _h(T[f"b{l}.f1_w"])
            t_ = rh(_tanh_w16(rh(rbf) @ w0.t() + b0), True)
            W = rh(t_ @ w1.t(), True)
        else:
            t_ = torch.tanh(rbf @ T[f"b{l}.f0_w"].t() + T[f"b{l}.f0_b"])
            W = t_ @ T[f"b{l}.f1_w"].t()
        msg = a[src] * W * C[:, None]
        m = torch.zeros_like(a).index_add(0, dst, msg)
        c = m @ T[f"b{l}.lin2_w"].t() + T[f"b{l}.lin2_b"]
        u = torch.tanh(c) @ T[f"b{l}.lin_w"].t() + T[f"b{l}.lin_b"]
        h = h + u
        if return_intermediates:
            inter[f"b{l}.a"], inter[f"b{l}.W"], inter[f"b{l}.m"], inter[f"b{l}.h"] = a, W, m, h
    y = h
    nl = P.num_out_layers
    for i in range(nl):
        w, b = T[f"out{i}_w"], T[f"out{i}_b"]
        if w16:
            y = rh(y, True) @ _h(w).t()
            if b is not None:
                y = y + _h(b)
            if i < nl - 1:
                y = rh(_tanh_w16(y), True)
        else:
            y = y @ w.t()
            if b is not None:
                y = y + b
            if i < nl - 1:
                y = torch.tanh(y)
    e_atom = y.flatten()
    energy = torch.zeros(n_mols, dtype=pos.dtype).index_add(0, batch, e_atom)
    if return_intermediates:
        inter["e_atom"] = e_atom
        return energy, inter
    return energy


def schnet_energy_forces(P: SchNetParams, pos, atom_types, batch, n_mols, edge_index, **kw):
    """GradientsOut.forward (models/gradients.py:227-290): forces = -d(sum E)/d(pos) by autograd."""
    pos = pos.detach().clone().requires_grad_(True)
    e = schnet_energy(P, pos, atom_types, batch, n_mols, edge_index, **kw)
    (g,) = torch.autograd.grad(e.sum(), pos)
    return e.detach(), -g


# --------------------------------------------------------------------------------------------
# priors
# --------------------------------------------------------------------------------------------


def feat_distance(pos, mapping):
    """geometry/internal_coordinates.py:73-101."""
    return (pos[mapping[1]] - pos[mapping[0]]).norm(p=2, dim=1)


def feat_angle_cos(pos, mapping):
    """geometry/internal_coordinates.py:140-170."""
    dr1 = pos[mapping[0]] - pos[mapping[1]]
    dr2 = pos[mapping[2]] - pos[mapping[1]]
    return (dr1 * dr2).sum(1) / (dr1.norm(p=2, dim=1) * dr2.norm(p=2, dim=1))


def feat_torsion(pos, mapping):
    """geometry/internal_coordinates.py:174-223 (MDTraj sign convention)."""
    n = torch.nn.functional.normalize
    dr1 = n(pos[mapping[1]] - pos[mapping[0]], dim=1)
    dr2 = n(pos[mapping[2]] - pos[mapping[1]], dim=1)
    dr3 = n(pos[mapping[3]] - pos[mapping[2]], dim=1)
    n1 = torch.cross(dr1, dr2, dim=1)
    n2 = torch.cross(dr2, dr3, dim=1)
    m1 = torch.cross(n1, dr2, dim=1)
    y = (m1 * n2).sum(-1)
    x = (n1 * n2).sum(-1)
    return torch.atan2(-y, x)


def prior_energy(kind: str, pos, mapping, mapping_batch, n_mols, params: Dict[str, torch.Tensor]):
    """Per-molecule energies [B] of one condensed prior term.
    kind: "bonds"/"angles" k (x-x0)^2 + V0 (prior/harmonic.py:122-123);
          "dihedrals" v0 + sum_n k1_n sin(n phi) + k2_n cos(n phi) (prior/fourier_series.py:154-192);
          "repulsion" (sigma/r)^6 (prior/repulsion.py:119-122)."""
    if kind == "bonds":
        x = feat_distance(pos, mapping)
        y = params["k"] * (x - params["x0"]) ** 2 + params.get("V0", 0.0)
    elif kind == "angles":
        x = feat_angle_cos(pos, mapping)
        y = params["k"] * (x - params["x0"]) ** 2 + params.get("V0", 0.0)
    elif kind == "dihedrals":
        phi = feat_torsion(pos, mapping)
        k1, k2 = params["k1s"], params["k2s"]
        n = torch.arange(1, k1.shape[1] + 1, dtype=pos.dtype)
        ang = phi[:, None] * n[None, :]
        y = (k1 * torch.sin(ang) + k2 * torch.cos(ang)).sum(1) + params["v_0"].flatten()
    elif kind == "repulsion":
        x = feat_distance(pos, mapping)
        rr = (params["sigma"] / x) * (params["sigma"] / x)
        y = rr * rr * rr
    else:
        raise ValueError(kind)
    return torch.zeros(n_mols, dtype=pos.dtype).index_add(0, mapping_batch, y)


def prior_energy_forces(kind, pos, mapping, mapping_batch, n_mols, params):
    pos = pos.detach().clone().requires_grad_(True)
    e = prior_energy(kind, pos, mapping, mapping_batch, n_mols, params)
    (g,) = torch.autograd.grad(e.sum(), pos)
    return e.detach(), -g


# --------------------------------------------------------------------------------------------
# integrator + replica exchange
# --------------------------------------------------------------------------------------------


def baoab_constants(dt: float, friction: float):
    """simulation/langevin.py:76-77."""
    vscale = np.exp(-dt * friction)
    noisescale = np.sqrt(1 - vscale * vscale)
    return vscale, noisescale


def baoab_pre(pos, vel, forces, masses, beta_mass_ratio, noise, dt, vscale, noisescale):
    """B, A, O, A of simulation/langevin.py:137-157 (everything before the force evaluation)."""
    v = vel + 0.5 * dt * forces / masses[:, None]
    x = pos + v * dt * 0.5
    v = v * vscale + noisescale * (beta_mass_ratio * noise)
    x = x + v * dt * 0.5
    return x, v


def baoab_post(vel, forces, masses, dt):
    """final B of simulation/langevin.py:169."""
    return vel + 0.5 * dt * forces / masses[:, None]


def pt_exchange(energies: np.ndarray, betas: np.ndarray, pair_a: np.ndarray, pair_b: np.ndarray,
                uniforms: np.ndarray) -> np.ndarray:
    """Metropolis decision of simulation/parallel_tempering.py:385-394:
    p = exp((u_a-u_b)(beta_a-beta_b)); approved = rand < p   (given the uniforms)."""
    p = np.exp((energies[pair_a] - energies[pair_b]) * (betas[pair_a] - betas[pair_b]))
    return uniforms < p
