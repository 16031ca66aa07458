"""Fused-step engine: the CGSchNet force field (+ priors) and the BAOAB Langevin update as a fixed
sequence of sm_100a kernel launches over pre-allocated HBM buffers, replayable as one CUDA graph.

This is the fast path behind `LangevinSimulation.timestep` / `SumOut.forward` (reference call
stack: simulation/langevin.py:101-179 -> simulation/base.py:821-909 -> models/gradients.py:72-152,
227-290 -> models/schnet.py:177-369).  It replaces autograd by an explicit analytic backward:

  forward                                   backward (weights frozen, only dE/dpos)
  -------                                   --------
  radius graph -> sorted CSR, d_e           F_i = sum_{e in seg(i)} (g_d[e] + g_d[rev e]) u_e
  h0 = Emb[types]                           g_h <- g_h + g_a W1
  per block: a = h W1^T                     g_a = CFConv(g_m, W)          (same kernel, symmetric list)
     t = tanh(rbf Wf0^T + b), W = t Wf1^T   g_W = g_m[src] a[dst] C ; g_t = (g_W Wf1)(1-t^2) ; g_rbf = g_t Wf0
     m = CFConv(a, W)                       g_d += sum_k g_rbf drbf_k/dd + C'(d) sum_f g_m a W   (exact)
     c = tanh(m W2^T + b2)                  g_m = ((g_h Wl)(1-c^2)) W2
     h = h + c Wl^T + bl
  e_atom = MLP_out(h); E_b = sum e_atom     g_h = dE/dh through MLP_out

Because radius_graph's list is symmetric and sorted centre-major / neighbour-ascending, the
"dst-major" CSR of the reference is the same list read through the reverse-edge map, and
W(e) == W(rev e) bit-for-bit, so BOTH CFConv directions stream the filter rows contiguously:
m[i] = sum_{e in seg(i)} a[dst_e] W_e C_e.  No permutation gather, no atomics, deterministic.

Edge buffers have a fixed capacity; the live edge count stays on the device (`seg_ptr[N]`), so a
step has no host synchronisation and can be captured in a CUDA graph.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L


# ------------------------------------------------------------------------------------------------
# flat weights
# ------------------------------------------------------------------------------------------------


@dataclass
class SchNetWeights:
    """Device-resident flat weights.  `tensors` uses nn.Linear layout [out, in] and the key names of
    `oracle`-independent golden files: embedding, b{l}.lin1_w, b{l}.f0_w, b{l}.f0_b, b{l}.f1_w,
    b{l}.lin2_w, b{l}.lin2_b, b{l}.lin_w, b{l}.lin_b, out{i}_w, out{i}_b."""
    tensors: Dict[str, torch.Tensor]
    num_blocks: int
    num_out_layers: int
    cutoff: float
    num_rbf: int
    rbf_lower: float = 0.0
    rbf_centers: Optional[torch.Tensor] = None     # trained / non-default basis: explicit centres and coefficient
    rbf_gamma: Optional[float] = None

    def __post_init__(self):
        t = self.tensors
        dev = t["embedding"].device
        self.device = dev
        self.hidden = t["embedding"].shape[1]
        self.filters = t["b0.f0_w"].shape[0]
        # RBF parameters exactly as GaussianBasis._initial_params (reference radial_basis/gaussian.py:64-75)
        centers = torch.linspace(self.rbf_lower, self.cutoff, self.num_rbf)
        self.gamma = float(-0.5 / (centers[1] - centers[0]) ** 2)
        if self.rbf_centers is not None:
            centers = self.rbf_centers.detach().float().cpu().flatten()
            assert centers.numel() == self.num_rbf
        if self.rbf_gamma is not None:
            self.gamma = float(self.rbf_gamma)
        self.centers = centers.to(dev).float().contiguous()
        k: Dict[str, torch.Tensor] = {}
        for name, w in t.items():
            if w is None:
                continue
            w = w.detach().to(dev).float().contiguous()
            k[name] = w                                   # [out,in]  == [K',N'] for the backward GEMM
            if name.endswith("_w"):
                k[name + "T"] = w.t().contiguous()        # [in,out]  == [K,N] for the forward GEMM
        # fp16 copies for the W16A16 path (reference models/gptq.py:176-184: plain .half() cast)
        for name in list(k.keys()):
            if ".f0_" in name or ".f1_" in name or name.startswith("out"):
                k[name + ".h"] = k[name].half().contiguous()
        # tensor-core operand of the fused filter kernels: Wf0 [F, num_rbf] zero-padded to 64 columns
        if self.num_rbf <= 64:
            for l in range(self.num_blocks):
                wp = torch.zeros((self.filters, 64), dtype=torch.float16, device=dev)
                wp[:, :self.num_rbf] = k[f"b{l}.f0_w.h"]
                k[f"b{l}.f0_w.hp"] = wp.contiguous()
        k["emb_lin1"] = (k["embedding"].double() @ k["b0.lin1_w"].double().t()).float().contiguous()
        k["emb_lin1.h"] = k["emb_lin1"].half().contiguous()
        self.k = k
        # [K,N] weight tensor -> the same matrix stored [N,K] (its transpose twin), for fmd_linear_tc's fast staging
        self.twin = {}
        for name in list(k.keys()):
            base = name.replace("_wT", "_w")
            if "_wT" in name and base in k:
                self.twin[k[name].data_ptr()] = k[base]
                self.twin[k[base].data_ptr()] = k[name]
        self.ones_col = None

    @staticmethod
    def from_flat(tensors: Dict[str, torch.Tensor], cutoff: float, num_rbf: int, device, rbf_centers=None,
                  rbf_gamma=None) -> "SchNetWeights":
        nb = len([n for n in tensors if n.endswith(".lin1_w")])
        no = len([n for n in tensors if n.startswith("out") and n.endswith("_w")])
        tt = {n: (torch.as_tensor(v).to(device) if v is not None else None) for n, v in tensors.items()}
        return SchNetWeights(tt, nb, no, float(cutoff), int(num_rbf), rbf_centers=rbf_centers, rbf_gamma=rbf_gamma)


@dataclass
class PriorTerm:
    kind: int                      # L.PRIOR_*
    mapping: torch.Tensor          # [order, n_terms] int32
    mapping_batch: torch.Tensor    # [n_terms] int32
    p0: torch.Tensor
    p1: Optional[torch.Tensor] = None
    p2: Optional[torch.Tensor] = None
    n_degs: int = 1

    @property
    def n_terms(self):
        return self.mapping.shape[1]


class PriorCSR:
    """Static incidence CSR of all prior terms (built once): what fmd_priors_csr consumes.

    Two-body kinds (bonds, polynomial bonds, repulsion) become 8-byte incidence records + a deduplicated parameter
    table; all angle-like kinds share one [n,8] table with a per-term form code, improper-like kinds another, the
    Fourier dihedrals keep their k1 / k2 arrays."""

    PAIR_KINDS = (L.PRIOR_BONDS, L.PRIOR_REPULSION, L.PRIOR_POLY_BONDS)

    def __init__(self, priors: List[PriorTerm], n_nodes: int, device):
        i32 = torch.int32
        self.n_nodes = n_nodes
        own, others, rec = [], [], []
        ang = [p for p in priors if p.kind in L.ANGLE_FORM]
        imp = [p for p in priors if p.kind in L.IMPROPER_FORM]
        dih = self._merge([p for p in priors if p.kind == L.PRIOR_DIHEDRALS])
        known = set(self.PAIR_KINDS) | set(L.ANGLE_FORM) | set(L.IMPROPER_FORM) | {L.PRIOR_DIHEDRALS}
        for p in priors:
            if p.kind not in known:
                raise ValueError(f"unknown prior kind {p.kind}")
        for p in priors:
            if p.kind not in self.PAIR_KINDS:
                continue
            i, j = p.mapping[0].long(), p.mapping[1].long()
            assert int(p.mapping.max()) < (1 << 28)
            nt = p.n_terms
            par = torch.zeros((nt, 5), dtype=torch.float32, device=device)
            if p.kind == L.PRIOR_POLY_BONDS:       # p0 = ks [n_terms, n_degs <= 4], p2 = V0
                ks = p.p0.float().reshape(nt, -1)
                if ks.shape[1] > 4:
                    raise ValueError("polynomial bond priors up to degree 4 are supported by the fused step")
                par[:, :ks.shape[1]] = ks
                if p.p2 is not None:
                    par[:, 4] = p.p2.float().flatten()
            else:
                par[:, 0] = p.p0.float().flatten()
                if p.kind == L.PRIOR_BONDS:
                    if p.p1 is not None:
                        par[:, 1] = p.p1.float().flatten()
                    if p.p2 is not None:
                        par[:, 2] = p.p2.float().flatten()
            kcol = torch.full((nt, 1), float(p.kind), device=device)
            for a, b in ((i, j), (j, i)):
                own.append(a)
                others.append(b)
                rec.append(torch.cat([kcol, par], 1))                          # [kind, p0..p4]
        self.pair_ptr = self.pair_ent = self.pair_tab = None
        if own:
            own, others, rec = torch.cat(own), torch.cat(others), torch.cat(rec)
            order = torch.argsort(own, stable=True)
            rec, others = rec[order], others[order]
            key = rec.contiguous().view(torch.int32)                         # (kind, p0..p4) bit patterns
            uniq, inv = torch.unique(key, dim=0, return_inverse=True)
            tab = torch.zeros((uniq.shape[0], 8), dtype=torch.float32, device=device)
            tab[:, :5] = uniq[:, 1:].contiguous().view(torch.float32)
            self.pair_tab = tab.contiguous()
            kind = rec[:, 0].long()
            self.pair_ent = torch.stack([(others | (kind << 28)).to(i32), inv.to(i32)], 1).contiguous()
            self.pair_ptr = self._ptr(own, n_nodes)
        # ---- angle-like terms: one table {p0..p5, V0, form}
        self.ang_map = self.ang_par = None
        if ang:
            maps, pars = [], []
            for p in ang:
                nt = p.n_terms
                row = torch.zeros((nt, 8), dtype=torch.float32, device=device)
                if p.kind in (L.PRIOR_ANGLES, L.PRIOR_RAW_ANGLES):
                    row[:, 0], row[:, 1] = p.p0.float().flatten(), p.p1.float().flatten()
                else:                                   # polynomial (k1..k6) / restricted (a, b, c, d, k): p0 = [n_terms, m]
                    ks = p.p0.float().reshape(nt, -1)
                    if ks.shape[1] > 6:
                        raise ValueError("polynomial angle priors up to degree 6 are supported by the fused step")
                    row[:, :ks.shape[1]] = ks
                if p.p2 is not None:
                    row[:, 6] = p.p2.float().flatten()
                row[:, 7] = torch.full((nt,), L.ANGLE_FORM[p.kind], dtype=i32, device=device).view(torch.float32)
                maps.append(p.mapping)
                pars.append(row)
            self.ang_map, self.ang_par = torch.cat(maps, 1).to(i32).contiguous(), torch.cat(pars, 0).contiguous()
        self.imp_map = self.imp_par = None
        if imp:
            maps, pars = [], []
            for p in imp:
                nt = p.n_terms
                row = torch.zeros((nt, 4), dtype=torch.float32, device=device)
                row[:, 0], row[:, 1] = p.p0.float().flatten(), p.p1.float().flatten()
                if p.p2 is not None:
                    row[:, 2] = p.p2.float().flatten()
                row[:, 3] = torch.full((nt,), L.IMPROPER_FORM[p.kind], dtype=i32, device=device).view(torch.float32)
                maps.append(p.mapping)
                pars.append(row)
            self.imp_map, self.imp_par = torch.cat(maps, 1).to(i32).contiguous(), torch.cat(pars, 0).contiguous()
        own, ent = [], []
        for table, mapping in ((0, self.ang_map), (1, dih[0].mapping if dih else None), (2, self.imp_map)):
            if mapping is None:
                continue
            nt = mapping.shape[1]
            assert nt < (1 << 28)
            t = torch.arange(nt, device=device, dtype=torch.int64)
            for r in range(mapping.shape[0]):
                own.append(mapping[r].long())
                ent.append((t | (r << 28) | (table << 30)).to(i32))
        self.mb_ptr = self.mb_ent = None
        if own:
            own, ent = torch.cat(own), torch.cat(ent)
            order = torch.argsort(own, stable=True)
            self.mb_ent = ent[order].contiguous()
            self.mb_ptr = self._ptr(own, n_nodes)
        self.dih = dih[0] if dih else None
        self.e_atom = torch.zeros(n_nodes, dtype=torch.float32, device=device)

    @staticmethod
    def _merge(terms: List[PriorTerm]) -> List[PriorTerm]:
        """Several Fourier-dihedral tables (no specialize_priors) -> one concatenated table."""
        if len(terms) <= 1:
            return terms
        if len({t.n_degs for t in terms}) != 1:
            raise ValueError("dihedral priors with different numbers of Fourier terms cannot share one table")

        def cat(name):
            vs = [getattr(t, name) for t in terms]
            if all(v is None for v in vs):
                return None
            vs = [v if v is not None else torch.zeros_like(t.p0[:, 0] if name == "p2" and t.p0.dim() > 1 else t.p0)
                  for v, t in zip(vs, terms)]
            return torch.cat(vs, 0).contiguous()
        return [PriorTerm(terms[0].kind, torch.cat([t.mapping for t in terms], 1).contiguous(),
                          torch.cat([t.mapping_batch for t in terms], 0).contiguous(), cat("p0"), cat("p1"), cat("p2"),
                          terms[0].n_degs)]

    @staticmethod
    def _ptr(own, n_nodes):
        cnt = torch.bincount(own, minlength=n_nodes)
        ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=own.device)
        ptr[1:] = torch.cumsum(cnt, 0)
        return ptr.to(torch.int32).contiguous()

    def launch(self, pos, forces, accumulate, st):
        d = self.dih
        L.call("fmd_priors_csr", L.ptr(pos), self.n_nodes, L.ptr(self.pair_ptr), L.ptr(self.pair_ent), L.ptr(self.pair_tab),
               L.ptr(self.mb_ptr), L.ptr(self.mb_ent),
               L.ptr(self.ang_map), self.ang_map.shape[1] if self.ang_map is not None else 0, L.ptr(self.ang_par),
               L.ptr(d.mapping) if d else None, d.n_terms if d else 0, L.ptr(d.p0) if d else None, L.ptr(d.p1) if d else None,
               L.ptr(d.p2) if d else None, d.n_degs if d else 1,
               L.ptr(self.imp_map), self.imp_map.shape[1] if self.imp_map is not None else 0, L.ptr(self.imp_par),
               L.ptr(self.e_atom), L.ptr(forces), int(accumulate), st)


# ------------------------------------------------------------------------------------------------
# force field
# ------------------------------------------------------------------------------------------------


class ForceField:
    """energy [B], forces [N,3] = f(pos) for SchNet (+ priors), all buffers pre-allocated."""

    def __init__(self, weights: Optional[SchNetWeights], priors: List[PriorTerm], atom_types: torch.Tensor,
                 mol_ptr: torch.Tensor, precision: str = "fp32", exact_cutoff_grad: bool = True,
                 edge_capacity: Optional[int] = None, max_num_neighbors: int = 1000, use_tensor_cores: bool = True):
        L.load()
        assert precision in ("fp32", "w16a16")
        self.w = weights
        self.priors = priors
        self.precision = precision
        self.exact = bool(exact_cutoff_grad)
        self.max_nn = int(max_num_neighbors)
        dev = atom_types.device
        if dev.type != "cuda":
            raise RuntimeError("flashmd.engine.ForceField needs CUDA tensors (no CPU fallback)")
        self.device = dev
        self.N = int(atom_types.numel())
        self.B = int(mol_ptr.numel() - 1)
        self.types = atom_types.to(torch.int32).contiguous()
        self.mol_ptr = mol_ptr.to(torch.int32).contiguous()
        sizes = (mol_ptr[1:] - mol_ptr[:-1])
        self.max_mol = int(sizes.max().item()) if self.B > 0 else 0
        if weights is not None and self.max_mol - 1 > self.max_nn:
            # radius_graph truncates per centre: the list would lose its symmetry (rev = -1 on dropped reverse edges) and
            # the analytic backward (CFConv over the same list, g_d[e] + g_d[rev e]) would no longer be the gradient
            raise ValueError(f"a molecule has {self.max_mol} beads but max_num_neighbors = {self.max_nn}: the fused step "
                             "needs a symmetric neighbour list (max_num_neighbors >= beads per molecule - 1)")
        N, B = self.N, self.B
        f32, i32 = torch.float32, torch.int32
        self.energy = torch.zeros(B, dtype=f32, device=dev)
        self.forces = torch.zeros((N, 3), dtype=f32, device=dev)
        self.energy_terms: Dict[str, torch.Tensor] = {}
        self.prior_csr = PriorCSR(priors, N, dev) if priors else None
        if weights is None:
            return
        F, R = weights.filters, weights.num_rbf
        H = weights.hidden
        assert F % 4 == 0 and H % 4 == 0
        if edge_capacity is None:
            # worst case: every pair inside each molecule (bounded by max_num_neighbors)
            per = torch.clamp(sizes - 1, max=self.max_nn).to(torch.int64) * sizes.to(torch.int64)
            edge_capacity = int(per.sum().item())
        self.cap = cap = max(int(edge_capacity), 1)
        wdt = torch.float16 if precision == "w16a16" else f32
        # W16A16 with the default widths runs the fused tcgen05 kernels (no [E,F] tensor in HBM);
        # other widths / the fp32 parity path use the materialised SIMT kernels.
        # (num_rbf <= 63: one padded column of the radial-basis operand carries the bias through the first GEMM)
        self.fused_tc = (precision == "w16a16" and F == 128 and H == 128 and R <= 63 and use_tensor_cores)
        # fp32 parity path: dense layers as fp32-accurate BF16x3 (fp32-emulation) tensor-core GEMMs (fmd_linear_x3) instead of SIMT FMA
        self.x3 = precision == "fp32" and use_tensor_cores and os.environ.get("FMD_X3", "1") == "1"
        # node-level layers stay on the true-fp32 FMA kernel by default: they are 6 % of the step and the tensor-core
        # accumulation (round-toward-zero) leaves a coherent -1e-5 bias in the per-molecule energies (forces equal)
        self.x3_nodes = os.environ.get("FMD_X3_NODES", "0") == "1"
        self.deg = torch.zeros(N, dtype=i32, device=dev)
        self.seg_ptr = torch.zeros(N + 1, dtype=i32, device=dev)
        self.n_edges_dev = self.seg_ptr[N:]                      # int32[1] view: live edge count
        self.max_edges_dev = torch.zeros(1, dtype=i32, device=dev)    # sticky high-water mark (fused path: fmd_nl_step)
        self.scan_ws = torch.zeros(N // 1024 + 4, dtype=i32, device=dev)
        self.src = torch.zeros(cap, dtype=i32, device=dev)
        self.dst = torch.zeros(cap, dtype=i32, device=dev)
        self.rev = torch.zeros(cap, dtype=i32, device=dev)
        self.dist = torch.zeros(cap, dtype=f32, device=dev)
        nb = weights.num_blocks
        if self.fused_tc:
            self.part = torch.zeros(((cap + 127) // 128, F), dtype=f32, device=dev)   # per-tile head partials
            # undirected pair list (edges with dst > src): the backward edge kernel runs over pairs, half the tiles
            pcap = self.pair_cap = cap // 2 + 1
            self.pair_cnt = torch.zeros(N, dtype=i32, device=dev)
            self.pair_ptr = torch.zeros(N + 1, dtype=i32, device=dev)
            self.n_pairs_dev = self.pair_ptr[N:]
            self.pair_own = torch.zeros(pcap, dtype=i32, device=dev)
            self.pair_nbr = torch.zeros(pcap, dtype=i32, device=dev)
            self.pair_dist = torch.zeros(pcap, dtype=f32, device=dev)
            self.pidx = torch.zeros(cap, dtype=i32, device=dev)
            self.g_pair = torch.zeros(pcap, dtype=f32, device=dev)
        else:
            self.rbf = torch.zeros((cap, R), dtype=f32, device=dev)
            self.t = [torch.zeros((cap, F), dtype=wdt, device=dev) for _ in range(nb)]
            self.W = [torch.zeros((cap, F), dtype=wdt, device=dev) for _ in range(nb)]
            self.gW = torch.zeros((cap, F), dtype=wdt, device=dev)
            self.gT = torch.zeros((cap, F), dtype=wdt, device=dev)
            self.g_rbf = torch.zeros((cap, R), dtype=f32, device=dev)
        self.g_d = torch.zeros(cap, dtype=f32, device=dev)
        self.h = [torch.zeros((N, H), dtype=f32, device=dev) for _ in range(nb + 1)]
        # the operands the fused edge kernels gather per edge (a, g_m) are stored as fp16 rows: half the L2 traffic of the
        # gathers, 16-byte cp.async staging (the reference multiplies fp32 x by the fp16 filter: ours rounds x as well)
        gdt = torch.float16 if self.fused_tc else f32
        self.a = [torch.zeros((N, F), dtype=gdt, device=dev) for _ in range(nb)]
        self.m = torch.zeros((N, F), dtype=f32, device=dev)
        self.c = [torch.zeros((N, H), dtype=f32, device=dev) for _ in range(nb)]
        self.g_h = [torch.zeros((N, H), dtype=f32, device=dev) for _ in range(2)]
        self.g_c = torch.zeros((N, H), dtype=f32, device=dev)
        self.g_m = torch.zeros((N, F), dtype=gdt, device=dev)
        self.g_a = torch.zeros((N, F), dtype=f32, device=dev)
        odt = torch.float16 if precision == "w16a16" else f32
        widths = [weights.tensors[f"out{i}_w"].shape[0] for i in range(weights.num_out_layers)]
        self.y = [torch.zeros((N, wd), dtype=(odt if i < len(widths) - 1 else f32), device=dev)
                  for i, wd in enumerate(widths)]
        self.g_y = [torch.zeros((N, wd), dtype=odt, device=dev) for wd in widths[:-1]]
        self.ones = torch.ones((N, 1), dtype=odt, device=dev)
        self.launches_per_eval = 0
        self._embedded = False
        self._side = None
        self.serial_priors = False     # True: prior kernel on the main stream instead of the forked one (per-kernel timing)

    # -- helpers ---------------------------------------------------------------------------------
    def _lin(self, x, w, bias, y, m_dev=None, **kw):
        M, K = x.shape
        N = w.shape[1]
        tc = self.fused_tc and K in (64, 128) and N in (64, 128)
        args = (L.ptr(x), L.dt_code(x), L.ptr(w), L.dt_code(w), L.ptr(bias), L.ptr(y), L.dt_code(y), M, N,
                K, L.ptr(m_dev), kw.get("pro_act", 0), int(kw.get("x_round", False)), kw.get("epi_act", 0),
                L.ptr(kw.get("aux")), L.dt_code(kw["aux"]) if kw.get("aux") is not None else 0, L.ptr(kw.get("res")))
        f32 = torch.float32
        x3 = (self.x3 and (self.x3_nodes or m_dev is not None) and x.dtype == f32 and w.dtype == f32 and y.dtype == f32 and K <= 128 and N <= 128 and K % 2 == 0
              and not kw.get("pro_act", 0) and not kw.get("x_round", False)
              and (kw.get("aux") is None or kw["aux"].dtype == f32) and (bias is None or bias.dtype == f32))
        if tc:
            wt = self.w.twin.get(w.data_ptr())
            if wt is not None:
                args = args[:2] + (L.ptr(wt),) + args[3:]
            L.call("fmd_linear_tc", *args, int(wt is not None), self._st)
        elif x3:
            # fp32 parity path: BF16x3 fp32-emulation GEMM on the tensor cores (fp32-accurate), HBM-bound streaming kernel
            L.call("fmd_linear_x3", L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(y), M, N, K, L.ptr(m_dev),
                   kw.get("epi_act", 0), L.ptr(kw.get("aux")), L.ptr(kw.get("res")), self._st)
        else:
            L.call("fmd_linear", *args, self._st)
        self._n += 1

    def _chain(self, x, stages, pro_act=0, x_round=False):
        """Consecutive node-level layers in ONE launch (fmd_linear_chain_tc).  stages: dicts with W ([N,K] tensor),
        bias, epi, aux, res, Y (tensor or None), round (feed the next stage with the fp16-rounded value)."""
        arr = (L.DenseStage * len(stages))()
        for s_, d in zip(arr, stages):
            W = d["W"]
            s_.W, s_.bias, s_.wdt, s_.N = L.ptr(W), L.ptr(d.get("bias")), L.dt_code(W), W.shape[0]
            s_.epi_act = d.get("epi", 0)
            aux, res, Y = d.get("aux"), d.get("res"), d.get("Y")
            s_.aux, s_.auxdt = L.ptr(aux), (L.dt_code(aux) if aux is not None else 0)
            s_.res, s_.Y, s_.ydt = L.ptr(res), L.ptr(Y), (L.dt_code(Y) if Y is not None else 0)
            s_.round_f16 = int(bool(d.get("round", False)))
        L.call("fmd_linear_chain_tc", L.ptr(x), L.dt_code(x), x.shape[0], x.shape[1], pro_act, int(x_round), arr,
               len(stages), self._st)
        self._n += 1

    def _cfconv(self, x, W, out):
        L.call("fmd_cfconv_csr", L.ptr(x), L.ptr(W), L.dt_code(W), L.ptr(self.dist), L.ptr(self.dst),
               L.ptr(self.seg_ptr), None, 4, self.N, self.cap, x.shape[1], self.w.cutoff, L.ptr(out), self._st)
        self._n += 1

    def _filter_cfconv(self, l, x, out):
        """out[i] = sum_{e in seg(i)} W_l(d_e) * x[dst_e] * C(d_e): filter network + CFConv fused on tensor cores
        (x: fp16 [N,128])."""
        assert x.dtype == torch.float16
        w, k = self.w, self.w.k
        L.call("fmd_filter_cfconv_fwd", L.ptr(self.dist), L.ptr(self.src), L.ptr(self.dst), L.ptr(self.seg_ptr), self.N,
               self.cap, L.ptr(self.n_edges_dev), L.ptr(k[f"b{l}.f0_w.hp"]), L.ptr(k[f"b{l}.f0_b.h"]),
               L.ptr(k[f"b{l}.f1_w.h"]), L.ptr(w.centers), w.num_rbf, w.gamma, w.cutoff, L.ptr(x), w.filters,
               L.ptr(out), L.ptr(self.part), self._st)
        self._n += 2

    def _filter_cfconv_bwd(self, l, a, g_m, accumulate=True):
        """g_pair[p] += d/dd_p of sum_f W_l(d_p)[f] C(d_p) (g_m[i,f] a[j,f] + g_m[j,f] a[i,f]) over the undirected pairs
        p = (i < j): the two directed-edge gradients of the reference's backward in one (fused, tensor cores)."""
        w, k = self.w, self.w.k
        L.call("fmd_filter_cfconv_bwd", L.ptr(self.pair_dist), L.ptr(self.pair_own), L.ptr(self.pair_nbr), self.pair_cap,
               L.ptr(self.n_pairs_dev), L.ptr(k[f"b{l}.f0_w.hp"]), L.ptr(k[f"b{l}.f0_b.h"]), L.ptr(k[f"b{l}.f1_w.h"]),
               L.ptr(w.centers), w.num_rbf, w.gamma, w.cutoff, L.ptr(a), L.ptr(g_m), w.filters, L.ptr(self.g_pair),
               int(accumulate), int(self.exact), self._st)
        self._n += 1

    # -- neighbour list --------------------------------------------------------------------------
    def build_neighbor_list(self, pos):
        st, w = self._st, self.w
        if self.fused_tc:
            # sorted symmetric edge list + reverse map + undirected pair list in four launches
            L.call("fmd_nl_step", L.ptr(pos), L.ptr(self.mol_ptr), self.B, self.N, self.max_mol, w.cutoff, self.max_nn,
                   L.ptr(self.deg), L.ptr(self.seg_ptr), self.cap, L.ptr(self.src), L.ptr(self.dst), L.ptr(self.dist),
                   L.ptr(self.rev), L.ptr(self.pair_cnt), L.ptr(self.pair_ptr), self.pair_cap, L.ptr(self.pair_own),
                   L.ptr(self.pair_nbr), L.ptr(self.pair_dist), L.ptr(self.pidx), L.ptr(self.max_edges_dev), st)
            self._n += 4
            return
        L.call("fmd_nl_count", L.ptr(pos), L.ptr(self.mol_ptr), self.B, self.N, self.max_mol, w.cutoff, self.max_nn,
               L.ptr(self.deg), st)
        L.call("fmd_exclusive_scan_i32", L.ptr(self.deg), L.ptr(self.seg_ptr), self.N, L.ptr(self.scan_ws), st)
        L.call("fmd_nl_fill", L.ptr(pos), L.ptr(self.mol_ptr), self.B, self.N, self.max_mol, w.cutoff, self.max_nn,
               L.ptr(self.seg_ptr), self.cap, L.ptr(self.src), L.ptr(self.dst), 4, L.ptr(self.dist), st)
        L.call("fmd_nl_reverse", L.ptr(self.seg_ptr), L.ptr(self.src), L.ptr(self.dst), 4, self.N, self.cap,
               L.ptr(self.rev), st)
        self._n += 6

    def _side_work(self, pos):
        """Work the forward kernels do not depend on, forked onto a second stream: ALL prior terms (they depend on
        positions only and write the force buffer, the SchNet forces are accumulated on top at the end).  In the captured
        graph this is a parallel branch: the prior kernel runs in the shadow of the first persistent forward kernel, which
        leaves issue slots idle (forked before the neighbour list it only slows that one down: measured)."""
        if self.prior_csr is None:
            return
        if self.serial_priors:          # per-kernel timing (bench.py): nothing overlaps, every kernel is timed alone
            self.prior_csr.launch(pos, self.forces, False, self._st)
            self._n += 1
            return
        if self._side is None:
            self._side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            self.prior_csr.launch(pos, self.forces, False, self._side.cuda_stream)
            self._n += 1

    def _join_side(self):
        if self.prior_csr is not None and not self.serial_priors:
            torch.cuda.current_stream().wait_stream(self._side)

    def num_edges(self) -> int:
        """Host read of the live edge count (synchronises)."""
        return int(self.n_edges_dev.item())

    # -- SchNet ----------------------------------------------------------------------------------
    def _chain_ok(self):
        w, k = self.w, self.w.k
        no = w.num_out_layers
        widths = [w.tensors[f"out{i}_w"].shape[0] for i in range(no)]
        return (self.fused_tc and w.hidden == 128 and w.filters == 128 and no >= 2 and widths[-1] == 1
                and all(x in (64, 128) for x in widths[:-1]) and k.get(f"out{no - 1}_b.h") is None
                and 2 + (no - 1) <= L.MAX_CHAIN)

    def _schnet_tc(self, pos):
        """W16A16 step with the default widths: 5 fused edge launches + 3 backward edge launches, the node layers
        as 6 chained tensor-core launches (fmd_linear_chain_tc), nothing else but neighbour list / head / forces."""
        w, k, st = self.w, self.w.k, self._st
        nb, no = w.num_blocks, w.num_out_layers
        T, TC = L.ACT_TANH, L.ACT_TANH_CLAMPED
        self.build_neighbor_list(pos)
        self._side_work(pos)
        if not self._embedded:
            # bead types never change during a run: h_0 = Emb[types] and a_0 = (Emb W1^T)[types] are written once
            # (a_0 as fp16 rows: 256 B = 64 floats for the row-copy kernel)
            L.call("fmd_embedding", L.ptr(k["embedding"]), L.ptr(self.types), 4, self.N, w.hidden, L.ptr(self.h[0]), st)
            L.call("fmd_embedding", L.ptr(k["emb_lin1.h"]), L.ptr(self.types), 4, self.N, w.filters // 2, L.ptr(self.a[0]), st)
            self._embedded = not torch.cuda.is_current_stream_capturing()
        for l in range(nb):
            self._filter_cfconv(l, self.a[l], self.m)
            last = l == nb - 1
            stages = [dict(W=k[f"b{l}.lin2_w"], bias=k[f"b{l}.lin2_b"], epi=T, Y=self.c[l]),
                      dict(W=k[f"b{l}.lin_w"], bias=k[f"b{l}.lin_b"], res=self.h[l], Y=self.h[l + 1], round=last)]
            if not last:
                stages.append(dict(W=k[f"b{l + 1}.lin1_w"], Y=self.a[l + 1]))
            else:   # output network except its last layer (fp16 operands, clamped tanh, fp16 storage)
                for i in range(no - 1):
                    stages.append(dict(W=k[f"out{i}_w.h"], bias=k.get(f"out{i}_b.h"), epi=TC, Y=self.y[i], round=True))
            self._chain(self.m, stages)
        L.call("fmd_out_head", L.ptr(self.y[no - 2]), L.ptr(k[f"out{no - 1}_w.h"]), L.F16, self.N, self.y[no - 2].shape[1],
               L.ptr(self.y[-1]), L.ptr(self.g_y[no - 2]), st)
        L.call("fmd_segment_sum", L.ptr(self.y[-1]), L.ptr(self.mol_ptr), self.B, L.ptr(self.energy), 0, st)
        self._n += 2
        # ---------------- backward.  [N,K] layout of a backward GEMM's weight == the forward weight's transpose.
        self._join_side()      # pair list (+ prior forces in self.forces) from the side stream
        gh_cur, gh_nxt = self.g_h[0], self.g_h[1]
        stages = []
        for i in range(no - 2, 0, -1):      # g_y[i-1] = (g_y[i] @ out_i_w) * (1 - y[i-1]^2), stored as fp16 in the reference
            stages.append(dict(W=k[f"out{i}_wT.h"], aux=self.y[i - 1], round=True))
        stages.append(dict(W=k["out0_wT.h"], Y=gh_cur))
        x = self.g_y[no - 2]
        for l in range(nb - 1, -1, -1):
            stages.append(dict(W=k[f"b{l}.lin_wT"], aux=self.c[l]))          # g_c = (g_h @ Wl) * (1 - c^2)
            stages.append(dict(W=k[f"b{l}.lin2_wT"], Y=self.g_m))           # g_m = g_c @ W2
            self._chain(x, stages)
            self._filter_cfconv_bwd(l, self.a[l], self.g_m, accumulate=(l < nb - 1))   # the first call overwrites g_pair
            if l > 0:   # dE/dh_0 does not enter the forces: no grad_x / g_h update below block 0
                self._filter_cfconv(l, self.g_m, self.g_a)
                stages = [dict(W=k[f"b{l}.lin1_wT"], res=gh_cur, Y=gh_nxt)]  # g_h <- g_h + g_a @ W1
                gh_cur, gh_nxt = gh_nxt, gh_cur
                x = self.g_a
        L.call("fmd_edge_grad_to_forces_csr", L.ptr(pos), L.ptr(self.seg_ptr), L.ptr(self.dst), L.ptr(self.pidx),
               L.ptr(self.dist), L.ptr(self.g_pair), self.N, self.cap, 1.0, L.ptr(self.forces),
               int(self.prior_csr is not None), 1, st)      # on top of the prior forces written by the side stream
        self._n += 1
        if self.prior_csr is not None:
            L.call("fmd_segment_sum", L.ptr(self.prior_csr.e_atom), L.ptr(self.mol_ptr), self.B, L.ptr(self.energy), 1, st)
            self._n += 1
        self._priors_done = True

    def _schnet(self, pos):
        if self._chain_ok():
            return self._schnet_tc(pos)
        w, k, st = self.w, self.w.k, self._st
        w16 = self.precision == "w16a16"
        nb, ned = w.num_blocks, self.n_edges_dev
        self.build_neighbor_list(pos)
        tc = self.fused_tc
        if not tc:
            # rbf [E,R] (distances were written by the neighbour-list fill)
            L.call("fmd_dist_rbf_cutoff_fwd", L.ptr(pos), L.ptr(self.src), L.ptr(self.dst), 4, self.cap, L.ptr(ned),
                   L.ptr(w.centers), w.num_rbf, w.gamma, w.cutoff, None, L.ptr(self.rbf), st)
            self._n += 1
        L.call("fmd_embedding", L.ptr(k["embedding"]), L.ptr(self.types), 4, self.N, w.hidden, L.ptr(self.h[0]), st)
        self._n += 1
        tanh_f = L.ACT_TANH_CLAMPED if w16 else L.ACT_TANH
        sfx = ".h" if w16 else ""
        for l in range(nb):
            if l == 0:
                # a_0 = Emb[types] W1^T = (Emb W1^T)[types]: the [n_types, F] product is precomputed once
                if tc:
                    L.call("fmd_embedding", L.ptr(k["emb_lin1.h"]), L.ptr(self.types), 4, self.N, w.filters // 2, L.ptr(self.a[0]), st)
                else:
                    L.call("fmd_embedding", L.ptr(k["emb_lin1"]), L.ptr(self.types), 4, self.N, w.filters, L.ptr(self.a[0]), st)
                self._n += 1
            else:
                self._lin(self.h[l], k[f"b{l}.lin1_wT"], None, self.a[l])
            if tc:
                self._filter_cfconv(l, self.a[l], self.m)
            else:
                self._lin(self.rbf, k[f"b{l}.f0_wT{sfx}"], k[f"b{l}.f0_b{sfx}"], self.t[l], m_dev=ned, x_round=w16,
                          epi_act=tanh_f)
                self._lin(self.t[l], k[f"b{l}.f1_wT{sfx}"], None, self.W[l], m_dev=ned)
                self._cfconv(self.a[l], self.W[l], self.m)
            self._lin(self.m, k[f"b{l}.lin2_wT"], k[f"b{l}.lin2_b"], self.c[l], epi_act=L.ACT_TANH)
            self._lin(self.c[l], k[f"b{l}.lin_wT"], k[f"b{l}.lin_b"], self.h[l + 1], res=self.h[l])
        # output network
        no = w.num_out_layers
        x = self.h[nb]
        # the last layer (hidden -> 1, no bias) and the first backward step are one pass over the last hidden layer
        head = no >= 2 and k.get(f"out{no - 1}_b{sfx}") is None and self.y[-1].shape[1] == 1
        for i in range(no - 1 if head else no):
            last = i == no - 1
            b = k.get(f"out{i}_b{sfx}")
            self._lin(x, k[f"out{i}_wT{sfx}"], b, self.y[i], x_round=(w16 and x.dtype == torch.float32),
                      epi_act=(L.ACT_NONE if last else tanh_f))
            x = self.y[i]
        if head:
            L.call("fmd_out_head", L.ptr(x), L.ptr(k[f"out{no - 1}_w{sfx}"]), L.dt_code(x), self.N, x.shape[1],
                   L.ptr(self.y[-1]), L.ptr(self.g_y[no - 2]), st)
            self._n += 1
        assert self.y[-1].shape[1] == 1
        L.call("fmd_segment_sum", L.ptr(self.y[-1]), L.ptr(self.mol_ptr), self.B, L.ptr(self.energy), 0, st)
        self._n += 1
        # ---------------- backward
        # dE/d(e_atom) = 1  ->  through the output MLP
        g = self.g_y[no - 2] if head else self.ones
        for i in range(no - 2 if head else no - 1, 0, -1):
            # g_y[i-1] = (g @ out_i_w[out,in]) * (1 - y[i-1]^2)
            self._lin(g, k[f"out{i}_w{sfx}"], None, self.g_y[i - 1], aux=self.y[i - 1], x_round=(w16 and g.dtype == torch.float32))
            g = self.g_y[i - 1]
        gh_cur, gh_nxt = self.g_h[0], self.g_h[1]
        self._lin(g, k[f"out0_w{sfx}"], None, gh_cur, x_round=(w16 and g.dtype == torch.float32))
        if not tc:
            self.g_d.zero_()
            self._n += 1
        for l in range(nb - 1, -1, -1):
            # h_{l+1} = h_l + c Wl^T + bl ; c = tanh(m W2^T + b2)
            self._lin(gh_cur, k[f"b{l}.lin_w"], None, self.g_c, aux=self.c[l])
            self._lin(self.g_c, k[f"b{l}.lin2_w"], None, self.g_m)
            if tc:
                self._filter_cfconv_bwd(l, self.a[l], self.g_m, accumulate=(l < nb - 1))
                if l > 0:      # dE/dh_0 (the embedding gradient) is not needed for forces: skip g_a / g_h of block 0
                    self._filter_cfconv(l, self.g_m, self.g_a)
                    self._lin(self.g_a, k[f"b{l}.lin1_w"], None, gh_nxt, res=gh_cur)
                    gh_cur, gh_nxt = gh_nxt, gh_cur
                continue
            # m[i] = sum_{e in seg(i)} a[dst_e] W_e C_e
            if l > 0:
                self._cfconv(self.g_m, self.W[l], self.g_a)
            L.call("fmd_cfconv_grad_filter", L.ptr(self.g_m), L.ptr(self.a[l]), L.ptr(self.dist), L.ptr(self.src),
                   L.ptr(self.dst), 4, self.cap, L.ptr(ned), w.filters, w.cutoff, L.ptr(self.gW), L.dt_code(self.gW),
                   L.ptr(self.W[l]) if self.exact else None, L.dt_code(self.W[l]),
                   L.ptr(self.g_d) if self.exact else None, 1, st)
            self._n += 1
            # g_t = (g_W @ f1_w[f,j]) * (1 - t^2) ; g_rbf = g_t @ f0_w[j,k]
            self._lin(self.gW, k[f"b{l}.f1_w{sfx}"], None, self.gT, m_dev=ned, aux=self.t[l])
            if self.x3 and w.filters <= 128 and w.filters % 2 == 0 and w.num_rbf <= 128:
                # g_rbf = gT @ f0_w contracted with d rbf / d d in the GEMM's epilogue: [E,R] never leaves the SM
                L.call("fmd_linear_x3_rbf_bwd", L.ptr(self.gT), L.ptr(k[f"b{l}.f0_w"]), self.cap, w.num_rbf, w.filters,
                       L.ptr(ned), L.ptr(self.dist), L.ptr(w.centers), w.gamma, w.cutoff, L.ptr(self.g_d), 1, st)
                self._n += 1
            else:
                self._lin(self.gT, k[f"b{l}.f0_w{sfx}"], None, self.g_rbf, m_dev=ned)
                L.call("fmd_rbf_bwd", L.ptr(self.dist), L.ptr(self.g_rbf), None, self.cap, L.ptr(ned), L.ptr(w.centers),
                       w.num_rbf, w.gamma, w.cutoff, L.ptr(self.g_d), 1, st)
                self._n += 1
            # g_h <- g_h + g_a @ lin1_w[f,h]   (not needed below block 0: dE/dh_0 does not enter the forces)
            if l > 0:
                self._lin(self.g_a, k[f"b{l}.lin1_w"], None, gh_nxt, res=gh_cur)
                gh_cur, gh_nxt = gh_nxt, gh_cur
        if tc:
            L.call("fmd_edge_grad_to_forces_csr", L.ptr(pos), L.ptr(self.seg_ptr), L.ptr(self.dst), L.ptr(self.pidx),
                   L.ptr(self.dist), L.ptr(self.g_pair), self.N, self.cap, 1.0, L.ptr(self.forces), 0, 1, st)
        else:
            L.call("fmd_edge_grad_to_forces_csr", L.ptr(pos), L.ptr(self.seg_ptr), L.ptr(self.dst), L.ptr(self.rev),
                   L.ptr(self.dist), L.ptr(self.g_d), self.N, self.cap, 1.0, L.ptr(self.forces), 0, 0, st)
        self._n += 1

    # -- public ----------------------------------------------------------------------------------
    def compute(self, pos: torch.Tensor):
        """Fill self.energy [B] and self.forces [N,3] for `pos` [N,3] (fp32, CUDA, contiguous)."""
        assert pos.is_cuda and pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape == (self.N, 3)
        self._st = L.stream_ptr()
        self._n = 0
        self._priors_done = False
        if self.w is not None:
            self._schnet(pos)      # writes self.energy (per-molecule SchNet energy) and self.forces
        elif self.prior_csr is None:
            self.energy.zero_()
            self.forces.zero_()
            self._n += 2
        if self.prior_csr is not None and not self._priors_done:
            # all prior classes in one owner-computes launch, then the per-molecule energy reduction
            self.prior_csr.launch(pos, self.forces, self.w is not None, self._st)
            L.call("fmd_segment_sum", L.ptr(self.prior_csr.e_atom), L.ptr(self.mol_ptr), self.B, L.ptr(self.energy),
                   int(self.w is not None), self._st)
            self._n += 2
        self.launches_per_eval = self._n
        return self.energy, self.forces


# ------------------------------------------------------------------------------------------------
# Langevin engine
# ------------------------------------------------------------------------------------------------


class LangevinEngine:
    """BAOAB steps (reference simulation/langevin.py:101-179) over a ForceField, optionally replayed
    as a CUDA graph.  State (pos, vel, forces) lives in fixed device buffers."""

    def __init__(self, ff: ForceField, pos: torch.Tensor, vel: torch.Tensor, masses: torch.Tensor,
                 beta: torch.Tensor, dt: float, friction: float, seed: int = 0, use_graph: bool = True,
                 noise_mode: str = "philox", node_offset: int = 0):
        self.ff = ff
        dev = ff.device
        self.pos = pos.detach().to(dev).float().contiguous().clone()
        self.vel = vel.detach().to(dev).float().contiguous().clone()
        masses = masses.detach().to(dev).float().contiguous()
        self.masses = masses
        self.inv_mass = (1.0 / masses).contiguous()
        sizes = (ff.mol_ptr[1:] - ff.mol_ptr[:-1]).long()
        beta_atom = beta.detach().to(dev).float().repeat_interleave(sizes)
        self.beta = beta.detach().to(dev).float().contiguous()
        # beta_mass_ratio (reference simulation/langevin.py:211-215)
        self.noise_std = torch.sqrt(1.0 / beta_atom / masses).contiguous()
        self.dt = float(dt)
        self.vscale = float(np.exp(-dt * friction))                 # langevin.py:76
        self.noisescale = float(np.sqrt(1 - self.vscale * self.vscale))   # langevin.py:77
        self.seed = int(seed) & ((1 << 64) - 1)
        # global index of this shard's first bead: the Philox counter is (step, node_offset + local bead), so a batch
        # sharded over several GPUs draws the noise of the unsharded run and no two shards share a stream
        self.node_offset = int(node_offset)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ke = torch.zeros(ff.B, dtype=torch.float32, device=dev)
        self.use_graph = use_graph
        self.graph = None
        # "philox": counter-based noise generated inside the kernel; "buffer": N(0,1) numbers the caller writes
        # into `noise_buf` before every step (e.g. torch's generator, for step-exact reference comparisons)
        assert noise_mode in ("philox", "buffer")
        self.noise_buf = torch.zeros((ff.N, 3), dtype=torch.float32, device=dev) if noise_mode == "buffer" else None
        self.n_steps_done = 0
        self.launches_per_step = 0
        ff.compute(self.pos)   # initial forces (reference simulation/base.py:525-526)

    def _step_body(self, noise=None):
        ff, st = self.ff, L.stream_ptr()
        if noise is None:
            noise = self.noise_buf
        L.call("fmd_baoab_pre", L.ptr(self.pos), L.ptr(self.vel), L.ptr(ff.forces), L.ptr(self.inv_mass),
               L.ptr(self.noise_std), L.ptr(noise), self.seed, 0, L.ptr(self.step_dev), self.node_offset, ff.N, self.dt, self.vscale,
               self.noisescale, st)
        ff.compute(self.pos)
        L.call("fmd_baoab_post", L.ptr(self.vel), L.ptr(ff.forces), L.ptr(self.inv_mass), ff.N, self.dt,
               L.ptr(ff.mol_ptr), ff.B, L.ptr(self.ke), L.ptr(self.step_dev), st)      # (+ the Philox step counter)
        self.launches_per_step = ff.launches_per_eval + 3

    def step(self, noise: Optional[torch.Tensor] = None):
        """One BAOAB step.  `noise` [N,3] replaces the Philox stream (step-exact parity tests)."""
        if noise is not None or not self.use_graph:
            self._step_body(noise)
        else:
            if self.graph is None:
                self._capture()
            self.graph.replay()
        self.n_steps_done += 1

    def _capture(self):
        # warm-up on a side stream (allocations, lazy module loads), then capture one step
        # the state is saved BEFORE the side stream's wait point: clones enqueued after s.wait_stream() would race with
        # the warm-up step (the restored state was then a partially advanced one, different from run to run)
        saved = (self.pos.clone(), self.vel.clone(), self.ff.forces.clone(), self.step_dev.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step_body()
        torch.cuda.synchronize()
        # restore the state from before warm-up + capture-time execution (capture does not execute)
        self.pos.copy_(saved[0]); self.vel.copy_(saved[1]); self.ff.forces.copy_(saved[2]); self.step_dev.copy_(saved[3])
        self.graph = g

    def run(self, n_steps: int):
        for _ in range(n_steps):
            self.step()

    def kinetic_energy(self):
        return self.ke

    def temperature_kT(self):
        """<2 KE / (3 n)> per molecule in energy units (equipartition: kT = 2 KE / dof)."""
        sizes = (self.ff.mol_ptr[1:] - self.ff.mol_ptr[:-1]).float()
        return 2.0 * self.ke / (3.0 * sizes)


class OverdampedEngine(LangevinEngine):
    """Overdamped Langevin steps (reference simulation/langevin.py:315-420) over a ForceField on the same fixed buffers /
    CUDA graph as LangevinEngine:  x += F D dt + sqrt(2 D dt) xi,  D = 1 / (beta friction) per bead.  No velocities."""

    def __init__(self, ff: ForceField, pos: torch.Tensor, beta: torch.Tensor, dt: float, friction: float, seed: int = 0,
                 use_graph: bool = True, noise_mode: str = "philox", node_offset: int = 0):
        n = ff.N
        super().__init__(ff, pos, torch.zeros((n, 3)), torch.ones(n), beta, dt, friction, seed=seed, use_graph=use_graph,
                         noise_mode=noise_mode, node_offset=node_offset)
        sizes = (ff.mol_ptr[1:] - ff.mol_ptr[:-1]).long()
        beta_atom = self.beta.repeat_interleave(sizes)
        self.dtau = (self.dt / (beta_atom * float(friction))).contiguous()     # D dt

    def _step_body(self, noise=None):
        ff, st = self.ff, L.stream_ptr()
        if noise is None:
            noise = self.noise_buf
        L.call("fmd_overdamped_step", L.ptr(self.pos), L.ptr(ff.forces), L.ptr(self.dtau), L.ptr(noise), self.seed, 0,
               L.ptr(self.step_dev), self.node_offset, ff.N, st)
        L.call("fmd_increment_u64", L.ptr(self.step_dev), st)
        ff.compute(self.pos)
        self.launches_per_step = ff.launches_per_eval + 2


# ------------------------------------------------------------------------------------------------
# builders for synthetic systems (benchmarks / tests)
# ------------------------------------------------------------------------------------------------


def prior_terms_from_system(system: dict, n_mols: int, device) -> List[PriorTerm]:
    """Collated, condensed prior tables (what reference simulation/specialize_prior.py:112-207 plus
    collate produce) for `n_mols` copies of the synthetic molecule of `flashmd.synthetic`."""
    ty = system["atom_types"]
    n = ty.shape[0]
    st = system["stats"]
    out = []

    def collate(m):
        nt = m.shape[1]
        mp = np.concatenate([m + b * n for b in range(n_mols)], axis=1).astype(np.int32)
        mb = np.repeat(np.arange(n_mols, dtype=np.int32), nt)
        return torch.from_numpy(mp).to(device).contiguous(), torch.from_numpy(mb).to(device).contiguous()

    def rep(v):
        return torch.from_numpy(np.concatenate([np.asarray(v, dtype=np.float32)] * n_mols, 0)).to(device).contiguous()

    m = system["bonds"]
    tt = (ty[m[0]], ty[m[1]])
    mp, mb = collate(m)
    out.append(PriorTerm(L.PRIOR_BONDS, mp, mb, rep(st["bonds"]["k"][tt]), rep(st["bonds"]["x_0"][tt])))
    m = system["angles"]
    tt = (ty[m[0]], ty[m[1]], ty[m[2]])
    mp, mb = collate(m)
    out.append(PriorTerm(L.PRIOR_ANGLES, mp, mb, rep(st["angles"]["k"][tt]), rep(st["angles"]["x_0"][tt])))
    m = system["dihedrals"]
    c = (ty[m[1]], ty[m[2]])
    nd = st["dihedrals"]["n_degs"]
    k1 = np.stack([st["dihedrals"]["k1_central"][d][c] for d in range(nd)], 1)
    k2 = np.stack([st["dihedrals"]["k2_central"][d][c] for d in range(nd)], 1)
    mp, mb = collate(m)
    out.append(PriorTerm(L.PRIOR_DIHEDRALS, mp, mb, rep(k1), rep(k2), rep(st["dihedrals"]["v0_central"][c]), nd))
    m = system["nonbonded"]
    tt = (ty[m[0]], ty[m[1]])
    mp, mb = collate(m)
    out.append(PriorTerm(L.PRIOR_REPULSION, mp, mb, rep(st["repulsion"]["sigma"][tt])))
    return out


def random_schnet_tensors(seed: int, num_rbf: int = 50, hidden: int = 128, filters: int = 128, num_blocks: int = 3,
                          out_widths=(128, 64), embedding_size: int = 25) -> Dict[str, torch.Tensor]:
    """Random-init CGSchNet weights with the reference's initialisation (Xavier-uniform weights,
    zero biases: models/_module_init.py:4-28; N(0,1) embedding: torch.nn.Embedding default)."""
    g = torch.Generator().manual_seed(seed)

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand((o, i), generator=g) * 2 - 1) * a

    t = {"embedding": torch.randn((embedding_size, hidden), generator=g)}
    for l in range(num_blocks):
        t[f"b{l}.lin1_w"] = xavier(filters, hidden)
        t[f"b{l}.f0_w"] = xavier(filters, num_rbf)
        t[f"b{l}.f0_b"] = torch.zeros(filters)
        t[f"b{l}.f1_w"] = xavier(filters, filters)
        t[f"b{l}.lin2_w"] = xavier(hidden, filters)
        t[f"b{l}.lin2_b"] = torch.zeros(hidden)
        t[f"b{l}.lin_w"] = xavier(hidden, hidden)
        t[f"b{l}.lin_b"] = torch.zeros(hidden)
    widths = [hidden] + list(out_widths) + [1]
    for i in range(len(widths) - 1):
        t[f"out{i}_w"] = xavier(widths[i + 1], widths[i])
        if i < len(widths) - 2:
            t[f"out{i}_b"] = torch.zeros(widths[i + 1])
    return t


def _step_host(self, pos_h, vel_h, forces_h, energy_h):
    """End-to-end step with HOST state: pinned pos/vel/forces -> device, one BAOAB step, new
    pos/vel/forces + per-molecule potential back to the pinned host buffers (synchronises).

    With pinned buffers the copies are part of the captured graph (one graph per set of host buffers): one launch and one
    synchronisation per step instead of seven copy calls around the step graph."""
    bufs = (pos_h, vel_h, forces_h, energy_h)

    def body():
        self.pos.copy_(pos_h, non_blocking=True)
        self.vel.copy_(vel_h, non_blocking=True)
        self.ff.forces.copy_(forces_h, non_blocking=True)
        self._step_body()
        pos_h.copy_(self.pos, non_blocking=True)
        vel_h.copy_(self.vel, non_blocking=True)
        forces_h.copy_(self.ff.forces, non_blocking=True)
        energy_h.copy_(self.ff.energy, non_blocking=True)

    if not (self.use_graph and all(b.is_pinned() for b in bufs)):
        self.pos.copy_(pos_h, non_blocking=True)
        self.vel.copy_(vel_h, non_blocking=True)
        self.ff.forces.copy_(forces_h, non_blocking=True)
        self.step()
        pos_h.copy_(self.pos, non_blocking=True)
        vel_h.copy_(self.vel, non_blocking=True)
        forces_h.copy_(self.ff.forces, non_blocking=True)
        energy_h.copy_(self.ff.energy, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return
    key = tuple(b.data_ptr() for b in bufs)
    graphs = self.__dict__.setdefault("_host_graphs", {})
    g = graphs.get(key)
    if g is None:
        if self.graph is None:
            self._capture()          # warm-up (allocations, lazy module loads) happens there
        torch.cuda.synchronize()
        saved = (self.pos.clone(), self.vel.clone(), self.ff.forces.clone(), self.step_dev.clone())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        torch.cuda.synchronize()     # capture does not execute: device state and host buffers are untouched, restore anyway
        self.pos.copy_(saved[0]); self.vel.copy_(saved[1]); self.ff.forces.copy_(saved[2]); self.step_dev.copy_(saved[3])
        graphs[key] = g
    g.replay()
    torch.cuda.current_stream().synchronize()
    self.n_steps_done += 1


LangevinEngine.step_host = _step_host
