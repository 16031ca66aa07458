"""Time the UNMODIFIED reference (its Triton GPU path) on the same B200 and the same synthetic workload as
bench.py — the "reference Triton path" baseline of BASELINE.json's north_star.  Not part of bench.py's
contract (the --impl reference arm is the CPU path); results go to profiles/.

  * reference = `pip install --no-index --no-deps --target baseline/_ref /root/reference` (git-ignored, travels
    to the GPU box), imported behind oracle/shims (nvtx, torch_geometric, jsonargparse, ruamel.yaml);
  * `torch_cluster.radius_graph` (un-vendored CUDA kernel, absent here) is replaced by OUR radius-graph kernels
    called through the C ABI — i.e. the reference is not charged for a slow stand-in;
  * `math` is injected into flashmd.kernels.cfconv_kernels (the shipped fused-RBF backward raises NameError);
  * timing = the reference's own second-half throughput metric (simulation/base.py:748-787).
Usage: python scripts/bench_triton_reference.py [--batch 128] [--n-beads 269] [--steps 40] [--gptq w16a16|none]
       [--compile 0|1]"""
import argparse
import ctypes
import json
import math
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(1, os.path.join(ROOT, "baseline", "_ref"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--n-beads", type=int, default=269)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--gptq", default="w16a16")
ap.add_argument("--compile", type=int, default=0)
ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"],
                help="cpu = the reference's --disable_optim path (MLCG_*=0, gptq=None, no compile) on the host cores")
ap.add_argument("--threads", type=int, default=0)
args = ap.parse_args()
CPU = args.device == "cpu"
if CPU:
    # scripts/nvt_langevin.py:6-17 (--disable_optim): every MLCG_* toggle off BEFORE flashmd is imported
    for k in ("MLCG_USE_TRITON_MESSAGE_PASSING", "MLCG_USE_FUSED_RBF", "MLCG_USE_FUSED_TANH_LINEAR", "MLCG_USE_CSR",
              "MLCG_USE_SRC_CSR_GRAD_X"):
        os.environ[k] = "0"
    args.gptq, args.compile = "none", 0
    torch.set_num_threads(args.threads or (os.cpu_count() or 1))

# ---- torch_cluster stand-in backed by libfmd_b200.so (raw ctypes: the drop-in package must not be imported here,
#      it has the same top-level name as the reference)
_lib = None if CPU else ctypes.CDLL(os.path.join(ROOT, "flash-molecular-dynamics_b200", "csrc", "libfmd_b200.so"))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
if not CPU:
  _lib.fmd_nl_count.argtypes = [vp, vp, ci, ci, ci, cf, ci, vp, vp]
  _lib.fmd_exclusive_scan_i32.argtypes = [vp, vp, ci, vp, vp]
  _lib.fmd_nl_fill.argtypes = [vp, vp, ci, ci, ci, cf, ci, vp, ci, vp, vp, ci, vp, vp]


def radius_graph_cuda(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target", num_workers=1,
                      batch_size=None):
    assert not loop and x.is_cuda
    N = x.shape[0]
    counts = torch.bincount(batch, minlength=int(batch[-1]) + 1)
    ptr = torch.zeros(counts.numel() + 1, dtype=torch.int32, device=x.device)
    ptr[1:] = torch.cumsum(counts, 0)
    B, max_mol = counts.numel(), int(counts.max())
    st = torch.cuda.current_stream().cuda_stream
    pos = x.contiguous().float()
    deg = torch.empty(N, dtype=torch.int32, device=x.device)
    seg = torch.zeros(N + 1, dtype=torch.int32, device=x.device)
    ws = torch.empty(N // 1024 + 4, dtype=torch.int32, device=x.device)
    assert _lib.fmd_nl_count(pos.data_ptr(), ptr.data_ptr(), B, N, max_mol, float(r), int(max_num_neighbors),
                             deg.data_ptr(), st) == 0
    assert _lib.fmd_exclusive_scan_i32(deg.data_ptr(), seg.data_ptr(), N, ws.data_ptr(), st) == 0
    E = int(seg[N])
    src = torch.empty(E, dtype=torch.int64, device=x.device)
    dst = torch.empty(E, dtype=torch.int64, device=x.device)
    assert _lib.fmd_nl_fill(pos.data_ptr(), ptr.data_ptr(), B, N, max_mol, float(r), int(max_num_neighbors),
                            seg.data_ptr(), E, src.data_ptr(), dst.data_ptr(), 8, None, st) == 0
    ei = torch.stack([src, dst])
    return ei if flow == "target_to_source" else ei.flip(0)


import torch_cluster  # noqa: E402  (the shim; on the CPU its masked-distance radius_graph is used as is)
if not CPU:
    torch_cluster.radius_graph = radius_graph_cuda

import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("fmd_synthetic", os.path.join(ROOT, "flash-molecular-dynamics_b200", "flashmd", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

import flashmd  # noqa: E402  (the reference)
assert "baseline/_ref" in flashmd.__file__, flashmd.__file__
import flashmd.kernels.cfconv_kernels as _ck  # noqa: E402
_ck.math = math
import flashmd.neighbor_list.torch_impl as _ti  # noqa: E402
if hasattr(_ti, "radius_graph") and not CPU:
    _ti.radius_graph = radius_graph_cuda
from flashmd.data import AtomicData  # noqa: E402
from flashmd.models import CosineCutoff, GaussianBasis, GradientsOut, StandardSchNet, SumOut  # noqa: E402
from flashmd.neighbor_list.neighbor_list import make_neighbor_list  # noqa: E402
from flashmd.prior import Dihedral, HarmonicAngles, HarmonicBonds, Repulsion  # noqa: E402
from flashmd.simulation import LangevinSimulation  # noqa: E402

torch.set_float32_matmul_precision("high")     # scripts/nvt_langevin.py:38
B, n = args.batch, args.n_beads
n_distinct = min(B, 16)
system = syn.synthetic_system(n_distinct, n, seed=0)
ty, st_ = system["atom_types"], system["stats"]
sb, sa, sd, sr = {}, {}, {}, {}
for i, j in system["bonds"].T:
    sb[(int(ty[i]), int(ty[j]))] = {"k": float(st_["bonds"]["k"][ty[i], ty[j]]), "x_0": float(st_["bonds"]["x_0"][ty[i], ty[j]])}
for i, j, k in system["angles"].T:
    key = (int(ty[i]), int(ty[j]), int(ty[k]))
    sa[key] = {"k": float(st_["angles"]["k"][key]), "x_0": float(st_["angles"]["x_0"][key])}
nd = st_["dihedrals"]["n_degs"]
for i, j, k, l in system["dihedrals"].T:
    c = (int(ty[j]), int(ty[k]))
    sd[(int(ty[i]), int(ty[j]), int(ty[k]), int(ty[l]))] = {
        "k1s": {f"k1_{q + 1}": float(st_["dihedrals"]["k1_central"][(q,) + c]) for q in range(nd)},
        "k2s": {f"k2_{q + 1}": float(st_["dihedrals"]["k2_central"][(q,) + c]) for q in range(nd)},
        "v_0": float(st_["dihedrals"]["v0_central"][c])}
for i, j in system["nonbonded"].T:
    sr[(int(ty[i]), int(ty[j]))] = {"sigma": float(st_["repulsion"]["sigma"][ty[i], ty[j]])}
rc = system["cutoff"]
torch.manual_seed(0)
schnet = StandardSchNet(GaussianBasis(CosineCutoff(0.0, rc), num_rbf=50), CosineCutoff(0.0, rc),
                        output_hidden_layer_widths=[128, 64], hidden_channels=128, embedding_size=syn.N_BEAD_TYPES + 1,
                        num_filters=128, num_interactions=3)
model = SumOut(torch.nn.ModuleDict({
    "SchNet": GradientsOut(schnet), "bonds": GradientsOut(HarmonicBonds(sb)), "angles": GradientsOut(HarmonicAngles(sa)),
    "dihedrals": GradientsOut(Dihedral(sd, n_degs=nd)), "repulsion": GradientsOut(Repulsion(sr))}))
configs = []
for b in range(B):
    nls = {"bonds": make_neighbor_list("bonds", 2, torch.from_numpy(system["bonds"])),
           "angles": make_neighbor_list("angles", 3, torch.from_numpy(system["angles"])),
           "dihedrals": make_neighbor_list("dihedrals", 4, torch.from_numpy(system["dihedrals"])),
           "repulsion": make_neighbor_list("repulsion", 2, torch.from_numpy(system["nonbonded"]))}
    configs.append(AtomicData.from_points(pos=torch.from_numpy(system["pos"][b % n_distinct].copy()),
                                          atom_types=torch.from_numpy(ty), masses=torch.from_numpy(system["masses"]),
                                          neighborlist=nls))
tmp = tempfile.mkdtemp()
gptq = None if args.gptq.lower() == "none" else args.gptq
sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=args.steps, save_interval=args.steps, export_interval=args.steps,
                         random_seed=103838, device=args.device, dtype="single", filename="ref", output_dir=tmp,
                         specialize_priors=True, compile_model=bool(args.compile), gptq=gptq)
t0 = time.perf_counter()
sim.attach_model_and_configurations(model, configs, beta=1.67)
t_attach = time.perf_counter() - t0
t0 = time.perf_counter()
sim.simulate()
t_sim = time.perf_counter() - t0
m = sim.get_throughput_metrics()
out = {"impl": "reference-cpu-disable_optim" if CPU else "reference-triton", "threads": torch.get_num_threads(), "gptq": gptq, "compile_model": bool(args.compile), "batch": B, "n_beads": n,
       "steps": args.steps, "attach_s": t_attach, "simulate_s": t_sim,
       "notes": ("unmodified reference, --disable_optim semantics; torch_cluster.radius_graph = oracle/shims stand-in"
                 if CPU else "radius_graph = our CUDA kernel (torch_cluster absent); math injected into cfconv_kernels; TF32 matmul"),
       "metrics": {k: (float(v) if isinstance(v, (int, float, np.floating)) else str(v)) for k, v in (m or {}).items()}}
print(json.dumps(out))
