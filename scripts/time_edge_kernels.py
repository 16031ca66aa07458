"""HBM-bound edge kernels at the cfg2 shape (128 x 269 beads, E ~ 1.85 M, F = 128, R = 50): time per launch (CUDA events)
and achieved GB/s of the ALGORITHMIC bytes against the measured HBM peak."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L, synthetic
from flashmd.engine import ForceField, SchNetWeights, random_schnet_tensors
L.load()
DEV = "cuda"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.6
B, n, F, R = 128, 269, 128, 50
sysd = synthetic.synthetic_system(8, n, seed=0)
pos = torch.from_numpy(sysd["pos"]).repeat(B // 8, 1, 1).reshape(B * n, 3).to(DEV).contiguous()
types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(DEV)
ptr = (torch.arange(B + 1) * n).to(DEV)
w = SchNetWeights.from_flat(random_schnet_tensors(0), sysd["cutoff"], R, DEV)
ff = ForceField(w, [], types, ptr, precision="fp32")
ff._st = L.stream_ptr(); ff._n = 0
ff.build_neighbor_list(pos)
E, N = ff.num_edges(), B * n
st = L.stream_ptr()
x = torch.randn((N, F), device=DEV)
g = torch.randn((N, F), device=DEV)
out = torch.empty((N, F), device=DEV)
W32 = torch.randn((E, F), device=DEV)
W16 = W32.half()
gW = torch.empty((E, F), device=DEV)
grbf = torch.randn((E, R), device=DEV)
rbf = torch.empty((E, R), device=DEV)
gd = torch.zeros(E, device=DEV)
rc = w.cutoff

def timeit(name, fn, byts, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    z.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(z) / n
    print(f"{name:34s} {ms*1e3:8.1f} us  {byts/1e6:8.1f} MB  {byts/ms/1e6:7.0f} GB/s  {100*byts/ms/1e6/peak:5.1f} % of {peak:.0f}")
    return ms

print(f"E={E} N={N}")
for name, Wt, b in (("cfconv_csr fp32 filter", W32, 4), ("cfconv_csr fp16 filter", W16, 2)):
    timeit(name, lambda Wt=Wt: L.call("fmd_cfconv_csr", L.ptr(x), L.ptr(Wt), L.dt_code(Wt), L.ptr(ff.dist), L.ptr(ff.dst),
                                      L.ptr(ff.seg_ptr), None, 4, N, E, F, rc, L.ptr(out), st),
           E * F * b + 8 * N * F + 8 * E + 4 * (N + 1))
timeit("cfconv_grad_filter (+exact cut)", lambda: L.call("fmd_cfconv_grad_filter", L.ptr(g), L.ptr(x), L.ptr(ff.dist), L.ptr(ff.src),
       L.ptr(ff.dst), 4, E, None, F, rc, L.ptr(gW), 0, L.ptr(W32), 0, L.ptr(gd), 1, st), 2 * E * F * 4 + 8 * N * F + 16 * E)
timeit("cfconv_grad_filter (no cut term)", lambda: L.call("fmd_cfconv_grad_filter", L.ptr(g), L.ptr(x), L.ptr(ff.dist), L.ptr(ff.src),
       L.ptr(ff.dst), 4, E, None, F, rc, L.ptr(gW), 0, None, 0, None, 0, st), E * F * 4 + 8 * N * F + 12 * E)
timeit("dist_rbf_cutoff_fwd", lambda: L.call("fmd_dist_rbf_cutoff_fwd", L.ptr(pos), L.ptr(ff.src), L.ptr(ff.dst), 4, E, None,
       L.ptr(w.centers), R, w.gamma, rc, None, L.ptr(rbf), st), 12 * N + 8 * E + 4 * E * R)
timeit("rbf_bwd (tiled)", lambda: L.call("fmd_rbf_bwd", L.ptr(ff.dist), L.ptr(grbf), None, E, None, L.ptr(w.centers), R, w.gamma, rc,
       L.ptr(gd), 1, st), 4 * E * R + 12 * E)
timeit("edge_grad_to_forces_csr", lambda: L.call("fmd_edge_grad_to_forces_csr", L.ptr(pos), L.ptr(ff.seg_ptr), L.ptr(ff.dst), L.ptr(ff.rev),
       L.ptr(ff.dist), L.ptr(gd), N, E, 1.0, L.ptr(ff.forces), 0, 0, st), 16 * E + 24 * N)
