"""Per-shape timing of fmd_linear_x3 (BF16x3 fp32-emulation GEMM) at the edge-level shapes of the fp32 path; prints achieved GB/s
of the algorithmic traffic (X read + Y written + aux/res read)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L
L.load()
dev = "cuda"
E = 1_845_940
shapes = [("rbf->t tanh", E, 50, 128, True, False), ("t->W", E, 128, 128, False, False),
          ("gW->gT aux", E, 128, 128, False, True), ("gT->g_rbf", E, 128, 50, False, False),
          ("node", 34432, 128, 128, False, False)]
for name, M, K, N, tanh, use_aux in shapes:
    x = torch.randn((M, K), device=dev)
    w = torch.randn((K, N), device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    aux = torch.rand((M, N), device=dev) if use_aux else None
    y = torch.empty((M, N), device=dev)
    def run():
        L.call("fmd_linear_x3", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), M, N, K, None, L.ACT_TANH if tanh else 0,
               L.ptr(aux), None, L.stream_ptr())
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    a.record()
    for _ in range(n): run()
    z.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(z) / n
    byts = 4 * M * (K + N + (N if use_aux else 0))
    print(f"{name:14s} M={M} K={K} N={N}: {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s  {2*M*K*N/ms/1e9:6.1f} TF/s(fp32-equiv)")
