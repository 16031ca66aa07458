import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L
L.load()
dev = "cuda"

def run(x, w, b=None):
    M, K = x.shape; N = w.shape[1]
    y = torch.full((M, N), float("nan"), device=dev)
    L.call("fmd_linear_tc", L.ptr(x), 0, L.ptr(w), 0, L.ptr(b), L.ptr(y), 0, M, N, K, None, 0, 0, 0, None, 0, None, 0, L.stream_ptr())
    torch.cuda.synchronize()
    return y

for (M, K, N) in [(128, 64, 64), (128, 128, 128), (1000, 128, 128), (300, 64, 128)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn((M, K), generator=g).to(dev); w = torch.randn((K, N), generator=g).to(dev)
    y = run(x, w)
    ref = x.double() @ w.double()
    print(M, K, N, "rel", float((y - ref).norm() / ref.norm()), "absmax y", float(y.abs().max()), "nan", int(torch.isnan(y).sum()))
    # which input does the output correlate with?
    if M == 128 and K == 64:
        xe = torch.zeros((M, K), device=dev); xe[5, 3] = 1.0
        we = torch.eye(K, N, device=dev)
        ye = run(xe, we)
        print("nonzeros", ye.nonzero().tolist()[:10], ye[ye != 0][:10].tolist())
for M in (34432,):
    x = torch.randn((M, 128), device=dev); w = torch.randn((128, 128), device=dev)
    for _ in range(3): run(x, w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = torch.empty((M, 128), device=dev)
        L.call("fmd_linear_tc", L.ptr(x), 0, L.ptr(w), 0, None, L.ptr(y), 0, M, 128, 128, None, 0, 0, 0, None, 0, None, 0, L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print("ms per call [K,N]", e0.elapsed_time(e1) / 20)
    wt = w.t().contiguous(); y = torch.empty((M, 128), device=dev)
    e0.record()
    for _ in range(20):
        L.call("fmd_linear_tc", L.ptr(x), 0, L.ptr(wt), 0, None, L.ptr(y), 0, M, 128, 128, None, 0, 0, 0, None, 0, None, 1, L.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print("ms per call [N,K]", e0.elapsed_time(e1) / 20)
