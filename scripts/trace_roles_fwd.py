"""Role timeline of the pipelined forward kernel (GPU box)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L
from flashmd.neighbor_list.torch_impl import radius_graph_csr
L.load()
dev = "cuda"
rng = np.random.default_rng(0)
sizes = [269] * 128
pos = torch.from_numpy(np.concatenate([rng.uniform(0, 24.0, size=(s, 3)) for s in sizes]).astype(np.float32)).to(dev)
ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)])).to(dev)
rc, R, F = 10.5, 50, 128
g = radius_graph_csr(pos, ptr, rc, idx_dtype=torch.int32)
src, dst, dist, seg = g["edge_index"][0].contiguous(), g["edge_index"][1].contiguous(), g["dist"], g["src_ptr"].contiguous()
E, N = src.numel(), pos.shape[0]
gen = torch.Generator().manual_seed(0)
wf0p = torch.zeros((F, 64), dtype=torch.float16); wf0p[:, :R] = (torch.rand((F, R), generator=gen) - 0.5).half()
wf0p = wf0p.to(dev); bf0h = torch.zeros(F, dtype=torch.float16, device=dev)
wf1h = ((torch.rand((F, F), generator=gen) - 0.5) * 0.3).half().to(dev)
x = torch.randn((N, F), generator=gen).to(dev)
centers = torch.linspace(0, rc, R).to(dev); gamma = float(-0.5 / (centers[1] - centers[0]) ** 2)
out = torch.zeros((N, F), device=dev); part = torch.zeros(((E + 127) // 128, F), device=dev)
trace = torch.zeros(9 * 64 * 3, dtype=torch.int64, device=dev)
def run():
    L.call("fmd_filter_cfconv_fwd2", L.ptr(dist), L.ptr(src), L.ptr(dst), L.ptr(seg), N, E, None, L.ptr(wf0p), L.ptr(bf0h),
           L.ptr(wf1h), L.ptr(centers), R, gamma, rc, L.ptr(x), F, L.ptr(out), L.ptr(part), L.stream_ptr())
for _ in range(3): run()
L.call("fmd_debug_set_trace", L.ptr(trace)); run(); torch.cuda.synchronize(); L.call("fmd_debug_set_trace", None)
t = trace.cpu().numpy().reshape(9, 64, 3).astype(np.int64)
names = ["P", "T", "E0", "E1", "M1", "M2"]
for r, nm in enumerate(names):
    rows = [i for i in range(8, 40) if t[r, i, 0] > 0 and t[r, i, 1] > 0]
    if not rows: continue
    if nm.startswith("M"):
        print(f"{nm:4s} wait {np.mean([t[r, i, 1] - t[r, i, 0] for i in rows]):8.0f}")
    else:
        print(f"{nm:4s} wait {np.mean([t[r, i, 1] - t[r, i, 0] for i in rows]):8.0f}  work {np.mean([t[r, i, 2] - t[r, i, 1] for i in rows]):8.0f}")
print("tile period (cycles):", np.mean(np.diff(t[0, 8:40, 2])))
