"""Diagnostics for the tcgen05 fused filter-network x CFConv forward kernel (GPU box only):
compares the dumped t, W and the reduced messages with a torch restatement of the W16A16 rounding
model (oracle/fmd_oracle.py schnet_energy, precision="w16a16")."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L  # noqa: E402
from flashmd.neighbor_list.torch_impl import radius_graph_csr  # noqa: E402


def run(sizes, box, rc, R=50, seed=0, verbose=True):
    dev = "cuda"
    rng = np.random.default_rng(seed)
    pos = torch.from_numpy(np.concatenate([rng.uniform(0, box, size=(s, 3)) for s in sizes]).astype(np.float32)).to(dev)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(sizes)])).to(dev)
    N = pos.shape[0]
    g = radius_graph_csr(pos, ptr, rc, idx_dtype=torch.int32)
    src, dst = g["edge_index"][0].contiguous(), g["edge_index"][1].contiguous()
    seg, dist = g["src_ptr"].contiguous(), g["dist"]
    E = src.numel()
    gen = torch.Generator().manual_seed(seed)
    F = 128
    wf0 = (torch.rand((F, R), generator=gen) * 2 - 1) * (6.0 / (F + R)) ** 0.5
    bf0 = (torch.rand(F, generator=gen) - 0.5) * 0.2
    wf1 = (torch.rand((F, F), generator=gen) * 2 - 1) * (6.0 / (2 * F)) ** 0.5
    x = torch.randn((N, F), generator=gen).to(dev)
    centers = torch.linspace(0.0, rc, R)
    gamma = float(-0.5 / (centers[1] - centers[0]) ** 2)
    wf0p = torch.zeros((F, 64), dtype=torch.float16)
    wf0p[:, :R] = wf0.half()
    wf0p, bf0h, wf1h, centers = wf0p.to(dev), bf0.half().to(dev), wf1.half().contiguous().to(dev), centers.to(dev)
    out = torch.full((N, F), float("nan"), device=dev)
    ntile = (max(E, 1) + 127) // 128
    part = torch.zeros((ntile, F), device=dev)
    dbg_t = torch.zeros((max(E, 1), F), dtype=torch.float16, device=dev)
    dbg_w = torch.zeros((max(E, 1), F), dtype=torch.float16, device=dev)
    L.call("fmd_filter_cfconv_fwd", L.ptr(dist), L.ptr(src), L.ptr(dst), L.ptr(seg), N, E, None, L.ptr(wf0p),
           L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc), L.ptr(x), F, L.ptr(out), L.ptr(part),
           L.ptr(dbg_t), L.ptr(dbg_w), L.stream_ptr())
    torch.cuda.synchronize()
    # torch restatement
    C = 0.5 * (torch.cos(dist * np.pi / rc) + 1.0) * (dist < rc)
    rbf = torch.exp(gamma * (dist[:, None] - centers[None, :]) ** 2) * C[:, None]
    pre = rbf.half().float() @ wf0p[:, :R].float().t() + bf0h.float()
    t = torch.tanh(pre).half()
    W = t.float() @ wf1h.float().t()
    m = torch.zeros((N, F), device=dev).index_add_(0, src.long(), x[dst.long()] * W * C[:, None])

    def rel(a, b):
        return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30))
    out2 = torch.full((N, F), float("nan"), device=dev)
    part2 = torch.zeros((ntile, F), device=dev)

    def fwd2():
        L.call("fmd_filter_cfconv_fwd2", L.ptr(dist), L.ptr(src), L.ptr(dst), L.ptr(seg), N, E, None, L.ptr(wf0p),
               L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc), L.ptr(x), F, L.ptr(out2), L.ptr(part2),
               L.stream_ptr())
    fwd2()
    torch.cuda.synchronize()
    res = {"E": E, "m2": rel(out2, m), "nan2": int(torch.isnan(out2).sum()), "t": rel(dbg_t[:E], t), "W": rel(dbg_w[:E], W), "m": rel(out, m),
           "t_max": float((dbg_t[:E].float() - t.float()).abs().max()) if E else 0.0,
           "nan": int(torch.isnan(out).sum())}
    if E > 100000:
        for _ in range(3):
            fwd2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fwd2()
        e1.record()
        torch.cuda.synchronize()
        res["fwd2_ms"] = e0.elapsed_time(e1) / 20
        out3 = out2.clone()
        fwd2()
        torch.cuda.synchronize()
        res["deterministic"] = bool(torch.equal(out2, out3))
    if verbose:
        print(sizes[:4], "rc", rc, res)
    return res


if __name__ == "__main__":
    L.load()
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        run([269] * 128, 24.0, 10.5)
        sys.exit(0)
    run([54] * 4, 14.0, 6.0)
    run([1, 2, 33, 7, 130, 64], 10.0, 3.5)
    run([300], 6.0, 50.0)          # degree 299: segments straddle 3 tiles
    run([269] * 128, 24.0, 7.5)
    run([269] * 128, 24.0, 10.5)
    run([100], 5.0, 3.0)
