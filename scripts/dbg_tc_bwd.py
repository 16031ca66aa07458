"""Diagnostics for the tcgen05 fused backward kernel (GPU box only): g_d against torch autograd of the
fp32 restatement of the same edge function."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
from flashmd import _lib as L  # noqa: E402
from flashmd.neighbor_list.torch_impl import radius_graph_csr  # noqa: E402


def run(sizes, box, rc, R=50, seed=0, timing=False):
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
    a = torch.randn((N, F), generator=gen).to(dev)
    gm = (torch.randn((N, F), generator=gen) * 0.1).to(dev)
    centers = torch.linspace(0.0, rc, R)
    gamma = float(-0.5 / (centers[1] - centers[0]) ** 2)
    wf0p = torch.zeros((F, 64), dtype=torch.float16)
    wf0p[:, :R] = wf0.half()
    wf0p, bf0h, wf1h, centers = wf0p.to(dev), bf0.half().to(dev), wf1.half().contiguous().to(dev), centers.to(dev)
    res = {"E": E}
    for exact in (1, 0):
        g_d = torch.full((max(E, 1),), float("nan"), device=dev)
        L.call("fmd_filter_cfconv_bwd", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None, L.ptr(wf0p), L.ptr(bf0h),
               L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc), L.ptr(a), L.ptr(gm), F, L.ptr(g_d), 0, exact,
               L.stream_ptr())
        torch.cuda.synchronize()
        d = dist.clone().requires_grad_(True)
        C = 0.5 * (torch.cos(d * np.pi / rc) + 1.0) * (d < rc)
        rbf = torch.exp(gamma * (d[:, None] - centers[None, :]) ** 2) * C[:, None]
        t = torch.tanh(rbf @ wf0p[:, :R].float().t() + bf0h.float())
        W = t @ wf1h.float().t()
        Cm = C if exact else C.detach()
        loss = (gm[src.long()] * a[dst.long()] * W * Cm[:, None]).sum()
        (ref,) = torch.autograd.grad(loss, d)
        g_d2 = torch.full((max(E, 1),), float("nan"), device=dev)
        L.call("fmd_filter_cfconv_bwd2", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None, L.ptr(wf0p), L.ptr(bf0h),
               L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc), L.ptr(a), L.ptr(gm), F, L.ptr(g_d2), 0, exact,
               L.stream_ptr())
        torch.cuda.synchronize()
        res[f"v2_exact{exact}"] = float((g_d2[:E] - ref).norm() / ref.norm().clamp_min(1e-30))
        res[f"v2_nan{exact}"] = int(torch.isnan(g_d2[:E]).sum())
        res[f"exact{exact}"] = float((g_d[:E] - ref).norm() / ref.norm().clamp_min(1e-30))
        res[f"nan{exact}"] = int(torch.isnan(g_d[:E]).sum())
    if timing:
        part = torch.zeros(((E + 127) // 128, F), device=dev)
        out = torch.zeros((N, F), device=dev)
        g_d = torch.zeros(E, device=dev)
        for name, fn in (
            ("fwd", lambda: L.call("fmd_filter_cfconv_fwd", L.ptr(dist), L.ptr(src), L.ptr(dst), L.ptr(seg), N, E, None,
                                   L.ptr(wf0p), L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc), L.ptr(a),
                                   F, L.ptr(out), L.ptr(part), None, None, L.stream_ptr())),
            ("bwd_exact", lambda: L.call("fmd_filter_cfconv_bwd", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None,
                                         L.ptr(wf0p), L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc),
                                         L.ptr(a), L.ptr(gm), F, L.ptr(g_d), 1, 1, L.stream_ptr())),
            ("bwd2_exact", lambda: L.call("fmd_filter_cfconv_bwd2", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None,
                                          L.ptr(wf0p), L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc),
                                          L.ptr(a), L.ptr(gm), F, L.ptr(g_d), 1, 1, L.stream_ptr())),
            ("bwd2_compat", lambda: L.call("fmd_filter_cfconv_bwd2", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None,
                                           L.ptr(wf0p), L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc),
                                           L.ptr(a), L.ptr(gm), F, L.ptr(g_d), 1, 0, L.stream_ptr())),
            ("bwd_compat", lambda: L.call("fmd_filter_cfconv_bwd", L.ptr(dist), L.ptr(src), L.ptr(dst), E, None,
                                          L.ptr(wf0p), L.ptr(bf0h), L.ptr(wf1h), L.ptr(centers), R, gamma, float(rc),
                                          L.ptr(a), L.ptr(gm), F, L.ptr(g_d), 1, 0, L.stream_ptr()))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name + "_ms"] = e0.elapsed_time(e1) / 20
    print(sizes[:3], "rc", rc, res)
    return res


if __name__ == "__main__":
    L.load()
    run([54] * 4, 14.0, 6.0)
    run([1, 2, 33, 7, 130, 64], 10.0, 3.5)
    run([300], 6.0, 50.0)
    run([269] * 128, 24.0, 10.5, timing=True)
