"""Tools (GPU box): the fused forward kernel on the bench's engine state - determinism, time per call, and the role
timeline of CTA 0 (traced instantiation of the kernel: wait / work cycles per role and tile; for the two MMA issuer rows
the columns are wait-for-operand / wait-for-accumulator)."""
import torch
from _bench_state import L, engine_state, print_timeline, time_ms
ff, w = engine_state()
k, l = w.k, 1
xh = ff.a[l]
out = torch.zeros_like(ff.m)
def fwd():
    L.call("fmd_filter_cfconv_fwd", L.ptr(ff.dist), L.ptr(ff.src), L.ptr(ff.dst), L.ptr(ff.seg_ptr), ff.N, ff.cap, L.ptr(ff.n_edges_dev),
           L.ptr(k[f"b{l}.f0_w.hp"]), L.ptr(k[f"b{l}.f0_b.h"]), L.ptr(k[f"b{l}.f1_w.h"]), L.ptr(w.centers), w.num_rbf, w.gamma, w.cutoff,
           L.ptr(xh), w.filters, L.ptr(out), L.ptr(ff.part), ff._st)
fwd(); torch.cuda.synchronize()
o1 = out.clone(); fwd(); torch.cuda.synchronize()
print("edges", ff.num_edges(), "deterministic:", torch.equal(out, o1), "finite:", bool(torch.isfinite(out).all()))
print("forward %.4f ms per call (includes the fix-up launch)" % time_ms(fwd))
trace = torch.zeros(9 * 64 * 3, dtype=torch.int64, device="cuda")
L.call("fmd_debug_set_trace_fwd", L.ptr(trace)); fwd(); torch.cuda.synchronize(); L.call("fmd_debug_set_trace_fwd", None)
print_timeline(trace, ["P", "T", "E0", "E1", "-", "-", "-", "M1", "M2"])
