"""Tools (GPU box): the fused backward kernel on the bench's engine state - determinism, time per call (exact and
Triton-compatible cut-off gradient), and the role timeline of CTA 0 (traced instantiation of the kernel)."""
import torch
from _bench_state import L, engine_state, print_timeline, time_ms
for exact in (1, 0):
    ff, w = engine_state(exact=bool(exact))
    k, l = w.k, 1
    ah, gh = ff.a[l], ff.g_m
    gd = torch.zeros_like(ff.g_pair)
    def bwd():
        L.call("fmd_filter_cfconv_bwd", L.ptr(ff.pair_dist), L.ptr(ff.pair_own), L.ptr(ff.pair_nbr), ff.pair_cap, L.ptr(ff.n_pairs_dev), L.ptr(k[f"b{l}.f0_w.hp"]),
               L.ptr(k[f"b{l}.f0_b.h"]), L.ptr(k[f"b{l}.f1_w.h"]), L.ptr(w.centers), w.num_rbf, w.gamma, w.cutoff, L.ptr(ah), L.ptr(gh),
               w.filters, L.ptr(gd), 0, exact, ff._st)
    bwd(); torch.cuda.synchronize()
    g1 = gd.clone(); bwd(); torch.cuda.synchronize()
    print("exact", exact, "edges", ff.num_edges(), "pairs", int(ff.n_pairs_dev.item()), "deterministic:", torch.equal(gd, g1), "finite:", bool(torch.isfinite(gd).all()),
          "backward %.4f ms per call" % time_ms(bwd))
trace = torch.zeros(9 * 64 * 3, dtype=torch.int64, device="cuda")
L.call("fmd_debug_set_trace_bwd", L.ptr(trace)); bwd(); torch.cuda.synchronize(); L.call("fmd_debug_set_trace_bwd", None)
print_timeline(trace, ["P", "E4", "G", "Gmid", "A", "B", "M1", "M3", "M4"])
