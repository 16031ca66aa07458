"""Tools (GPU box): the node-level chain launches of one step in isolation (same stage lists as ForceField._schnet_tc):
time per launch, and a bit-compare of the chained result with the single-layer kernel."""
import torch
from _bench_state import L, engine_state, time_ms
ff, w = engine_state()
k = w.k
T, TC = L.ACT_TANH, L.ACT_TANH_CLAMPED
nb, no = w.num_blocks, w.num_out_layers
def fwd_chain(l):
    last = l == nb - 1
    stages = [dict(W=k[f"b{l}.lin2_w"], bias=k[f"b{l}.lin2_b"], epi=T, Y=ff.c[l]),
              dict(W=k[f"b{l}.lin_w"], bias=k[f"b{l}.lin_b"], res=ff.h[l], Y=ff.h[l + 1], round=last)]
    if not last:
        stages.append(dict(W=k[f"b{l + 1}.lin1_w"], Y=ff.a[l + 1]))
    else:
        for i in range(no - 1):
            stages.append(dict(W=k[f"out{i}_w.h"], bias=k.get(f"out{i}_b.h"), epi=TC, Y=ff.y[i], round=True))
    ff._chain(ff.m, stages)
def bwd_chain_first():
    stages = []
    for i in range(no - 2, 0, -1):
        stages.append(dict(W=k[f"out{i}_wT.h"], aux=ff.y[i - 1], round=True))
    stages.append(dict(W=k["out0_wT.h"], Y=ff.g_h[0]))
    l = nb - 1
    stages.append(dict(W=k[f"b{l}.lin_wT"], aux=ff.c[l]))
    stages.append(dict(W=k[f"b{l}.lin2_wT"], Y=ff.g_m))
    ff._chain(ff.g_y[no - 2], stages)
def bwd_chain(l):
    stages = [dict(W=k[f"b{l + 1}.lin1_wT"], res=ff.g_h[0], Y=ff.g_h[1]), dict(W=k[f"b{l}.lin_wT"], aux=ff.c[l]),
              dict(W=k[f"b{l}.lin2_wT"], Y=ff.g_m)]
    ff._chain(ff.g_a, stages)
tot = 0.0
for name, fn in [("fwd block 0 (3 stages)", lambda: fwd_chain(0)), ("fwd block 1 (3 stages)", lambda: fwd_chain(1)),
                 (f"fwd block 2 + output net ({2 + no - 1} stages)", lambda: fwd_chain(2)),
                 ("bwd out net + block 2", bwd_chain_first), ("bwd block 1 (3 stages)", lambda: bwd_chain(1)),
                 ("bwd block 0 (3 stages)", lambda: bwd_chain(0))]:
    t = time_ms(fn, 50)
    tot += t
    print(f"{name:40s} {t * 1000:7.1f} us")
print(f"sum of the 6 chain launches of a step: {tot:.4f} ms")
