import sys, os, time, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/flash-molecular-dynamics_b200")
import bench
from flashmd.engine import ForceField, LangevinEngine, SchNetWeights, random_schnet_tensors, prior_terms_from_system
from flashmd.neighbor_list import radius_graph_csr
class A: pass
args = A(); args.n_beads = 269; args.batch = 128; args.blocks = 3
dev = torch.device("cuda")
sysd, pos_np = bench.build_system(args, seed=0)
B, n = 128, 269
for label, sl in (("all four priors", slice(0, 4)), ("without repulsion", slice(0, 3)), ("no priors", slice(0, 0))):
    pos = torch.from_numpy(pos_np).reshape(B * n, 3).to(dev).contiguous()
    types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(dev)
    mol_ptr = (torch.arange(B + 1) * n).to(dev)
    w = SchNetWeights.from_flat(random_schnet_tensors(0, num_blocks=3), sysd["cutoff"], 50, dev)
    pri = prior_terms_from_system(sysd, B, dev)[sl]
    e0 = radius_graph_csr(pos, mol_ptr, sysd["cutoff"], idx_dtype=torch.int32)["edge_index"].shape[1]
    ff = ForceField(w, pri, types, mol_ptr, precision="w16a16", edge_capacity=int(1.35 * e0) + 4096)
    masses = torch.from_numpy(sysd["masses"]).repeat(B)
    eng = LangevinEngine(ff, pos, torch.zeros((B * n, 3)), masses, torch.full((B,), 1.67), 0.004, 1.0, seed=1, use_graph=True)
    for _ in range(5): eng.step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): eng.step()     # few steps: without repulsion the chains collapse, keep the edge count comparable
    b.record(); torch.cuda.synchronize()
    print(f"{label:22s} {a.elapsed_time(b) / 20:.4f} ms/step  edges {ff.num_edges()}")
    # the same with the prior kernel on the main stream (no fork / join in the graph)
    ff.serial_priors = True
    eng.graph = None
    for _ in range(3): eng.step()
    torch.cuda.synchronize(); a.record()
    for _ in range(20): eng.step()
    b.record(); torch.cuda.synchronize()
    print(f"{'   serial priors':22s} {a.elapsed_time(b) / 20:.4f} ms/step  edges {ff.num_edges()}")
    # per-call times of an eager step
    from flashmd import _lib as L
    import flashmd.engine as E_mod
    tim = {}
    orig = L.call
    def timed(name, *aa):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig(name, *aa); e1.record(); tim.setdefault(name, []).append((e0, e1))
    E_mod.L.call = timed
    try:
        for _ in range(3): eng._step_body()
        torch.cuda.synchronize()
    finally:
        E_mod.L.call = orig
    print("   ", {k: round(sum(x.elapsed_time(y) for x, y in v) / 3, 4) for k, v in tim.items() if k in ("fmd_priors_csr", "fmd_segment_sum", "fmd_edge_grad_to_forces_csr", "fmd_filter_cfconv_fwd", "fmd_baoab_post")})
