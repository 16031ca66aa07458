"""Shared by the tools scripts (GPU box): the bench's own engine state (269 beads x 128 molecules, 3 blocks, W16A16)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
import bench
from flashmd import _lib as L
from flashmd.engine import ForceField, SchNetWeights, random_schnet_tensors
from flashmd.neighbor_list import radius_graph_csr


def engine_state(exact=True, batch=128, n_beads=269):
    class A: pass
    args = A(); args.n_beads = n_beads; args.batch = batch; args.blocks = 3
    dev = torch.device("cuda")
    sysd, pos_np = bench.build_system(args, seed=0)
    B, n = batch, n_beads
    pos = torch.from_numpy(pos_np).reshape(B * n, 3).to(dev).contiguous()
    types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(dev)
    mol_ptr = (torch.arange(B + 1) * n).to(dev)
    w = SchNetWeights.from_flat(random_schnet_tensors(0, num_blocks=3), sysd["cutoff"], 50, dev)
    e0 = radius_graph_csr(pos, mol_ptr, sysd["cutoff"], idx_dtype=torch.int32)["edge_index"].shape[1]
    ff = ForceField(w, [], types, mol_ptr, precision="w16a16", edge_capacity=int(1.35 * e0) + 4096, exact_cutoff_grad=exact)
    ff.compute(pos); torch.cuda.synchronize()
    ff._st = L.stream_ptr(); ff._n = 0
    return ff, w


def time_ms(fn, reps=30):
    for _ in range(5): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def print_timeline(trace, names):
    tr = trace.cpu().numpy().reshape(9, 64, 3).astype(np.int64)
    line = []
    for r, nm in enumerate(names):
        rows = [i for i in range(8, 40) if tr[r, i, 0] > 0 and tr[r, i, 2] > 0]
        if not rows or nm == "-": continue
        a = np.mean([tr[r, i, 1] - tr[r, i, 0] for i in rows]); b = np.mean([tr[r, i, 2] - tr[r, i, 1] for i in rows])
        line.append(f"{nm} {a:5.0f}/{b:5.0f}")
    print(f"tile period {np.mean(np.diff(tr[0, 8:40, 2])):6.0f} cycles | wait/work per role: " + "  ".join(line))
    t0 = tr[0, 10, 0]
    for i in range(10, 15):
        print(f" tile {i}: " + "  ".join(f"{nm}:{tr[r, i, 0] - t0}/{tr[r, i, 1] - t0}/{tr[r, i, 2] - t0}"
                                        for r, nm in enumerate(names) if nm != "-" and tr[r, i, 0] > 0))
