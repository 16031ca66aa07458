"""W16A16 fused path against the oracle's W16A16 rounding model and against fp32, golden system + cfg2-shaped system."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from oracle import fmd_oracle as O
from helpers import golden_params, golden_system, load_golden, rel_l2
from flashmd.engine import ForceField, SchNetWeights
g = load_golden("schnet_n54_b4.npz")
pos, types, batch, ptr, B, n = golden_system(g)
tensors = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")}
w = SchNetWeights.from_flat(tensors, float(g["sys.cutoff"]), int(g["meta.hparams"][2]), "cuda")
P = golden_params(g)
ei = torch.from_numpy(O.radius_graph(pos.numpy(), ptr, float(g["sys.cutoff"])))
ff = ForceField(w, [], types.cuda(), torch.from_numpy(ptr).cuda(), precision="w16a16")
e, f = ff.compute(pos.cuda().contiguous())
for prec in ("w16a16", "fp32"):
    e_ref, f_ref = O.schnet_energy_forces(P, pos, types, batch, B, ei, precision=prec)
    print(f"fused W16A16 vs oracle {prec}: rel_f {rel_l2(f.cpu(), f_ref):.2e} rel_e {rel_l2(e.cpu(), e_ref):.2e}")
