"""fp32 parity path at cfg2 shape: energy / force error of the first molecules against the fp64 oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from oracle import fmd_oracle as O
from helpers import rel_l2
from flashmd import synthetic
from flashmd.engine import ForceField, SchNetWeights, random_schnet_tensors
DEV = "cuda"
B, n, NM = 16, 269, 4
sysd = synthetic.synthetic_system(8, n, seed=0)
pos = torch.from_numpy(sysd["pos"]).repeat(B // 8, 1, 1).reshape(B * n, 3).to(DEV).contiguous()
types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(DEV)
ptr = (torch.arange(B + 1) * n).to(DEV)
w = SchNetWeights.from_flat(random_schnet_tensors(0), sysd["cutoff"], 50, DEV)
P = O.SchNetParams({k: (v if v is None else v.clone()) for k, v in random_schnet_tensors(0).items()} | {"out2_b": None}, 3, 3, sysd["cutoff"], 50)
refs = []
for m in range(NM):
    p0 = pos[m * n:(m + 1) * n].cpu()
    ei = torch.from_numpy(O.radius_graph(p0.numpy(), np.array([0, n]), sysd["cutoff"]))
    refs.append(O.schnet_energy_forces(P.to(torch.float64), p0.double(), types[:n].cpu(), torch.zeros(n, dtype=torch.long), 1, ei))
for label, tc, nodes in (("fma", False, "1"), ("x3 all", True, "1"), ("x3 edge-level only", True, "0")):
    os.environ["FMD_X3_NODES"] = nodes
    ff = ForceField(w, [], types, ptr, precision="fp32", use_tensor_cores=tc)
    e, f = ff.compute(pos)
    for m in range(NM):
        e0, f0 = refs[m]
        print(f"{label:20s} mol {m}: E={float(e[m]):.6f} ref={float(e0):.6f} rel_e={rel_l2(e[m:m+1].cpu(), e0):.2e} rel_f={rel_l2(f[m*n:(m+1)*n].cpu(), f0):.2e}")
