"""Bitwise run-to-run determinism of the force field (fp32 and W16A16 paths) at a mid-size shape."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from flashmd import synthetic
from flashmd.engine import ForceField, SchNetWeights, random_schnet_tensors
DEV = "cuda"
for (B, n) in ((4, 54), (32, 269)):
    sysd = synthetic.synthetic_system(min(B, 8), n, seed=0)
    pos = torch.from_numpy(sysd["pos"]).repeat(B // min(B, 8), 1, 1).reshape(B * n, 3).to(DEV).contiguous()
    types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(DEV)
    ptr = (torch.arange(B + 1) * n).to(DEV)
    w = SchNetWeights.from_flat(random_schnet_tensors(0), sysd["cutoff"], 50, DEV)
    for prec in ("fp32", "w16a16"):
        ff = ForceField(w, [], types, ptr, precision=prec)
        e0, f0 = [t.clone() for t in ff.compute(pos)]
        bad = 0
        for it in range(20):
            e, f = ff.compute(pos)
            torch.cuda.synchronize()
            if not (torch.equal(e, e0) and torch.equal(f, f0)):
                bad += 1
                if bad == 1:
                    print(f"   first mismatch at iteration {it}: max|df|={float((f - f0).abs().max()):.3e} max|de|={float((e - e0).abs().max()):.3e} "
                          f"n_diff_f={int((f != f0).sum())}")
        print(f"B={B} n={n} {prec}: {bad}/20 runs differ from the first")
