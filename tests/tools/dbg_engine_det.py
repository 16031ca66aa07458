"""Bisect run-to-run differences: fresh ForceField + LangevinEngine from identical inputs, 20 NVE steps."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from copy import deepcopy
from helpers import dropin_model_from_golden, load_golden
from flashmd.simulation import LangevinSimulation
from flashmd.simulation.lowering import lower
from flashmd.engine import LangevinEngine
g = load_golden("schnet_n54_b4.npz")
model, _, configs = dropin_model_from_golden(g)
model = model.to("cuda")
data = LangevinSimulation.collate(deepcopy(configs)).to("cuda")
pos0 = data.pos.clone()
gen = torch.Generator().manual_seed(103)
vel0 = (torch.randn(pos0.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * data.masses.cpu()))[:, None]).cuda()
B = 4
for prec, x3 in (("fp32", "0"), ("fp32", "1")):
    os.environ["FMD_X3"] = x3
    for with_priors in (False,):
        for graph in (True,):
            res = []
            for rep in range(3):
                junk = torch.randn(64 * 1024 * 1024, device="cuda")   # dirty the allocator's free blocks
                del junk
                ff = lower(model, data, prec, True)
                if not with_priors:
                    ff.priors, ff.prior_csr = [], None
                f_init = None
                eng = LangevinEngine(ff, pos0, vel0, data.masses, torch.full((B,), 1.67), 0.001, 0.0, seed=1, use_graph=graph)
                f_init = ff.forces.clone()
                traj = []
                for k in range(20):
                    eng.step()
                    traj.append((eng.pos.clone(), ff.forces.clone(), ff.num_edges()))
                torch.cuda.synchronize()
                res.append((eng.pos.clone(), f_init, ff.energy.clone(), traj))
            for r in res[1:]:
                for k in range(20):
                    if not torch.equal(res[0][3][k][0], r[3][k][0]) or not torch.equal(res[0][3][k][1], r[3][k][1]):
                        print(f"   first differing step {k}: pos equal {torch.equal(res[0][3][k][0], r[3][k][0])} forces equal "
                              f"{torch.equal(res[0][3][k][1], r[3][k][1])} edges {res[0][3][k][2]} vs {r[3][k][2]} max|df| {float((res[0][3][k][1]-r[3][k][1]).abs().max()):.3e}")
                        break
            eq = [torch.equal(res[0][0], r[0]) for r in res[1:]]
            eqf = [torch.equal(res[0][1], r[1]) for r in res[1:]]
            print(f"x3={x3} {prec:7s} priors={with_priors!s:5s} graph={graph!s:5s}: pos equal {eq}, initial forces equal {eqf}, max|dpos| "
                  f"{max(float((res[0][0]-r[0]).abs().max()) for r in res[1:]):.3e}")
