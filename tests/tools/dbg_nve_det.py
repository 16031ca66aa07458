"""Is an NVE run through the fused engine bitwise reproducible (same process, same inputs)?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from helpers import dropin_model_from_golden, load_golden
from flashmd.simulation import NVESimulation
g = load_golden("schnet_n54_b4.npz")
for gptq in (None, "w16a16"):
    outs = []
    for rep in range(3):
        model, _, configs = dropin_model_from_golden(g)
        torch.manual_seed(3)
        sim = NVESimulation(dt=0.001, n_timesteps=400, save_interval=20, save_energies=True, random_seed=1, device="cuda", gptq=gptq)
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(103)
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]).to(d.pos.device)
        v0 = d.velocities.clone().cpu()
        sim.simulate()
        outs.append((np.array(sim.simulated_coords).copy(), np.array(sim.simulated_potential).copy(), v0))
    for rep in (1, 2):
        same_x = np.array_equal(outs[0][0], outs[rep][0]); same_v0 = torch.equal(outs[0][2], outs[rep][2])
        dx = np.abs(outs[0][0] - outs[rep][0])
        first = int(np.argmax(dx.reshape(dx.shape[0], dx.shape[1], -1).max(axis=(0, 2)) > 0)) if not same_x else -1
        print(f"gptq={gptq} rep {rep} vs 0: v0 equal {same_v0}; coords equal {same_x}; first differing frame {first}; max |dx| {dx.max():.3e}; U[0,:3] {outs[rep][1][0,:3]}")
