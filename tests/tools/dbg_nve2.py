"""Energy drift of NVE runs vs time step (second-order integrator: drift ~ dt^2 if forces = -grad E)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from helpers import dropin_model_from_golden, load_golden
from flashmd.simulation import NVESimulation
g = load_golden("schnet_n54_b4.npz")
for gptq in (None, "w16a16"):
    for seed in (100, 103):
        out = []
        for dt, n in ((0.001, 2000), (0.0005, 4000), (0.00025, 8000)):
            model, _, configs = dropin_model_from_golden(g)
            sim = NVESimulation(dt=dt, n_timesteps=n, save_interval=n // 100, save_energies=True, random_seed=1, device="cuda", gptq=gptq)
            sim.attach_model_and_configurations(model, configs, beta=1.67)
            d = sim.initial_data
            gen = torch.Generator().manual_seed(seed)
            d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]).to(d.pos.device)
            sim.simulate()
            e_tot = sim.simulated_potential + sim.simulated_kinetic_energies
            ke = sim.simulated_kinetic_energies.mean()
            out.append((np.abs(e_tot - e_tot[:, :1]).max() / ke, np.abs(e_tot[:, -10:].mean(axis=1) - e_tot[:, :10].mean(axis=1)).max() / ke))
        print(f"gptq={gptq} seed {seed}: (drift, slope) at dt, dt/2, dt/4 = " + ", ".join(f"({a:.5f}, {b:.5f})" for a, b in out))
