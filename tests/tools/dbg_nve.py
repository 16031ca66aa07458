"""NVE energy conservation of the fp32 path (tests/test_gpu_simulation.py::test_nve_energy_conservation_fused_engine)
for the GEMM variants: prints drift / slope in units of the mean kinetic energy."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from helpers import dropin_model_from_golden, load_golden
from flashmd.simulation import NVESimulation
g = load_golden("schnet_n54_b4.npz")
for label, env in (("fma", {"FMD_X3": "0"}), ("x3 edge-level", {"FMD_X3": "1", "FMD_X3_NODES": "0"})):
    os.environ.update(env)
    for rep in range(5):
        model, _, configs = dropin_model_from_golden(g)
        torch.manual_seed(3)
        sim = NVESimulation(dt=0.001, n_timesteps=2000, save_interval=20, save_energies=True, random_seed=1, device="cuda", gptq=None)
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(100 + rep)
        scale = torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * scale).to(d.pos.device)
        sim.simulate()
        e_tot = sim.simulated_potential + sim.simulated_kinetic_energies
        ke = sim.simulated_kinetic_energies.mean()
        drift = np.abs(e_tot - e_tot[:, :1]).max() / ke
        slope = np.abs(e_tot[:, -10:].mean(axis=1) - e_tot[:, :10].mean(axis=1)).max() / ke
        print(f"{label:14s} rep {rep}: drift {drift:.5f} slope {slope:.5f}  e_tot[0,:3]={e_tot[0,:3]} v0sum={float(np.abs(sim.simulated_kinetic_energies[:,0]).sum()):.6f}")
