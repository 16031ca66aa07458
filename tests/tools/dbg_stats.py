"""KE fluctuation statistic of the 10^4-step statistics test for several pinned velocity seeds."""
import os, sys
from copy import deepcopy
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-molecular-dynamics_b200"), os.path.join(ROOT, "tests")]
from helpers import dropin_model_from_golden, load_golden
from flashmd import synthetic
from flashmd.simulation import LangevinSimulation
ref = load_golden("stats_n54_10k.npz")
g = dict(load_golden("schnet_n54_b4.npz"))
n_ref = int(ref["n_mols"])
system = synthetic.synthetic_system(n_ref, 54, seed=0, target_degree=30.0)
g["sys.pos"], g["sys.cutoff"] = system["pos"], np.float64(system["cutoff"])
dt, friction, beta, n_steps, save_interval = (float(v) for v in ref["params"])
print("params", dt, friction, beta, n_steps, save_interval)
canon = ref["ke"][:, 100:].mean() * np.sqrt(2.0 / (3 * 54))
for gptq in ("w16a16", None):
    for seed in range(4):
        model, _, configs0 = dropin_model_from_golden(g)
        configs = [deepcopy(configs0[b]) for rep in range(8) for b in range(n_ref)]
        sim = LangevinSimulation(friction=friction, dt=dt, n_timesteps=int(n_steps), save_interval=int(save_interval),
                                 save_energies=True, random_seed=99, device="cuda", gptq=gptq)
        sim.attach_model_and_configurations(model, configs, beta=beta)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(seed)
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (beta * d.masses.cpu()))[:, None]).to(d.pos.device)
        sim.simulate()
        ke = sim.simulated_kinetic_energies[:, 100:]
        per_frame = ke.std(axis=0, ddof=1).mean()
        mol_mean = ke.mean(axis=1)
        print(f"gptq={gptq} seed {seed}: per_frame/canon {per_frame/canon:.3f} mean KE {ke.mean():.2f} mol-mean min/max {mol_mean.min():.1f}/{mol_mean.max():.1f} "
              f"median time-std {np.median(ke.std(axis=1)):.2f} max KE {ke.max():.1f} per-frame std q10/q50/q90 {np.quantile(ke.std(axis=0, ddof=1), [0.1,0.5,0.9]).round(2)}")
