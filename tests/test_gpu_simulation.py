"""GPU tests of the reference-facing objects: model modules through the operator-level kernels
(`flashmd.kernels`, autograd), and LangevinSimulation / PTSimulation through the fused engine."""
import os
from copy import deepcopy

import numpy as np
import pytest
import torch

from helpers import dropin_model_from_golden, load_golden, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["schnet_n54_b4.npz", "schnet_n24_b3_l2.npz"])
def test_module_path_on_gpu_uses_kernels_and_matches_reference(name):
    """SumOut(GradientsOut(StandardSchNet), priors) on CUDA: fused RBF, CSR CFConv (+ exact cut-off gradient),
    fused tanh-linear operators with autograd; fp32 energies/forces within 1e-5 of the reference."""
    from flashmd.simulation import LangevinSimulation
    g = load_golden(name)
    model, schnet, configs = dropin_model_from_golden(g)
    data = LangevinSimulation.collate(deepcopy(configs)).to(DEV)
    model = model.to(DEV).eval()
    for p_ in model.parameters():          # weights are frozen in simulations (reference base.py:357-358);
        p_.requires_grad_(False)           # the operators provide d/d(pos) only
    torch.backends.cuda.matmul.allow_tf32 = False
    data = model(data)
    assert rel_l2(data.out["SchNet"]["energy"].cpu(), g["ref64.energy.SchNet"]) < 1e-5
    assert rel_l2(data.out["SchNet"]["forces"].cpu(), g["ref64.forces.SchNet"]) < 1e-5
    assert rel_l2(data.out["forces"].cpu(), g["ref64.forces.total"]) < 3e-5
    assert rel_l2(data.out["energy"].cpu(), g["ref64.energy.total"]) < 1e-5


def test_gptq_modules_on_gpu():
    from oracle import fmd_oracle as O
    from helpers import golden_params, golden_system
    from flashmd.models import GradientsOut, apply_gptq_w16a16_to_model, validate_gptq_w16a16
    from flashmd.simulation import LangevinSimulation
    g = load_golden("schnet_n54_b4.npz")
    model, schnet, configs = dropin_model_from_golden(g)
    net = GradientsOut(apply_gptq_w16a16_to_model(schnet)).to(DEV).eval()
    for p_ in net.parameters():
        p_.requires_grad_(False)
    assert validate_gptq_w16a16(net)
    data = LangevinSimulation.collate(deepcopy(configs)).to(DEV)
    data = net(data)
    pos, types, batch, ptr, B, n = golden_system(g)
    e_ref, f_ref = O.schnet_energy_forces(golden_params(g), pos, types, batch, B,
                                          torch.from_numpy(g["ref.schnet_edge_index"]), precision="w16a16")
    assert rel_l2(data.out["SchNet"]["forces"].cpu(), f_ref) < 1e-2
    assert rel_l2(data.out["SchNet"]["energy"].cpu(), e_ref) < 1e-2


@pytest.mark.parametrize("gptq", [None, "w16a16"])
def test_langevin_simulation_fused_engine(tmp_path, gptq):
    """Same object protocol as on CPU; the step runs as one CUDA-graph replay of the fused engine."""
    from flashmd.simulation import LangevinSimulation
    g = load_golden("schnet_n54_b4.npz")
    t = load_golden("langevin_n54_b4.npz")
    model, _, configs = dropin_model_from_golden(g)
    torch.manual_seed(1234)
    sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=10, save_interval=1, export_interval=10,
                             save_forces=True, save_energies=True, random_seed=103838, device=DEV, dtype="single",
                             filename="g", output_dir=str(tmp_path), specialize_priors=True, gptq=gptq,
                             noise_source="torch", create_checkpoints=True)
    sim.attach_model_and_configurations(model, configs, beta=1.67)
    sim.initial_data.velocities = torch.from_numpy(t["v0"]).to(DEV)
    noise = torch.from_numpy(t["noise"]).to(DEV)
    step = {"i": 0}

    def inject(eng):          # the reference drew its noise from a CPU generator: replay exactly those numbers
        eng.noise_buf.copy_(noise[step["i"]])
        step["i"] += 1
        eng.step()
    sim._engine_timestep = inject
    sim.simulate()
    assert sim.get_throughput_metrics()["path"] == "fused-engine"
    coords = np.load(tmp_path / "g_coords_0000.npy")
    tol = 1e-5 if gptq is None else 1e-4
    assert coords.shape == (4, 10, 54, 3) and rel_l2(coords, t["coords"]) < tol
    assert rel_l2(np.load(tmp_path / "g_potential_0000.npy"), t["potential"]) < (1e-4 if gptq is None else 1e-2)
    assert rel_l2(np.load(tmp_path / "g_kineticenergy_0000.npy"), t["kinetic"]) < (1e-4 if gptq is None else 1e-3)
    assert rel_l2(np.load(tmp_path / "g_forces_0000.npy"), t["forces"]) < (1e-3 if gptq is None else 2e-2)
    assert os.path.exists(tmp_path / "g_checkpoint_0000.pt")


def test_langevin_simulation_philox_temperature(tmp_path):
    """Default noise (in-kernel Philox): equipartition <KE> = 3 N / (2 beta) after equilibration."""
    from flashmd.simulation import LangevinSimulation
    g = load_golden("schnet_n54_b4.npz")
    model, _, configs = dropin_model_from_golden(g)
    cfgs = [deepcopy(configs[i % 4]) for i in range(32)]
    sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=3000, save_interval=10, save_energies=True,
                             random_seed=5, device=DEV, gptq="w16a16")
    sim.attach_model_and_configurations(model, cfgs, beta=1.67)
    sim.simulate()
    ke = sim.simulated_kinetic_energies[:, 100:]            # [n_sims, frames]
    expect = 1.5 * 54 / 1.67
    assert abs(float(ke.mean()) / expect - 1.0) < 0.03, (float(ke.mean()), expect)
    assert np.isfinite(sim.simulated_coords).all()


def test_pt_simulation_fused_engine(tmp_path):
    from flashmd.simulation import PTSimulation
    g = load_golden("schnet_n54_b4.npz")
    model, _, configs = dropin_model_from_golden(g)
    betas = [1.67, 1.42, 1.16]
    sim = PTSimulation(friction=1.0, dt=0.004, n_timesteps=200, save_interval=10, export_interval=100,
                       exchange_interval=20, save_energies=True, random_seed=7, device=DEV, filename="pt",
                       output_dir=str(tmp_path), gptq="w16a16", exchange_rng="philox")
    sim.attach_model_and_configurations(model, configs, betas=betas)
    sim.simulate()
    assert sim.get_throughput_metrics()["path"] == "fused-engine"
    assert sim.exchange_summary["attempted"] == 10 * 4 and 0 <= sim.exchange_summary["ratio"] <= 1
    # like the reference, the acceptance file is written BEFORE the exchange of the same step and is numbered
    # one higher than the coordinate file of that export: 4 rounds, then 5 rounds, x 4 pairs
    acc = np.load(tmp_path / "pt_acceptance_0001.npy")
    assert acc.shape == (3, 3) and acc.sum() == 4 * 4
    assert np.load(tmp_path / "pt_acceptance_0002.npy").sum() == 5 * 4
    assert np.load(tmp_path / "pt_coords_0001.npy").shape == (12, 10, 54, 3)
    assert np.isfinite(np.load(tmp_path / "pt_kineticenergy_0001.npy")).all()


@pytest.mark.parametrize("gptq", ["w16a16", None])
def test_trajectory_statistics_vs_reference_10k_steps(gptq):
    """north_star: temperature and RMSD statistics agree with the reference over 10^4-step trajectories.
    Reference side: tests/golden/stats_n54_10k.npz, 8 molecules x 10^4 BAOAB steps of the UNMODIFIED reference
    (oracle/make_golden.py --stats).  Ours: the same model, 64 replicas (8 copies of each start structure), own
    Philox noise -> only distributions can agree; tolerances are a few standard errors of the reference sample."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    from tests_helpers_kabsch import kabsch_rmsd, radius_of_gyration
    from flashmd import synthetic
    from flashmd.simulation import LangevinSimulation
    ref = load_golden("stats_n54_10k.npz")
    g = load_golden("schnet_n54_b4.npz")
    n_ref = int(ref["n_mols"])
    system = synthetic.synthetic_system(n_ref, 54, seed=0, target_degree=30.0)
    # the statistics run used the same seeded weights / bead types / prior tables as the golden model, on the
    # 8-molecule synthetic system (its own start structures and tuned cutoff)
    g = dict(g)
    g["sys.pos"], g["sys.cutoff"] = system["pos"], np.float64(system["cutoff"])
    assert np.array_equal(system["atom_types"], g["sys.atom_types"])
    model, _, configs0 = dropin_model_from_golden(g)
    configs = [deepcopy(configs0[b]) for rep in range(8) for b in range(n_ref)]
    dt, friction, beta, n_steps, save_interval = (float(v) for v in ref["params"])
    sim = LangevinSimulation(friction=friction, dt=dt, n_timesteps=int(n_steps), save_interval=int(save_interval),
                             save_energies=True, random_seed=99, device=DEV, gptq=gptq)
    sim.attach_model_and_configurations(model, configs, beta=beta)
    # pinned initial velocities (attach draws them from the global RNG, i.e. from whatever ran before this test)
    d = sim.initial_data
    gen = torch.Generator().manual_seed(0)
    d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (beta * d.masses.cpu()))[:, None]).to(d.pos.device)
    sim.simulate()
    assert sim.get_throughput_metrics()["path"] == "fused-engine"
    x, ke, pe = sim.simulated_coords, sim.simulated_kinetic_energies, sim.simulated_potential
    assert x.shape == (64, 1000, 54, 3) and np.isfinite(x).all()
    burn = 100
    # --- temperature: <KE> and its fluctuation (equipartition: mean 3N/2beta, relative std sqrt(2/(3N)))
    ke_ref, ke_our = ref["ke"][:, burn:], ke[:, burn:]
    # equipartition 3N/(2 beta) = 48.50; the 8-molecule reference sample (48.83) carries ~0.7 % standard error
    # (velocity autocorrelation ~ 1/friction = 25 saved frames), ours 64 molecules
    assert abs(ke_our.mean() / (1.5 * 54 / beta) - 1.0) < 0.01, ke_our.mean()
    assert abs(ke_our.mean() / ke_ref.mean() - 1.0) < 0.025, (ke_our.mean(), ke_ref.mean())
    # canonical fluctuation: std = mean * sqrt(2 / (3 N)) = 5.43.  The molecules keep relaxing during the run (RMSD
    # still grows at step 10^4), so the mean drifts and the GLOBAL sample std is inflated by a run-dependent 0-15 %
    # (reference sample: 6.01; ours 5.3-6.3 over repeated runs).  The fluctuation is therefore measured per saved
    # frame ACROSS the 64 independent molecules (insensitive to the drift) and averaged over frames.
    # This synthetic model at dt = 0.004 has rare heating events (a repulsive contact integrated with a large step): the
    # UNMODIFIED reference's own sample holds KE spikes up to 139 (mean 48.8, 34 of 8000 samples above 75) and ours, with
    # 8x the molecules, up to ~10^3 for some velocity seeds (tests/tools/dbg_stats.py).  Means and ROBUST widths are compared:
    # the median over frames of the cross-molecule std, and the MAD-based width of the whole sample.
    canon = ke_ref.mean() * np.sqrt(2.0 / (3 * 54))
    per_frame = np.median(ke_our.std(axis=0, ddof=1))
    assert abs(per_frame / canon - 1.0) < 0.06, (per_frame, canon)

    def mad_std(a):
        return 1.4826 * np.median(np.abs(a - np.median(a)))
    assert abs(mad_std(ke_our) / mad_std(ke_ref) - 1.0) < 0.12, (mad_std(ke_our), mad_std(ke_ref))
    # --- potential energy level
    se = ref["pe"][:, burn:].mean(axis=1).std() / np.sqrt(n_ref)
    assert abs(pe[:, burn:].mean() - ref["pe"][:, burn:].mean()) < 5 * se + 0.01 * abs(ref["pe"][:, burn:].mean())
    # --- structure: RMSD from the start structure at several times, radius of gyration
    x0 = system["pos"]
    for frame in (99, 299, 599, 999):
        ours = np.array([kabsch_rmsd(x[i, frame], x0[i % n_ref]) for i in range(64)])
        r = ref["rmsd"][:, frame]
        tol = 4.0 * r.std() / np.sqrt(n_ref) + 0.05 * r.mean()
        assert abs(ours.mean() - r.mean()) < tol, (frame, ours.mean(), r.mean(), tol)
    rg_our, rg_ref = radius_of_gyration(x[:, burn:]), ref["rg"][:, burn:]
    assert abs(rg_our.mean() / rg_ref.mean() - 1.0) < 0.05, (rg_our.mean(), rg_ref.mean())


@pytest.mark.parametrize("gptq,drift_bar,slope_bar", [(None, 2e-3, 1e-3), ("w16a16", 5e-3, 3e-3)])
def test_nve_energy_conservation_fused_engine(gptq, drift_bar, slope_bar):
    """Velocity Verlet through the fused engine: total energy is conserved, which checks that the analytic
    forces are the gradient of the energy the kernels report.
    Initial velocities are pinned (seeded CPU generator): attach draws them from the global RNG like the reference,
    and the size of the integration error depends strongly on the initial condition (close repulsive contacts): with
    the TRUE-fp32 FMA GEMMs the drift at dt = 0.001 is 1e-4 ... 2e-2 of the mean kinetic energy over five seeds, and it
    vanishes with the time step (seed 100: 1.8e-2, 7.7e-3, 8e-4 at dt, dt/2, dt/4; tests/tools/dbg_nve.py, dbg_nve2.py),
    which is the signature of integrator error, not of inconsistent forces.  The test therefore runs at dt/4."""
    from flashmd.simulation import NVESimulation
    g = load_golden("schnet_n54_b4.npz")
    for seed in (100, 103):
        model, _, configs = dropin_model_from_golden(g)
        sim = NVESimulation(dt=0.00025, n_timesteps=8000, save_interval=80, save_energies=True, random_seed=1, device=DEV,
                            gptq=gptq)
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(seed)
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]).to(d.pos.device)
        sim.simulate()
        assert sim.get_throughput_metrics()["path"] == "fused-engine"
        e_tot = sim.simulated_potential + sim.simulated_kinetic_energies      # [n_sims, frames]
        ke_scale = sim.simulated_kinetic_energies.mean()
        drift = np.abs(e_tot - e_tot[:, :1]).max() / ke_scale
        assert drift < drift_bar, (seed, drift)     # bounded fluctuation of the shadow Hamiltonian
        slope = np.abs(e_tot[:, -10:].mean(axis=1) - e_tot[:, :10].mean(axis=1)).max() / ke_scale
        assert slope < slope_bar, (seed, slope)     # no systematic drift


def test_fused_engine_graph_path_is_reproducible():
    """Two fresh engines from identical inputs give bitwise identical trajectories through the CUDA-graph path (the
    state saved around the capture used to race with the warm-up step)."""
    from copy import deepcopy
    from flashmd.engine import LangevinEngine
    from flashmd.simulation import LangevinSimulation
    from flashmd.simulation.lowering import lower
    g = load_golden("schnet_n54_b4.npz")
    model, _, configs = dropin_model_from_golden(g)
    model = model.to(DEV)
    data = LangevinSimulation.collate(deepcopy(configs)).to(DEV)
    gen = torch.Generator().manual_seed(5)
    vel = (torch.randn(data.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * data.masses.cpu()))[:, None]).to(DEV)
    for prec in ("fp32", "w16a16"):
        outs = []
        for rep in range(3):
            junk = torch.randn(16 * 1024 * 1024, device=DEV)   # dirty the allocator's free blocks between runs
            del junk
            ff = lower(model, data, prec, True)
            eng = LangevinEngine(ff, data.pos, vel, data.masses, torch.full((4,), 1.67), 0.004, 1.0, seed=11, use_graph=True)
            eng.run(25)
            torch.cuda.synchronize()
            outs.append((eng.pos.clone(), eng.vel.clone(), ff.energy.clone()))
        for o in outs[1:]:
            assert all(torch.equal(a, b) for a, b in zip(outs[0], o)), prec


@pytest.mark.parametrize("gptq", [None, "w16a16"])
def test_long_run_is_bitwise_reproducible(gptq):
    """Race detector for the warp-specialised tensor-core kernels: 64 molecules x 1500 BAOAB steps (Philox noise, pinned
    velocities) twice through the CUDA-graph path must end in bitwise identical states.  (A generic->async proxy fence
    moved from the writing threads to the MMA-issuing thread passed every parity test but failed this kind of check.)"""
    from flashmd import synthetic
    from flashmd.simulation import LangevinSimulation
    g = dict(load_golden("schnet_n54_b4.npz"))
    system = synthetic.synthetic_system(8, 54, seed=0, target_degree=30.0)
    g["sys.pos"], g["sys.cutoff"] = system["pos"], np.float64(system["cutoff"])
    outs = []
    for rep in range(2):
        model, _, configs0 = dropin_model_from_golden(g)
        configs = [deepcopy(configs0[b]) for r_ in range(8) for b in range(8)]
        sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=1500, save_interval=500, save_energies=True,
                                 random_seed=5, device=DEV, gptq=gptq)
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(17)
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]).to(d.pos.device)
        sim.simulate()
        assert sim.get_throughput_metrics()["path"] == "fused-engine"
        outs.append((np.array(sim.simulated_coords).copy(), np.array(sim.simulated_potential).copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_async_save_points_are_consistent_snapshots():
    """Save points of the fused path are copied to the host asynchronously (device snapshot -> pinned staging on a side
    stream) while the next steps already run: saving EVERY step must give the same frames as saving every 5th step, and
    forces / energies / kinetic energies must belong to the same step as the coordinates."""
    from flashmd.simulation import LangevinSimulation
    g = load_golden("schnet_n54_b4.npz")
    runs = {}
    for si in (1, 5):
        model, _, configs = dropin_model_from_golden(g)
        sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=60, save_interval=si, export_interval=None,
                                 save_forces=True, save_energies=True, random_seed=21, device=DEV, gptq="w16a16")
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        d = sim.initial_data
        gen = torch.Generator().manual_seed(1)
        d.velocities = (torch.randn(d.pos.shape, generator=gen) * torch.sqrt(1.0 / (1.67 * d.masses.cpu()))[:, None]).to(d.pos.device)
        sim.simulate()
        assert sim.get_throughput_metrics()["path"] == "fused-engine"
        runs[si] = (np.array(sim.simulated_coords), np.array(sim.simulated_forces), np.array(sim.simulated_potential),
                    np.array(sim.simulated_kinetic_energies))
    for a, b in zip(runs[1], runs[5]):
        assert a.shape[1] == 60 and b.shape[1] == 12
        assert np.array_equal(a[:, 4::5], b)
    # consecutive frames differ (the snapshots are not all the same late state)
    assert np.abs(np.diff(runs[1][0], axis=1)).max(axis=(0, 2, 3)).min() > 0


def test_sharded_noise_is_keyed_by_the_global_bead_index():
    """Two half batches with node_offset = 0 / half draw exactly the noise of the full batch (replicas sharded
    over ranks used to share one stream); a wrong offset gives a different trajectory."""
    from flashmd.engine import LangevinEngine
    from test_gpu_parity import _engine_from_golden
    g = load_golden("schnet_n54_b4.npz")
    B, n = 4, 54
    masses = torch.from_numpy(g["sys.masses"]).repeat(B)
    ff, pos = _engine_from_golden(g, "fp32", priors=True)
    full = LangevinEngine(ff, pos, torch.zeros(B * n, 3), masses, torch.full((B,), 1.67), 0.004, 1.0, seed=5, use_graph=False)
    full.run(5)
    gh = {k: (v[2:] if k == "sys.pos" else v) for k, v in g.items()}     # molecules 2, 3
    ffh, posh = _engine_from_golden(gh, "fp32", priors=True)
    for off, same in ((2 * n, True), (0, False)):
        half = LangevinEngine(ffh, posh, torch.zeros(2 * n, 3), masses[: 2 * n], torch.full((2,), 1.67), 0.004, 1.0, seed=5,
                              use_graph=False, node_offset=off)
        half.run(5)
        d = float((half.pos - full.pos[2 * n:]).abs().max())
        assert (d < 1e-5) == same, (off, d)


def test_nccl_sharded_exchange_two_gpus():
    """Sharded replica exchange over NCCL (device-side decisions, static peer exchange) == the single-process reference
    exchange, and a sharded PTSimulation on the fused engine; needs two GPUs (skipped on a one-GPU box)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29531", os.path.join(root, "tests", "tools", "dist_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded PTSimulation ok on all ranks: True" in r.stdout


def test_extra_prior_classes_run_on_the_fused_step():
    """GeneralBonds / GeneralAngles / Polynomial (bond lengths) / QuarticAngles / RestrictedQuartic / shifted impropers are
    lowered to the owner-computes prior kernel (one launch for all prior classes) instead of dropping the model to the
    module path; energies and forces against the UNMODIFIED reference (tests/golden/extra_priors_n54_b4.npz), alone and
    all six together next to the benchmark's four."""
    from helpers import extra_prior_configs, extra_prior_objects
    from flashmd.models import GradientsOut, SumOut
    from flashmd.simulation import LangevinSimulation
    from flashmd.simulation.lowering import lower
    t = load_golden("extra_priors_n54_b4.npz")
    priors, system = extra_prior_objects()
    e_sum, f_sum, nls = 0.0, 0.0, {}
    for name, (prior, mapping, order) in priors.items():
        configs = extra_prior_configs(prior, mapping, order, system)
        data = LangevinSimulation.collate(configs).to(DEV)
        model = SumOut(torch.nn.ModuleDict({prior.name: GradientsOut(prior)})).to(DEV)
        ff = lower(model, data, "fp32")
        e, f = ff.compute(data.pos.float().contiguous())
        for what, val in (("energy", e), ("forces", f)):
            r32, r64 = t[f"ref32.{name}.{what}"], t[f"ref64.{name}.{what}"]
            tol = max(1e-5, 2.0 * rel_l2(r32, r64))
            assert rel_l2(val.cpu(), r64) < tol, (name, what, rel_l2(val.cpu(), r64), tol)
        e_sum, f_sum = e_sum + t[f"ref64.{name}.energy"], f_sum + t[f"ref64.{name}.forces"]
        nls[prior.name] = (mapping, order)
    # all six in ONE model (two angle-like classes share the angle table, impropers next to nothing else)
    from flashmd.data import AtomicData
    from flashmd.neighbor_list import make_neighbor_list
    ty = system["atom_types"]
    configs = [AtomicData.from_points(pos=torch.from_numpy(system["pos"][b].copy()), atom_types=torch.from_numpy(ty),
                                      masses=torch.from_numpy(system["masses"]),
                                      neighborlist={nm: make_neighbor_list(nm, o, torch.from_numpy(m)) for nm, (m, o) in nls.items()})
               for b in range(4)]
    data = LangevinSimulation.collate(configs).to(DEV)
    model = SumOut(torch.nn.ModuleDict({p.name: GradientsOut(p) for p, _, _ in priors.values()})).to(DEV)
    ff = lower(model, data, "fp32")
    e, f = ff.compute(data.pos.float().contiguous())
    assert rel_l2(e.cpu(), e_sum) < 3e-5 and rel_l2(f.cpu(), f_sum) < 3e-5


def test_overdamped_integrator_on_the_fused_engine(tmp_path):
    """OverdampedSimulation's update (reference simulation/langevin.py:361-414: D = 1 / (beta friction), x += F D dt +
    sqrt(2 D dt) xi) on the fused engine, step for step against the UNMODIFIED reference's trajectory
    (tests/golden/integrators_n54_b4.npz) with the reference's own noise stream (torch.Generator(seed) on the CPU);
    then through OverdampedSimulation itself (Philox noise, CUDA graph): fused path, finite, diffusing."""
    from flashmd.engine import OverdampedEngine
    from flashmd.simulation import OverdampedSimulation
    from test_gpu_parity import _engine_from_golden
    g = load_golden("schnet_n54_b4.npz")
    t = load_golden("integrators_n54_b4.npz")
    _, dt_od, fr_od, beta, seed, _ = (float(v) for v in t["params"])
    B, n = 4, 54
    ff, pos = _engine_from_golden(g, "fp32", priors=True)
    eng = OverdampedEngine(ff, pos, torch.full((B,), beta), dt_od, fr_od, use_graph=False)
    gen = torch.Generator(device="cpu").manual_seed(int(seed))
    for s_ in range(t["overdamped.coords"].shape[1]):
        noise = torch.empty((B * n, 3)).normal_(generator=gen)
        eng.step(noise=noise.to(DEV).contiguous())
        assert rel_l2(eng.pos.view(B, n, 3).cpu(), t["overdamped.coords"][:, s_]) < 1e-5, s_
        assert rel_l2(ff.energy.cpu(), t["overdamped.potential"][:, s_]) < 1e-4, s_
    model, _, configs = dropin_model_from_golden(g)
    sim = OverdampedSimulation(dt=dt_od, friction=fr_od, n_timesteps=200, save_interval=20, export_interval=200,
                               save_energies=True, random_seed=3, device=DEV, filename="od", output_dir=str(tmp_path))
    sim.attach_model_and_configurations(model, configs, beta=beta)
    sim.simulate()
    assert sim.get_throughput_metrics()["path"] == "fused-engine"
    x = np.load(tmp_path / "od_coords_0000.npy")           # [n_sims, frames, n_atoms, 3]
    assert x.shape == (4, 10, 54, 3) and np.isfinite(x).all() and np.abs(x[:, -1] - x[:, 0]).max() > 1e-2
