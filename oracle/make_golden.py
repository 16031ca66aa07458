"""Generate golden vectors by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE ONLY.

Run (in the build container, where /root/reference exists):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference/src/flashmd) is imported as-is behind the stand-ins in
`oracle/shims/` for the packages that are absent from this image (nvtx, torch_geometric,
torch_cluster, jsonargparse, ruamel.yaml).  It runs its pure-PyTorch path (`--disable_optim`
semantics: all MLCG_* toggles "0", gptq=None, compile off — scripts/nvt_langevin.py:6-17,40-60),
the only path that executes without a GPU.  Nothing from the reference is copied into the repo;
only small input/output arrays are stored.

Files written:
  schnet_n54_b4.npz    StandardSchNet energies/forces + every prior term + totals, 4 x 54 beads
  schnet_n24_b3_l2.npz a second shape (3 x 24 beads, 2 interaction blocks, F=64, R=20, biases != 0)
  langevin_n54_b4.npz  10 BAOAB steps of LangevinSimulation with the noise that was drawn
  pt_n24.npz           PTSimulation exchange bookkeeping (pairs, swaps) for 3 betas x 2 configs
  known_answers.npz    cutoff / basis known-answer values from the reference's own unit tests
  schnet_n40_b2_l5.npz (--l5) cfg5's depth: 5 interaction blocks, default widths, 2 x 40 beads
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for k in ("MLCG_USE_TRITON_MESSAGE_PASSING", "MLCG_USE_FUSED_RBF", "MLCG_USE_FUSED_TANH_LINEAR",
          "MLCG_USE_CSR", "MLCG_USE_SRC_CSR_GRAD_X"):
    os.environ[k] = "0"
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(1, "/root/reference/src")

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(4)

spec = importlib.util.spec_from_file_location(
    "fmd_synthetic", os.path.join(ROOT, "flash-molecular-dynamics_b200", "flashmd", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)

import flashmd  # noqa: E402  (the reference)
from flashmd.data import AtomicData  # noqa: E402
from flashmd.models import (CosineCutoff, GaussianBasis, GradientsOut, StandardSchNet, SumOut)  # noqa: E402
from flashmd.neighbor_list.neighbor_list import make_neighbor_list  # noqa: E402
from flashmd.prior import Dihedral, HarmonicAngles, HarmonicBonds, Repulsion  # noqa: E402
from flashmd.prior import (GeneralAngles, GeneralBonds, HarmonicImpropers, Polynomial, QuarticAngles,  # noqa: E402
                           RestrictedQuartic)
from flashmd.prior.harmonic import HarmonicAnglesRaw, ShiftedPeriodicHarmonicImpropers  # noqa: E402
from flashmd.simulation import LangevinSimulation, PTSimulation  # noqa: E402
from flashmd.simulation import NVESimulation, OverdampedSimulation  # noqa: E402

assert flashmd.__file__.startswith("/root/reference"), flashmd.__file__
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def stats_from_tables(system):
    """Reference-format statistics dicts (keys = bead-type tuples that occur in the molecule)."""
    ty = system["atom_types"]
    st = system["stats"]
    bonds, angles, dihedrals, nonbonded = (system[k] for k in ("bonds", "angles", "dihedrals", "nonbonded"))
    sb, sa, sd, sr = {}, {}, {}, {}
    for i, j in bonds.T:
        key = (int(ty[i]), int(ty[j]))
        sb[key] = {"k": float(st["bonds"]["k"][key]), "x_0": float(st["bonds"]["x_0"][key])}
    for i, j, k in angles.T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]))
        sa[key] = {"k": float(st["angles"]["k"][key]), "x_0": float(st["angles"]["x_0"][key])}
    nd = st["dihedrals"]["n_degs"]
    for i, j, k, l in dihedrals.T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]), int(ty[l]))
        c = (int(ty[j]), int(ty[k]))
        sd[key] = {
            "k1s": {f"k1_{n + 1}": float(st["dihedrals"]["k1_central"][(n,) + c]) for n in range(nd)},
            "k2s": {f"k2_{n + 1}": float(st["dihedrals"]["k2_central"][(n,) + c]) for n in range(nd)},
            "v_0": float(st["dihedrals"]["v0_central"][c]),
        }
    for i, j in nonbonded.T:
        key = (int(ty[i]), int(ty[j]))
        sr[key] = {"sigma": float(st["repulsion"]["sigma"][key])}
    return sb, sa, sd, sr, nd


def build_reference(system, hidden, filters, num_rbf, num_blocks, out_widths, seed, bias_scale=0.0):
    rc = system["cutoff"]
    torch.manual_seed(seed)
    schnet = StandardSchNet(GaussianBasis(CosineCutoff(0.0, rc), num_rbf=num_rbf), CosineCutoff(0.0, rc),
                            output_hidden_layer_widths=list(out_widths), hidden_channels=hidden,
                            embedding_size=syn.N_BEAD_TYPES + 1, num_filters=filters,
                            num_interactions=num_blocks)
    if bias_scale:
        with torch.no_grad():
            for n_, p in schnet.named_parameters():
                if n_.endswith("bias"):
                    p.normal_(0.0, bias_scale)
    sb, sa, sd, sr, nd = stats_from_tables(system)
    models = torch.nn.ModuleDict({
        "SchNet": GradientsOut(schnet),
        "bonds": GradientsOut(HarmonicBonds(sb)),
        "angles": GradientsOut(HarmonicAngles(sa)),
        "dihedrals": GradientsOut(Dihedral(sd, n_degs=nd)),
        "repulsion": GradientsOut(Repulsion(sr)),
    })
    model = SumOut(models)
    configs = []
    for b in range(system["pos"].shape[0]):
        nls = {
            "bonds": make_neighbor_list("bonds", 2, torch.from_numpy(system["bonds"])),
            "angles": make_neighbor_list("angles", 3, torch.from_numpy(system["angles"])),
            "dihedrals": make_neighbor_list("dihedrals", 4, torch.from_numpy(system["dihedrals"])),
            "repulsion": make_neighbor_list("repulsion", 2, torch.from_numpy(system["nonbonded"])),
        }
        configs.append(AtomicData.from_points(
            pos=torch.from_numpy(system["pos"][b].copy()), atom_types=torch.from_numpy(system["atom_types"]),
            masses=torch.from_numpy(system["masses"]), neighborlist=nls))
    return model, schnet, configs


def schnet_weights(schnet):
    """Flat, framework-neutral weight arrays (nn.Linear layout [out, in])."""
    w = {"embedding": schnet.embedding_layer.weight}
    for l, blk in enumerate(schnet.interaction_blocks):
        cf = blk.conv
        w[f"b{l}.lin1_w"] = cf.lin1.weight
        w[f"b{l}.f0_w"] = cf.filter_network.layers[0].weight
        w[f"b{l}.f0_b"] = cf.filter_network.layers[0].bias
        w[f"b{l}.f1_w"] = cf.filter_network.layers[2].weight
        w[f"b{l}.lin2_w"] = cf.lin2.weight
        w[f"b{l}.lin2_b"] = cf.lin2.bias
        w[f"b{l}.lin_w"] = blk.lin.weight
        w[f"b{l}.lin_b"] = blk.lin.bias
    lin = [m for m in schnet.output_network.layers if isinstance(m, torch.nn.Linear)]
    for i, m in enumerate(lin):
        w[f"out{i}_w"] = m.weight
        if m.bias is not None:
            w[f"out{i}_b"] = m.bias
    return {"w." + k: v.detach().numpy().copy() for k, v in w.items()}


def eval_reference(model, configs, dtype=torch.float32):
    data = LangevinSimulation.collate(configs)
    data.pos = data.pos.to(dtype)
    model = model.to(dtype)
    data.out = {}
    data = model(data)
    out = {}
    for name in model.models.keys():
        out[f"energy.{name}"] = data.out[name]["energy"].detach().numpy().copy()
        out[f"forces.{name}"] = data.out[name]["forces"].detach().numpy().copy()
    out["energy.total"] = data.out["energy"].detach().numpy().copy()
    out["forces.total"] = data.out["forces"].detach().numpy().copy()
    return out, data


def system_arrays(system):
    return {"sys." + k: v for k, v in system.items() if k not in ("stats",)} | {
        "stats.bonds.k": system["stats"]["bonds"]["k"], "stats.bonds.x_0": system["stats"]["bonds"]["x_0"],
        "stats.angles.k": system["stats"]["angles"]["k"], "stats.angles.x_0": system["stats"]["angles"]["x_0"],
        "stats.dihedrals.k1_central": system["stats"]["dihedrals"]["k1_central"],
        "stats.dihedrals.k2_central": system["stats"]["dihedrals"]["k2_central"],
        "stats.dihedrals.v0_central": system["stats"]["dihedrals"]["v0_central"],
        "stats.repulsion.sigma": system["stats"]["repulsion"]["sigma"],
    }


def golden_static(fname, n_mols, n_beads, seed, hidden, filters, num_rbf, num_blocks, out_widths,
                  target_degree, bias_scale=0.0):
    system = syn.synthetic_system(n_mols, n_beads, seed=seed, target_degree=target_degree)
    model, schnet, configs = build_reference(system, hidden, filters, num_rbf, num_blocks, out_widths, seed,
                                             bias_scale)
    out32, data = eval_reference(model, configs, torch.float32)
    # the SchNet neighbour list the reference built this forward (through the torch_cluster stand-in)
    nl = schnet.neighbor_list(data, system["cutoff"], 1000)["SchNet"]["index_mapping"].numpy()
    out64, _ = eval_reference(model, configs, torch.float64)
    model.to(torch.float32)
    arrs = system_arrays(system) | schnet_weights(schnet)
    arrs |= {"ref32." + k: v for k, v in out32.items()} | {"ref64." + k: v for k, v in out64.items()}
    arrs["ref.schnet_edge_index"] = nl
    arrs["meta.hparams"] = np.array([hidden, filters, num_rbf, num_blocks] + list(out_widths), dtype=np.int64)
    arrs["meta.rbf_centers"] = schnet.rbf_layer.offset.numpy().copy()
    arrs["meta.rbf_gamma"] = np.array(schnet.rbf_layer.coeff.item())
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, {k: v.shape for k, v in out32.items() if k.startswith("energy")}, "E=", nl.shape[1])
    return system, model, schnet, configs


def golden_langevin(system, model, schnet, configs, fname, n_steps=10):
    import tempfile
    tmp = tempfile.mkdtemp()
    seed = 103838
    torch.manual_seed(1234)  # Maxwell-Boltzmann initial velocities use the global RNG (langevin.py:79-99)
    sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=n_steps, save_interval=1,
                             export_interval=n_steps, save_forces=True, save_energies=True,
                             random_seed=seed, device="cpu", dtype="single", filename="g", output_dir=tmp,
                             specialize_priors=True, compile_model=False, gptq=None)
    sim.attach_model_and_configurations(model, configs, beta=1.67)
    v0 = sim.initial_data.velocities.numpy().copy()
    sim.simulate()
    n = system["pos"].shape[0] * system["pos"].shape[1]
    g = torch.Generator(device="cpu").manual_seed(seed)
    noise = np.stack([torch.empty(n, 3).normal_(generator=g).numpy().copy() for _ in range(n_steps)])
    arrs = {
        "v0": v0, "noise": noise,
        "coords": np.load(os.path.join(tmp, "g_coords_0000.npy")),
        "forces": np.load(os.path.join(tmp, "g_forces_0000.npy")),
        "potential": np.load(os.path.join(tmp, "g_potential_0000.npy")),
        "kinetic": np.load(os.path.join(tmp, "g_kineticenergy_0000.npy")),
        "params": np.array([0.004, 1.0, 1.67, seed], dtype=np.float64),
        "files": np.array(sorted(os.listdir(tmp))),
    }
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, arrs["coords"].shape, arrs["potential"].shape, list(arrs["files"]))


def golden_integrators(system, model, configs, fname, n_steps=12):
    """The other integrators of the reference, CPU module path, saved every step:
    NVESimulation (simulation/velocity_verlet.py:12-95; deterministic given the Maxwell-Boltzmann velocities drawn from
    the global RNG) and OverdampedSimulation (simulation/langevin.py:315-420; noise from torch.Generator(seed))."""
    import tempfile
    arrs = {}
    for name, cls, kw in (("nve", NVESimulation, dict(dt=0.001)), ("overdamped", OverdampedSimulation, dict(dt=0.002, friction=2.0)),
                          ("langevin_double", LangevinSimulation, dict(dt=0.004, friction=1.0, dtype="double"))):
        tmp = tempfile.mkdtemp()
        torch.manual_seed(4321)
        kw = dict(dtype="single") | kw
        sim = cls(n_timesteps=n_steps, save_interval=1, export_interval=n_steps, save_forces=True, save_energies=True,
                  random_seed=777, device="cpu", filename="g", output_dir=tmp, specialize_priors=True,
                  compile_model=False, gptq=None, **kw)
        sim.attach_model_and_configurations(model, configs, beta=1.67)
        if name == "nve":
            arrs["nve.v0"] = sim.initial_data.velocities.numpy().copy()
        sim.simulate()
        arrs[f"{name}.coords"] = np.load(os.path.join(tmp, "g_coords_0000.npy"))
        arrs[f"{name}.forces"] = np.load(os.path.join(tmp, "g_forces_0000.npy"))
        arrs[f"{name}.potential"] = np.load(os.path.join(tmp, "g_potential_0000.npy"))
        arrs[f"{name}.files"] = np.array(sorted(os.listdir(tmp)))
    arrs["params"] = np.array([0.001, 0.002, 2.0, 1.67, 777, 4321], dtype=np.float64)   # dt_nve, dt_od, friction_od, beta, seed, global seed
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, {k: v.shape for k, v in arrs.items() if k.endswith("coords")})


def golden_extra_priors(system, fname):
    """Energies (per molecule) and forces of the prior classes the benchmark system does not use, evaluated by the
    UNMODIFIED reference (GradientsOut autograd, fp32 and fp64) on the golden 4 x 54-bead system; statistics from
    oracle/extra_prior_stats.py so that the drop-in test can rebuild the same objects."""
    import extra_prior_stats as X
    ty = system["atom_types"]
    kb, ka, kd = X.type_keys(ty, system["bonds"]), X.type_keys(ty, system["angles"]), X.type_keys(ty, system["dihedrals"])
    priors = {
        "gbonds": (GeneralBonds(X.harmonic_stats(kb, 3.6, 4.0), "gbonds"), system["bonds"], 2),
        "gangles": (GeneralAngles(X.harmonic_stats(ka, -0.6, 0.2), "gangles"), system["angles"], 3),
        # HarmonicAnglesRaw cannot be constructed in the reference (its __init__ omits Harmonic's `order`, harmonic.py:287)
        # HarmonicImpropers.forward raises in the reference (data2features is a @staticmethod taking self, harmonic.py:313)
        "poly_bonds": (Polynomial(X.polynomial_stats(kb), "poly_bonds", order=2, n_degs=4), system["bonds"], 2),
        "quartic_angles": (QuarticAngles(X.polynomial_stats(ka), name="quartic_angles"), system["angles"], 3),
        "restricted": (RestrictedQuartic(X.restricted_quartic_stats(ka), name="restricted"), system["angles"], 3),
        "shifted_impropers": (ShiftedPeriodicHarmonicImpropers(X.harmonic_stats(kd, -0.5, 0.5)), system["dihedrals"], 4),
    }
    # Polynomial itself has no feature function (the reference's users subclass it): distances for the bond set
    from flashmd.geometry import compute_distances
    priors["poly_bonds"][0].data2features = lambda data, p=priors["poly_bonds"][0]: compute_distances(
        data.pos, data.neighbor_list[p.name]["index_mapping"])
    arrs = {}
    for name, (prior, mapping, order) in priors.items():
        nl_name = prior.name
        configs = [AtomicData.from_points(pos=torch.from_numpy(system["pos"][b].copy()), atom_types=torch.from_numpy(ty),
                                          masses=torch.from_numpy(system["masses"]),
                                          neighborlist={nl_name: make_neighbor_list(nl_name, order, torch.from_numpy(mapping))})
                   for b in range(system["pos"].shape[0])]
        model = SumOut(torch.nn.ModuleDict({nl_name: GradientsOut(prior)}))
        for dt, tag in ((torch.float32, "ref32"), (torch.float64, "ref64")):
            out, _ = eval_reference(model, configs, dt)
            arrs[f"{tag}.{name}.energy"] = out["energy.total"]
            arrs[f"{tag}.{name}.forces"] = out["forces.total"]
        model.to(torch.float32)
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, sorted(k for k in arrs if k.startswith("ref64") and k.endswith("energy")))


def golden_pt(fname):
    """PTSimulation bookkeeping: pair sets, one Metropolis decision + swap with recorded uniforms
    (parallel_tempering.py:256-284, 368-481), and a short full run for the file set."""
    import tempfile
    system = syn.synthetic_system(2, 24, seed=5, target_degree=12)
    model, schnet, configs = build_reference(system, 32, 32, 16, 1, (16,), 5)
    tmp = tempfile.mkdtemp()
    torch.manual_seed(99)
    betas = [1.67, 1.42, 1.16]
    sim = PTSimulation(friction=1.0, dt=0.004, n_timesteps=40, save_interval=10, export_interval=40,
                       exchange_interval=10, save_energies=True, random_seed=7, device="cpu",
                       dtype="single", filename="pt", output_dir=tmp, specialize_priors=True,
                       compile_model=False, gptq=None)
    sim.attach_model_and_configurations(model, configs, betas=betas)
    arrs = {
        "betas": np.array(betas), "n_indep": np.array(sim.n_indep_sims),
        "beta_per_sim": sim.beta.numpy().copy(),
        "even_a": sim._even_pairs[0].numpy().copy(), "even_b": sim._even_pairs[1].numpy().copy(),
        "odd_a": sim._odd_pairs[0].numpy().copy(), "odd_b": sim._odd_pairs[1].numpy().copy(),
        "pair_to_beta_idx": sim.pair_to_beta_idx.numpy().copy(),
    }
    # one decision + swap on hand-made energies (two rounds: even pairs then odd pairs)
    from copy import deepcopy
    data = deepcopy(sim.initial_data)
    sim._set_up_simulation(overwrite=True)
    g = torch.Generator().manual_seed(5)
    for rnd in range(2):
        energies = torch.randn(sim.n_sims, generator=g) * 2.0
        data.out = {"energy": energies}
        x0 = data.pos.clone(); v0 = data.velocities.clone()
        torch.manual_seed(1000 + rnd)
        n_pairs = len(sim._even_pairs[0] if rnd == 0 else sim._odd_pairs[0])
        uniforms = torch.rand(n_pairs)
        torch.manual_seed(1000 + rnd)
        data = sim.detect_and_exchange_replicas(data)
        arrs[f"round{rnd}.energies"] = energies.numpy().copy()
        arrs[f"round{rnd}.uniforms"] = uniforms.numpy().copy()
        arrs[f"round{rnd}.x_before"] = x0.numpy().copy(); arrs[f"round{rnd}.v_before"] = v0.numpy().copy()
        arrs[f"round{rnd}.x_after"] = data.pos.numpy().copy(); arrs[f"round{rnd}.v_after"] = data.velocities.numpy().copy()
        arrs[f"round{rnd}.acceptance_matrix"] = sim.acceptance_matrix.numpy().copy()
    sim._propose_even_pairs = True
    sim.simulate(overwrite=True)
    arrs["files"] = np.array(sorted(os.listdir(tmp)))
    arrs["coords_shape"] = np.array(np.load(os.path.join(tmp, "pt_coords_0000.npy")).shape)
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, {k: v.shape for k, v in arrs.items()}, list(arrs["files"]))


def golden_known_answers(fname):
    """Known-answer values of the reference's own unit tests (tests/models/test_cutoff.py:27-57,
    tests/models/radial_basis/test_radial_basis.py:13-18) evaluated with the reference classes."""
    from flashmd.models.cutoff import ShiftedCosineCutoff
    d = torch.tensor([0.0, 1.0, 2.5, 4.999, 5.0, 7.5, 10.0])
    arrs = {
        "d": d.numpy(),
        "cos_0_5": CosineCutoff(0, 5)(d).numpy(),
        "cos_0_10": CosineCutoff(0, 10)(d).numpy(),
        "cos_5_10": CosineCutoff(5, 10)(d).numpy(),
        "shifted_5_05": ShiftedCosineCutoff(5, 0.5)(d).numpy(),
    }
    gb = GaussianBasis(CosineCutoff(0.0, 10.0), num_rbf=50)
    arrs["gauss_centers"] = gb.offset.numpy().copy()
    arrs["gauss_gamma"] = np.array(gb.coeff.item())
    arrs["gauss_val"] = gb(d).numpy()
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, "ok")


if __name__ == "__main__" and "--integrators" in sys.argv:
    # only the integrator vectors (the static vectors are regenerated in memory, nothing else is rewritten)
    _system = syn.synthetic_system(4, 54, seed=0, target_degree=30.0)
    _model, _schnet, _configs = build_reference(_system, 128, 128, 50, 3, (128, 64), 0, 0.0)
    golden_integrators(_system, _model, _configs, "integrators_n54_b4.npz", n_steps=12)
    sys.exit(0)

if __name__ == "__main__" and "--l5" in sys.argv:
    # cfg5's depth (5 interaction blocks, default widths) on a small system: 2 x 40 beads
    golden_static("schnet_n40_b2_l5.npz", 2, 40, 7, 128, 128, 50, 5, (128, 64), 20.0)
    sys.exit(0)

if __name__ == "__main__" and "--extra-priors" in sys.argv:
    golden_extra_priors(syn.synthetic_system(4, 54, seed=0, target_degree=30.0), "extra_priors_n54_b4.npz")
    sys.exit(0)

if __name__ == "__main__" and "--stats" not in sys.argv:
    golden_known_answers("known_answers.npz")
    system, model, schnet, configs = golden_static("schnet_n54_b4.npz", 4, 54, 0, 128, 128, 50, 3, (128, 64), 30.0)
    golden_langevin(system, model, schnet, configs, "langevin_n54_b4.npz", n_steps=10)
    golden_integrators(system, model, configs, "integrators_n54_b4.npz", n_steps=12)
    golden_static("schnet_n24_b3_l2.npz", 3, 24, 3, 64, 64, 20, 2, (32,), 10.0, bias_scale=0.2)
    golden_pt("pt_n24.npz")


def golden_stats(fname, n_steps=10000, n_mols=8, save_interval=10):
    """Trajectory statistics of the UNMODIFIED reference (CPU path) over 10^4 Langevin steps: kinetic
    energy (temperature), RMSD from the start structure, radius of gyration, potential energy.  The GPU tests
    compare the fused engine's own 10^4-step trajectories with these distributions (north_star)."""
    import tempfile
    system = syn.synthetic_system(n_mols, 54, seed=0, target_degree=30.0)
    model, schnet, configs = build_reference(system, 128, 128, 50, 3, (128, 64), 0)
    tmp = tempfile.mkdtemp()
    torch.manual_seed(4321)
    sim = LangevinSimulation(friction=1.0, dt=0.004, n_timesteps=n_steps, save_interval=save_interval,
                             export_interval=n_steps, save_energies=True, random_seed=2024, device="cpu",
                             dtype="single", filename="s", output_dir=tmp, specialize_priors=True,
                             compile_model=False, gptq=None)
    sim.attach_model_and_configurations(model, configs, beta=1.67)
    sim.simulate()
    x = np.load(os.path.join(tmp, "s_coords_0000.npy"))            # [B, frames, n, 3]
    ke = np.load(os.path.join(tmp, "s_kineticenergy_0000.npy"))    # [B, frames]
    pe = np.load(os.path.join(tmp, "s_potential_0000.npy"))
    x0 = system["pos"][:n_mols]
    from tests_helpers_kabsch import kabsch_rmsd, radius_of_gyration
    rmsd = np.stack([[kabsch_rmsd(x[b, f], x0[b]) for f in range(x.shape[1])] for b in range(n_mols)])
    rg = radius_of_gyration(x)
    arrs = {"ke": ke, "pe": pe, "rmsd": rmsd, "rg": rg, "params": np.array([0.004, 1.0, 1.67, n_steps, save_interval]),
            "n_mols": np.array(n_mols)}
    np.savez_compressed(os.path.join(OUT, fname), **arrs)
    print(fname, "KE mean", ke[:, 100:].mean(), "expected", 1.5 * 54 / 1.67, "rmsd end", rmsd[:, -1].mean())


if __name__ == "__main__" and "--stats" in sys.argv:
    golden_stats("stats_n54_10k.npz")
