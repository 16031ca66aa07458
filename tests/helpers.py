"""Shared test helpers: load golden fixtures into oracle-side structures."""
import os

import numpy as np
import torch

from oracle import fmd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_params(g, dtype=torch.float32) -> O.SchNetParams:
    hp = g["meta.hparams"]
    num_blocks = int(hp[3])
    t = {k[2:]: torch.from_numpy(v).to(dtype) for k, v in g.items() if k.startswith("w.")}
    n_out = len([k for k in t if k.startswith("out") and k.endswith("_w")])
    for i in range(n_out):
        t.setdefault(f"out{i}_b", None)
    return O.SchNetParams(t, num_blocks, n_out, float(g["sys.cutoff"]), int(hp[2]))


def golden_system(g, dtype=torch.float32):
    pos = torch.from_numpy(g["sys.pos"]).to(dtype)
    B, n, _ = pos.shape
    types = torch.from_numpy(g["sys.atom_types"]).repeat(B)
    batch = torch.arange(B).repeat_interleave(n)
    ptr = np.arange(B + 1) * n
    return pos.reshape(B * n, 3), types, batch, ptr, B, n


def prior_tables(g, kind, B, n, dtype=torch.float32):
    """Flat per-term parameter vectors + collated mapping (what condense_* produces,
    simulation/specialize_prior.py:112-207)."""
    ty = g["sys.atom_types"]
    key = {"bonds": "sys.bonds", "angles": "sys.angles", "dihedrals": "sys.dihedrals", "repulsion": "sys.nonbonded"}[kind]
    m = g[key]
    tt = tuple(ty[m[i]] for i in range(m.shape[0]))
    if kind in ("bonds", "angles"):
        p = {"k": g[f"stats.{kind}.k"][tt], "x0": g[f"stats.{kind}.x_0"][tt]}
    elif kind == "dihedrals":
        c = (tt[1], tt[2])
        k1 = g["stats.dihedrals.k1_central"]
        p = {"k1s": np.stack([k1[d][c] for d in range(k1.shape[0])], 1),
             "k2s": np.stack([g["stats.dihedrals.k2_central"][d][c] for d in range(k1.shape[0])], 1),
             "v_0": g["stats.dihedrals.v0_central"][c]}
    else:
        p = {"sigma": g["stats.repulsion.sigma"][tt]}
    nt = m.shape[1]
    mapping = np.concatenate([m + b * n for b in range(B)], axis=1)
    mbatch = np.repeat(np.arange(B), nt)
    params = {k: torch.from_numpy(np.concatenate([v] * B, 0)).to(dtype) for k, v in p.items()}
    return torch.from_numpy(mapping), torch.from_numpy(mbatch), params


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ------------------------------------------------------------------------------------------------
# building reference-style model objects with the drop-in `flashmd` package from a golden file
# ------------------------------------------------------------------------------------------------
def golden_statistics(g):
    """Reference-format statistics dicts (keys = bead-type tuples) from the dense tables of a golden file."""
    ty = g["sys.atom_types"]
    sb, sa, sd, sr = {}, {}, {}, {}
    for i, j in g["sys.bonds"].T:
        key = (int(ty[i]), int(ty[j]))
        sb[key] = {"k": float(g["stats.bonds.k"][key]), "x_0": float(g["stats.bonds.x_0"][key])}
    for i, j, k in g["sys.angles"].T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]))
        sa[key] = {"k": float(g["stats.angles.k"][key]), "x_0": float(g["stats.angles.x_0"][key])}
    k1, k2, v0 = g["stats.dihedrals.k1_central"], g["stats.dihedrals.k2_central"], g["stats.dihedrals.v0_central"]
    nd = k1.shape[0]
    for i, j, k, l in g["sys.dihedrals"].T:
        key = (int(ty[i]), int(ty[j]), int(ty[k]), int(ty[l]))
        c = (int(ty[j]), int(ty[k]))
        sd[key] = {"k1s": {f"k1_{n + 1}": float(k1[(n,) + c]) for n in range(nd)},
                   "k2s": {f"k2_{n + 1}": float(k2[(n,) + c]) for n in range(nd)}, "v_0": float(v0[c])}
    for i, j in g["sys.nonbonded"].T:
        key = (int(ty[i]), int(ty[j]))
        sr[key] = {"sigma": float(g["stats.repulsion.sigma"][key])}
    return sb, sa, sd, sr, nd


def dropin_model_from_golden(g, embedding_size=None):
    """(model, schnet, configs) built exactly like oracle/make_golden.py builds the reference objects, but with
    the drop-in package; weights copied from the golden file."""
    from flashmd.data import AtomicData
    from flashmd.models import CosineCutoff, GaussianBasis, GradientsOut, StandardSchNet, SumOut
    from flashmd.neighbor_list import make_neighbor_list
    from flashmd.prior import Dihedral, HarmonicAngles, HarmonicBonds, Repulsion
    hp = [int(v) for v in g["meta.hparams"]]
    hidden, filters, num_rbf, nblocks, widths = hp[0], hp[1], hp[2], hp[3], hp[4:]
    rc = float(g["sys.cutoff"])
    schnet = StandardSchNet(GaussianBasis(CosineCutoff(0.0, rc), num_rbf=num_rbf), CosineCutoff(0.0, rc),
                            output_hidden_layer_widths=list(widths), hidden_channels=hidden,
                            embedding_size=g["w.embedding"].shape[0], num_filters=filters, num_interactions=nblocks)
    W = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w.")}
    with torch.no_grad():
        schnet.embedding_layer.weight.copy_(W["embedding"])
        for l, blk in enumerate(schnet.interaction_blocks):
            cf = blk.conv
            cf.lin1.weight.copy_(W[f"b{l}.lin1_w"])
            cf.filter_network.layers[0].weight.copy_(W[f"b{l}.f0_w"])
            cf.filter_network.layers[0].bias.copy_(W[f"b{l}.f0_b"])
            cf.filter_network.layers[2].weight.copy_(W[f"b{l}.f1_w"])
            cf.lin2.weight.copy_(W[f"b{l}.lin2_w"]); cf.lin2.bias.copy_(W[f"b{l}.lin2_b"])
            blk.lin.weight.copy_(W[f"b{l}.lin_w"]); blk.lin.bias.copy_(W[f"b{l}.lin_b"])
        lin = [m for m in schnet.output_network.layers if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lin):
            m.weight.copy_(W[f"out{i}_w"])
            if m.bias is not None:
                m.bias.copy_(W[f"out{i}_b"])
    sb, sa, sd, sr, nd = golden_statistics(g)
    model = SumOut(torch.nn.ModuleDict({
        "SchNet": GradientsOut(schnet), "bonds": GradientsOut(HarmonicBonds(sb)),
        "angles": GradientsOut(HarmonicAngles(sa)), "dihedrals": GradientsOut(Dihedral(sd, n_degs=nd)),
        "repulsion": GradientsOut(Repulsion(sr))}))
    configs = []
    for b in range(g["sys.pos"].shape[0]):
        nls = {"bonds": make_neighbor_list("bonds", 2, torch.from_numpy(g["sys.bonds"])),
               "angles": make_neighbor_list("angles", 3, torch.from_numpy(g["sys.angles"])),
               "dihedrals": make_neighbor_list("dihedrals", 4, torch.from_numpy(g["sys.dihedrals"])),
               "repulsion": make_neighbor_list("repulsion", 2, torch.from_numpy(g["sys.nonbonded"]))}
        configs.append(AtomicData.from_points(pos=torch.from_numpy(g["sys.pos"][b].copy()),
                                              atom_types=torch.from_numpy(g["sys.atom_types"]),
                                              masses=torch.from_numpy(g["sys.masses"]), neighborlist=nls))
    return model, schnet, configs


def extra_prior_objects():
    """The prior classes beyond the benchmark's four, built on the golden 4 x 54-bead system exactly like
    oracle/make_golden.py --extra-priors builds the reference objects: {name: (prior, mapping, order)} and the system."""
    import sys
    sys.path.insert(0, os.path.join(ROOT_DIR, "oracle"))
    import extra_prior_stats as X
    from flashmd import synthetic
    from flashmd.geometry import compute_distances
    from flashmd.prior import (GeneralAngles, GeneralBonds, Polynomial, QuarticAngles, RestrictedQuartic,
                               ShiftedPeriodicHarmonicImpropers)
    system = synthetic.synthetic_system(4, 54, seed=0, target_degree=30.0)
    ty = system["atom_types"]
    kb, ka, kd = X.type_keys(ty, system["bonds"]), X.type_keys(ty, system["angles"]), X.type_keys(ty, system["dihedrals"])
    poly = Polynomial(X.polynomial_stats(kb), "poly_bonds", order=2, n_degs=4)
    poly.compute_features = staticmethod(compute_distances)
    priors = {
        "gbonds": (GeneralBonds(X.harmonic_stats(kb, 3.6, 4.0), "gbonds"), system["bonds"], 2),
        "gangles": (GeneralAngles(X.harmonic_stats(ka, -0.6, 0.2), "gangles"), system["angles"], 3),
        "poly_bonds": (poly, system["bonds"], 2),
        "quartic_angles": (QuarticAngles(X.polynomial_stats(ka), name="quartic_angles"), system["angles"], 3),
        "restricted": (RestrictedQuartic(X.restricted_quartic_stats(ka), name="restricted"), system["angles"], 3),
        "shifted_impropers": (ShiftedPeriodicHarmonicImpropers(X.harmonic_stats(kd, -0.5, 0.5)), system["dihedrals"], 4),
    }
    return priors, system


def extra_prior_configs(prior, mapping, order, system):
    from flashmd.data import AtomicData
    from flashmd.neighbor_list import make_neighbor_list
    ty = system["atom_types"]
    return [AtomicData.from_points(pos=torch.from_numpy(system["pos"][b].copy()), atom_types=torch.from_numpy(ty),
                                   masses=torch.from_numpy(system["masses"]),
                                   neighborlist={prior.name: make_neighbor_list(prior.name, order, torch.from_numpy(mapping))})
            for b in range(4)]
