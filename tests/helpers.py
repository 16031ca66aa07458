"""Shared test helpers: load golden fixtures into oracle-side structures."""
import os

import numpy as np
import torch

from oracle import fmd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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
