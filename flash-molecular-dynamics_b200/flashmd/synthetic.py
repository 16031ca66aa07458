"""Synthetic 1ENH-shaped coarse-grained systems (no real structures/weights are available offline).

Stand-alone on purpose (numpy only, no package-relative imports): the golden-vector generator
loads this file by path next to the *reference* `flashmd` package, and the benchmark / tests use
it through `flashmd.synthetic`.

What it produces mirrors what the reference's `--structure_file` / `--model_file` carry
(/root/reference/src/flashmd/simulation/cli.py:120-128): per-molecule positions, bead types,
masses, and the bonded / non-bonded index mappings consumed by the prior terms
(prior/harmonic.py:79-123, prior/fourier_series.py:97-192, prior/repulsion.py:80-122).
"""
from __future__ import annotations

import numpy as np

BOND_LENGTH = 3.8          # Angstrom, CA-CA virtual bond
N_BEAD_TYPES = 24          # bead types 1..24 (embedding_size 25)
MASS = 12.0 / 418.4        # 12 amu in the reference's kcal/mol-Angstrom-ps unit system (simulation/base.py:58-59)


def make_chains(n_mols: int, n_beads: int, seed: int = 0, bond: float = BOND_LENGTH,
                min_dist: float = 3.4, compactness: float = 1.05) -> np.ndarray:
    """Batched self-avoiding random walks confined to a sphere -> float32 [n_mols, n_beads, 3].

    The confining radius is chosen so that bead density is protein-like; all molecules are
    generated simultaneously (vectorised over `n_mols`).
    """
    rng = np.random.default_rng(seed)
    radius = compactness * bond * (n_beads ** (1.0 / 3.0))
    pos = np.zeros((n_mols, n_beads, 3), dtype=np.float64)
    pos[:, 0] = rng.normal(scale=0.3 * radius, size=(n_mols, 3))
    for i in range(1, n_beads):
        todo = np.ones(n_mols, dtype=bool)
        best = np.zeros((n_mols, 3))
        best_score = np.full(n_mols, -np.inf)
        for _ in range(40):
            idx = np.nonzero(todo)[0]
            if idx.size == 0:
                break
            d = rng.normal(size=(idx.size, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            cand = pos[idx, i - 1] + bond * d
            dist = np.linalg.norm(pos[idx, :i] - cand[:, None, :], axis=2)
            # the bonded neighbour is at exactly `bond`; only earlier beads constrain
            dmin = dist[:, : i - 1].min(axis=1) if i > 1 else np.full(idx.size, np.inf)
            inside = np.linalg.norm(cand, axis=1) <= radius
            score = np.minimum(dmin, min_dist) - (~inside) * 1.0
            better = score > best_score[idx]
            best[idx[better]] = cand[better]
            best_score[idx[better]] = score[better]
            ok = (dmin >= min_dist) & inside
            todo[idx[ok]] = False
        pos[:, i] = best
    pos -= pos.mean(axis=1, keepdims=True)
    return pos.astype(np.float32)


def bead_types(n_beads: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed + 7919)
    return rng.integers(1, N_BEAD_TYPES + 1, size=n_beads).astype(np.int64)


def chain_topology(n_beads: int, nonbonded_min_sep: int = 3):
    """Index mappings of a linear chain: bonds [2,nb], angles [3,na], dihedrals [4,nd],
    non-bonded pairs [2,np] with |i-j| >= nonbonded_min_sep, each unordered pair once."""
    i = np.arange(n_beads - 1)
    bonds = np.stack([i, i + 1])
    i = np.arange(n_beads - 2)
    angles = np.stack([i, i + 1, i + 2])
    i = np.arange(n_beads - 3)
    dihedrals = np.stack([i, i + 1, i + 2, i + 3])
    a, b = np.triu_indices(n_beads, k=nonbonded_min_sep)
    nonbonded = np.stack([a, b])
    return (bonds.astype(np.int64), angles.astype(np.int64), dihedrals.astype(np.int64),
            nonbonded.astype(np.int64))


def prior_statistics(seed: int = 0, n_degs: int = 3, bond: float = BOND_LENGTH):
    """Per-bead-type-tuple parameter tables with mild constants (10^4-step trajectories stay bounded).

    Returns dense numpy tables indexed by bead type (size N_BEAD_TYPES+1 per axis):
      bonds:  k[t,t], x0[t,t]          E = k (d - x0)^2
      angles: k[t,t,t], x0[t,t,t]      E = k (cos(theta) - x0)^2
      dihedrals: k1[n,t,t,t,t], k2[...], v0[t,t,t,t]
      repulsion: sigma[t,t]            E = (sigma / d)^6
    """
    rng = np.random.default_rng(seed + 104729)
    T = N_BEAD_TYPES + 1
    out = {}
    kb = 20.0 + 5.0 * rng.random((T, T))
    kb = 0.5 * (kb + kb.T)
    x0b = bond + 0.05 * (rng.random((T, T)) - 0.5)
    x0b = 0.5 * (x0b + x0b.T)
    out["bonds"] = {"k": kb.astype(np.float32), "x_0": x0b.astype(np.float32)}
    ka = 3.0 + 2.0 * rng.random((T, T, T))
    x0a = -0.3 + 0.4 * (rng.random((T, T, T)) - 0.5)
    out["angles"] = {"k": ka.astype(np.float32), "x_0": x0a.astype(np.float32)}
    # dihedral tables are keyed on the two central bead types only (broadcast over the outer ones)
    k1c = 0.3 * (rng.random((n_degs, T, T)) - 0.5)
    k2c = 0.3 * (rng.random((n_degs, T, T)) - 0.5)
    v0c = 0.1 * rng.random((T, T))
    out["dihedrals"] = {"k1_central": k1c.astype(np.float32), "k2_central": k2c.astype(np.float32),
                        "v0_central": v0c.astype(np.float32), "n_degs": n_degs}
    sg = 3.0 + 0.4 * rng.random((T, T))
    sg = 0.5 * (sg + sg.T)
    out["repulsion"] = {"sigma": sg.astype(np.float32)}
    return out


def mean_degree(pos: np.ndarray, rc: float) -> float:
    """Mean number of neighbours within rc (strict <), averaged over all beads of all molecules."""
    tot = 0
    for p in pos:
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        tot += (d2 < rc * rc).sum() - p.shape[0]
    return tot / (pos.shape[0] * pos.shape[1])


def tune_cutoff(pos: np.ndarray, target_degree: float, lo: float = 4.0, hi: float = 60.0) -> float:
    """Bisection on the cutoff radius so the initial mean degree is ~target (rounded to 0.05 A)."""
    n = pos.shape[1]
    target = min(target_degree, n - 1.0)
    sample = pos[: min(8, pos.shape[0])]
    if target >= n - 1.0:
        return float(np.ceil(max(np.linalg.norm(p[:, None] - p[None], axis=2).max() for p in sample) * 1.25))
    for _ in range(30):
        mid = 0.5 * (lo + hi)
        if mean_degree(sample, mid) < target:
            lo = mid
        else:
            hi = mid
    return float(np.round(0.5 * (lo + hi) / 0.05) * 0.05)


def synthetic_system(n_mols: int, n_beads: int, seed: int = 0, target_degree: float = 55.0):
    """Everything needed to build configs + model for one benchmark/test system."""
    pos = make_chains(n_mols, n_beads, seed=seed)
    types = bead_types(n_beads, seed=seed)
    bonds, angles, dihedrals, nonbonded = chain_topology(n_beads)
    rc = tune_cutoff(pos, target_degree)
    return {
        "pos": pos,                       # [B, n, 3] float32
        "atom_types": types,              # [n] int64
        "masses": np.full(n_beads, MASS, dtype=np.float32),
        "bonds": bonds, "angles": angles, "dihedrals": dihedrals, "nonbonded": nonbonded,
        "cutoff": rc,
        "stats": prior_statistics(seed=seed),
    }
