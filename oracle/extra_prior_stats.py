"""TEST INFRASTRUCTURE.  Deterministic synthetic statistics for the prior classes that the synthetic benchmark system
does not use (GeneralBonds / GeneralAngles / HarmonicAnglesRaw / Polynomial / QuarticAngles / RestrictedQuartic /
HarmonicImpropers): both oracle/make_golden.py (UNMODIFIED reference) and tests/test_dropin_api.py (drop-in package) build
their prior objects from these tables, so only the reference's OUTPUTS need to be stored in the golden file."""
import numpy as np


def _u(key, salt):
    """pseudo-random but reproducible number in [0, 1) from a type tuple"""
    h = 1469598103934665603
    for v in tuple(key) + (salt,):
        h = ((h ^ (int(v) + 1)) * 1099511628211) % (1 << 64)
    return (h % 100003) / 100003.0


def type_keys(atom_types, mapping):
    ty = np.asarray(atom_types)
    return sorted({tuple(int(ty[i]) for i in col) for col in np.asarray(mapping).T})


def harmonic_stats(keys, x0_lo, x0_hi, k_lo=1.0, k_hi=5.0):
    return {k: {"k": k_lo + (k_hi - k_lo) * _u(k, 1), "x_0": x0_lo + (x0_hi - x0_lo) * _u(k, 2)} for k in keys}


def polynomial_stats(keys, n_degs=4):
    return {k: {"ks": {f"k_{n}": (-1.0) ** n * (0.2 + _u(k, 10 + n)) / n for n in range(1, n_degs + 1)}, "v_0": _u(k, 3) - 0.5}
            for k in keys}


def restricted_quartic_stats(keys):
    return {k: {"a": 0.5 + _u(k, 20), "b": _u(k, 21) - 0.5, "c": 0.3 + _u(k, 22), "d": _u(k, 23) - 0.5,
                "k": 0.05 + 0.1 * _u(k, 24), "v_0": _u(k, 25)} for k in keys}
