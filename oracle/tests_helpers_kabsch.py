"""Small structure-analysis helpers shared by the golden generator and the tests (numpy)."""
import numpy as np


def kabsch_rmsd(a: np.ndarray, b: np.ndarray) -> float:
    """Minimum RMSD between two [n,3] structures after optimal rigid superposition."""
    a = a - a.mean(0)
    b = b - b.mean(0)
    h = a.T @ b
    u, s, vt = np.linalg.svd(h)
    d = np.sign(np.linalg.det(vt.T @ u.T))
    s[-1] *= d
    e0 = (a * a).sum() + (b * b).sum()
    return float(np.sqrt(max(e0 - 2.0 * s.sum(), 0.0) / a.shape[0]))


def radius_of_gyration(x: np.ndarray) -> np.ndarray:
    """[..., n, 3] -> [...]"""
    c = x - x.mean(axis=-2, keepdims=True)
    return np.sqrt((c * c).sum(-1).mean(-1))
