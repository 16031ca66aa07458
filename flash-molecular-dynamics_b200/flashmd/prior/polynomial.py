"""Polynomial priors V = V0 + sum_{n=1}^{n_degs} k_n x^n on an internal coordinate (reference prior/polynomial.py:13-186);
QuarticAngles = degree-4 polynomial in cos(theta)."""
from typing import Dict, Optional

import torch

from ..geometry import compute_angles_cos
from .base import _Prior, type_table


class Polynomial(_Prior):
    @property
    def kernel_kind(self):
        """Fused-step kind (include/fmd_b200.h): polynomial in the bond length (order 2, degree <= 4; the lowering checks
        that the attached feature function measures distances) or in cos(theta) (QuarticAngles, degree <= 6)."""
        if isinstance(self, QuarticAngles):
            return 5 if self.n_degs <= 6 else None
        if self.order == 2 and self.n_degs <= 4 and hasattr(self, "compute_features"):
            return 4
        return None

    def __init__(self, statistics: Dict, name: str, order: Optional[int] = None, n_degs: int = 4) -> None:
        super().__init__()
        self.allowed_interaction_keys = list(statistics.keys())
        self.name = name
        self.order = order
        lens = {len(st["ks"]) for st in statistics.values()}
        assert len(lens) == 1, "ks in the statistics dictionary must be of the same size for all the keys"
        assert lens == {n_degs}, f"length of parameters {lens} doesn't match degrees {n_degs}"
        self.n_degs = n_degs
        self.k_names = [f"k_{i}" for i in range(1, n_degs + 1)]
        self.register_buffer("ks", torch.stack([type_table(statistics, order, lambda st, n=n: st["ks"][n])
                                                for n in self.k_names]))
        self.register_buffer("v_0", type_table(statistics, order, "v_0"))

    def data2parameters(self, data) -> Dict:
        tt = self.types_of_terms(data)
        return {"ks": torch.stack([self.ks[i][tt] for i in range(self.n_degs)], 1), "v_0s": self.v_0[tt]}

    def term_energies(self, data):
        p = self.data2parameters(data)
        return Polynomial.compute(self.data2features(data).flatten(), p["ks"].t(), p["v_0s"])

    @staticmethod
    def compute(x: torch.Tensor, ks, V0) -> torch.Tensor:
        """ks[n-1] multiplies x^n; the powers are built by repeated multiplication (the reference's order of operations)."""
        v = ks[0] * x
        xp = x
        for k in ks[1:]:
            xp = xp * x
            v = v + k * xp
        return v + V0


class QuarticAngles(Polynomial):
    def __init__(self, statistics, name="angles", n_degs: int = 4) -> None:
        super().__init__(statistics, name, order=3, n_degs=n_degs)

    @staticmethod
    def compute_features(pos, mapping):
        return compute_angles_cos(pos, mapping)
