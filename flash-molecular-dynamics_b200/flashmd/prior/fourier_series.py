"""v0 + sum_n k1_n sin(n phi) + k2_n cos(n phi) on torsions (reference prior/fourier_series.py:16-260)."""
from typing import Dict

import torch

from ..geometry import compute_torsions
from .base import _Prior, type_table


class FourierSeries(_Prior):
    def __init__(self, statistics: Dict, name: str = "", n_degs: int = 6, order: int = 4) -> None:
        super().__init__()
        self.allowed_interaction_keys = list(statistics.keys())
        self.name, self.order, self.n_degs = name, order, n_degs
        self.k1_names = [f"k1_{i}" for i in range(1, n_degs + 1)]
        self.k2_names = [f"k2_{i}" for i in range(1, n_degs + 1)]
        self.register_buffer("k1s", torch.stack([type_table(statistics, order, lambda s, n=n: s["k1s"][n])
                                                 for n in self.k1_names]))
        self.register_buffer("k2s", torch.stack([type_table(statistics, order, lambda s, n=n: s["k2s"][n])
                                                 for n in self.k2_names]))
        self.register_buffer("v_0", type_table(statistics, order, "v_0"))

    def data2parameters(self, data) -> Dict:
        tt = self.types_of_terms(data)
        k1 = torch.stack([self.k1s[i][tt] for i in range(self.n_degs)], dim=1)   # [n_terms, n_degs]
        k2 = torch.stack([self.k2s[i][tt] for i in range(self.n_degs)], dim=1)
        return {"k1s": k1, "k2s": k2, "v_0": self.v_0[tt].view(-1, 1)}

    @staticmethod
    def compute(theta, v_0, k1s, k2s):
        n = torch.arange(1, k1s.shape[1] + 1, dtype=theta.dtype, device=theta.device)
        ang = theta.view(-1, 1) * n.view(1, -1)
        return (k1s * torch.sin(ang) + k2s * torch.cos(ang)).sum(dim=1) + v_0.flatten()

    def term_energies(self, data):
        p = self.data2parameters(data)
        return FourierSeries.compute(self.data2features(data), p["v_0"], p["k1s"], p["k2s"])


class Dihedral(FourierSeries):
    name = "dihedrals"
    kernel_kind = 2

    _order = 4

    def __init__(self, statistics, n_degs: int = 3, name: str = "dihedrals") -> None:    # reference defaults (:448-456)
        super().__init__(statistics, name=name, n_degs=n_degs, order=self._order)

    @staticmethod
    def compute_features(pos, mapping):
        return compute_torsions(pos, mapping)

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(Dihedral.name, 4, topology)
