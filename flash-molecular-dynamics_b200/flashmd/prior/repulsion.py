"""(sigma / r)^6 excluded-volume prior on a given pair list (reference prior/repulsion.py:14-122)."""
from typing import Dict

import torch

from ..geometry import compute_distances
from .base import _Prior, type_table


class Repulsion(_Prior):
    name = "repulsion"
    order = 2
    kernel_kind = 3

    def __init__(self, statistics: Dict) -> None:
        super().__init__()
        self.allowed_interaction_keys = list(statistics.keys())
        self.name, self.order = Repulsion.name, 2
        self.register_buffer("sigma", type_table(statistics, 2, "sigma"))

    def data2parameters(self, data) -> Dict:
        return {"sigma": self.sigma[self.types_of_terms(data)]}

    @staticmethod
    def compute_features(pos, mapping):
        return compute_distances(pos, mapping)

    @staticmethod
    def compute(x, sigma):
        rr = (sigma / x) * (sigma / x)
        return rr * rr * rr

    def term_energies(self, data):
        return Repulsion.compute(self.data2features(data), self.data2parameters(data)["sigma"])

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(Repulsion.name, 2, topology)
