"""Restricted-bending angle prior (reference prior/restricted_bending.py:13-238; J. Chem. Theory Comput. 2013, 9, 3282):
    V(theta) = a cos^4 + b cos^3 + c cos^2 + d cos + k / sin^2(theta) + V0
a quartic in cos(theta) plus a term that keeps the angle away from 0 and pi."""
from typing import Dict

import torch

from ..geometry import compute_angles_raw
from .base import _Prior, type_table


class RestrictedQuartic(_Prior):
    _fields = ("a", "b", "c", "d", "k", "v_0")
    kernel_kind = 6

    def __init__(self, statistics: Dict, name: str = "angles") -> None:
        super().__init__()
        self.allowed_interaction_keys = list(statistics.keys())
        self.name = name
        self.order = 3
        for f in self._fields:
            self.register_buffer(f, type_table(statistics, 3, f))

    def data2parameters(self, data) -> Dict[str, torch.Tensor]:
        tt = self.types_of_terms(data)
        return {f: getattr(self, f)[tt].flatten() for f in self._fields}

    def term_energies(self, data):
        return RestrictedQuartic.compute(self.data2features(data).flatten(), **self.data2parameters(data))

    @staticmethod
    def compute_features(pos, mapping):
        return compute_angles_raw(pos, mapping)      # theta in radians

    @staticmethod
    def compute(x, a, b, c, d, k, v_0):
        cos, sin = torch.cos(x), torch.sin(x)
        quart = a * torch.pow(cos, 4) + b * torch.pow(cos, 3) + c * torch.pow(cos, 2) + d * cos
        return quart + k / (sin ** 2) + v_0
