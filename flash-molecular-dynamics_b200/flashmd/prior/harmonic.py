"""k (x - x0)^2 priors on distances and cos(angle) (reference prior/harmonic.py:23-330)."""
import math
from typing import Dict

import torch

from ..geometry import compute_angles_cos, compute_angles_raw, compute_distances, compute_torsions
from .base import _Prior, type_table


class Harmonic(_Prior):
    def __init__(self, statistics: Dict, name: str, order: int) -> None:
        super().__init__()
        self.allowed_interaction_keys = list(statistics.keys())
        self.name = name
        self.order = order
        self.register_buffer("x_0", type_table(statistics, order, "x_0"))
        self.register_buffer("k", type_table(statistics, order, "k"))

    def data2parameters(self, data) -> Dict:
        tt = self.types_of_terms(data)
        return {"x0": self.x_0[tt], "k": self.k[tt]}

    @staticmethod
    def compute(x, x0, k, V0=0):
        return k * (x - x0) ** 2 + V0

    def term_energies(self, data):
        p = self.data2parameters(data)
        return Harmonic.compute(self.data2features(data), p["x0"], p["k"])


class HarmonicBonds(Harmonic):
    name = "bonds"
    kernel_kind = 0

    def __init__(self, statistics) -> None:
        super().__init__(statistics, HarmonicBonds.name, order=2)

    @staticmethod
    def compute_features(pos, mapping):
        return compute_distances(pos, mapping)

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(HarmonicBonds.name, 2, topology)


class HarmonicAngles(Harmonic):
    """Harmonic in cos(theta)."""
    name = "angles"
    kernel_kind = 1

    def __init__(self, statistics) -> None:
        super().__init__(statistics, HarmonicAngles.name, order=3)

    @staticmethod
    def compute_features(pos, mapping):
        return compute_angles_cos(pos, mapping)

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(HarmonicAngles.name, 3, topology)


class HarmonicImpropers(Harmonic):
    name = "impropers"
    _order = 4
    kernel_kind = 8

    def __init__(self, statistics) -> None:
        super().__init__(statistics, HarmonicImpropers.name, order=4)

    @staticmethod
    def compute_features(pos, mapping):
        return compute_torsions(pos, mapping)

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(HarmonicImpropers.name, 4, topology)


class HarmonicAnglesRaw(Harmonic):
    """Harmonic in theta itself (radians) (reference prior/harmonic.py:267-300; there the constructor forgets Harmonic's
    `order` argument and raises - same signature here, working)."""
    name = "angles"
    kernel_kind = 7

    def __init__(self, statistics, name="angles") -> None:
        super().__init__(statistics, HarmonicAnglesRaw.name, order=3)
        self.name = name

    @staticmethod
    def compute_features(pos, mapping):
        return compute_angles_raw(pos, mapping)

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(HarmonicAnglesRaw.name, 3, topology)


class GeneralBonds(Harmonic):
    """Harmonic bonds registered under a caller-chosen name (several bond sets in one model; reference
    prior/harmonic.py:407-428)."""
    _order = 2
    kernel_kind = 0

    def __init__(self, statistics, name) -> None:
        super().__init__(statistics, HarmonicBonds.name, order=GeneralBonds._order)
        self.name = name

    @staticmethod
    def compute_features(pos, mapping):
        return compute_distances(pos, mapping)


class GeneralAngles(Harmonic):
    """Harmonic cos(angle) terms registered under a caller-chosen name (reference prior/harmonic.py:430-450)."""
    _order = 3
    kernel_kind = 1

    def __init__(self, statistics, name) -> None:
        super().__init__(statistics, HarmonicAngles.name, order=GeneralAngles._order)
        self.name = name

    @staticmethod
    def compute_features(pos, mapping):
        return compute_angles_cos(pos, mapping)


class ShiftedPeriodicHarmonicImpropers(Harmonic):
    """Harmonic improper torsions for distributions centred on +-pi (e.g. omega): torsions below zero are shifted by 2 pi
    and pi is subtracted, so the harmonic well sits at the discontinuity (reference prior/harmonic.py:327-405)."""
    name = "impropers"
    _order = 4
    kernel_kind = 9

    def __init__(self, statistics) -> None:
        super().__init__(statistics, ShiftedPeriodicHarmonicImpropers.name, order=4)

    @staticmethod
    def compute_features(pos, mapping):
        phi = compute_torsions(pos, mapping)
        # the reference's pi is torch.tensor(pi), i.e. the FLOAT32 value 3.14159274..., also on its fp64 path (harmonic.py:20)
        pi32 = float(torch.tensor(math.pi, dtype=torch.float32))
        return torch.where(phi < 0, phi + 2 * pi32, phi) - pi32

    @staticmethod
    def neighbor_list(topology) -> Dict:
        return _Prior._nl(HarmonicImpropers.name, 4, topology)
