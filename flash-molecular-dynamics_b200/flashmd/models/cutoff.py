"""Smooth cut-off envelopes (math of reference models/cutoff.py:75-200)."""
import math
from typing import Union

import torch


class _Cutoff(torch.nn.Module):
    """Envelope on [cutoff_lower, cutoff_upper]."""

    cutoff_lower: float = 0.0
    cutoff_upper: float = 0.0

    def check_cutoff(self):
        if self.cutoff_upper < self.cutoff_lower:
            raise ValueError(f"Upper cutoff {self.cutoff_upper} is less than lower cutoff {self.cutoff_lower}")


class _OneSidedCutoff(_Cutoff):
    cutoff_lower = 0.0


class IdentityCutoff(_Cutoff):
    """All ones on the interval (used by GaussianBasis when only a radius is given)."""

    def __init__(self, cutoff_lower: float = 0.0, cutoff_upper: float = 10000.0):
        super().__init__()
        self.cutoff_lower, self.cutoff_upper = cutoff_lower, cutoff_upper
        self.check_cutoff()

    def forward(self, distances: torch.Tensor) -> torch.Tensor:
        return torch.ones_like(distances)


class CosineCutoff(_Cutoff):
    """0.5 (cos(pi d / r_hi) + 1) for d < r_hi when r_lo == 0; for r_lo > 0 the shifted two-sided form
    0.5 cos(pi (2 (d - r_lo)/(r_hi - r_lo) + 1)) + 0.5 on (r_lo, r_hi)."""

    def __init__(self, cutoff_lower: float = 0.0, cutoff_upper: float = 5.0):
        super().__init__()
        self.cutoff_lower, self.cutoff_upper = cutoff_lower, cutoff_upper
        self.check_cutoff()

    def forward(self, distances: torch.Tensor) -> torch.Tensor:
        lo, hi = self.cutoff_lower, self.cutoff_upper
        inside = (distances < hi).to(distances.dtype)
        if lo > 0:
            env = 0.5 * (torch.cos(math.pi * (2.0 * (distances - lo) / (hi - lo) + 1.0)) + 1.0)
            return env * inside * (distances > lo).to(distances.dtype)
        return 0.5 * (torch.cos(distances * math.pi / hi) + 1.0) * inside


class ShiftedCosineCutoff(_OneSidedCutoff):
    """1 below r_hi - sigma, cosine switch to 0 at r_hi."""

    def __init__(self, cutoff: Union[int, float] = 5.0, smooth_width: Union[int, float] = 0.5):
        super().__init__()
        self.cutoff_upper = cutoff
        self.smooth_width = smooth_width

    def forward(self, distances: torch.Tensor) -> torch.Tensor:
        start = self.cutoff_upper - self.smooth_width
        sw = 0.5 + 0.5 * torch.cos(math.pi * (distances - start) / self.smooth_width)
        out = torch.where(distances > start, sw, torch.ones_like(distances))
        return torch.where(distances > self.cutoff_upper, torch.zeros_like(distances), out).view(-1)
