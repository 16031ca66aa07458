"""Gaussian radial basis (math of reference models/radial_basis/gaussian.py:9-102):
centres mu = linspace(lower, upper, R), coeff gamma = -0.5 / (mu_1 - mu_0)^2,
rbf_k(d) = exp(gamma (d - mu_k)^2) * cutoff(d)."""
from typing import Union

import torch
from torch import nn

from ..cutoff import IdentityCutoff, _Cutoff


class _RadialBasis(nn.Module):
    pass


class GaussianBasis(_RadialBasis):
    def __init__(self, cutoff: Union[int, float, _Cutoff], num_rbf: int = 50, trainable: bool = False):
        super().__init__()
        if isinstance(cutoff, (int, float)):
            self.cutoff = IdentityCutoff(0, cutoff)
        elif isinstance(cutoff, _Cutoff):
            self.cutoff = cutoff
        else:
            raise TypeError(f"Supplied cutoff {cutoff} is neither a number nor a _Cutoff instance.")
        self.check_cutoff()
        self.num_rbf = num_rbf
        self.trainable = trainable
        offset, coeff = self._initial_params()
        if trainable:
            self.coeff = nn.Parameter(coeff)
            self.offset = nn.Parameter(offset)
        else:
            self.register_buffer("coeff", coeff)
            self.register_buffer("offset", offset)

    def check_cutoff(self):
        if self.cutoff.cutoff_upper < self.cutoff.cutoff_lower:
            raise ValueError("Upper cutoff is less than lower cutoff")

    def _initial_params(self):
        offset = torch.linspace(self.cutoff.cutoff_lower, self.cutoff.cutoff_upper, self.num_rbf)
        coeff = -0.5 / (offset[1] - offset[0]) ** 2
        return offset, coeff

    def reset_parameters(self):
        offset, coeff = self._initial_params()
        self.offset.data.copy_(offset)
        self.coeff.data.copy_(coeff)

    def forward(self, dist: torch.Tensor) -> torch.Tensor:
        diff = dist.unsqueeze(-1) - self.offset
        return self.cutoff(dist).unsqueeze(-1) * torch.exp(self.coeff * diff * diff)
