"""Energy -> forces wrappers (reference models/gradients.py:42-290): `GradientsOut` differentiates a
sub-model's summed energy w.r.t. positions with autograd, `SumOut` adds the sub-models' outputs.
(Simulations replace this by the analytic backward of the fused engine.)"""
from typing import List, Sequence

import torch

from ..data._keys import ENERGY_KEY, FORCE_KEY, POSITIONS_KEY


class SumOut(torch.nn.Module):
    name: str = "SumOut"

    def __init__(self, models: torch.nn.ModuleDict, targets: List[str] = None):
        super().__init__()
        self.targets = targets if targets is not None else [ENERGY_KEY, FORCE_KEY]
        self.models = models

    def forward(self, data):
        for t in self.targets:
            data.out[t] = 0.0
        for name, model in self.models.items():
            data = model(data)
            for t in self.targets:
                data.out[t] = data.out[t] + data.out[name][t]
        return data

    def neighbor_list(self, **kwargs):
        nl = {}
        for model in self.models.values():
            nl.update(model.neighbor_list(**kwargs))
        return nl


class EnergyOut(torch.nn.Module):
    name: str = "EnergyOut"

    def __init__(self, model: torch.nn.Module, targets: List[str] = None):
        super().__init__()
        self.model = model
        self.name = model.name
        self.targets = targets if targets is not None else [ENERGY_KEY]

    def forward(self, data):
        data = self.model(data)
        data.out[ENERGY_KEY] = data.out[self.name][ENERGY_KEY]
        return data


class GradientsOut(torch.nn.Module):
    _targets = {FORCE_KEY: ENERGY_KEY}

    def __init__(self, model: torch.nn.Module, targets=FORCE_KEY):
        super().__init__()
        self.model = model
        self.name = model.name
        if isinstance(targets, str):
            self.targets = [targets]
        elif isinstance(targets, Sequence):
            self.targets = list(targets)
        else:
            raise ValueError("targets must be a string or a sequence of strings")
        assert all(t in self._targets for t in self.targets)

    def forward(self, data):
        pos = data[POSITIONS_KEY]
        pos.requires_grad_(True)
        data = self.model(data)
        energy = data.out[self.name][ENERGY_KEY]
        if FORCE_KEY in self.targets:
            (grad,) = torch.autograd.grad(energy.sum(), pos, create_graph=self.training)
            data.out[self.name][FORCE_KEY] = -grad
        data.out[self.name][ENERGY_KEY] = energy if self.training else energy.detach()
        data[POSITIONS_KEY] = pos.detach() if not self.training else pos
        return data

    def neighbor_list(self, **kwargs):
        return self.model.neighbor_list(**kwargs)
