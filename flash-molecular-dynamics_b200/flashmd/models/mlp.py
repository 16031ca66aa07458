"""Linear/activation stack with Xavier-uniform weights and zero biases (reference models/mlp.py:6-57,
models/_module_init.py:4-28)."""
from typing import List, Optional

import torch


def init_xavier_uniform(module: torch.nn.Module) -> None:
    if isinstance(module, torch.nn.Linear):
        torch.nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            torch.nn.init.zeros_(module.bias)


class MLP(torch.nn.Module):
    """layer_widths = [in, h1, ..., out]; activation between layers, none after the last; `last_bias`
    controls the bias of the last Linear.  `self.layers` is an nn.Sequential as in the reference."""

    def __init__(self, layer_widths: List[int] = None, activation_func: torch.nn.Module = torch.nn.Tanh(),
                 last_bias: bool = True, use_triton: bool = True):
        super().__init__()
        if layer_widths is None:
            layer_widths = [10, 10, 1]
        mods = []
        n = len(layer_widths) - 1
        for i, (w_in, w_out) in enumerate(zip(layer_widths[:-1], layer_widths[1:])):
            last = i == n - 1
            mods.append(torch.nn.Linear(w_in, w_out, bias=(last_bias if last else True)))
            if not last:
                mods.append(activation_func)
        self.layers = torch.nn.Sequential(*mods)
        self.reset_parameters()

    def reset_parameters(self):
        self.layers.apply(init_xavier_uniform)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.layers(x)


class TypesMLP(torch.nn.Module):
    """Per-atom energy MLP, optionally with a separate set of weights for every atom species (reference
    models/mlp.py:60-121).  forward(features [N, in], data) -> [N, 1]."""
    name = "TypesMLP"

    def __init__(self, layer_widths: List[int], activation: torch.nn.Module = torch.nn.Tanh(),
                 species: Optional[torch.Tensor] = None):
        super().__init__()
        self.weights_per_species = species is not None
        if self.weights_per_species:
            self.register_buffer("species", torch.unique(species))
            self.mlp = torch.nn.ModuleList([MLP(layer_widths, activation) for _ in self.species])
        else:
            self.species = None
            self.mlp = MLP(layer_widths, activation)

    def reset_parameters(self):
        for mod in (self.mlp if self.weights_per_species else [self.mlp]):
            mod.reset_parameters()

    def forward(self, features, data):
        if not self.weights_per_species:
            return self.mlp(features)
        yi = torch.zeros((features.shape[0], 1), dtype=features.dtype, device=features.device)
        for sp, mlp in zip(self.species, self.mlp):
            mask = data.atom_types == sp
            yi[mask] = mlp(features[mask])
        return yi
