"""Linear/activation stack with Xavier-uniform weights and zero biases (reference models/mlp.py:6-57,
models/_module_init.py:4-28)."""
from typing import List

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
