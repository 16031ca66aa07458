"""Common machinery of the classical prior terms (reference prior/base.py, prior/harmonic.py:59-120):
per-type parameter tables indexed by the bead types of each term, features from internal coordinates,
per-molecule scatter with `mapping_batch`."""
from typing import Dict, Tuple

import torch

from ..data._keys import ATOM_TYPE_KEY, ENERGY_KEY, POSITIONS_KEY
from ..neighbor_list import make_neighbor_list


def type_table(statistics: Dict, order: int, field, default: float = 0.0) -> torch.Tensor:
    """Dense [T]*order table of one parameter; `field` is a key or a callable(stat_dict) -> float."""
    keys = list(statistics.keys())
    t_max = int(torch.tensor(keys).max()) + 1
    assert int(torch.tensor(keys).min()) >= 0
    table = torch.full((t_max,) * order, float(default))
    for key, st in statistics.items():
        table[tuple(key)] = float(field(st) if callable(field) else st[field])
    return table


class _Prior(torch.nn.Module):
    name: str = ""
    order: int = 2
    #: kind code of the CUDA prior kernels (include/fmd_b200.h FMD_PRIOR_*), None = not accelerated
    kernel_kind = None

    def types_of_terms(self, data) -> Tuple[torch.Tensor, ...]:
        mapping = data.neighbor_list[self.name]["index_mapping"]
        return tuple(data[ATOM_TYPE_KEY][mapping[i]] for i in range(self.order))

    def data2features(self, data) -> torch.Tensor:
        return self.compute_features(data[POSITIONS_KEY], data.neighbor_list[self.name]["index_mapping"])

    def term_energies(self, data) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, data):
        y = self.term_energies(data)
        mb = data.neighbor_list[self.name]["mapping_batch"]
        n_mol = data.ptr.numel() - 1 if "ptr" in data else 1
        energy = torch.zeros(n_mol, dtype=y.dtype, device=y.device).index_add(0, mb, y)
        data.out[self.name] = {ENERGY_KEY: energy}
        return data

    @staticmethod
    def _nl(name: str, order: int, topology_or_mapping) -> Dict:
        return {name: make_neighbor_list(name, order, torch.as_tensor(topology_or_mapping))}
