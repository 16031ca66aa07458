"""Neighbour-list dictionaries (reference neighbor_list/neighbor_list.py:6-163)."""
from typing import Dict, Optional

import torch


def make_neighbor_list(tag: str, order: int, index_mapping: torch.Tensor, mapping_batch: Optional[torch.Tensor] = None,
                       cell_shifts: Optional[torch.Tensor] = None, rcut: Optional[float] = None,
                       self_interaction: Optional[bool] = None) -> Dict:
    index_mapping = torch.as_tensor(index_mapping, dtype=torch.long)
    if index_mapping.shape[0] != order:
        raise RuntimeError(f"index_mapping shape does not match the order:{index_mapping.shape[0]} != {order}")
    if mapping_batch is None:
        mapping_batch = torch.zeros(index_mapping.shape[1] if index_mapping.numel() else 0, dtype=torch.long,
                                    device=index_mapping.device)
    return dict(tag=tag, order=order, index_mapping=index_mapping, cell_shifts=cell_shifts, rcut=rcut,
                self_interaction=self_interaction, mapping_batch=mapping_batch)


def validate_neighborlist(nl) -> bool:
    """True iff `nl` is a well-formed neighbour-list dict (reference :131-163)."""
    if not isinstance(nl, dict):
        return False
    for key in ("tag", "order", "index_mapping", "mapping_batch"):
        if key not in nl:
            return False
    im = nl["index_mapping"]
    return torch.is_tensor(im) and im.dim() == 2 and im.shape[0] == nl["order"]


def atomic_data2neighbor_list(data, rcut: float, self_interaction: bool = False, max_num_neighbors: int = 1000) -> Dict:
    """Radius graph of a (collated) AtomicData as a neighbour-list dict (reference :6-63): row 0 = centre,
    row 1 = neighbour, centre-major / neighbour-ascending."""
    from .torch_impl import torch_neighbor_list
    idx = torch_neighbor_list(data, rcut, self_interaction=self_interaction, max_num_neighbors=max_num_neighbors)
    batch = data["batch"] if "batch" in data else torch.zeros(data.pos.shape[0], dtype=torch.long, device=data.pos.device)
    return make_neighbor_list(tag="", order=2, index_mapping=idx, mapping_batch=batch[idx[0]], rcut=rcut,
                              self_interaction=self_interaction)
