"""PyG-free `AtomicData` + collate (reference data/atomic_data.py:21-237 subclasses
torch_geometric.data.Data; the container here is a thin attribute/mapping object with the same
fields, `.to()`, and the same collate rules: tensors whose key contains "index" are concatenated
along the last dim and offset by the running node count, `mapping_batch` is offset by one per
sample (atomic_data.py:96-103), `batch` / `ptr` are added)."""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, List, Optional

import torch

from ._keys import (ATOM_TYPE_KEY, ENERGY_KEY, FORCE_KEY, MASS_KEY, N_ATOMS_KEY, NEIGHBOR_LIST_KEY, POSITIONS_KEY,
                    TAG_KEY, VELOCITY_KEY)


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_tensors(v, fn) for v in obj)
    return obj


class AtomicData:
    """One configuration (or a collated batch of configurations)."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            self._store[k] = v
        self._store.setdefault("out", {})

    # ---- mapping / attribute protocol
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store and self._store[key] is not None

    def keys(self):
        return list(self._store.keys())

    def __deepcopy__(self, memo):
        return AtomicData(**deepcopy(self._store, memo))

    def clone(self):
        return deepcopy(self)

    def __repr__(self):
        fields = ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else type(v).__name__}"
                           for k, v in self._store.items())
        return f"AtomicData({fields})"

    def to(self, *args, **kwargs):
        def mv(t):
            if t.is_floating_point():
                return t.to(*args, **kwargs)
            dev_only = {k: v for k, v in kwargs.items() if k == "device"}
            a = [x for x in args if not isinstance(x, torch.dtype)]
            return t.to(*a, **dev_only) if (a or dev_only) else t
        for k in list(self._store.keys()):
            self._store[k] = _map_tensors(self._store[k], mv)
        return self

    @property
    def num_nodes(self):
        return self._store[POSITIONS_KEY].shape[0]

    # ---- constructors
    @staticmethod
    def from_ase(frame, energy_tag: str = ENERGY_KEY, force_tag: str = FORCE_KEY) -> "AtomicData":
        """From an ase.Atoms-like object (duck-typed: get_atomic_numbers / get_positions / get_masses / get_pbc /
        get_cell, .info, .arrays), as reference data/atomic_data.py:106-152: types = atomic numbers, tag / energy from
        `frame.info`, forces from `frame.arrays`."""
        pos = torch.from_numpy(frame.get_positions())
        return AtomicData.from_points(
            pos=pos, atom_types=torch.from_numpy(frame.get_atomic_numbers()), masses=torch.from_numpy(frame.get_masses()),
            pbc=torch.from_numpy(frame.get_pbc()), cell=torch.tensor(frame.get_cell().tolist(), dtype=pos.dtype),
            tag=frame.info.get("tag"), energy=frame.info.get(energy_tag), forces=frame.arrays.get(force_tag))

    @staticmethod
    def from_points(pos: torch.Tensor, atom_types: torch.Tensor, masses: Optional[torch.Tensor] = None,
                    pbc=None, cell=None, tag: Optional[str] = None, energy=None, forces=None,
                    velocities: Optional[torch.Tensor] = None,
                    neighborlist: Optional[Dict[str, Dict[str, Any]]] = None, **kwargs) -> "AtomicData":
        """Reference data/atomic_data.py:154-237."""
        pos = torch.as_tensor(pos)
        assert pos.dim() == 2 and pos.shape[1] == 3, "pos must be [n_atoms, 3]"
        atom_types = torch.as_tensor(atom_types).long()
        assert atom_types.shape == (pos.shape[0],)
        d = {POSITIONS_KEY: pos, ATOM_TYPE_KEY: atom_types, N_ATOMS_KEY: torch.tensor([pos.shape[0]])}
        if masses is not None:
            masses = torch.as_tensor(masses)
            assert masses.shape == (pos.shape[0],)
            d[MASS_KEY] = masses
        if velocities is not None:
            assert velocities.shape == pos.shape
            d[VELOCITY_KEY] = torch.as_tensor(velocities)
        if energy is not None:
            d[ENERGY_KEY] = torch.as_tensor(energy)
        if forces is not None:
            assert forces.shape == pos.shape
            d[FORCE_KEY] = torch.as_tensor(forces)
        if tag is not None:
            d[TAG_KEY] = tag
        if pbc is not None and bool(torch.as_tensor(pbc).any()):
            raise NotImplementedError("periodic systems are out of scope of the B200 hot path (DESIGN.md section 7)")
        # a non-periodic frame may still carry (ignored) cell vectors, e.g. from ASE
        d[NEIGHBOR_LIST_KEY] = neighborlist if neighborlist is not None else {}
        d.update(kwargs)
        return AtomicData(**d)


def _collate_value(key: str, values: List[Any], node_offsets: List[int]):
    v0 = values[0]
    if torch.is_tensor(v0):
        if "index" in key:
            return torch.cat([v + off for v, off in zip(values, node_offsets)], dim=-1)
        if key == "mapping_batch":
            return torch.cat([v + i for i, v in enumerate(values)], dim=0)
        if v0.dim() == 0:
            return torch.stack(values)
        return torch.cat(values, dim=0)
    if isinstance(v0, dict):
        return {k: _collate_value(k, [v[k] for v in values], node_offsets) for k in v0.keys()}
    if v0 is None or isinstance(v0, (int, float, bool, str)):
        return v0 if all(v == v0 for v in values) else list(values)
    return list(values)


def collate(data_list: List[AtomicData]) -> AtomicData:
    """Batch configurations like torch_geometric's collate(..., increment=True, add_batch=True) as used by
    reference simulation/base.py:985-997."""
    assert len(data_list) > 0
    sizes = [d.num_nodes for d in data_list]
    offsets = [0]
    for s in sizes[:-1]:
        offsets.append(offsets[-1] + s)
    out = {}
    for key in data_list[0].keys():
        if key == "out":
            continue
        out[key] = _collate_value(key, [d[key] for d in data_list], offsets)
    dev = data_list[0][POSITIONS_KEY].device
    out["batch"] = torch.repeat_interleave(torch.arange(len(sizes), device=dev), torch.tensor(sizes, device=dev))
    ptr = torch.zeros(len(sizes) + 1, dtype=torch.long, device=dev)
    ptr[1:] = torch.cumsum(torch.tensor(sizes, device=dev), 0)
    out["ptr"] = ptr
    out["out"] = {}
    return AtomicData(**out)
