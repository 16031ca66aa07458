"""Sparse storage of the per-type parameter tables of prior modules (reference models/utils.py): the tables are dense
[T]^order tensors that are mostly zeros for real force fields; checkpoints store them sparse."""
import warnings

import torch

_TABLES = {"FourierSeries": ("v_0", "k1s", "k2s"), "Harmonic": ("x_0", "k")}


def _table_names(module):
    from ..prior import FourierSeries, Harmonic
    if isinstance(module, FourierSeries):
        return _TABLES["FourierSeries"]
    if isinstance(module, Harmonic):
        return _TABLES["Harmonic"]
    return None


def sparsify_prior_module(module) -> torch.nn.Module:
    """In place: parameter buffers -> sparse COO tensors (Harmonic and Dihedral / FourierSeries priors)."""
    names = _table_names(module)
    if names is None:
        warnings.warn("Module is not supported for sparsification. It will be returned as is")
        return module
    for n in names:
        setattr(module, n, getattr(module, n).to_sparse())
    return module


def desparsify_prior_module(module) -> torch.nn.Module:
    """In place: the inverse of sparsify_prior_module."""
    for n in _table_names(module) or ():
        setattr(module, n, getattr(module, n).to_dense())
    return module
