import copy
import torch


class Data:
    """Attribute/dict style container (subset of torch_geometric.data.Data)."""

    def __init__(self, **kwargs):
        object.__setattr__(self, "_store", {})
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute access -------------------------------------------------
    def __getattr__(self, key):
        store = object.__getattribute__(self, "_store")
        if key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key.startswith("_") and key != "_store":
            object.__setattr__(self, key, value)
        else:
            self._store[key] = value

    def __delattr__(self, key):
        if key in self._store:
            del self._store[key]
        else:
            object.__delattr__(self, key)

    # mapping access ---------------------------------------------------
    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store and self._store[key] is not None

    def keys(self):
        return [k for k, v in self._store.items() if v is not None]

    def get(self, key, default=None):
        return self._store.get(key, default)

    @property
    def num_nodes(self):
        if "pos" in self._store and self._store["pos"] is not None:
            return self._store["pos"].shape[0]
        return None

    def __inc__(self, key, value, *args, **kwargs):
        return self.num_nodes if "index" in key else 0

    def __cat_dim__(self, key, value, *args, **kwargs):
        return -1 if "index" in key else 0

    def to(self, device=None, dtype=None, **kw):
        def mv(v):
            if isinstance(v, torch.Tensor):
                return v.to(device=device) if device is not None else v
            if isinstance(v, dict):
                return {k: mv(x) for k, x in v.items()}
            return v

        for k in list(self._store.keys()):
            self._store[k] = mv(self._store[k])
        return self

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        object.__setattr__(new, "_store", copy.deepcopy(self._store, memo))
        for k, v in self.__dict__.items():
            if k != "_store":
                object.__setattr__(new, k, copy.deepcopy(v, memo))
        return new

    def __getstate__(self):
        return self.__dict__

    def __setstate__(self, state):
        for k, v in state.items():
            object.__setattr__(self, k, v)
