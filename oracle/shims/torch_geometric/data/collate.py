import torch


def _collate_value(key, values, data_list, incs_fn):
    v0 = values[0]
    if isinstance(v0, torch.Tensor):
        cat_dim = data_list[0].__cat_dim__(key, v0)
        outs = []
        running = 0
        for d, v in zip(data_list, values):
            inc = d.__inc__(key, v)
            if v.dim() == 0:
                v = v.unsqueeze(0)
            outs.append(v + running if (inc != 0 and v.dtype in (torch.long, torch.int32)) else v)
            running = running + inc
        return torch.cat(outs, dim=cat_dim if outs[0].dim() > 0 else 0)
    if isinstance(v0, dict):
        return {k: _collate_value(k, [v[k] for v in values], data_list, incs_fn) for k in v0.keys()}
    if isinstance(v0, (int, float, str, bool)) or v0 is None:
        return v0 if all(v == v0 for v in values) else list(values)
    return list(values)


def collate(cls, data_list, increment=True, add_batch=True, **kwargs):
    out = cls.__new__(cls)
    object.__setattr__(out, "_store", {})
    keys = list(data_list[0]._store.keys())
    for key in keys:
        values = [d._store[key] for d in data_list]
        if key == "out":
            out._store[key] = {}
            continue
        out._store[key] = _collate_value(key, values, data_list, None)
    if add_batch:
        sizes = [d.num_nodes for d in data_list]
        out._store["batch"] = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(torch.tensor(sizes), 0)
        out._store["ptr"] = ptr
    return out, None, None
