def collate(cls, data_list, increment=True, add_batch=True, **kwargs):
    """(batch, slices, incs) like torch_geometric.data.collate.collate; the batch comes from the drop-in's own collate."""
    from flashmd.data import collate as _collate
    return _collate(list(data_list)), None, None
