"""TEST INFRASTRUCTURE: lets the reference's own test files (which import torch_geometric's collate) run against the
drop-in package, whose AtomicData is not a PyG Data object; collate is served by flashmd.data.collate."""
