"""Minimal stand-in for the absent `torch_geometric` (test infrastructure only).

Implements exactly what the reference touches on the CGSchNet path:
`data.Data`, `data.collate.collate`, `nn.MessagePassing`, `utils.scatter`.
"""
__version__ = "2.4.0"  # < 2.5 so the reference's inspector monkey-patch is skipped
from . import data, nn, utils  # noqa: F401
