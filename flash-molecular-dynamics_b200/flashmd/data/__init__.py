from .atomic_data import AtomicData, collate  # noqa: F401
from . import _keys  # noqa: F401
