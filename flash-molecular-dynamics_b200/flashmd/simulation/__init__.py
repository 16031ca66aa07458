from .base import _Simulation  # noqa: F401
from .langevin import LangevinSimulation, OverdampedSimulation  # noqa: F401
from .parallel_tempering import PTSimulation  # noqa: F401
from .velocity_verlet import NVESimulation  # noqa: F401
from .specialize_prior import condense_all_priors_for_simulation  # noqa: F401
from .cli import parse_simulation_config  # noqa: F401
