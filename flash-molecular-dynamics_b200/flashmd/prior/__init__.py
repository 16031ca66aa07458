from .base import _Prior  # noqa: F401
from .fourier_series import Dihedral, FourierSeries  # noqa: F401
from .harmonic import Harmonic, HarmonicAngles, HarmonicBonds, HarmonicImpropers  # noqa: F401
from .repulsion import Repulsion  # noqa: F401
