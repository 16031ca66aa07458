from .base import _Prior  # noqa: F401
from .fourier_series import Dihedral, FourierSeries  # noqa: F401
from .harmonic import (GeneralAngles, GeneralBonds, Harmonic, HarmonicAngles, HarmonicAnglesRaw, HarmonicBonds,  # noqa: F401
                       HarmonicImpropers, ShiftedPeriodicHarmonicImpropers)
from .polynomial import Polynomial, QuarticAngles  # noqa: F401
from .repulsion import Repulsion  # noqa: F401
from .restricted_bending import RestrictedQuartic  # noqa: F401
