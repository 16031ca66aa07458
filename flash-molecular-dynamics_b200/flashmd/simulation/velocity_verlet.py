"""NVE velocity-Verlet dynamics (reference simulation/velocity_verlet.py:12-95):
    v += dt/2 F/m ; x += dt v ; F = -grad U(x) ; v += dt/2 F/m
Initial velocities are Maxwell-Boltzmann at `beta`.  On a CUDA device the step is the fused engine's BAOAB
kernels with friction 0 (vscale = 1, noisescale = 0), which is exactly velocity Verlet: B, A(dt/2), A(dt/2), B."""
from typing import List, Union

import torch

from ..data._keys import MASS_KEY, POSITIONS_KEY, VELOCITY_KEY
from .langevin import LangevinSimulation


class NVESimulation(LangevinSimulation):
    def __init__(self, **kwargs):
        kwargs.pop("friction", None)
        super().__init__(friction=1.0, **kwargs)
        self.friction = 0.0
        self.vscale, self.noisescale = 1.0, 0.0

    def timestep(self, data, forces):
        v, x, m = data[VELOCITY_KEY], data[POSITIONS_KEY], data[MASS_KEY][:, None]
        v_half = v + 0.5 * self.dt * forces / m
        data[POSITIONS_KEY] = x + self.dt * v_half
        potential, forces = self.calculate_potential_and_forces(data)
        data[VELOCITY_KEY] = v_half + 0.5 * self.dt * forces / m
        return data, potential, forces
