"""BAOAB Langevin dynamics (reference simulation/langevin.py:30-310).

    B: v += dt/2 F/m   A: x += dt/2 v   O: v = v e^{-gamma dt} + sqrt(1 - e^{-2 gamma dt}) sqrt(1/(beta m)) xi
    A: x += dt/2 v     [forces at the new x]              B: v += dt/2 F/m

On a CUDA device the step is lowered to the fused engine: `fmd_baoab_pre` (B-A-O-A with in-kernel
counter-based Philox noise, or torch-generated noise when noise_source="torch"), the force field, and
`fmd_baoab_post`, replayed as one CUDA graph."""
from typing import Any, List, Union

import numpy as np
import torch

from ..data._keys import MASS_KEY, POSITIONS_KEY, VELOCITY_KEY
from .base import _Simulation


class LangevinSimulation(_Simulation):
    def __init__(self, friction: float = 1e-3, **kwargs: Any):
        super().__init__(**kwargs)
        assert friction > 0
        self.friction = friction
        self.vscale = np.exp(-self.dt * self.friction)
        self.noisescale = np.sqrt(1 - self.vscale * self.vscale)

    @staticmethod
    def sample_maxwell_boltzmann(betas: torch.Tensor, masses: torch.Tensor) -> torch.Tensor:
        """v ~ N(0, 1/(beta m)) per component; drawn with the GLOBAL torch RNG like the reference (:79-99)."""
        assert bool((masses > 0).all())
        scale = torch.sqrt(1.0 / (betas * masses))
        return torch.distributions.Normal(loc=0.0, scale=scale).sample((3,)).t()

    def _attach_configurations(self, configurations: List, beta: Union[float, List[float]]):
        super()._attach_configurations(configurations, beta)
        d = self.initial_data
        if VELOCITY_KEY not in d:
            d[VELOCITY_KEY] = LangevinSimulation.sample_maxwell_boltzmann(
                self.beta.repeat_interleave(self.n_atoms), d[MASS_KEY]).to(self.dtype)
        assert d[VELOCITY_KEY].shape == d[POSITIONS_KEY].shape
        self.beta_mass_ratio = torch.sqrt(1.0 / self.beta.repeat_interleave(self.n_atoms) / d[MASS_KEY])[:, None].to(self.dtype)

    # ------------------------------------------------------------------ module path (CPU / --disable_optim)
    def timestep(self, data, forces):
        v, x, m = data[VELOCITY_KEY], data[POSITIONS_KEY], data[MASS_KEY]
        hdt = 0.5 * self.dt
        v = v + hdt * forces / m[:, None]
        x = x + v * hdt
        noise = self._noise_buffer.normal_(generator=self.rng)
        v = v * self.vscale + self.noisescale * self.beta_mass_ratio * noise
        x = x + v * hdt
        data[POSITIONS_KEY] = x
        potential, forces = self.calculate_potential_and_forces(data)
        v = v + hdt * forces / m[:, None]
        data[VELOCITY_KEY] = v
        return data, potential, forces

    def _set_up_simulation(self, overwrite: bool = False):
        super()._set_up_simulation(overwrite)
        self.simulated_kinetic_energies = torch.zeros(self._save_size, self.n_sims) if self.save_energies else None
        self._noise_buffer = torch.empty((self.n_sims * self.n_atoms, self.n_dims), dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ fused path
    def _build_engine(self, data):
        if self.device.type != "cuda" or self.dtype != torch.float32 or getattr(self, "force_module_path", False):
            return None
        from ..engine import LangevinEngine
        from .lowering import NotLowerable, lower
        try:
            ff = lower(self.model, data, "w16a16" if self.gptq == "w16a16" else "fp32", self.exact_cutoff_grad)
        except NotLowerable as err:
            # no silent second code path: a CUDA fp32 run either uses the fused step or is asked for the module path
            raise RuntimeError(f"the model cannot be lowered to the fused CUDA step ({err}); set "
                               "`simulation.force_module_path = True` (CLI: --disable_optim) to run the PyTorch module "
                               "path on purpose") from err
        seed = self.random_seed if self.random_seed is not None else 0
        return LangevinEngine(ff, data[POSITIONS_KEY], data[VELOCITY_KEY], data[MASS_KEY], self.beta, self.dt,
                              self.friction, seed=seed, use_graph=True,
                              noise_mode="buffer" if self.noise_source == "torch" else "philox",
                              node_offset=self._shard_node_offset())

    def _shard_node_offset(self) -> int:
        """Global index of this rank's first bead when the replicas are sharded over ranks (equal shards): keys the
        Philox noise by the GLOBAL bead index, so shards never share a stream and match the unsharded run."""
        off = getattr(self, "_node_offset", None)
        if off is not None:
            return int(off)
        from .distributed import dist_info
        rank, world = dist_info()
        return rank * self.n_sims * self.n_atoms if world > 1 else 0

    def _engine_timestep(self, eng):
        if eng.noise_buf is not None:
            eng.noise_buf.normal_(generator=self.rng)
        eng.step()

    # ------------------------------------------------------------------ output
    def _extra_save_tensors(self) -> dict:
        return {"ke": self.engine.ke} if self.save_energies else {}

    def _traj_buffers(self) -> dict:
        b = super()._traj_buffers()
        b["ke"] = self.simulated_kinetic_energies
        return b

    def _store_extra(self, i: int, host: dict, bufs: dict):
        if self.save_energies:
            bufs["ke"][i] = host["ke"]

    def save(self, pos, vel, forces, potential, t: int):
        i = super().save(pos, vel, forces, potential, t)
        if self.save_energies and self.engine is None:      # module path (the fused path stores KE asynchronously)
            m = self.initial_data[MASS_KEY].view(self.n_sims, self.n_atoms)
            ke = 0.5 * (m[:, :, None] * vel.view(-1, self.n_atoms, self.n_dims) ** 2).sum(dim=(1, 2))
            self.simulated_kinetic_energies[i] = ke.cpu()
        return i

    def write(self):
        if self._saver is not None:
            self._saver.drain()
        if self.save_energies:
            np.save(f"{self.filename}_kineticenergy_{self._get_numpy_count()}.npy",
                    self._swap_and_export(self.simulated_kinetic_energies))
            self.simulated_kinetic_energies = torch.zeros(self._save_size, self.n_sims)
        super().write()

    def reshape_output(self):
        super().reshape_output()
        if self.save_energies:
            self.simulated_kinetic_energies = self._swap_and_export(self.simulated_kinetic_energies)


class OverdampedSimulation(_Simulation):
    """Overdamped Langevin (Brownian) dynamics with the reference's exact update (simulation/langevin.py:315-420):
        D = 1 / (beta friction) per bead,  x += F D dt + sqrt(2 D dt) xi
    (the drift carries no extra beta - the reference's own convention, reproduced for parity).  Masses and velocities
    are not used.  CUDA fp32: the fused engine (fmd_overdamped_step + the force field, one CUDA graph per step); otherwise
    the module path.  Pinned by tests/golden/integrators_n54_b4.npz."""

    def __init__(self, friction: float = 1.0, **kwargs: Any):
        super().__init__(**kwargs)
        assert friction > 0
        self.friction = friction

    def _attach_configurations(self, configurations: List, beta: Union[float, List[float]]):
        super()._attach_configurations(configurations, beta)
        if MASS_KEY in self.initial_data:
            import warnings
            warnings.warn("Masses were provided, but will not be used since an overdamped Langevin scheme is being "
                          "used for integration.")
        self.expanded_beta = self.beta.repeat_interleave(self.n_atoms)[:, None]
        self.diffusion = 1 / self.expanded_beta / self.friction
        self._dtau = self.diffusion * self.dt

    def _set_up_simulation(self, overwrite: bool = False):
        super()._set_up_simulation(overwrite)
        self._noise_buffer = torch.empty((self.n_sims * self.n_atoms, self.n_dims), dtype=self.dtype, device=self.device)

    # ------------------------------------------------------------------ fused path
    def _build_engine(self, data):
        if self.device.type != "cuda" or self.dtype != torch.float32 or getattr(self, "force_module_path", False):
            return None
        from ..engine import OverdampedEngine
        from .lowering import NotLowerable, lower
        try:
            ff = lower(self.model, data, "w16a16" if self.gptq == "w16a16" else "fp32", self.exact_cutoff_grad)
        except NotLowerable as err:
            raise RuntimeError(f"the model cannot be lowered to the fused CUDA step ({err}); set "
                               "`simulation.force_module_path = True` to run the PyTorch module path on purpose") from err
        from .distributed import dist_info
        rank, world = dist_info()
        return OverdampedEngine(ff, data[POSITIONS_KEY], self.beta, self.dt, self.friction,
                                seed=self.random_seed if self.random_seed is not None else 0, use_graph=True,
                                noise_mode="buffer" if self.noise_source == "torch" else "philox",
                                node_offset=rank * self.n_sims * self.n_atoms if world > 1 else 0)

    def _engine_timestep(self, eng):
        if eng.noise_buf is not None:
            eng.noise_buf.normal_(generator=self.rng)
        eng.step()

    def timestep(self, data, forces):
        noise = self._noise_buffer.normal_(generator=self.rng)
        dtau = self._dtau.to(device=forces.device, dtype=forces.dtype)
        data[POSITIONS_KEY] = data[POSITIONS_KEY].detach() + forces * dtau + torch.sqrt(2 * dtau) * noise
        potential, forces = self.calculate_potential_and_forces(data)
        return data, potential, forces
