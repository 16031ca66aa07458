"""Simulation driver with the reference's object protocol (reference simulation/base.py:185-1477):
constructor keywords, attach_model_and_configurations, simulate(), .npy / checkpoint outputs
(`{filename}_coords_NNNN.npy` shaped [n_sims, frames, n_atoms, 3], ...), throughput metrics of the second
half of the run.  Two execution paths:

  * CUDA device + lowerable model  -> the fused engine (flashmd/engine.py): one CUDA-graph replay per step;
  * otherwise (CPU, `--disable_optim`) -> the module path: `self.model(data)` with autograd forces,
    exactly the reference's semantics (this is what the golden vectors pin).
"""
import glob
import os
import time
import warnings
from copy import deepcopy
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from ..data import AtomicData, collate
from ..data._keys import ENERGY_KEY, FORCE_KEY, MASS_KEY, POSITIONS_KEY, VELOCITY_KEY
from .specialize_prior import condense_all_priors_for_simulation


class _AsyncSaver:
    """Save points of the fused path without stalling the graphed step: the tensors of a save
    point are snapshotted on the compute stream (device-to-device, in order), copied to pinned host memory on a side
    stream, and consumed (stored into the trajectory buffers, blow-up / capacity checks) when their slot comes round again,
    at the next export, or at the end of the run.  `n_slots` save points may be in flight."""

    def __init__(self, device, n_slots: int = 3):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.n_slots = n_slots
        self.pending = []          # FIFO of (done_event, host_tensors, device_snapshots, callback)
        self.pinned = [dict() for _ in range(n_slots)]
        self.k = 0

    def submit(self, tensors, callback):
        while len(self.pending) >= self.n_slots:
            self._finish_one()
        slot = self.k % self.n_slots
        self.k += 1
        snaps = {k: v.detach().clone() for k, v in tensors.items()}     # ordered after the step on the compute stream
        ready = torch.cuda.Event()
        ready.record()
        host = {}
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            for k, v in snaps.items():
                buf = self.pinned[slot].get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = self.pinned[slot][k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                buf.copy_(v, non_blocking=True)
                host[k] = buf
            done = torch.cuda.Event()
            done.record(self.stream)
        self.pending.append((done, host, snaps, callback))

    def _finish_one(self):
        done, host, snaps, callback = self.pending.pop(0)
        done.synchronize()
        del snaps
        callback(host)

    def drain(self):
        while self.pending:
            self._finish_one()


class _Simulation:
    def __init__(self, dt: float = 5e-4, save_forces: bool = False, save_energies: bool = False,
                 save_force_components: bool = False, save_energy_components: bool = False,
                 force_components=None, energy_components=None, n_timesteps: int = 100, save_interval: int = 10,
                 create_checkpoints: bool = False, read_checkpoint_file: Union[str, bool, None] = None,
                 random_seed: Optional[int] = 233, device: str = "cpu", dtype: str = "single",
                 export_interval: Optional[int] = None, log_interval: Optional[int] = None, log_type: str = "write",
                 filename: Optional[str] = None, add_timestamp: bool = False, output_dir: str = "./outputs",
                 specialize_priors: bool = False, tqdm_refresh: float = 10, sim_subroutine: Optional[Callable] = None,
                 sim_subroutine_interval: Optional[int] = None, save_subroutine: Optional[Callable] = None,
                 compile: bool = False, compile_mode: str = "default", force_compile: bool = False,
                 compile_model: bool = True, profile_start_step: Optional[int] = None,
                 profile_end_step: Optional[int] = None, print_shape: bool = False, print_shape_steps: int = 3,
                 dump_neighbor_list: bool = False, dump_neighbor_list_last_n: Optional[int] = None,
                 gptq: Optional[str] = "w16a16", noise_source: str = "philox", exact_cutoff_grad: bool = True):
        if gptq is not None and gptq not in ["w16a16"]:
            raise ValueError(f"Unsupported GPTQ mode: {gptq}. Supported: 'w16a16'")
        if log_type not in ["print", "write"]:
            raise ValueError("log_type can be either 'print' or 'write'")
        if dtype not in ("single", "double"):
            raise ValueError("dtype must be 'single' or 'double'")
        self.model, self.initial_data, self.gptq = None, None, gptq
        self.dt, self.n_timesteps, self.save_interval = dt, n_timesteps, save_interval
        self.save_forces, self.save_energies = save_forces, save_energies
        self.save_force_components, self.save_energy_components = save_force_components, save_energy_components
        if save_force_components or save_energy_components:
            raise NotImplementedError("per-component outputs are not part of the B200 hot path")
        self.dtype = torch.float32 if dtype == "single" else torch.float64
        self.device = torch.device(device)
        self.export_interval = n_timesteps if export_interval is None else export_interval
        self.log_interval, self.log_type = log_interval, log_type
        self.create_checkpoints = create_checkpoints
        self.read_checkpoint_file = None if read_checkpoint_file is False else read_checkpoint_file
        self.output_dir = output_dir
        if filename is not None:
            os.makedirs(output_dir, exist_ok=True)
            if add_timestamp:
                filename = f"{filename}_{time.strftime('%Y%m%d_%H%M%S')}"
            filename = os.path.join(output_dir, filename)
        self.filename = filename
        self.specialize_priors = specialize_priors
        self.sim_subroutine, self.sim_subroutine_interval = sim_subroutine, sim_subroutine_interval
        self.save_subroutine = save_subroutine
        self.tqdm_refresh = tqdm_refresh
        self.profile_start_step, self.profile_end_step = profile_start_step, profile_end_step
        # accepted for config compatibility; the fused step replaces torch.compile by a CUDA graph
        self.compile, self.compile_mode, self.force_compile = compile, compile_mode, force_compile
        self._compile_model_flag = compile_model
        self.noise_source = noise_source
        self.exact_cutoff_grad = exact_cutoff_grad
        self.input_option_checks()
        self.random_seed = random_seed
        # replicas sharded over ranks (torchrun) must not share a noise stream: the rank enters the generator seed
        # (the fused path keys its Philox counters by the global bead index instead, simulation/langevin.py)
        from .distributed import dist_info
        _rank, _world = dist_info()
        _seed = None if random_seed is None else random_seed + (7919 * _rank if _world > 1 else 0)
        self.rng = None if random_seed is None else torch.Generator(device=self.device).manual_seed(_seed)
        self._simulated = False
        self.checkpointed_data = None
        self.engine = None
        self._warmup_end_time = self._simulation_end_time = None
        self._saver = None
        self._post_warmup_steps = 0

    # ------------------------------------------------------------------ option checks
    def input_option_checks(self):
        if self.save_interval <= 0 or self.n_timesteps % self.save_interval != 0:
            raise ValueError("The save_interval must be a positive factor of the simulation length")
        if self.export_interval % self.save_interval != 0:
            raise ValueError("Numpy export_interval must be an integer multiple of save_interval")
        if self.export_interval is not None and self.filename is None:
            self.export_interval = None   # nothing to export to; results stay in memory
        if self.log_interval is not None and self.log_interval % self.save_interval != 0:
            raise ValueError("The log_interval must be a multiple of save_interval")
        if (self.sim_subroutine is None) != (self.sim_subroutine_interval is None):
            raise ValueError("sim_subroutine and sim_subroutine_interval must be given together")

    # ------------------------------------------------------------------ attach
    def attach_model_and_configurations(self, model: torch.nn.Module, configurations: List[AtomicData],
                                        beta: Union[float, List[float]], overdamped: bool = False):
        if self.specialize_priors:
            model, configurations = condense_all_priors_for_simulation(model, configurations)
        if self.filename is not None:
            torch.save((deepcopy(model), deepcopy(configurations)), f"{self.filename}_specialized_model_and_config.pt")
        self._attach_model(model)
        self._attach_configurations(configurations, beta)

    def attach_model(self, model: torch.nn.Module):
        warnings.warn("using 'attach_model' is deprecated, use 'attach_model_and_configurations' instead.", DeprecationWarning)
        self._attach_model(model)

    def attach_configurations(self, configurations: List[AtomicData], beta: Union[float, List[float]]):
        warnings.warn("using 'attach_configurations' is deprecated, use 'attach_model_and_configurations' instead.",
                      DeprecationWarning)
        self._attach_configurations(configurations, beta)

    def _attach_model(self, model: torch.nn.Module):
        self.model = deepcopy(model).eval().to(device=self.device, dtype=self.dtype)
        for p in self.model.parameters():
            p.requires_grad_(False)

    @staticmethod
    def collate(data_list: List[AtomicData]) -> AtomicData:
        return collate(data_list)

    def validate_data_list(self, data_list: List[AtomicData]):
        """All configurations must describe the same molecule (reference base.py:914-983)."""
        d0 = data_list[0]
        for d in data_list[1:]:
            if d.pos.shape != d0.pos.shape:
                raise ValueError("Configurations do not all have the same number of atoms")
            if not torch.equal(d.atom_types, d0.atom_types):
                raise ValueError("Configurations do not all have the same atom types")
            if MASS_KEY in d0 and not torch.equal(d.masses, d0.masses):
                raise ValueError("Configurations do not all have the same masses")
            if d.neighbor_list.keys() != d0.neighbor_list.keys():
                raise ValueError("Configurations do not all have the same neighbor lists")
            for k in d0.neighbor_list:
                if not torch.equal(d.neighbor_list[k]["index_mapping"], d0.neighbor_list[k]["index_mapping"]):
                    raise ValueError(f"neighbor list '{k}' differs between configurations")

    def _attach_configurations(self, configurations: List[AtomicData], beta: Union[float, List[float]]):
        self.validate_data_list(configurations)
        self._read_checkpoint()
        self.initial_data = self.collate([deepcopy(c) for c in configurations]).to(device=self.device)
        self.n_sims = len(configurations)
        self.n_atoms = len(configurations[0].atom_types)
        self.n_dims = configurations[0].pos.shape[1]
        self.initial_data[POSITIONS_KEY] = self.initial_data[POSITIONS_KEY].to(self.dtype)
        if MASS_KEY in self.initial_data:
            self.initial_data[MASS_KEY] = self.initial_data[MASS_KEY].to(self.dtype)
        if self.checkpointed_data is not None:
            self.initial_data[POSITIONS_KEY] = self.checkpointed_data[POSITIONS_KEY].to(self.device, self.dtype)
            if VELOCITY_KEY in self.checkpointed_data:
                self.initial_data[VELOCITY_KEY] = self.checkpointed_data[VELOCITY_KEY].to(self.device, self.dtype)
        if isinstance(beta, (int, float)):
            if beta <= 0:
                raise ValueError("beta must be positive")
            self.beta = torch.full((self.n_sims,), float(beta), dtype=self.dtype, device=self.device)
        else:
            if len(beta) != self.n_sims:
                raise ValueError("a list of betas must have one entry per configuration")
            self.beta = torch.tensor(beta, dtype=self.dtype, device=self.device)
        self.initial_pos_spread = torch.stack([c.pos.std() for c in configurations]).max().detach().cpu()

    # ------------------------------------------------------------------ checkpoints
    def _read_checkpoint(self):
        self.current_timestep = 0
        path = self.read_checkpoint_file
        if path is None:
            return
        if path is True:
            found = sorted(glob.glob(f"{self.filename}_checkpoint_[0-9]*.pt"))
            if not found:
                return
            path = found[-1]
        ck = torch.load(path, weights_only=False)
        self.checkpointed_data = ck
        self.current_timestep = int(ck.get("current_timestep", 0))

    # ------------------------------------------------------------------ forces
    def calculate_potential_and_forces(self, data: AtomicData):
        """(potential [n_sims], forces [N,3]) from the attached model (reference base.py:821-909)."""
        data.out = {}
        data = self.model(data)
        return data.out[ENERGY_KEY].detach(), data.out[FORCE_KEY].detach()

    def timestep(self, data, forces):
        raise NotImplementedError

    # ------------------------------------------------------------------ buffers / output
    def _set_up_simulation(self, overwrite: bool = False):
        if self._simulated and not overwrite:
            raise RuntimeError("Simulation results are already populated. Set overwrite=True to overwrite.")
        if self.filename is not None and not overwrite:
            existing = glob.glob(f"{self.filename}_coords_[0-9]*.npy")
            if existing and self.read_checkpoint_file is None:
                raise RuntimeError(f"{existing[0]} exists; set overwrite=True or choose another filename")
        interval = self.export_interval if self.export_interval is not None else self.n_timesteps
        self._save_size = interval // self.save_interval
        self._npy_file_index = self.current_timestep if self.export_interval is not None else 0
        self._alloc_buffers()
        self.checkpoint = None

    def _alloc_buffers(self):
        shp = (self._save_size, self.n_sims, self.n_atoms, self.n_dims)
        self.simulated_coords = torch.zeros(shp)
        self.simulated_forces = torch.zeros(shp) if self.save_forces else None
        self.simulated_potential = torch.zeros(self._save_size, self.n_sims) if self.save_energies else None

    def _get_numpy_count(self) -> str:
        return f"{self._npy_file_index:04d}"

    @staticmethod
    def _swap_and_export(t: torch.Tensor) -> np.ndarray:
        return t.detach().cpu().numpy().swapaxes(0, 1)

    def _extra_save_tensors(self) -> dict:
        """Subclasses: further device tensors of a save point on the fused path (name -> tensor)."""
        return {}

    def _store_extra(self, i: int, host: dict, bufs: dict):
        pass

    def _traj_buffers(self) -> dict:
        return {"coords": self.simulated_coords, "forces": self.simulated_forces, "potential": self.simulated_potential}

    def save(self, pos, vel, forces, potential, t: int):
        i = t // self.save_interval - self._npy_file_index * self._save_size
        x = pos.view(-1, self.n_atoms, self.n_dims)
        spread = x.std(dim=(1, 2))
        if self.engine is not None:
            # fused path: nothing here waits for the GPU
            eng = self.engine
            # edge count: the larger of the live count and the high-water mark since the run began (an overflow between two
            # save points is not missed)
            n_e = spread.new_zeros(())
            if eng.ff.w is not None:
                n_e = torch.maximum(eng.ff.n_edges_dev[0], eng.ff.max_edges_dev[0]).float()
            status = torch.stack([spread.max(), torch.isnan(spread).any().float(), n_e])
            tensors = {"pos": x, "status": status}
            if self.save_forces:
                tensors["forces"] = forces.view(-1, self.n_atoms, self.n_dims)
            if self.save_energies:
                tensors["potential"] = potential
            if self.create_checkpoints:
                tensors["vel"] = vel
            tensors.update(self._extra_save_tensors())
            bufs = self._traj_buffers()
            limit = 1e3 * float(self.initial_pos_spread)
            cap = eng.ff.cap if eng.ff.w is not None else None

            def done(host, i=i, t=t, bufs=bufs):
                smax, snan, n_edges = (float(v) for v in host["status"])
                if snan > 0 or not (smax <= limit):
                    raise RuntimeError(f"Simulation of trajectory blew up at #timestep={t}")
                if cap is not None and n_edges > cap:
                    raise RuntimeError(f"neighbour list overflow at #timestep={t}: {int(n_edges)} edges > capacity {cap}")
                bufs["coords"][i] = host["pos"]
                if self.save_forces:
                    bufs["forces"][i] = host["forces"]
                if self.save_energies:
                    bufs["potential"][i] = host["potential"]
                if self.create_checkpoints:
                    self.checkpoint = {POSITIONS_KEY: host["pos"].reshape(-1, self.n_dims).clone(),
                                       VELOCITY_KEY: host["vel"].clone()}
                self._store_extra(i, host, bufs)
            if self._saver is None:
                self._saver = _AsyncSaver(self.device)
            self._saver.submit(tensors, done)
            return i
        if bool(((spread.max() > 1e3 * self.initial_pos_spread.to(spread.device)) | torch.isnan(spread).any()).item()):
            raise RuntimeError(f"Simulation of trajectory blew up at #timestep={t}")
        self.simulated_coords[i] = x.cpu()
        if self.save_forces:
            self.simulated_forces[i] = forces.view(-1, self.n_atoms, self.n_dims).cpu()
        if self.save_energies:
            self.simulated_potential[i] = potential.cpu()
        if self.create_checkpoints:
            self.checkpoint = {POSITIONS_KEY: pos.detach().clone().cpu(), VELOCITY_KEY: vel.detach().clone().cpu()}
        return i

    def write(self):
        if self._saver is not None:
            self._saver.drain()            # every save point of this export interval has reached the host buffers
        key = self._get_numpy_count()
        np.save(f"{self.filename}_coords_{key}.npy", self._swap_and_export(self.simulated_coords))
        if self.save_forces:
            np.save(f"{self.filename}_forces_{key}.npy", self._swap_and_export(self.simulated_forces))
        if self.save_energies:
            np.save(f"{self.filename}_potential_{key}.npy", self._swap_and_export(self.simulated_potential))
        if self.create_checkpoints and self.checkpoint is not None:
            self.checkpoint.update(current_timestep=self._npy_file_index + 1, export_interval=self.export_interval,
                                   save_interval=self.save_interval, log_interval=self.log_interval)
            torch.save(self.checkpoint, f"{self.filename}_checkpoint_{key}.pt")
        self._last_exported = {"coords": self.simulated_coords, "forces": self.simulated_forces,
                               "potential": self.simulated_potential}
        self._alloc_buffers()
        self._npy_file_index += 1

    def reshape_output(self):
        self.simulated_coords = self._swap_and_export(self.simulated_coords)
        if self.save_forces:
            self.simulated_forces = self._swap_and_export(self.simulated_forces)
        if self.save_energies:
            self.simulated_potential = self._swap_and_export(self.simulated_potential)

    # ------------------------------------------------------------------ fused engine hooks (subclasses)
    def _build_engine(self, data):
        return None

    # ------------------------------------------------------------------ main loop
    def simulate(self, overwrite: bool = False, prof=None) -> np.ndarray:
        self._set_up_simulation(overwrite)
        self._file_index0 = self._npy_file_index
        data = deepcopy(self.initial_data).to(self.device)
        t_init = self.current_timestep * self.export_interval if self.export_interval is not None else 0
        if t_init >= self.n_timesteps:
            raise ValueError(f"Simulation has already been running for {t_init} steps, which is larger than the "
                             f"target number of steps {self.n_timesteps}")
        eng = self.engine = self._build_engine(data)
        if eng is None:
            _, forces = self.calculate_potential_and_forces(data)
        if self.create_checkpoints and t_init == 0 and self.filename is not None:
            torch.save({POSITIONS_KEY: data[POSITIONS_KEY].detach().clone().cpu(),
                        VELOCITY_KEY: data[VELOCITY_KEY].detach().clone().cpu() if VELOCITY_KEY in data else None,
                        "current_timestep": 0, "export_interval": self.export_interval,
                        "save_interval": self.save_interval, "log_interval": self.log_interval},
                       f"{self.filename}_checkpoint_init.pt")
        cuda = self.device.type == "cuda"
        halfway = self.n_timesteps // 2
        t = t_init - 1
        for t in range(t_init, self.n_timesteps):
            if self.profile_start_step is not None and t == self.profile_start_step and cuda:
                torch.cuda.cudart().cudaProfilerStart()
            if t == halfway and self._warmup_end_time is None:
                if cuda:
                    torch.cuda.synchronize()
                    torch.cuda.reset_peak_memory_stats()
                self._warmup_end_time = time.perf_counter()
            if eng is not None:
                self._engine_timestep(eng)
                pos, vel, forces, potential = eng.pos, eng.vel, eng.ff.forces, eng.ff.energy
            else:
                data, potential, forces = self.timestep(data, forces)
                pos, vel = data[POSITIONS_KEY], data[VELOCITY_KEY] if VELOCITY_KEY in data else None
            if (t + 1) % self.save_interval == 0:
                self.save(pos, vel, forces, potential, t)   # fused path: asynchronous (incl. the capacity check)
                if self.export_interval is not None and (t + 1) % self.export_interval == 0:
                    self.write()
                    if self.save_subroutine is not None:
                        self.save_subroutine(data, (t + 1) // self.save_interval)
                if self.log_interval is not None and (t + 1) % self.log_interval == 0:
                    self.log((t + 1) // self.save_interval)
            if self.sim_subroutine is not None and (t + 1) % self.sim_subroutine_interval == 0:
                if eng is not None:
                    self._engine_subroutine(eng)
                else:
                    data.out = {ENERGY_KEY: potential, FORCE_KEY: forces}
                    data = self.sim_subroutine(data)
            self._final_potential = potential
            if eng is None:
                data.out = {}
            if prof:
                prof.step()
            if self.profile_end_step is not None and t == self.profile_end_step and cuda:
                torch.cuda.cudart().cudaProfilerStop()
        if self._saver is not None:
            self._saver.drain()
        if cuda:
            torch.cuda.synchronize()
        self._simulation_end_time = time.perf_counter()
        self._post_warmup_steps = self.n_timesteps - halfway
        if cuda:
            self._second_half_peak_memory_allocated = torch.cuda.max_memory_allocated() / 1024 ** 3
            self._second_half_peak_memory_reserved = torch.cuda.max_memory_reserved() / 1024 ** 3
        else:
            self._second_half_peak_memory_allocated = self._second_half_peak_memory_reserved = 0
        if self.export_interval is not None and (t + 1) % self.export_interval > 0:
            self.write()
        if eng is not None:
            data[POSITIONS_KEY], data[VELOCITY_KEY] = eng.pos.clone(), eng.vel.clone()
        self.final_data = data
        self._simulated = True
        self.summary()
        self.reshape_output()
        return self.simulated_coords

    def _engine_timestep(self, eng):
        eng.step()

    def _engine_subroutine(self, eng):
        raise NotImplementedError

    # ------------------------------------------------------------------ reporting
    def log(self, iter_: int):
        msg = f"{iter_}/{self.n_timesteps // self.save_interval} time points saved"
        if self.log_type == "print":
            print(msg)
        elif self.filename is not None:
            with open(f"{self.filename}_log.txt", "a") as fh:
                fh.write(msg + "\n")

    def summary(self):
        pass

    def get_throughput_metrics(self) -> dict:
        """timestep*mol/s over the second half of the run: the reference's keys (base.py:748-787: second_half_elapsed_time,
        second_half_steps, throughput, ms_per_timestep, first_half_steps, n_sims, n_atoms, peak_memory_*_gb; None before a
        run) plus `path` (fused-engine | module) and the earlier names of this package as aliases."""
        if self._warmup_end_time is None or self._simulation_end_time is None:
            return None
        dt = self._simulation_end_time - self._warmup_end_time
        steps = self._post_warmup_steps
        thr = steps * self.n_sims / dt if (dt > 0 and steps > 0) else 0
        ms = 1e3 * dt / steps if (dt > 0 and steps > 0) else 0
        return {"second_half_elapsed_time": dt, "second_half_steps": steps, "throughput": thr, "ms_per_timestep": ms,
                "first_half_steps": self.n_timesteps // 2, "n_sims": self.n_sims, "n_atoms": self.n_atoms,
                "peak_memory_allocated_gb": self._second_half_peak_memory_allocated,
                "peak_memory_reserved_gb": self._second_half_peak_memory_reserved,
                "path": "fused-engine" if self.engine is not None else "module",
                "post_warmup_steps": steps, "post_warmup_time_s": dt, "throughput_timestep_mol_per_s": thr, "ms_per_step": ms}
