"""Command-line / YAML front end with the reference's surface (simulation/cli.py:22-245): `--config FILE`
(keys `simulation:`, `betas`, `model_file`, `structure_file`), dotted overrides `--simulation.<kwarg> VALUE`,
`--betas`, `--model_file`, `--structure_file`, `--batch_size`, `--profile`.  argparse + PyYAML instead of
jsonargparse / ruamel (absent from the image)."""
import argparse
import inspect
import os
from copy import deepcopy
from typing import Any, Dict, List, Tuple

import torch
import yaml

from .base import _Simulation
from .utils import dump_yaml, load_yaml


def _simulation_kwargs(simulation_class) -> Dict[str, Any]:
    out = {}
    for cls in reversed(simulation_class.__mro__):
        if cls is object or "__init__" not in cls.__dict__:
            continue
        for name, p in inspect.signature(cls.__init__).parameters.items():
            if name in ("self", "kwargs") or p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
                continue
            out[name] = p.default
    for k in ("sim_subroutine", "save_subroutine", "sim_subroutine_interval"):
        out.pop(k, None)
    return out


def _value(text: str):
    try:
        return yaml.safe_load(text)
    except yaml.YAMLError:
        return text


def build_parser(simulation_class, description: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--config", "-c", help="YAML configuration file")
    ap.add_argument("--betas", "-tm", type=_value, help="inverse temperature(s) 1/kBT, e.g. '[1.67]'")
    ap.add_argument("--model_file", "-mf", help="pickled torch.nn.Module (SumOut of GradientsOut models)")
    ap.add_argument("--structure_file", "-sf", help="pickled List[AtomicData] of initial configurations")
    ap.add_argument("--batch_size", type=int, default=None, help="use / duplicate configurations up to this count")
    ap.add_argument("--profile", default=None, help="directory for a torch.profiler trace")
    for name in _simulation_kwargs(simulation_class):
        ap.add_argument(f"--simulation.{name}", dest=f"simulation.{name}", type=_value, default=argparse.SUPPRESS)
    return ap


def load_model_file(path: str) -> torch.nn.Module:
    """Reference models/pyg_forward_compatibility.py:227-243 unpickles mlcg/PyG checkpoints; here the file must
    be a pickle of the drop-in module classes (same import paths)."""
    return torch.load(path, weights_only=False, map_location="cpu")


def expand_batch(initial_data_list: List, batch_size) -> List:
    """--batch_size semantics of the reference (cli.py:130-158): truncate or cyclically duplicate."""
    if batch_size is None:
        return initial_data_list
    if batch_size <= 0:
        raise ValueError(f"batch_size must be positive, got {batch_size}")
    n = len(initial_data_list)
    if batch_size <= n:
        return initial_data_list[:batch_size]
    return [deepcopy(initial_data_list[i % n]) for i in range(batch_size)]


def parse_simulation_config(simulation_class, description: str = "Simulation command line tool",
                            parser_kwargs: Dict[str, Any] = None, subclass_mode: bool = False, argv=None
                            ) -> Tuple[torch.nn.Module, List, Any, _Simulation, Any]:
    """-> (model, initial_data_list, betas, simulation, profile)"""
    args = vars(build_parser(simulation_class, description).parse_args(argv))
    config: Dict[str, Any] = {"simulation": {}}
    if args.get("config"):
        loaded = load_yaml(args["config"])
        config.update({k: v for k, v in loaded.items() if k != "simulation"})
        config["simulation"].update(loaded.get("simulation") or {})
    for k, v in args.items():
        if k.startswith("simulation."):
            config["simulation"][k.split(".", 1)[1]] = v
        elif k != "config" and v is not None:
            config[k] = v
    for req in ("model_file", "structure_file", "betas"):
        if config.get(req) is None:
            raise SystemExit(f"missing required option: {req}")
    sim_kwargs = dict(config["simulation"])
    allowed = _simulation_kwargs(simulation_class)
    unknown = [k for k in sim_kwargs if k not in allowed]
    if unknown:
        raise SystemExit(f"unknown simulation option(s): {unknown}")
    simulation = simulation_class(**sim_kwargs)
    if simulation.filename is not None:
        dump_yaml(f"{simulation.filename}_config.yaml", {k: v for k, v in config.items() if k != "profile"})
    model = load_model_file(config["model_file"])
    data_list = expand_batch(torch.load(config["structure_file"], weights_only=False), config.get("batch_size"))
    betas = config["betas"]
    if not isinstance(betas, (list, tuple)):
        betas = [betas]
    betas = [float(b) for b in betas]
    if len(betas) == 1:
        betas = betas[0]
    return model, data_list, betas, simulation, config.get("profile")
