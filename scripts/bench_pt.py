"""BASELINE config 4: parallel-tempering Langevin, 1ENH-shaped synthetic system, betas [1.67, 1.42, 1.16] x n_indep
replicas each, replica exchange every `exchange_interval` steps; replicas sharded over the ranks (torchrun), energies
all-gathered and accepted pairs swapped over NCCL (flashmd/simulation/distributed.py).  Through the drop-in API
(PTSimulation -> fused engine).  Prints one JSON line on rank 0: whole-job timestep*mol/s from the reference's own
second-half throughput metric (max time over ranks).
Usage: [torchrun --nproc-per-node N] python scripts/bench_pt.py [--n-indep 256] [--n-beads 269] [--steps 400]"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--n-indep", type=int, default=256)
ap.add_argument("--n-beads", type=int, default=269)
ap.add_argument("--steps", type=int, default=400)
ap.add_argument("--exchange-interval", type=int, default=100)
ap.add_argument("--gptq", default="w16a16")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

from flashmd import synthetic as syn  # noqa: E402
from flashmd.data import AtomicData  # noqa: E402
from flashmd.models import CosineCutoff, GaussianBasis, GradientsOut, StandardSchNet, SumOut  # noqa: E402
from flashmd.neighbor_list import make_neighbor_list  # noqa: E402
from flashmd.prior import Dihedral, HarmonicAngles, HarmonicBonds, Repulsion  # noqa: E402
from flashmd.simulation import PTSimulation  # noqa: E402

BETAS = [1.67, 1.42, 1.16]
n, n_indep = args.n_beads, args.n_indep
n_distinct = min(n_indep, 16)
system = syn.synthetic_system(n_distinct, n, seed=0)
ty, st_ = system["atom_types"], system["stats"]
sb, sa, sd, sr = {}, {}, {}, {}
for i, j in system["bonds"].T:
    sb[(int(ty[i]), int(ty[j]))] = {"k": float(st_["bonds"]["k"][ty[i], ty[j]]), "x_0": float(st_["bonds"]["x_0"][ty[i], ty[j]])}
for i, j, k in system["angles"].T:
    key = (int(ty[i]), int(ty[j]), int(ty[k]))
    sa[key] = {"k": float(st_["angles"]["k"][key]), "x_0": float(st_["angles"]["x_0"][key])}
nd = st_["dihedrals"]["n_degs"]
for i, j, k, l in system["dihedrals"].T:
    c = (int(ty[j]), int(ty[k]))
    sd[(int(ty[i]), int(ty[j]), int(ty[k]), int(ty[l]))] = {
        "k1s": {f"k1_{q + 1}": float(st_["dihedrals"]["k1_central"][(q,) + c]) for q in range(nd)},
        "k2s": {f"k2_{q + 1}": float(st_["dihedrals"]["k2_central"][(q,) + c]) for q in range(nd)},
        "v_0": float(st_["dihedrals"]["v0_central"][c])}
for i, j in system["nonbonded"].T:
    sr[(int(ty[i]), int(ty[j]))] = {"sigma": float(st_["repulsion"]["sigma"][ty[i], ty[j]])}
rc = system["cutoff"]
torch.manual_seed(0)
schnet = StandardSchNet(GaussianBasis(CosineCutoff(0.0, rc), num_rbf=50), CosineCutoff(0.0, rc),
                        output_hidden_layer_widths=[128, 64], hidden_channels=128, embedding_size=syn.N_BEAD_TYPES + 1,
                        num_filters=128, num_interactions=3)
model = SumOut(torch.nn.ModuleDict({
    "SchNet": GradientsOut(schnet), "bonds": GradientsOut(HarmonicBonds(sb)), "angles": GradientsOut(HarmonicAngles(sa)),
    "dihedrals": GradientsOut(Dihedral(sd, n_degs=nd)), "repulsion": GradientsOut(Repulsion(sr))}))
nls = {"bonds": make_neighbor_list("bonds", 2, torch.from_numpy(system["bonds"])),
       "angles": make_neighbor_list("angles", 3, torch.from_numpy(system["angles"])),
       "dihedrals": make_neighbor_list("dihedrals", 4, torch.from_numpy(system["dihedrals"])),
       "repulsion": make_neighbor_list("repulsion", 2, torch.from_numpy(system["nonbonded"]))}
configs = [AtomicData.from_points(pos=torch.from_numpy(system["pos"][b % n_distinct].copy()), atom_types=torch.from_numpy(ty),
                                  masses=torch.from_numpy(system["masses"]), neighborlist=nls) for b in range(n_indep)]
tmp = tempfile.mkdtemp()
sim = PTSimulation(friction=1.0, dt=0.004, n_timesteps=args.steps, save_interval=args.steps, export_interval=args.steps,
                   exchange_interval=args.exchange_interval, random_seed=103838, device=str(dev), dtype="single",
                   filename="pt", output_dir=tmp, gptq=None if args.gptq == "none" else args.gptq)
sim.attach_model_and_configurations(model, configs, betas=BETAS)
sim.simulate()
m = sim.get_throughput_metrics()
t = torch.tensor([float(m["second_half_elapsed_time"])], device=dev)
approved = torch.tensor([float(sim.exchange_summary["approved"])], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
n_total = len(BETAS) * n_indep
val = n_total * float(m["second_half_steps"]) / float(t)
if rank == 0:
    print(json.dumps({"metric": "timestep*mol/s, parallel tempering Langevin (config 4)", "value": val, "unit": "timestep*mol/s",
                      "n_gpus": world, "n_sims_total": n_total, "n_sims_per_gpu": n_total // world, "betas": BETAS,
                      "n_beads": n, "steps": args.steps, "exchange_interval": args.exchange_interval,
                      "second_half_steps": float(m["second_half_steps"]), "second_half_s_max_over_ranks": float(t),
                      "ms_per_step": 1e3 * float(t) / float(m["second_half_steps"]), "path": m.get("path"),
                      "exchanges_attempted": sim.exchange_summary["attempted"], "exchanges_approved_rank0": float(approved),
                      "precision": args.gptq}))
if world > 1:
    dist.destroy_process_group()
