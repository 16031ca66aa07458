#!/usr/bin/env python
"""Benchmark of the CGSchNet force-field + Langevin step (BASELINE.json metric: timestep·mol/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--precision w16a16|fp32] [--n-beads 269] [--batch 128]

One "step" = one BAOAB Langevin step (neighbour list -> SchNet energy -> analytic forces ->
priors -> integrator) of `batch` molecules per GPU.  N>1 (under torchrun): every rank runs its own
independent replica batch (no data-path collective; weak scaling), time = max over ranks.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "timestep*mol/s, CGSchNet 1ENH batch128 Langevin"
UNIT = "timestep*mol/s"
DT, FRICTION, BETA, SEED = 0.004, 1.0, 1.67, 103838     # reference examples/langevin.yaml:2-9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="w16a16", choices=["w16a16", "fp32"])
    ap.add_argument("--n-beads", type=int, default=269)
    ap.add_argument("--batch", type=int, default=128, help="molecules per GPU")
    ap.add_argument("--blocks", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-triton-baseline", action="store_true",
                    help="skip the reference's Triton GPU path (timed on the same GPU after our own run, N = 1 only)")
    ap.add_argument("--min-seconds", type=float, default=0.5,
                    help="the timed region replays the K-step block until it is at least this long (per-step numbers divide back)")
    ap.add_argument("--pt", action="store_true",
                    help="BASELINE config 4: parallel tempering, betas [1.67, 1.42, 1.16] x 256 replicas (768 simulations "
                         "sharded over the ranks), replica exchange every 100 steps (NCCL energy all-gather + peer swaps)")
    ap.add_argument("--profile-kernels", action="store_true", help="print per-kernel event timings to stderr")
    return ap.parse_args()


def peaks():
    """(hbm GB/s, bf16 TF/s sustained, bf16 TF/s burst, source) - measured on this pool's B200s by the driver."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return (float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), float(d["bf16_tflops"]),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_system(args, seed):
    from flashmd import synthetic
    n_distinct = min(args.batch, 16)
    sysd = synthetic.synthetic_system(n_distinct, args.n_beads, seed=seed)
    reps = -(-args.batch // n_distinct)
    pos = np.concatenate([sysd["pos"]] * reps, 0)[: args.batch]
    return sysd, pos


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------


def cpu_oracle_throughput(args, n_mols, n_steps, threads):
    """BAOAB steps of `n_mols` molecules with the CPU oracle (fp32 PyTorch path = the reference's
    --disable_optim semantics).  Returns (timestep*mol/s, seconds)."""
    from oracle import fmd_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(threads)
    from flashmd import synthetic
    from flashmd.engine import random_schnet_tensors
    sysd = synthetic.synthetic_system(n_mols, args.n_beads, seed=0)
    n = args.n_beads
    pos = torch.from_numpy(sysd["pos"]).reshape(n_mols * n, 3)
    types = torch.from_numpy(sysd["atom_types"]).repeat(n_mols)
    batch = torch.arange(n_mols).repeat_interleave(n)
    ptr = np.arange(n_mols + 1) * n
    t = random_schnet_tensors(0, num_blocks=args.blocks)
    t.setdefault("out2_b", None)
    P = O.SchNetParams(t, args.blocks, 3, sysd["cutoff"], 50)
    masses = torch.from_numpy(sysd["masses"]).repeat(n_mols)
    bmr = torch.sqrt(1.0 / (BETA * masses))[:, None]
    vs, ns = O.baoab_constants(DT, FRICTION)
    pri = _oracle_priors(sysd, n_mols)

    def force(x):
        ei = torch.from_numpy(O.radius_graph(x.numpy(), ptr, sysd["cutoff"]))
        e, f = O.schnet_energy_forces(P, x, types, batch, n_mols, ei)
        for kind, (m, mb, pr) in pri.items():
            ek, fk = O.prior_energy_forces(kind, x, m, mb, n_mols, pr)
            e, f = e + ek, f + fk
        return e, f

    x, v = pos, torch.zeros_like(pos)
    _, f = force(x)
    g = torch.Generator().manual_seed(SEED)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        noise = torch.empty_like(x).normal_(generator=g)
        x, v = O.baoab_pre(x, v, f, masses, bmr, noise, DT, vs, ns)
        x, v = x.float(), v.float()
        e, f = force(x)
        v = O.baoab_post(v, f, masses, DT).float()
    dtm = time.perf_counter() - t0
    return n_mols * n_steps / dtm, dtm


def _oracle_priors(sysd, B):
    ty, st, n = sysd["atom_types"], sysd["stats"], sysd["atom_types"].shape[0]
    out = {}

    def col(m):
        return (torch.from_numpy(np.concatenate([m + b * n for b in range(B)], 1)),
                torch.from_numpy(np.repeat(np.arange(B), m.shape[1])))

    def rep(v):
        return torch.from_numpy(np.concatenate([np.asarray(v, np.float32)] * B, 0))
    m = sysd["bonds"]; tt = (ty[m[0]], ty[m[1]])
    out["bonds"] = (*col(m), {"k": rep(st["bonds"]["k"][tt]), "x0": rep(st["bonds"]["x_0"][tt])})
    m = sysd["angles"]; tt = (ty[m[0]], ty[m[1]], ty[m[2]])
    out["angles"] = (*col(m), {"k": rep(st["angles"]["k"][tt]), "x0": rep(st["angles"]["x_0"][tt])})
    m = sysd["dihedrals"]; c = (ty[m[1]], ty[m[2]]); nd = st["dihedrals"]["n_degs"]
    out["dihedrals"] = (*col(m), {
        "k1s": rep(np.stack([st["dihedrals"]["k1_central"][d][c] for d in range(nd)], 1)),
        "k2s": rep(np.stack([st["dihedrals"]["k2_central"][d][c] for d in range(nd)], 1)),
        "v_0": rep(st["dihedrals"]["v0_central"][c])})
    m = sysd["nonbonded"]; tt = (ty[m[0]], ty[m[1]])
    out["repulsion"] = (*col(m), {"sigma": rep(st["repulsion"]["sigma"][tt])})
    return out


def cpu_reference_throughput(args, n_mols, n_steps, threads):
    """The UNMODIFIED reference (pip-installed into baseline/_ref, imported behind oracle/shims) on the host cores with
    its --disable_optim semantics (MLCG_*=0, gptq=None, no compile: scripts/nvt_langevin.py:6-17,40-60), in a
    subprocess because its package name clashes with the drop-in.  Timed with the reference's own second-half
    throughput metric (simulation/base.py:748-787) over 2*n_steps timesteps.  None when baseline/_ref is absent."""
    import subprocess
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "flashmd")) or args.blocks != 3:
        return None
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_triton_reference.py"), "--device", "cpu", "--batch",
           str(n_mols), "--n-beads", str(args.n_beads), "--steps", str(2 * n_steps), "--threads", str(threads)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        rec = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        m = rec["metrics"]
        return float(m["throughput"]), float(m["second_half_elapsed_time"]), int(m["second_half_steps"]), int(rec["threads"])
    except Exception as e:  # noqa: BLE001  (fall back to the port, say why)
        print(f"[bench] reference CPU run failed ({e!r}); falling back to the oracle port", file=sys.stderr)
        return None


def cpu_baseline_entry(args, n_mols, n_steps):
    """cpu_baseline object: the reference itself when it is installed (kind "reference"), else the oracle port."""
    cores = os.cpu_count() or 1
    ref = cpu_reference_throughput(args, n_mols, n_steps, cores)
    if ref is not None:
        val, secs, steps, thr = ref
        return {"value": val, "unit": UNIT, "cores": thr, "kind": "reference",
                "sample": f"{n_mols} molecules x {args.n_beads} beads, second half ({steps} BAOAB steps, {secs:.1f} s) of a "
                          f"{2 * steps}-step run of the UNMODIFIED reference (baseline/_ref) with --disable_optim semantics "
                          f"(MLCG_*=0, gptq=None, no compile) on {thr} host threads; torch_cluster.radius_graph served by "
                          f"the oracle/shims stand-in"}, val, secs, steps
    cpu_oracle_throughput(args, n_mols, 1, cores)     # warm-up (thread pool, allocator)
    val, secs = cpu_oracle_throughput(args, n_mols, n_steps, cores)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_mols} molecules x {args.n_beads} beads, {n_steps} BAOAB steps ({secs:.1f} s), oracle port of the "
                      f"reference's --disable_optim fp32 PyTorch path (oracle/fmd_oracle.py), {cores} threads"}, val, secs, n_steps


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_mols = 16
    steps, warm = max(1, min(args.steps, 10)), max(0, min(args.warmup, 2))
    cb, val, secs, steps = cpu_baseline_entry(args, n_mols, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": f"CGSchNet 1ENH-shaped synthetic CG protein, {args.n_beads} beads/molecule, batch "
                        f"{args.batch} per GPU, {args.blocks} interaction blocks, F=128, R=50, Langevin beta={BETA} "
                        f"dt={DT}; priors: bonds+angles+dihedrals+repulsion",
            "n_beads": args.n_beads, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "precision_path": args.precision, "parallelism": f"replica-sharded x{world} (no data-path collective)",
            "l2": "inputs larger than L2: one step streams ~0.5 GB of distinct buffers (node activations of all blocks, "
                  "edge list, prior incidence lists) against a 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        # NCCL's own log (communicator size, NVLS / ring choice) is kept, but away from stdout (one JSON line there)
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fmd_nccl_%h_%p.log")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from flashmd import _lib as L
    from flashmd.engine import (ForceField, LangevinEngine, SchNetWeights, prior_terms_from_system,
                                random_schnet_tensors)
    L.load()
    if args.pt:
        return run_pt(args, rank, world, local, dev, dist)
    # the SAME molecules on every rank (weak scaling with identical work per GPU); velocities and noise differ per rank
    sysd, pos_np = build_system(args, seed=0)
    B, n = args.batch, args.n_beads
    pos = torch.from_numpy(pos_np).reshape(B * n, 3).to(dev).contiguous()
    types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(dev)
    mol_ptr = (torch.arange(B + 1) * n).to(dev)
    w = SchNetWeights.from_flat(random_schnet_tensors(0, num_blocks=args.blocks), sysd["cutoff"], 50, dev)
    priors = prior_terms_from_system(sysd, B, dev)
    # edge capacity: 1.35x the initial edge count (density stays bounded by the repulsion/bond priors)
    from flashmd.neighbor_list import radius_graph_csr
    e0 = radius_graph_csr(pos, mol_ptr, sysd["cutoff"], idx_dtype=torch.int32)["edge_index"].shape[1]
    cap = int(1.35 * e0) + 4096
    ff = ForceField(w, priors, types, mol_ptr, precision=args.precision, edge_capacity=cap)
    masses = torch.from_numpy(sysd["masses"]).repeat(B)
    g = torch.Generator().manual_seed(1234 + rank)
    v0 = torch.randn((B * n, 3), generator=g) * torch.sqrt(1.0 / (BETA * masses))[:, None]
    eng = LangevinEngine(ff, pos, v0, masses, torch.full((B,), BETA), DT, FRICTION, seed=SEED, use_graph=True,
                         node_offset=rank * B * n)     # Philox noise keyed by the global bead index

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(max(args.warmup, 3)):
        eng.step()
    edges_start = ff.num_edges()
    # state at the start of the timed region: the end-to-end leg below restarts from it, so both legs see the same phase of
    # the trajectory (the synthetic chains relax and lose edges while they run)
    pos_t0, vel_t0, frc_t0 = eng.pos.clone(), eng.vel.clone(), ff.forces.clone()
    # the timed region is `repeats` back-to-back blocks of exactly K steps, long enough (--min-seconds) that launch jitter
    # and the barrier do not show in the max-over-ranks time; per-step numbers divide back
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    est = (time.perf_counter() - t0) / 3
    repeats = max(1, int(np.ceil(args.min_seconds / max(est * args.steps, 1e-9))))
    if dist is not None:
        rt = torch.tensor([repeats], device=dev)
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        repeats = int(rt.item())
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(repeats * args.steps):
        eng.step()
    ev1.record()
    barrier()
    ms_rank = ev0.elapsed_time(ev1)
    ms = ms_rank
    clocks = sampler.stop() if rank == 0 else None
    edges_now = ff.num_edges()
    per_rank = {"ms_per_step": [ms_rank / (repeats * args.steps)], "edges_start": [edges_start], "edges_end": [edges_now]}
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        mine = torch.tensor([ms_rank / (repeats * args.steps), float(edges_start), float(edges_now)], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [float(a[0]) for a in allr], "edges_start": [int(a[1]) for a in allr],
                    "edges_end": [int(a[2]) for a in allr]}
    assert edges_now <= cap, f"edge capacity overflow: {edges_now} > {cap}"
    assert torch.isfinite(eng.pos).all(), "trajectory diverged"
    n_timed = repeats * args.steps
    value = world * B * n_timed / (ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (H2D + step + D2H inside the timed region)
    ph = torch.empty((B * n, 3), dtype=torch.float32).pin_memory()
    vh = torch.empty((B * n, 3), dtype=torch.float32).pin_memory()
    fh = torch.empty((B * n, 3), dtype=torch.float32).pin_memory()
    eh = torch.empty(B, dtype=torch.float32).pin_memory()
    ph.copy_(pos_t0); vh.copy_(vel_t0); fh.copy_(frc_t0)
    for _ in range(3):
        eng.step_host(ph, vh, fh, eh)
    e2e_edges_start = ff.num_edges()
    barrier()
    ev0.record()
    for _ in range(n_timed):          # the same number of steps as the device-timed region: same trajectory phase
        eng.step_host(ph, vh, fh, eh)
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_val = world * B * n_timed / (ms_e2e * 1e-3)
    e2e_edges_end = ff.num_edges()
    h2d = 3 * B * n * 3 * 4
    d2h = 3 * B * n * 3 * 4 + B * 4

    # ---- roofline of the dominant kernel (CFConv CSR segment reduce), CUDA events on the launch stream
    roof, kern_table = kernel_roofline(ff, eng, args, edges_now)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / n_timed, "higher_is_better": True,
            "timed_steps": n_timed, "repeats_of_k_steps": repeats,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 filter-network operands / tf32 node layers / f32 accumulate" if args.precision == "w16a16" else "f32",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / n_timed, "steps": n_timed, "edges_start": e2e_edges_start,
                    "edges_end": e2e_edges_end,
                    "note": "restarts from the state at the start of the device-timed region and runs the same number of steps"},
            "gpu_launches": eng.launches_per_step * n_timed,
            "launches_per_step": eng.launches_per_step,
            "edges": edges_now, "edges_start": edges_start, "nodes": B * n,
            "per_rank": per_rank,
            "roofline": roof, "kernels_ms_per_step": kern_table,
        }
        if dist is not None:
            line["nccl"] = {"nranks": world, "version": ".".join(str(v) for v in torch.cuda.nccl.version()),
                            "data_path_collectives_per_step": 0, "log": os.environ.get("NCCL_DEBUG_FILE"),
                            "log_excerpt": nccl_log_excerpt()}
        if world == 1 and not args.no_triton_baseline and args.precision == "w16a16" and args.blocks == 3:
            line["triton_baseline"] = triton_baseline_entry(args, value)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(args, 16, 8)[0]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def nccl_log_excerpt():
    """Communicator lines of this process's NCCL log (NCCL_DEBUG_FILE), so that the bench line itself shows the ranks."""
    import glob
    import socket
    pat = os.environ.get("NCCL_DEBUG_FILE", "")
    if not pat or pat.startswith("/dev/"):
        return None
    path = pat.replace("%h", socket.gethostname()).replace("%p", str(os.getpid()))
    files = [path] if os.path.exists(path) else sorted(glob.glob(pat.replace("%h", "*").replace("%p", "*")))[-1:]
    out = []
    for f in files:
        try:
            for ln in open(f, errors="replace"):
                if "nranks" in ln or "Init COMPLETE" in ln or "NVLS" in ln:
                    out.append(ln.strip()[-220:])
        except OSError:
            pass
    return out[:6] or None


def triton_baseline_entry(args, our_value):
    """The UNMODIFIED reference's default GPU path (Triton kernels, gptq="w16a16", all MLCG_*=1) timed on the SAME GPU and
    the same synthetic workload right after our own run, in a subprocess (its package name clashes with the drop-in):
    scripts/bench_triton_reference.py, the reference's own second-half throughput metric (simulation/base.py:748-787).
    compile_model=True (the reference's default, simulation/base.py:362-368) is tried first; with torch 2.11 + triton 3.6
    inductor rejects the reference's own CSR kernel, so the error is recorded and the eager run is the baseline."""
    script = os.path.join(ROOT, "scripts", "bench_triton_reference.py")
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "flashmd")):
        return {"unavailable": "baseline/_ref (pip install --target of the reference) is absent"}
    env = dict(os.environ, CXX=os.environ.get("FMD_REF_CXX", "/usr/bin/g++"))    # the image's default g++ wrapper lacks libgomp.spec
    out = {"nl": "radius_graph served by this repo's CUDA kernel (torch_cluster is not installable)", "steps": 100}

    def run(compile_flag, timeout):
        cmd = [sys.executable, script, "--device", "cuda", "--batch", str(args.batch), "--n-beads", str(args.n_beads),
               "--steps", "100", "--gptq", "w16a16", "--compile", str(compile_flag)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        recs = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not recs:
            errs = [ln for ln in (r.stderr or "").splitlines() if "Error" in ln or "error" in ln]
            return None, (errs[-1] if errs else (r.stderr or "")[-300:]).strip()[:400]
        return json.loads(recs[-1]), None
    try:
        rec, err = (None, "skipped (FMD_REF_COMPILE=0)") if os.environ.get("FMD_REF_COMPILE", "1") == "0" else run(1, 600)
        out["compile_model"] = rec is not None
        if rec is None:
            out["compile_model_error"] = err
            rec, err = run(0, 600)
        if rec is None:
            out["unavailable"] = err
            return out
        m = rec["metrics"]
        out.update({"value": float(m["throughput"]), "unit": UNIT, "ms_per_step": float(m["ms_per_timestep"]),
                    "second_half_steps": int(m["second_half_steps"]), "attach_s": rec.get("attach_s"),
                    "ours_over_triton": our_value / float(m["throughput"]), "notes": rec.get("notes")})
    except Exception as e:  # noqa: BLE001
        out["unavailable"] = repr(e)[:300]
    return out


def run_pt(args, rank, world, local, dev, dist):
    """BASELINE config 4 on the fused engine: 3 betas x 256 replicas = 768 simulations of the 269-bead molecule sharded
    contiguously over the ranks (reference layout: sim = beta_index * n_indep + replica), Langevin steps replayed as a
    CUDA graph, replica exchange every 100 steps (simulation/distributed.py: NCCL all-gather of the energies, device-side
    decisions, static peer exchange).  One "step" = one BAOAB step of all 768 simulations; strong scaling."""
    from flashmd import _lib as L
    from flashmd.engine import ForceField, LangevinEngine, SchNetWeights, prior_terms_from_system, random_schnet_tensors
    from flashmd.neighbor_list import radius_graph_csr
    from flashmd.simulation.distributed import ShardedExchange, shard_range
    from flashmd.simulation.parallel_tempering import adjacent_pairs
    L.load()
    betas, n_indep, interval = [1.67, 1.42, 1.16], 256, 100
    n_total = len(betas) * n_indep
    lo, hi = shard_range(n_total, rank, world)
    B, n = hi - lo, args.n_beads
    from flashmd import synthetic
    sysd = synthetic.synthetic_system(16, n, seed=0)
    pos_np = np.stack([sysd["pos"][(s % n_indep) % 16] for s in range(lo, hi)])
    beta_all = torch.tensor([b for b in betas for _ in range(n_indep)], dtype=torch.float32)
    pos = torch.from_numpy(pos_np).reshape(B * n, 3).to(dev).contiguous()
    types = torch.from_numpy(sysd["atom_types"]).repeat(B).to(dev)
    mol_ptr = (torch.arange(B + 1) * n).to(dev)
    w = SchNetWeights.from_flat(random_schnet_tensors(0, num_blocks=args.blocks), sysd["cutoff"], 50, dev)
    priors = prior_terms_from_system(sysd, B, dev)
    e0 = radius_graph_csr(pos, mol_ptr, sysd["cutoff"], idx_dtype=torch.int32)["edge_index"].shape[1]
    cap = int(1.35 * e0) + 4096
    ff = ForceField(w, priors, types, mol_ptr, precision=args.precision, edge_capacity=cap)
    masses = torch.from_numpy(sysd["masses"]).repeat(B)
    beta_loc = beta_all[lo:hi]
    g = torch.Generator().manual_seed(1234 + rank)
    v0 = torch.randn((B * n, 3), generator=g) * torch.sqrt(1.0 / (beta_loc.repeat_interleave(n) * masses))[:, None]
    eng = LangevinEngine(ff, pos, v0, masses, beta_loc, DT, FRICTION, seed=SEED, use_graph=True, node_offset=lo * n)
    ex = ShardedExchange(beta_all, n, rank, world)
    even, odd = adjacent_pairs(len(betas), n_indep)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    accs = []

    def block(k):
        for _ in range(interval):
            eng.step()
        pa, pb = even if k % 2 == 0 else odd
        accs.append(ex.exchange(eng.pos, eng.vel, ff.energy, pa, pb, None, SEED, k))   # counter-based decisions on the device

    for k in range(2):
        block(k)
    edges_start = ff.num_edges()
    n_blocks = max(1, -(-max(args.steps, interval) // interval))
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for k in range(2, 2 + n_blocks):
        block(k)
    ev1.record()
    barrier()
    ms_rank = ev0.elapsed_time(ev1)
    ms = ms_rank
    clocks = sampler.stop() if rank == 0 else None
    edges_now = ff.num_edges()
    steps = n_blocks * interval
    per_rank = {"ms_per_step": [ms_rank / steps], "edges_start": [edges_start], "edges_end": [edges_now]}
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        mine = torch.tensor([ms_rank / steps, float(edges_start), float(edges_now)], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [float(a[0]) for a in allr], "edges_start": [int(a[1]) for a in allr],
                    "edges_end": [int(a[2]) for a in allr]}
    assert edges_now <= cap and torch.isfinite(eng.pos).all()
    n_acc = int(sum(int(a.sum()) for a in accs[2:]))
    if rank == 0:
        cfg = workload_config(args, world)
        cfg.update({"workload": f"parallel tempering (BASELINE config 4): betas {betas} x {n_indep} replicas = {n_total} simulations "
                                f"of the {n}-bead molecule, exchange every {interval} steps, sharded over {world} GPU(s)",
                    "batch_per_gpu": B, "global_batch": n_total,
                    "parallelism": f"replicas sharded x{world}; per exchange: NCCL all-gather of {n_total} energies + static peer swap"})
        print(json.dumps({
            "metric": METRIC.replace("batch128 Langevin", "parallel tempering 3x256"), "value": n_total * steps / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 2 * interval, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 filter-network operands / tf32 node layers / f32 accumulate", "data": "synthetic", "config": cfg,
            "clocks": clocks, "gpu_launches": eng.launches_per_step * steps, "launches_per_step": eng.launches_per_step,
            "edges": edges_now, "edges_start": edges_start, "nodes": B * n, "per_rank": per_rank,
            "exchanges": {"n": n_blocks, "pairs_proposed": int(sum(a.numel() for a in accs[2:])), "pairs_accepted": n_acc,
                          "rng": "Philox keyed by (seed, exchange index, pair), identical on every rank"},
            "nccl": None if dist is None else {"nranks": world, "version": ".".join(str(v) for v in torch.cuda.nccl.version()),
                                               "log_excerpt": nccl_log_excerpt()}}))
    if dist is not None:
        dist.destroy_process_group()


def kernel_roofline(ff, eng, args, E):
    """Per-kernel-class device time of one eager step (CUDA events on the launching stream, 3 reps), and the
    roofline entry of the dominant kernel.  Algorithmic work per launch (DESIGN.md "Kernels"):
      fused filter-network x CFConv forward (tensor-bound): 2*E*F*(R+F) flop, R = 50 (unpadded);
      fused backward: 2*E*F*(F+R) flop for g_t = g_W Wf1, g_rbf = g_t Wf0 (the recomputation of t is NOT counted);
      materialised CFConv CSR (HBM-bound, fp32 path): E*F*b + 8*N*F + 8*E + 4*(N+1) bytes."""
    from flashmd import _lib as L
    hbm, tf_sus, tf_burst, which = peaks()
    N, F, R = ff.N, ff.w.filters, ff.w.num_rbf
    b = 2 if args.precision == "w16a16" else 4
    timings = {}
    L.load()
    orig_call = L.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        timings.setdefault(name, []).append((e0, e1))

    import flashmd.engine as E_mod
    reps = 3
    e_before = ff.num_edges()
    E_mod.L.call = timed_call
    ff.serial_priors = True      # every kernel alone on the stream: the forked prior kernel would overlap the first forward launch
    try:
        for _ in range(reps):
            eng._step_body()
        torch.cuda.synchronize()
    finally:
        E_mod.L.call = orig_call
        ff.serial_priors = False
    E = 0.5 * (e_before + ff.num_edges())      # live edge count of the steps that were timed
    table = {}
    for name, evs in timings.items():
        tot = sum(a.elapsed_time(bb) for a, bb in evs)
        table[name] = {"ms_per_step": tot / reps, "launches_per_step": len(evs) // reps}

    def avg_ms(name):
        ev = timings.get(name, [])
        return (sum(a.elapsed_time(bb) for a, bb in ev) / len(ev)) if ev else float("nan")

    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = {}
    how = ("CUDA events around each launch of an eager (non-graph) step, every kernel alone on the stream (the prior kernel, "
           "forked in the real step, runs serially here), averaged over 3 steps")
    if "fmd_filter_cfconv_fwd" in timings:
        def tensor_roof(cname, kname, flops):
            ms = avg_ms(cname)
            ach = flops / (ms * 1e-3) / 1e12
            return {"kernel": kname, "bound": "tensor", "achieved": ach, "peak": tf_sus, "unit": "TFLOP/s",
                    "frac": ach / tf_sus, "traffic": traffic.get(kname), "algorithmic_flops_per_launch": flops,
                    "edges": E,
                    "avg_launch_ms": ms, "launches_per_step": len(timings[cname]) // reps,
                    "peak_source": which + " bf16_tflops_sustained (kernel timed inside a long step); burst " +
                                   f"{tf_burst:.0f}",
                    "how": how}
        flops = 2.0 * E * F * (R + F)
        roof = tensor_roof("fmd_filter_cfconv_fwd", "filter_cfconv_fwd_kernel", flops)
        roof["also"] = [tensor_roof("fmd_filter_cfconv_bwd", "filter_cfconv_bwd_kernel", flops)]
        # by time per step the backward kernel is the dominant one: report it first
        if roof["also"][0]["avg_launch_ms"] * roof["also"][0]["launches_per_step"] > roof["avg_launch_ms"] * roof["launches_per_step"]:
            first = roof.pop("also")[0]
            first["also"] = [roof]
            roof = first
        roof["note"] = ("fp16 tcgen05 GEMMs fused with the tanh / gather / segment-reduce epilogues; the kernels are bound by "
                        "their SIMT roles (E x F tanh on the MUFU pipe, E x F multiply-accumulate, role-to-role hand-off "
                        "latency), not by the tensor pipe: see DESIGN.md; avg_launch_ms of the forward includes its fix-up launch")
    else:
        cf_ms = avg_ms("fmd_cfconv_csr")
        alg = E * F * b + 4 * N * F + 4 * N * F + 4 * E + 4 * E + 4 * (N + 1)
        achieved = alg / (cf_ms * 1e-3) / 1e9
        kname = "cfconv_csr128_kernel<float>" if (b == 4 and F == 128) else "cfconv_csr_kernel"
        roof = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic.get(kname),
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": cf_ms,
                "launches_per_step": len(timings.get("fmd_cfconv_csr", [])) // reps,
                "peak_source": which + " hbm_gbs (burst copy)", "how": how}
        if "fmd_linear_x3" in timings:
            # fp32-emulation GEMMs of the filter network (edge level, HBM-bound streaming kernels): algorithmic bytes per
            # interaction block = X read + Y written (+ aux read): rbf->t 4E(R+F), t->W 8EF, gW->gT 12EF, and the fused
            # gT -> g_rbf -> g_d launch 4EF + 12E
            nb = ff.w.num_blocks
            alg3 = nb * (4.0 * E * (R + F) + 8.0 * E * F + 12.0 * E * F)
            ms3 = sum(a.elapsed_time(bb) for a, bb in timings["fmd_linear_x3"]) / reps
            n_node = sum(1 for _ in timings["fmd_linear_x3"]) // reps - 3 * nb
            algr = nb * (4.0 * E * F + 12.0 * E)
            msr = sum(a.elapsed_time(bb) for a, bb in timings.get("fmd_linear_x3_rbf_bwd", [])) / reps
            roof["also"] = [{"kernel": "linear_x3_kernel (BF16x3 fp32-emulation GEMM, edge-level launches)", "bound": "hbm",
                             "achieved": (alg3 + algr) / ((ms3 + msr) * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": (alg3 + algr) / ((ms3 + msr) * 1e-3) / 1e9 / hbm, "traffic": traffic.get("linear_x3_kernel"),
                             "algorithmic_bytes_per_step": alg3 + algr, "ms_per_step": ms3 + msr,
                             "launches_per_step": 4 * nb, "node_level_launches_included_in_time": max(n_node, 0),
                             "peak_source": which + " hbm_gbs (burst copy)", "how": how}]
    return roof, table


if __name__ == "__main__":
    main()
