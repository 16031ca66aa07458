"""Multi-GPU checks (run under torchrun on the GPU box, NCCL):
  1. ShardedExchange on CUDA tensors == single-process reference exchange;
  2. PTSimulation sharded over the ranks runs through the fused engine with NCCL exchanges."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "flash-molecular-dynamics_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from flashmd.simulation.distributed import ShardedExchange, exchange_uniforms, shard_range
    from flashmd.simulation.parallel_tempering import adjacent_pairs
    from test_distributed_cpu import _reference_exchange
    betas, n_indep, n_atoms = [1.67, 1.42, 1.16, 1.0], world, 7
    n_total = len(betas) * n_indep
    beta_all = torch.tensor([b for b in betas for _ in range(n_indep)])
    g = torch.Generator().manual_seed(0)
    x_all = torch.randn((n_total * n_atoms, 3), generator=g)
    v_all = torch.randn((n_total * n_atoms, 3), generator=g)
    lo, hi = shard_range(n_total, rank, world)
    x, v = x_all[lo * n_atoms:hi * n_atoms].to(dev), v_all[lo * n_atoms:hi * n_atoms].to(dev)
    ex = ShardedExchange(beta_all, n_atoms, rank, world)
    even, odd = adjacent_pairs(len(betas), n_indep)
    x_ref, v_ref = x_all.clone(), v_all.clone()
    for k in range(6):
        e_all = torch.randn(n_total, generator=g) * 3.0
        pa, pb = (even if k % 2 == 0 else odd)
        u = exchange_uniforms(11, k, len(pa))
        acc = ex.exchange(x, v, e_all[lo:hi].to(dev), pa, pb, u)
        acc_ref = u < torch.exp((e_all[pa] - e_all[pb]) * (beta_all[pa] - beta_all[pb]))
        assert torch.equal(acc.cpu().bool(), acc_ref)
        x_ref, v_ref = _reference_exchange(x_ref, v_ref, beta_all, pa, pb, acc_ref, n_atoms)
        assert torch.equal(x.cpu(), x_ref[lo * n_atoms:hi * n_atoms]), (rank, k)
        assert torch.allclose(v.cpu(), v_ref[lo * n_atoms:hi * n_atoms], rtol=1e-6, atol=0), (rank, k)
    # counter-based decisions (no uniforms at all): identical on every rank without communication
    acc = ex.exchange(x, v, e_all[lo:hi].to(dev), even[0], even[1], None, seed=5, exchange_index=3)
    allacc = [torch.empty_like(acc) for _ in range(world)]
    dist.all_gather(allacc, acc)
    assert all(torch.equal(a, allacc[0]) for a in allacc)
    if rank == 0:
        print(f"[dist_check] NCCL sharded exchange == reference on {world} GPUs (device-side decide / select, no host read)")
    # ---- PT simulation sharded over the ranks
    from helpers import dropin_model_from_golden, load_golden
    from flashmd.simulation import PTSimulation
    gold = load_golden("schnet_n54_b4.npz")
    model, _, configs = dropin_model_from_golden(gold)
    out = f"/tmp/pt_dist_{rank}"
    os.makedirs(out, exist_ok=True)
    sim = PTSimulation(friction=1.0, dt=0.004, n_timesteps=200, save_interval=10, export_interval=100,
                       exchange_interval=20, save_energies=True, random_seed=7, device=str(dev), filename="pt",
                       output_dir=out, gptq="w16a16")
    sim.attach_model_and_configurations(model, [configs[i % len(configs)] for i in range(world)],
                                        betas=[1.67, 1.42, 1.16, 1.0])
    assert sim.n_sims == 4 * world // world
    assert sim._node_offset == rank * sim.n_sims * sim.n_atoms      # Philox noise keyed by the global bead index
    sim.simulate()
    # shards must not share a noise stream: the velocities of local replica 0 differ between ranks
    v0 = sim.engine.vel[: sim.n_atoms].clone()
    vs = [torch.empty_like(v0) for _ in range(world)]
    dist.all_gather(vs, v0)
    if world > 1:
        assert not torch.allclose(vs[0], vs[1])
    m = sim.get_throughput_metrics()
    ok = np.isfinite(sim.simulated_coords).all() and m["path"] == "fused-engine"
    t = torch.tensor([int(ok), sim.exchange_summary["approved"]], device=dev)
    dist.all_reduce(t)
    if rank == 0:
        print(f"[dist_check] sharded PTSimulation ok on all ranks: {int(t[0]) == world}; exchanges attempted "
              f"{sim.exchange_summary['attempted']} approved(sum over ranks)/world {int(t[1]) // world}")
    assert int(t[0]) == world
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
