"""world_size-2 gloo tests of the multi-GPU host logic on CPU: sharding, energy all-gather, identical
decisions on every rank, cross-rank swaps equal to the single-process reference exchange."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "flash-molecular-dynamics_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _reference_exchange(x, v, beta, pair_a, pair_b, accepted, n_atoms):
    x, v = x.clone().view(-1, n_atoms, 3), v.clone().view(-1, n_atoms, 3)
    a, b = pair_a[accepted], pair_b[accepted]
    xa, xb, va, vb = x[a].clone(), x[b].clone(), v[a].clone(), v[b].clone()
    s = torch.sqrt(beta[a] / beta[b])[:, None, None]
    x[a], x[b] = xb, xa
    v[a], v[b] = vb * s, va / s
    return x.view(-1, 3), v.view(-1, 3)


def _worker(rank, world, port, n_indep, n_atoms, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flashmd.simulation.distributed import ShardedExchange, exchange_uniforms, shard_range
    from flashmd.simulation.parallel_tempering import adjacent_pairs
    betas = [1.67, 1.42, 1.16, 1.0]
    n_total = len(betas) * n_indep
    beta_all = torch.tensor([b for b in betas for _ in range(n_indep)])
    g = torch.Generator().manual_seed(0)
    x_all = torch.randn((n_total * n_atoms, 3), generator=g)
    v_all = torch.randn((n_total * n_atoms, 3), generator=g)
    lo, hi = shard_range(n_total, rank, world)
    x, v = x_all[lo * n_atoms:hi * n_atoms].clone(), v_all[lo * n_atoms:hi * n_atoms].clone()
    ex = ShardedExchange(beta_all, n_atoms, rank, world)
    even, odd = adjacent_pairs(len(betas), n_indep)
    x_ref, v_ref = x_all.clone(), v_all.clone()
    for k in range(6):
        e_all = torch.randn(n_total, generator=g) * 3.0
        pa, pb = (even if k % 2 == 0 else odd)
        u = exchange_uniforms(11, k, len(pa))
        acc = ex.exchange(x, v, e_all[lo:hi].clone(), pa, pb, u)
        acc_ref = u < torch.exp((e_all[pa] - e_all[pb]) * (beta_all[pa] - beta_all[pb]))
        assert torch.equal(acc, acc_ref)
        x_ref, v_ref = _reference_exchange(x_ref, v_ref, beta_all, pa, pb, acc_ref, n_atoms)
        assert torch.equal(x, x_ref[lo * n_atoms:hi * n_atoms]), (rank, k)
        assert torch.allclose(v, v_ref[lo * n_atoms:hi * n_atoms], rtol=1e-6, atol=0), (rank, k)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([int(acc.sum())]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_indep", [1, 3])
def test_sharded_replica_exchange_world2(tmp_path, n_indep):
    # n_indep=1: every adjacent-beta pair of the 4 replicas that straddles the shard boundary crosses ranks;
    # n_indep=3: a mix of rank-local and cross-rank pairs (6 sims per rank)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_indep, 5, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok_0.npy") and os.path.exists(tmp_path / "ok_1.npy")


def test_shard_helpers():
    from flashmd.simulation.distributed import exchange_uniforms, shard_configurations, shard_range
    assert shard_range(8, 1, 2) == (4, 8)
    with pytest.raises(ValueError):
        shard_range(7, 0, 2)
    assert shard_configurations(list(range(8)), rank=3, world=4) == [6, 7]
    assert torch.equal(exchange_uniforms(3, 5, 7), exchange_uniforms(3, 5, 7))
    assert not torch.equal(exchange_uniforms(3, 5, 7), exchange_uniforms(3, 6, 7))
