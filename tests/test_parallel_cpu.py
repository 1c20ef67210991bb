"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2: path sharding,
the packed [gradient | loss] all-reduce, and the fact that sharded sums reproduce the single-process result
(the oracle stands in for the per-rank gradient kernels, which need a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dnnpde_b200 as pde
from dnnpde_b200 import parallel


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 65536, 10 ** 9):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [parallel.shard_range(100, r, 8)[1] - parallel.shard_range(100, r, 8)[0] for r in range(8)] == \
        [13, 13, 13, 13, 12, 12, 12, 12]
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import fbsnn_oracle as orc
        torch.set_num_threads(2)
        torch.manual_seed(21)
        np.random.seed(21)
        D, M, N = 6, 11, 5                      # uneven split: 6 + 5 paths
        layers = [D + 1, 16, 16, 1]
        Xi = np.random.uniform(0.5, 1.5, (1, D))
        sol = orc.OracleSolver("bsptest", Xi, 1.0, M, N, D, layers, "FC", "Sine", squeeze_quirk=False)
        t, W = sol.fetch_minibatch()            # every rank draws the same global minibatch (same NumPy seed)
        lo, hi = parallel.shard_range(M, rank, world)
        loss, _, _, _, grads = sol.grads(t[lo:hi], W[lo:hi])
        flat = torch.cat([g.reshape(-1) for g in grads.values()])
        loss_t = loss.reshape(1).clone()
        parallel.allreduce_grads_and_loss(flat, loss_t)
        full_loss, _, _, _, full_grads = sol.grads(t, W)
        full_flat = torch.cat([g.reshape(-1) for g in full_grads.values()])
        assert torch.allclose(flat, full_flat, rtol=1e-4, atol=1e-5), float((flat - full_flat).abs().max())
        assert abs(float(loss_t) - float(full_loss)) <= 1e-5 * abs(float(full_loss))
        # Monte-Carlo partial sums: (sum, sum of squares) add up over ranks
        sums = torch.tensor([float(rank + 1), float((rank + 1) ** 2)], dtype=torch.float64)
        parallel.allreduce_sums(sums)
        assert sums.tolist() == [3.0, 5.0]
        assert parallel.is_distributed() and parallel.world_size() == 2 and parallel.rank() == rank
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_single_process_helpers_are_noops():
    g = torch.arange(5.0)
    l = torch.tensor([2.0])
    parallel.allreduce_grads_and_loss(g, l)
    assert g.tolist() == [0, 1, 2, 3, 4] and float(l) == 2.0
    assert not parallel.is_distributed() and parallel.world_size() == 1 and parallel.rank() == 0
