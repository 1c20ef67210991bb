#!/usr/bin/env python
"""Multi-GPU check of the fused NVLink all-reduce (fbsnn_peer_allreduce_adam), run under torchrun on N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/peer_check.py

K training iterations of a BSB problem with (a) the peer-memory kernel, (b) the NCCL all-reduce and, on rank 0,
(c) the whole batch on one GPU; parameters after K steps must agree (a == b bit-for-bit at world 2, where both
sum two addends; a ~ c within fp32 reassociation), and all ranks must hold bit-identical parameters."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import dnnpde_b200 as pde


def run(collective, data_parallel, M, K, precision):
    D, N = 100, 20
    layers = [D + 1, 256, 256, 256, 256, 1]
    torch.manual_seed(3)
    Xi = np.array([1.0, 0.5] * (D // 2))[None, :]
    sol = pde.BlackScholesBarenblatt(Xi, 1.0, M, N, D, layers, "FC", "Sine", precision=precision,
                                     data_parallel=data_parallel, collective=collective)
    np.random.seed(17)                       # every rank draws the same global minibatches
    sol.train(K, 1e-3)
    torch.cuda.synchronize()
    return sol._fp.flat.detach().clone(), np.array(sol.last_losses), sol.collective


def run_philox(data_parallel, M, K1, K2, cuda_graph=True):
    """Two consecutive train() calls with in-kernel increments (brownian='philox'): -> per-iteration losses of both."""
    D, N = 100, 20
    layers = [D + 1, 256, 256, 256, 256, 1]
    torch.manual_seed(3)
    Xi = np.array([1.0, 0.5] * (D // 2))[None, :]
    sol = pde.BlackScholesBarenblatt(Xi, 1.0, M, N, D, layers, "FC", "Sine", precision="fp32", brownian="philox", seed=41,
                                     data_parallel=data_parallel, collective="peer", cuda_graph=cuda_graph)
    sol.train(K1, 1e-3)
    l1 = np.array(sol.last_losses)
    sol.train(K2, 1e-3)
    l2 = np.array(sol.last_losses)
    torch.cuda.synchronize()
    return l1, l2, sol._fp.flat.detach().clone()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # ---- in-kernel increments on several GPUs: the persistent device counter keys Philox -------------
    a1, a2, pa = run_philox(True, 203, 4, 4, cuda_graph=True)       # data-parallel, graph-captured peer step
    e1, e2, pe = run_philox(True, 203, 4, 4, cuda_graph=False)      # data-parallel, eager
    fresh = not np.allclose(a1, a2, rtol=1e-3)                       # second train() call draws NEW noise
    graph_eq = bool(np.allclose(a1, e1, rtol=1e-6) and np.allclose(a2, e2, rtol=1e-6) and
                    float((pa - pe).abs().max()) <= 1e-6)
    msg = f"[philox] second train() draws fresh noise={fresh} graph==eager={graph_eq}"
    ok &= fresh and graph_eq
    if rank == 0:
        s1, s2, _ = run_philox(False, 203, 4, 4)                     # the same two calls on ONE GPU
        rl = float(max(np.max(np.abs(a1 - s1) / np.abs(s1)), np.max(np.abs(a2 - s2) / np.abs(s2))))
        msg += f" max rel loss diff vs 1 GPU={rl:.3e}"
        ok &= rl <= 2e-4
        print(msg, flush=True)
    for precision, tol in (("fp32", 2e-5), ("tf32x3", 5e-5)):
        M, K = 203, 6                        # uneven shards
        p_peer, l_peer, used = run("peer", True, M, K, precision)
        p_nccl, l_nccl, _ = run("nccl", True, M, K, precision)
        # identical parameters on every rank
        gathered = [torch.empty_like(p_peer) for _ in range(world)]
        dist.all_gather(gathered, p_peer)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        d_pn = float((p_peer - p_nccl).abs().max())
        msg = f"[{precision}] collective={used} ranks_identical={same} max|peer-nccl|={d_pn:.3e}"
        ok &= same and used == "peer" and d_pn <= (0.0 if world == 2 else 1e-5)
        if rank == 0:
            # (c) undistributed reference on one GPU
            p_one, l_one, _ = run("nccl", False, M, K, precision)
            d_po = float((p_peer - p_one).abs().max())
            rl = float(np.max(np.abs(l_peer - l_one) / np.abs(l_one)))
            msg += f" max|peer-single|={d_po:.3e} max rel loss diff={rl:.3e}"
            ok &= d_po <= 2e-4 and rl <= tol * 10
            print(msg, flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_CHECK", "OK" if int(flag) else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
