#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: BSB-100D FBSNN training iterations/s (FC-Sine 4x256, N=50)
and Monte-Carlo basket paths/s, on N GPUs of one box.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (CPU oracle port) on host cores

One "step" = one training iteration (Brownian minibatch already resident in HBM for `value`; copied from pinned
host memory inside the timed region for `e2e`).  Paths are sharded over ranks (strong scaling: the global batch
M is fixed); the only collective is one sum-allreduce of [gradient | loss] per iteration.  Rank 0 prints ONE JSON
line.  See DESIGN.md section "Measurement" for the algorithmic FLOP count (2.671 MFLOP per (path, step) row).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

D, NSTEPS, H, NLAYERS = 100, 50, 256, 4
LAYERS = [D + 1] + NLAYERS * [H] + [1]
FLOP_PER_ROW = 2.0 * 3 * ((D + 1) * H + 3 * H * H + H + D * H + 3 * H * H + H)   # 3(F + A) MACs, SURVEY section 8d
METRIC = "BSB-100D FBSDE train iters/s"


def xi_bsb():
    return np.array([1.0, 0.5] * (D // 2))[None, :]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_train_rate(paths_global, sample_paths, steps, warmup):
    """iters/s of the reference algorithm for a global batch of `paths_global`, measured on a bounded sample of
    `sample_paths` paths per step (cost is linear in the number of paths: every path is an independent row
    block) and scaled.  Runs the oracle's faithful op sequence (autograd double-backward, diag_embed sigma)."""
    from oracle import fbsnn_oracle as orc
    torch.manual_seed(1234)
    np.random.seed(1234)
    sol = orc.OracleSolver("bsb", xi_bsb(), 1.0, sample_paths, NSTEPS, D, LAYERS, "FC", "Sine")
    sol.make_optimizer(1e-3)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        t, W = sol.fetch_minibatch()
        sol.train_step(t, W)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return (sample_paths / sec) / paths_global, sec


def cpu_mc_rate(n_sample):
    from oracle import mc_oracle as mco
    np.random.seed(0)
    corr = mco.random_correlation(D)
    t0 = time.perf_counter()
    mco.mc_price(np.ones(D), 0.05, 0.2, corr, True, np.ones(D) / D, 1.0, 1.0, NSTEPS, n_sample)
    return n_sample / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sample = args.cpu_sample_paths
    value, sec = cpu_train_rate(args.paths, sample, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of {args.paths} paths per step ({sec:.2f} s/step measured), scaled "
                                   f"linearly in paths; oracle port of DeepBSDE.py train loop, anomaly detection off"},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"BSB-100D FBSNN FC-Sine 4x256, M={args.paths} paths (global), N={NSTEPS} steps, Adam",
            "paths": args.paths, "time_steps": NSTEPS, "dim": D, "precision": args.precision,
            "l2_policy": "inputs larger than L2: each step reads a different resident minibatch "
                         "(1.3 GB at M=65536) and streams GBs of sweep arrays"}


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup(args):
    import torch.distributed as dist

    import dnnpde_b200 as pde
    from dnnpde_b200 import parallel

    c = Ctx()
    c.pde, c.parallel, c.dist = pde, parallel, dist
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=c.dev)
    c.lib = pde._lib.load()

    def barrier():
        if c.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if c.world > 1:
            tt = torch.tensor([x], dtype=torch.float64, device=c.dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt)
        return x

    c.barrier, c.max_over_ranks = barrier, max_over_ranks
    return c


def make_solver(c, args, batches=None):
    """Solver + `nbatch` resident synthetic minibatches (reference layout), generated on the device by Philox."""
    import ctypes
    pde = c.pde
    M = args.paths
    torch.manual_seed(1234)
    sol = pde.BlackScholesBarenblatt(xi_bsb(), 1.0, M, NSTEPS, D, LAYERS, "FC", "Sine", precision=args.precision,
                                     data_parallel=True, collective=args.collective)
    lo, hi = c.parallel.shard_range(M, c.rank, c.world)
    m_loc = hi - lo
    sp = sol._spec()
    ws = sol._workspace(c.lib, sp, m_loc, True)
    if batches is not None:
        return sol, batches, m_loc
    batches = []
    for b in range(2):
        t = torch.empty(m_loc, NSTEPS + 1, 1, device=c.dev)
        W = torch.empty(m_loc, NSTEPS + 1, D, device=c.dev)
        rc = c.lib.fbsnn_fetch_minibatch(ctypes.byref(sp), 1.0, m_loc, lo, 777, b, None, ctypes.c_void_p(ws.data_ptr()),
                                         ws.numel(), ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(W.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        pde._lib.check(rc, "fbsnn_fetch_minibatch")
        batches.append((t, W))
    return sol, batches, m_loc


def measure_value(c, args, sol, batches):
    """K training iterations, minibatches resident in HBM, CUDA events on the launching stream, max over ranks."""
    nb = len(batches)
    loss_buf = torch.zeros(args.warmup + args.steps + 1, device=c.dev)
    sol.begin_training(1e-3)
    l0 = c.lib.fbsnn_launch_count()
    sol.training_step(*batches[0], loss_buf[0:1])           # eager: counts the kernels of one iteration
    launches = c.lib.fbsnn_launch_count() - l0
    for i in range(args.warmup):                             # warm-up (captures one CUDA graph per resident batch)
        sol._step(*batches[i % nb], loss_buf[i:i + 1], False, i, alias_inputs=True)
    c.barrier()
    clocks = ClockSampler(c.local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        sol._step(*batches[i % nb], loss_buf[args.warmup + i:args.warmup + i + 1], False, i, alias_inputs=True)
    e1.record()
    c.barrier()
    ms = c.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop()
    return ms / args.steps, int(launches), clk, float(loss_buf[args.warmup + args.steps - 1])


def measure_roofline(c, args, sol, batches, ms_per_step, m_loc):
    """Dominant kernel = the dense-layer GEMM: per-launch CUDA events around every dense launch of one more step."""
    import ctypes
    pk = peaks()
    loss = torch.zeros(1, device=c.dev)
    nrep = 3                                        # a few consecutive eager steps, every dense launch timed
    c.lib.fbsnn_dense_timing(1)
    for i in range(nrep):
        sol.training_step(*batches[i % len(batches)], loss)
    torch.cuda.synchronize()
    out = (ctypes.c_double * 8)()
    c.pde._lib.check(c.lib.fbsnn_dense_timing_read(out), "timing")
    c.lib.fbsnn_dense_timing(0)
    n_dense, dense_ms, dense_flops, dense_bytes = int(out[0]) // nrep, out[1] / nrep, out[2] / nrep, out[6] / nrep
    traffic = None
    tf = os.path.join(ROOT, "profiles", "r01_traffic.json")   # dram bytes per dense launch from the ncu --set full capture
    if os.path.exists(tf):
        with open(tf) as f:
            tj = json.load(f)
        key = f"{args.precision}_M{args.paths // c.world}"
        traffic = tj.get(key, {}).get("dram_bytes_per_dense_launch")
    tflops = dense_flops / (dense_ms * 1e-3) / 1e12 if dense_ms > 0 else 0.0
    gbs = dense_bytes / (dense_ms * 1e-3) / 1e9 if dense_ms > 0 else 0.0
    tf32_peak = pk["bf16"] * 0.5          # kind::tf32 issues at half the bf16 rate (nominal 1.1 vs 2.25 PFLOP/s)
    is_tc = out[3] > 0
    # The dense-layer kernel moves 3-5 row arrays per 2*256*256 FLOP per row (26-43 FLOP/B, ridge ~130 FLOP/B for
    # TF32), so on the tcgen05 variant it is HBM-bound: report it against the measured copy bandwidth.  The SIMT
    # fp32 variant is FMA-issue bound; it is reported against the same HBM peak for comparability, with the FLOP
    # rates alongside ("Roofline").
    return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
            "traffic": traffic, "algorithmic_bytes_per_launch": dense_bytes / max(n_dense, 1),
            "kernel": (("gemm_tc_kernel (tcgen05 kind::tf32 x3 hi/lo split, TMA, TMEM; sweeps) + gemm_tc2_kernel "
                        "(cta_group::2 pair; weight gradients)") if args.precision == "tf32x3" else
                       "gemm_tc_kernel (tcgen05 kind::tf32, TMA, TMEM)") if is_tc else "gemm_simt_kernel (fp32 FMA)",
            "launches_per_step": n_dense, "dense_ms_per_step": dense_ms, "dense_share_of_step": dense_ms / ms_per_step,
            "algorithmic_bytes_per_step": dense_bytes, "tflops": tflops, "tf32_peak_tflops": tf32_peak,
            "tensor_frac": tflops / tf32_peak, "tc_launches": int(out[3]) // nrep,
            "peak_source": f"{pk['source']}: HBM copy {pk['hbm']} GB/s; bf16 {pk['bf16']} TFLOP/s x 0.5 for tf32",
            "step_tflops": FLOP_PER_ROW * m_loc * (NSTEPS + 1) / (ms_per_step * 1e-3) / 1e12}


def measure_e2e(c, args, sol, batches):
    """The public train() API with the minibatch copied from pinned host memory every step (double-buffered on a
    copy stream, every byte crosses PCIe inside the timed region) and the per-step losses read back at the end."""
    import contextlib
    import io
    dev = c.dev
    host = [(t.cpu().pin_memory(), W.cpu().pin_memory()) for t, W in batches]
    copy_stream = torch.cuda.Stream(device=dev)
    state = {"i": 0, "pending": None}

    def issue(i):
        th, Wh = host[i % len(host)]
        with torch.cuda.stream(copy_stream):
            td = torch.empty(th.shape, device=dev)
            Wd = torch.empty(Wh.shape, device=dev)
            td.copy_(th, non_blocking=True)
            Wd.copy_(Wh, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return td, Wd, ev

    def fetch():
        if state["pending"] is None:
            state["pending"] = issue(state["i"])
        td, Wd, ev = state["pending"]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        td.record_stream(cur), Wd.record_stream(cur)
        state["i"] += 1
        state["pending"] = issue(state["i"])
        return td, Wd

    sol.fetch_minibatch = fetch
    with contextlib.redirect_stdout(io.StringIO()):
        sol.train(max(1, args.warmup), 1e-3)
        c.barrier()
        t0 = time.perf_counter()
        sol.train(args.steps, 1e-3)
        c.barrier()
        sec = c.max_over_ranks(time.perf_counter() - t0)
    h2d = int((host[0][0].numel() + host[0][1].numel()) * 4)
    return {"value": args.steps / sec, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
            "api": "BlackScholesBarenblatt.train(K, lr), host-supplied Brownian minibatches (pinned, double-buffered "
                   "H2D on a copy stream), losses and Y0 read back"}


def measure_mc(c, args):
    pde = c.pde
    np.random.seed(0)
    model = pde.BlackScholesModel(0.05, 0.2, D, True)
    pr = pde.MonteCarloPricer(model, pde.BasketOption(np.ones(D) / D, 1.0), 1.0, NSTEPS, args.mc_paths, seed=7,
                              data_parallel=True)
    n_lo, n_hi = c.parallel.shard_range(args.mc_paths, c.rank, c.world)
    pr.price_async(np.ones(D), max(1, (n_hi - n_lo) // 16), n_lo, 7)      # warm-up
    c.barrier()
    m0 = c.lib.mc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sums = pr.price_async(np.ones(D), n_hi - n_lo, n_lo, 7)
    e1.record()
    c.barrier()
    mc_ms = c.max_over_ranks(e0.elapsed_time(e1))
    c.parallel.allreduce_sums(sums)
    s, q = (float(v) for v in sums.cpu())
    mean = s / args.mc_paths
    se = float(np.sqrt(max(q / args.mc_paths - mean * mean, 0.0) / args.mc_paths))
    pps = args.mc_paths / (mc_ms * 1e-3)
    return {"metric": "MC basket paths/s", "value": pps, "unit": "paths/s", "paths": args.mc_paths, "ms": mc_ms,
            "price": mean, "stderr": se, "normals_per_s": pps * NSTEPS * D,
            "gpu_launches": int(c.lib.mc_launch_count() - m0),
            "form": "N*D Philox/Box-Muller normals per path, Cholesky matvec hoisted (L sum_t z_t)"}


def run_ours(args):
    c = setup(args)
    sol, batches, m_loc = make_solver(c, args)
    ms_per_step, launches, clk, final_loss = measure_value(c, args, sol, batches)
    line = {
        "metric": METRIC, "value": 1e3 / ms_per_step, "unit": "iters/s", "n_gpus": c.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "data": "synthetic",
        "dtype": {"tf32": "tf32", "tf32x3": "tf32x3 (fp32-grade: hi/lo-split TF32 MMAs, fp32 accumulate)",
                  "fp32": "f32"}[args.precision],
        "config": workload_config(args), "clocks": clk, "gpu_launches": launches, "final_loss": final_loss,
    }
    if c.world > 1:
        line["collective"] = ("fbsnn_peer_allreduce_adam (P2P loads over NVLink, fused with clip norm + Adam)"
                              if sol.collective == "peer" else "NCCL all-reduce of [grad | loss]")
    line["roofline"] = measure_roofline(c, args, sol, batches, ms_per_step, m_loc)
    if not args.skip_e2e:
        line["e2e"] = measure_e2e(c, args, sol, batches)
    del sol
    torch.cuda.empty_cache()
    if not args.skip_fp32:
        # the other arithmetic variants on the same workload and minibatches (tests/test_parity_gpu.py states
        # each variant's tolerance): single-pass TF32 (faster, looser) and SIMT fp32 (slower, tightest)
        notes = {"tf32": "tcgen05 kind::tf32 single pass; tolerance TOL_TF32 (loss 1e-2, Z 5e-2)",
                 "tf32x3": "tcgen05 3xTF32 (hi/lo split), fp32-grade; tolerance TOL_X3 (loss 1e-4, Z 5e-5)",
                 "fp32": "gemm_simt_kernel fp32 FMA; tolerance TOL (loss 2e-5, Z 1e-5)"}
        line["variants"] = {args.precision: {"value": 1e3 / ms_per_step, "ms_per_step": ms_per_step,
                                             "note": notes[args.precision]}}
        for prec in ("tf32x3", "tf32", "fp32"):
            if prec == args.precision:
                continue
            a2 = argparse.Namespace(**vars(args))
            a2.precision, a2.steps = prec, (max(2, args.steps // 3) if prec == "fp32" else args.steps)
            sol2, _, _ = make_solver(c, a2, batches)
            ms2, _, _, _ = measure_value(c, a2, sol2, batches)
            line["variants"][prec] = {"value": 1e3 / ms2, "ms_per_step": ms2, "steps": a2.steps, "note": notes[prec]}
            del sol2
            torch.cuda.empty_cache()
    del batches
    torch.cuda.empty_cache()
    if not args.skip_small and args.paths != 100:
        # the other end of BASELINE.json's range (configs[1]: M=100): launch-latency bound, replayed as one CUDA graph
        a3 = argparse.Namespace(**vars(args))
        a3.paths, a3.steps, a3.warmup = 100, 200, 20
        sol3, b3, _ = make_solver(c, a3)
        ms3, l3, _, _ = measure_value(c, a3, sol3, b3)
        line["small_m"] = {"metric": METRIC, "value": 1e3 / ms3, "unit": "iters/s", "ms_per_step": ms3, "paths": 100,
                           "steps": a3.steps, "warmup": a3.warmup, "precision": a3.precision, "gpu_launches": int(l3),
                           "note": "M=100 (5 100 rows, L2-resident): bound by the latency of ~38 dependent launches"}
        del sol3, b3
        torch.cuda.empty_cache()
    if not args.skip_mc:
        line["mc"] = measure_mc(c, args)
    # CPU baseline (rank 0, N = 1 only): oracle port on the host cores, bounded sample
    if c.rank == 0 and c.world == 1 and not args.skip_cpu:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        v, sec = cpu_train_rate(args.paths, args.cpu_sample_paths, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "iters/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_sample_paths} of {args.paths} paths per step ({sec:.2f} s/step), "
                                          "scaled linearly in paths"}
        if not args.skip_mc:
            line["mc"]["cpu_baseline"] = {"value": cpu_mc_rate(20000), "unit": "paths/s", "cores": 1, "kind": "port",
                                          "sample": "20000 paths, D=100, N=50"}
    if c.rank == 0:
        print(json.dumps(line))
    if c.world > 1:
        c.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--paths", type=int, default=65536, help="global number of Brownian paths M")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32", "tf32x3"],
                    help="tf32x3 (default): tcgen05 tensor cores with hi/lo operand split, fp32-grade results; "
                         "tf32: single-pass tcgen05 (looser stated tolerance); fp32: SIMT FMA (tightest parity)")
    ap.add_argument("--mc-paths", type=int, default=1 << 28)
    ap.add_argument("--cpu-sample-paths", type=int, default=256)
    ap.add_argument("--skip-mc", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-small", action="store_true", help="do not also time the M=100 configuration")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU gradient all-reduce: fused NVLink peer-memory kernel (default) or NCCL")
    ap.add_argument("--skip-fp32", action="store_true", help="do not also time the fp32 SIMT variant")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only (the JSON line then has no e2e)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
