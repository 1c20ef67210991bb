#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: FBSNN training iterations/s and Monte-Carlo basket paths/s on N
GPUs of one box.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the UNMODIFIED reference (baseline/_ref) on the host cores
    python bench.py --workload basket_nais --dim 100 --act ReLU --paths 16384    # one workload of configs 3-4 as headline

Headline (`value`): BSB-100D, FC-Sine 4x256, N = 50, M = 65 536 global paths (BASELINE.json configs[1]), default
precision tf32x3; one "step" = one training iteration with the Brownian minibatch already resident in HBM.  `e2e` is the
same metric through the public train() API with host (pinned) minibatches copied inside the timed region.  Paths are
sharded over ranks (strong scaling: the global batch is fixed); the only collective is one sum-all-reduce of
[gradient | loss] per iteration.  Rank 0 prints ONE JSON line.  Sub-objects: `roofline` (tensor-core fraction of the
whole step against a TF32 peak measured in this process, per-launch table, DRAM traffic), `variants`, `small_m` /
`mid_m` (M = 100 / 4 096, the launch-bound end), `workloads` (NAIS-Net basket D = 5/10/50/100, correlated 100-D,
HJB-100D: BASELINE.json configs 3-4), `mc` (10^9-path correlated basket, with its generator roofline) and
`cpu_baseline` (the unmodified reference timed on the host cores in this run).  DESIGN.md section 6 has the definitions.
"""
from __future__ import annotations

import os
import sys

if "--impl=reference" in sys.argv or ("--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""    # the reference moves itself to cuda:0 when it sees one (DeepBSDE.py:143)

import argparse
import contextlib
import io
import json
import subprocess
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

NSTEPS, H, NHIDDEN = 50, 256, 4
METRIC = "BSB-100D FBSDE train iters/s"


def flop_per_row(net: str, D: int, Hw: int = H, L: int = NHIDDEN) -> float:
    """Algorithmic FLOPs of one (path, step) row: 3 (F + A) MACs (3xTF32 counts them once)."""
    d = D + 1
    if net == "FC":
        F, A = d * Hw + (L - 1) * Hw * Hw + Hw, D * Hw + (L - 1) * Hw * Hw + Hw
    else:   # NAIS-Net, L - 1 stable blocks, every block re-reads the input
        F, A = L * d * Hw + (L - 1) * Hw * Hw + Hw, L * D * Hw + (L - 1) * Hw * Hw + Hw
    return 2.0 * 3.0 * (F + A)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.t_begin = self.t_end = None

    def start(self):
        """Starts sampling (call before the warm-up: nvidia-smi takes a few hundred ms to deliver its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0:   # nvidia-smi needs a few hundred ms for its first line
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()

        def parse(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
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
            return sm, mx, reasons

        window = "timed region"
        inside = [x for x in self.lines if self.t_begin is not None and self.t_begin <= x[0] <= (self.t_end or 1e30) + 0.05]
        sm, mx, reasons = parse(inside)
        if not sm:   # a timed region shorter than one sampling period: the samples under load since the warm-up
            window = "warm-up + timed region (the timed region was shorter than one sampling period)"
            sm, mx, reasons = parse(self.lines[-8:])
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs 2-4)
# ------------------------------------------------------------------------------------------------------------
def workload_spec(kind: str, dim: int, act: str):
    """-> dict(label, net, problem class name, Xi, ctor kwargs) of one benchmarked FBSNN configuration."""
    layers = [dim + 1] + NHIDDEN * [H] + [1]
    if kind == "bsb":
        return dict(label=f"BSB-{dim}D FBSNN FC-{act} 4x256", net="FC", cls="BlackScholesBarenblatt",
                    xi=np.array([1.0, 0.5] * (dim // 2) + [1.0] * (dim % 2))[None, :], args=(layers, "FC", act), kw={})
    if kind == "basket_nais":
        return dict(label=f"basket call {dim}D NAIS-Net-{act} (3 stable blocks x 256), clip 1.0", net="NAIS",
                    cls="BasketCallOption", xi=np.ones((1, dim)), args=(None, layers, "Naisnet", act, "no_correlation"), kw={})
    if kind == "corr_nais":
        return dict(label=f"correlated basket call {dim}D NAIS-Net-{act}, Cholesky-correlated increments, clip 1.0",
                    net="NAIS", cls="BasketCallOption", xi=np.ones((1, dim)),
                    args=(None, layers, "Naisnet", act, "random_correlation"), kw={}, unit_diag_corr=True)
    if kind == "hjb":
        return dict(label=f"HJB-{dim}D NAIS-Net-{act}, clip 1.0", net="NAIS", cls="HamiltonJacobiBellman",
                    xi=np.zeros((1, dim)), args=(layers, "Naisnet", act), kw={})
    raise ValueError(kind)


def workload_config(args):
    w = workload_spec(args.workload, args.dim, args.act)
    return {"workload": f"{w['label']}, M={args.paths} paths (global), N={NSTEPS} steps, Adam",
            "paths": args.paths, "time_steps": NSTEPS, "dim": args.dim, "precision": args.precision,
            "l2_policy": "inputs larger than L2: each step reads a different resident minibatch "
                         "(1.3 GB at M=65536) and streams GBs of sweep arrays"}


# ------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference (baseline/_ref) on the host cores
# ------------------------------------------------------------------------------------------------------------
def ref_subprocess(argv, timeout=900):
    """Run baseline/ref_bench.py in its own process (CUDA hidden there) and parse its JSON line."""
    try:
        # all the host threads the reference can use: torchrun exports OMP_NUM_THREADS=1 to every rank, which would time the
        # reference on ONE core (7.0 instead of 0.9 s per 1000-path iteration)
        env = dict(os.environ)
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            env[k] = str(os.cpu_count() or 1)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_bench.py")] + argv, capture_output=True,
                           text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"available": False, "error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:   # noqa: BLE001
        return {"available": False, "error": repr(e)[:300]}


def port_train_rate(sample_paths, steps, warmup, dim=100):
    """Fallback when baseline/_ref is absent: the oracle port of the reference loop (kind = "port")."""
    from oracle import fbsnn_oracle as orc
    torch.manual_seed(1234)
    np.random.seed(1234)
    layers = [dim + 1] + NHIDDEN * [H] + [1]
    sol = orc.OracleSolver("bsb", np.array([1.0, 0.5] * (dim // 2))[None, :], 1.0, sample_paths, NSTEPS, dim, layers,
                           "FC", "Sine")
    sol.make_optimizer(1e-3)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        t, W = sol.fetch_minibatch()
        sol.train_step(t, W)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def cpu_baseline_block(args, steps, warmup, suite):
    """cpu_baseline object: the reference's train() at a bounded batch (sample) on all host cores, measured, plus the
    labelled extrapolation to the full batch; optionally the BASELINE.md section 3 table (suite)."""
    cores = os.cpu_count()
    sample = args.cpu_sample_paths
    res = ref_subprocess(["train", "--paths", str(sample), "--steps", str(steps), "--warmup", str(warmup)])
    if res.get("available") and "sec_per_iter" in res:
        sec, kind = res["sec_per_iter"], "reference"
        how = "DeepBSDE.BlackScholesBarenblatt.train() of the unmodified reference (baseline/_ref), anomaly detection off"
    else:
        sec, kind = port_train_rate(min(sample, 256), steps, warmup), "port"
        sample = min(sample, 256)
        how = "oracle port of the DeepBSDE.py train loop (baseline/_ref not installed: " + str(res.get("error", ""))[:80] + ")"
    cores = res.get("torch_threads") or cores            # the threads the reference actually ran on
    out = {"value": 1.0 / sec, "unit": "iters/s", "cores": cores, "kind": kind,
           "sample": f"{sample}-path minibatches of the {args.paths}-path workload, {steps} measured iterations "
                     f"({sec:.3f} s each); {how}",
           "sample_paths": sample, "sec_per_iter": sec,
           "extrapolated_full_batch": {"iters_per_s": (sample / sec) / args.paths,
                                       "note": f"LABELLED EXTRAPOLATION: linear in paths from the measured {sample}-path rate; "
                                               "the reference cannot run M=65536 (about 262 GB of retained diag_embed "
                                               "tensors, SURVEY.md section 6.2)"},
           "host": {k: res.get(k) for k in ("cpu_model", "cpu_count", "torch_threads")}}
    if suite:
        out["suite"] = ref_subprocess(["suite", "--budget", str(args.cpu_suite_budget)], timeout=1800).get("rows")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    blk = cpu_baseline_block(args, args.steps, args.warmup, suite=False)
    sec = blk["sec_per_iter"]
    line = {
        "impl": "reference", "metric": METRIC, "value": blk["value"], "unit": "iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args), reference_step=f"one iteration of a {blk['sample_paths']}-path minibatch "
                       "(bounded sample of the workload; value and ms_per_step are MEASURED for that batch, "
                       "not scaled)"),
        "cpu_baseline": blk,
        "e2e": {"value": blk["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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


def build_solver(c, kind, dim, act, paths, precision, collective, **extra):
    w = workload_spec(kind, dim, act)
    cls = getattr(c.pde, w["cls"])
    torch.manual_seed(1234)
    np.random.seed(1234)
    sol = cls(w["xi"], 1.0, paths, NSTEPS, dim, *w["args"], precision=precision, data_parallel=True, collective=collective,
              **w["kw"], **extra)
    if w.get("unit_diag_corr"):
        # benchmark with a unit-diagonal correlation (the reference's own FBSNN generator has a diagonal of ~135 at
        # D = 100; that matrix is covered by the parity fixture basket100_nais_relu_corr)
        np.random.seed(0)
        sol.correlation_matrix = c.pde.CorrelationMatrix(dim).matrix
        sol._chol_dev = None
    return sol, w


def make_batches(c, sol, m_loc, lo, dim, n=2):
    """`n` resident synthetic minibatches (reference layout) drawn on the device by Philox."""
    import ctypes
    sp = sol._spec()
    ws = sol._workspace(c.lib, sp, m_loc, True)
    batches = []
    for b in range(n):
        t = torch.empty(m_loc, NSTEPS + 1, 1, device=c.dev)
        W = torch.empty(m_loc, NSTEPS + 1, dim, device=c.dev)
        chol = sol._chol_device()
        rc = c.lib.fbsnn_fetch_minibatch(ctypes.byref(sp), 1.0, m_loc, lo, 777, b,
                                         None if chol is None else ctypes.c_void_p(chol.data_ptr()),
                                         ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(t.data_ptr()),
                                         ctypes.c_void_p(W.data_ptr()),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        c.pde._lib.check(rc, "fbsnn_fetch_minibatch")
        batches.append((t, W))
    return batches


def time_steps(c, sol, batches, steps, warmup, sample_clocks=False):
    """`steps` training iterations after `warmup`, CUDA events on the launching stream, max over ranks.  batches = list
    of resident (t, W) pairs cycled through, or None: Brownian increments drawn in-kernel (Philox) inside the step."""
    loss_buf = torch.zeros(warmup + steps + 1, device=c.dev)
    sol.begin_training(1e-3)
    l0 = c.lib.fbsnn_launch_count()
    first = batches[0] if batches else (None, None)
    sol.training_step(*first, loss_buf[0:1])                 # eager: counts the kernels of one iteration
    launches = c.lib.fbsnn_launch_count() - l0

    def step(i, k):
        b = batches[i % len(batches)] if batches else (None, None)
        sol._step(*b, loss_buf[k:k + 1], False, i, alias_inputs=bool(batches))

    clocks = ClockSampler(c.local) if sample_clocks else None
    if clocks:
        clocks.start()
    for i in range(warmup):                                  # warm-up (captures one CUDA graph per resident batch)
        step(i, i)
    c.barrier()
    if clocks:
        clocks.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i, warmup + i)
    e1.record()
    c.barrier()
    if clocks:
        clocks.mark_end()
    ms = c.max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if clocks else None
    return ms / steps, int(launches), clk, float(loss_buf[warmup + steps - 1])


def measure_tf32_peak(dev, seconds=1.5):
    """TF32 tensor-core peak measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with TF32 inputs
    (cuBLAS), best of 10 (burst) and back to back for `seconds` (sustained, power-capped)."""
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(seconds * 1e3 / best))
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
        del a, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    fl = 2.0 * n ** 3
    return {"burst_tflops": fl / (best * 1e-3) / 1e12, "sustained_tflops": fl / (sustained * 1e-3) / 1e12,
            "how": f"torch.matmul fp32 {n}^3 with allow_tf32 (cuBLAS TF32): best of 10 and {reps} back to back"}


def measure_roofline(c, args, sol, batches, ms_per_step, m_loc, net, tf32_peak):
    """Whole-step tensor-core fraction + CUDA-event table of every dense launch of a step."""
    import ctypes
    pk = peaks()
    loss = torch.zeros(1, device=c.dev)
    nrep = 3
    c.lib.fbsnn_dense_timing(1)
    for i in range(nrep):
        b = batches[i % len(batches)] if batches else (None, None)
        sol.training_step(*b, loss)
    torch.cuda.synchronize()
    rows_tab = {}
    order = []
    out4 = (ctypes.c_double * 4)()
    i = 0
    while True:
        tag = c.lib.fbsnn_dense_timing_entry(i, out4)
        if tag is None:
            break
        tag = tag.decode()
        if tag not in rows_tab:
            rows_tab[tag] = dict(launches=0, ms=0.0, flops=0.0, bytes=0.0, tc=bool(out4[3]))
            order.append(tag)
        r = rows_tab[tag]
        r["launches"] += 1
        r["ms"] += out4[0]
        r["flops"] += out4[1]
        r["bytes"] += out4[2]
        i += 1
    c.lib.fbsnn_dense_timing(0)
    table = []
    dense_ms = 0.0
    for tag in order:
        r = rows_tab[tag]
        ms = r["ms"] / nrep
        dense_ms += ms
        table.append({"sweep": tag, "launches_per_step": r["launches"] // nrep, "ms_per_step": ms,
                      "tflops": r["flops"] / nrep / (ms * 1e-3) / 1e12 if ms > 0 else 0.0,
                      "algorithmic_gbs": r["bytes"] / nrep / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                      "tensor_cores": r["tc"]})
    R = m_loc * (NSTEPS + 1)
    fpr = flop_per_row(net, args.dim)
    step_tflops = fpr * R / (ms_per_step * 1e-3) / 1e12
    peak = tf32_peak["sustained_tflops"] if tf32_peak else pk["bf16_sustained"] * 0.5
    traffic = None
    tf = os.path.join(ROOT, "profiles", "r02_traffic.json")   # DRAM bytes of one step from the ncu --set full capture
    if os.path.exists(tf):
        with open(tf) as f:
            traffic = json.load(f).get(f"{args.workload}_{args.precision}_M{m_loc}", {}).get("dram_bytes_per_step")
    # algorithmic HBM lower bound of SURVEY.md section 8(d): increments in, X/Y stacks out, parameter state
    alg_bytes = 4.0 * m_loc * NSTEPS * args.dim + 4.0 * R * (args.dim + 1) + 28.0 * sol._fp.n
    dom = max(table, key=lambda r: r["ms_per_step"]) if table else None
    return {"bound": "tensor", "achieved": step_tflops, "peak": peak, "unit": "TFLOP/s", "frac": step_tflops / peak,
            "traffic": traffic, "algorithmic_bytes": alg_bytes,
            "definition": f"R x {fpr / 1e6:.3f} MFLOP x iters/s (R = {R} rows; 3xTF32 counts algorithmic FLOPs once) / "
                          "sustained TF32 peak measured in this process",
            "peak_source": (tf32_peak or {}).get("how", pk["source"] + ": 0.5 x bf16 sustained"),
            "tf32_peak_burst_tflops": (tf32_peak or {}).get("burst_tflops"),
            "hbm": {"traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                    "hbm_peak_gbs": pk["hbm"], "peak_source": pk["source"],
                    "traffic_gbs": (traffic / (ms_per_step * 1e-3) / 1e9) if traffic else None},
            "kernel": ("chaint_kernel (layer-chained tcgen05 sweeps: operand of the next MMA written to TMEM by the fused "
                       "epilogue, TMA-fed row-array I/O ring) + gemm_tc2g_kernel (weight gradients, cta_group::2, A operand "
                       "in TMEM)") if any(r["sweep"].endswith("*") for r in table)
            else "gemm_tc*_kernel (tcgen05, one launch per dense layer)" if any(r["tensor_cores"] for r in table)
            else "gemm_simt_kernel (fp32 FMA)",
            "dominant_launch": dom, "launch_table": table, "dense_ms_per_step": dense_ms,
            "dense_share_of_step": dense_ms / ms_per_step}


def measure_e2e_host(c, args, sol, batches):
    """The public train() API with the minibatch copied from pinned host memory every step (double-buffered on a
    copy stream, every byte crosses PCIe inside the timed region) and the per-step losses read back at the end."""
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
            "api": "train(K, lr) with host-supplied Brownian minibatches (pinned, double-buffered H2D on a copy "
                   "stream), losses and Y0 read back"}


def measure_e2e_product(c, args, brownian, steps, warmup):
    """Un-patched public API: Solver(..., brownian=...).train(K, lr) -- "philox" draws the increments in-kernel (no host
    RNG, no H2D); "numpy" is the reference-compatible default (host MT19937 stream, fp64 cumsum, H2D every step)."""
    sol, _ = build_solver(c, args.workload, args.dim, args.act, args.paths, args.precision, args.collective,
                          brownian=brownian)
    with contextlib.redirect_stdout(io.StringIO()):
        if warmup:
            sol.train(warmup, 1e-3)
        c.barrier()
        t0 = time.perf_counter()
        sol.train(steps, 1e-3)
        c.barrier()
        sec = c.max_over_ranks(time.perf_counter() - t0)
    del sol
    torch.cuda.empty_cache()
    return steps / sec


def measure_mc(c, args):
    import ctypes
    pde = c.pde
    D = 100
    np.random.seed(0)
    model = pde.BlackScholesModel(0.05, 0.2, D, True)
    pr = pde.MonteCarloPricer(model, pde.BasketOption(np.ones(D) / D, 1.0), 1.0, NSTEPS, args.mc_paths, seed=7,
                              data_parallel=True)
    n_lo, n_hi = c.parallel.shard_range(args.mc_paths, c.rank, c.world)
    warm = pr.price_async(np.ones(D), max(1, (n_hi - n_lo) // 64), n_lo, 7)      # warm-up (+ NCCL communicator)
    c.parallel.allreduce_sums(warm)
    c.barrier()
    m0 = c.lib.mc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sums = pr.price_async(np.ones(D), n_hi - n_lo, n_lo, 7)
    c.parallel.allreduce_sums(sums)           # the scalar all-reduce is inside the timed region
    e1.record()
    c.barrier()
    mc_ms = c.max_over_ranks(e0.elapsed_time(e1))
    launches = int(c.lib.mc_launch_count() - m0)
    s, q = (float(v) for v in sums.cpu())
    mean = s / args.mc_paths
    se = float(np.sqrt(max(q / args.mc_paths - mean * mean, 0.0) / args.mc_paths))
    pps = args.mc_paths / (mc_ms * 1e-3)
    # roofline denominator: the generator alone on the same GPU(s)
    scratch = torch.empty(c.lib.mc_scratch_bytes(), dtype=torch.uint8, device=c.dev)
    drawn = ctypes.c_uint64(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    c.lib.mc_normal_rate_probe(1 << 30, 7, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(drawn), st)
    torch.cuda.synchronize()
    e0.record()
    rc = c.lib.mc_normal_rate_probe(1 << 35, 7, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(drawn), st)
    e1.record()
    torch.cuda.synchronize()
    probe = (drawn.value / (e0.elapsed_time(e1) * 1e-3)) * c.world if rc == 0 else None
    normals = pps * NSTEPS * D
    fp32_peak = 148 * 128 * 2 * 1.965e9 * c.world
    return {"metric": "MC basket paths/s", "value": pps, "unit": "paths/s", "paths": args.mc_paths, "ms": mc_ms,
            "price": mean, "stderr": se, "normals_per_s": normals, "gpu_launches": launches,
            "config": {"workload": f"correlated GBM basket call, D={D}, N={NSTEPS}, {args.mc_paths} paths, "
                                   "Philox4x32-10 keyed by global path id; timed region includes the scalar all-reduce"},
            "roofline": {"bound": "issue", "achieved": normals, "peak": probe, "unit": "normals/s",
                         "frac": (normals / probe) if probe else None, "traffic": None,
                         "definition": "normals/s of the pricer / normals/s of the generator alone (Philox4x32-10 + MUFU "
                                       "Box-Muller, one add per normal: mc_normal_rate_probe) on the same GPU(s)",
                         "as_written_flops": {"tflops": pps * 2.0 * NSTEPS * D * D / 1e12,
                                              "fp32_simt_peak_tflops": fp32_peak / 1e12,
                                              "frac": pps * 2.0 * NSTEPS * D * D / fp32_peak,
                                              "note": "form (i) of SURVEY 8d: the reference's per-step matvec (2 N D^2 "
                                                      "FLOP per path); the kernel applies L once to the summed normals, "
                                                      "so this counts work it does not do"}},
            "form": "N*D Philox/Box-Muller normals per path, Cholesky matvec hoisted (L sum_t z_t)"}


def small_point(c, args, paths, steps, warmup, note):
    a2 = argparse.Namespace(**vars(args))
    a2.paths = paths
    sol, w = build_solver(c, args.workload, args.dim, args.act, paths, args.precision, args.collective)
    lo, hi = c.parallel.shard_range(paths, c.rank, c.world)
    batches = make_batches(c, sol, hi - lo, lo, args.dim) if hi > lo else None
    ms, launches, _, _ = time_steps(c, sol, batches, steps, warmup)
    R = paths * (NSTEPS + 1)
    out = {"metric": METRIC, "value": 1e3 / ms, "unit": "iters/s", "ms_per_step": ms, "paths": paths, "steps": steps,
           "warmup": warmup, "precision": args.precision, "gpu_launches": int(launches),
           "step_tflops": flop_per_row(w["net"], args.dim) * R / (ms * 1e-3) / 1e12, "note": note}
    del sol, batches
    torch.cuda.empty_cache()
    return out


def measure_workloads(c, args, tf32_peak):
    """BASELINE.json configs 3-4: NAIS-Net basket D = 5/10/50/100 (Sine, ReLU), correlated 100-D, HJB-100D; at the
    reference's M (100; 16 for HJB) and at a large M.  Brownian increments are drawn in-kernel inside every step."""
    peak = tf32_peak["sustained_tflops"] if tf32_peak else peaks()["bf16_sustained"] * 0.5
    cases = [("basket_nais", d, act) for d in (5, 10, 50, 100) for act in ("Sine", "ReLU")]
    cases += [("corr_nais", 100, "Sine"), ("hjb", 100, "ReLU")]
    out = []
    for kind, dim, act in cases:
        for paths, steps, warm in ((16 if kind == "hjb" else 100, 50, 10), (args.workload_paths, max(3, args.steps // 2), 3)):
            if paths < c.world:
                continue
            try:
                sol, w = build_solver(c, kind, dim, act, paths, args.precision, args.collective, brownian="philox")
                ms, launches, _, _ = time_steps(c, sol, None, steps, warm)
                R = paths * (NSTEPS + 1)
                tfl = flop_per_row(w["net"], dim) * R / (ms * 1e-3) / 1e12
                out.append({"workload": w["label"], "kind": kind, "dim": dim, "act": act, "paths": paths,
                            "iters_per_s": 1e3 / ms, "ms_per_step": ms, "gpu_launches": int(launches),
                            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peak, "unit": "TFLOP/s",
                                         "frac": tfl / peak, "mflop_per_row": flop_per_row(w["net"], dim) / 1e6}})
                del sol
            except Exception as e:   # noqa: BLE001 -- keep the headline line even if one side workload fails
                out.append({"kind": kind, "dim": dim, "act": act, "paths": paths, "error": repr(e)[:200]})
            torch.cuda.empty_cache()
    return out


def run_ours(args):
    c = setup(args)
    tf32_peak = None
    if not args.skip_peak:
        tf32_peak = measure_tf32_peak(c.dev)
    sol, w = build_solver(c, args.workload, args.dim, args.act, args.paths, args.precision, args.collective)
    lo, hi = c.parallel.shard_range(args.paths, c.rank, c.world)
    m_loc = hi - lo
    batches = make_batches(c, sol, m_loc, lo, args.dim)
    ms_per_step, launches, clk, final_loss = time_steps(c, sol, batches, args.steps, args.warmup, sample_clocks=True)
    line = {
        "metric": METRIC if args.workload == "bsb" and args.dim == 100 else f"{w['label']} train iters/s",
        "value": 1e3 / ms_per_step, "unit": "iters/s", "n_gpus": c.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "data": "synthetic",
        "dtype": {"tf32": "tf32", "tf32x3": "tf32x3 (fp32-grade: hi/lo-split TF32 MMAs, fp32 accumulate)",
                  "fp32": "f32"}[args.precision],
        "config": workload_config(args), "clocks": clk, "gpu_launches": launches, "final_loss": final_loss,
    }
    if c.world > 1:
        line["collective"] = ("fbsnn_peer_allreduce_adam (P2P loads over NVLink, fused with clip norm + Adam; the whole "
                              "iteration replays from one CUDA graph)" if sol.collective == "peer"
                              else "NCCL all-reduce of [grad | loss]")
    line["roofline"] = measure_roofline(c, args, sol, batches, ms_per_step, m_loc, w["net"], tf32_peak)
    if not args.skip_e2e:
        line["e2e"] = measure_e2e_host(c, args, sol, batches)
    del sol
    torch.cuda.empty_cache()
    if not args.skip_e2e:
        # the un-patched product path: increments drawn in-kernel, nothing crosses PCIe but the logged scalars
        v = measure_e2e_product(c, args, "philox", args.steps, max(1, args.warmup))
        line["e2e_philox"] = {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                              "api": "Solver(..., brownian='philox').train(K, lr), un-patched"}
        if c.world == 1 and not args.skip_numpy:
            v = measure_e2e_product(c, args, "numpy", 1, 0)
            line["e2e_numpy_default"] = {
                "value": v, "unit": "iters/s", "h2d_bytes_per_step": int(args.paths * (NSTEPS + 1) * (args.dim + 1) * 4),
                "note": "class default brownian='numpy' reproduces the reference's MT19937 stream on the host: at this "
                        "batch the iteration is bound by np.random.normal + fp64 cumsum of M*N*D doubles, not by the GPU; "
                        "one measured iteration"}
    if not args.skip_variants:
        notes = {"tf32": "tcgen05 kind::tf32 single pass; tolerance TOL_TF32 (loss 1e-2, Z 5e-2)",
                 "tf32x3": "tcgen05 3xTF32 (hi/lo split), fp32-grade; tolerance TOL_X3 (loss 1e-4, Z 5e-5)",
                 "fp32": "gemm_simt_kernel fp32 FMA; tolerance TOL (loss 2e-5, Z 1e-5)"}
        line["variants"] = {args.precision: {"value": 1e3 / ms_per_step, "ms_per_step": ms_per_step,
                                             "note": notes[args.precision]}}
        for prec in ("tf32x3", "tf32", "fp32"):
            if prec == args.precision:
                continue
            steps2 = max(2, args.steps // 3) if prec == "fp32" else args.steps
            sol2, _ = build_solver(c, args.workload, args.dim, args.act, args.paths, prec, args.collective)
            ms2, _, _, _ = time_steps(c, sol2, batches, steps2, args.warmup)
            R = args.paths * (NSTEPS + 1)
            line["variants"][prec] = {"value": 1e3 / ms2, "ms_per_step": ms2, "steps": steps2, "note": notes[prec],
                                      "step_tflops": flop_per_row(w["net"], args.dim) * R / (ms2 * 1e-3) / 1e12}
            del sol2
            torch.cuda.empty_cache()
    del batches
    torch.cuda.empty_cache()
    if not args.skip_small and args.paths > 4096:
        line["small_m"] = small_point(c, args, 100, 200, 20,
                                      "M=100, the configuration the reference ships (DeepBSDE.py:432): 5 100 rows, L2-resident, "
                                      "launch-latency bound; one CUDA-graph replay per iteration")
        line["mid_m"] = small_point(c, args, 4096, 30, 5, "M=4096: 209k rows")
    if not args.skip_workloads:
        line["workloads"] = measure_workloads(c, args, tf32_peak)
    if not args.skip_mc:
        line["mc"] = measure_mc(c, args)
    if c.rank == 0 and c.world == 1 and not args.skip_cpu:
        line["cpu_baseline"] = cpu_baseline_block(args, 5, 1, suite=args.cpu_suite)
        if not args.skip_mc:
            r = ref_subprocess(["mc", "--paths", "20000"])
            if r.get("available") and "paths_per_s" in r:
                line["mc"]["cpu_baseline"] = {"value": r["paths_per_s"], "unit": "paths/s", "cores": os.cpu_count(),
                                              "kind": "reference", "sample": "20000 paths, D=100, N=50: "
                                              "MonteCarloPricer.price of the unmodified reference (NumPy)"}
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
    ap.add_argument("--workload", default="bsb", choices=["bsb", "basket_nais", "corr_nais", "hjb"])
    ap.add_argument("--dim", type=int, default=100)
    ap.add_argument("--act", default=None, help="activation (default: Sine; ReLU for hjb)")
    ap.add_argument("--paths", type=int, default=65536, help="global number of Brownian paths M")
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32", "tf32x3"],
                    help="tf32x3 (default): tcgen05 tensor cores with hi/lo operand split, fp32-grade results; "
                         "tf32: single-pass tcgen05 (looser stated tolerance); fp32: SIMT FMA (tightest parity)")
    ap.add_argument("--mc-paths", type=int, default=10 ** 9)
    ap.add_argument("--workload-paths", type=int, default=16384, help="large-M point of the `workloads` sub-objects")
    ap.add_argument("--cpu-sample-paths", type=int, default=1000,
                    help="batch size of the reference arm's measured iterations (bounded sample of the workload)")
    ap.add_argument("--cpu-suite", action="store_true", help="also time the BASELINE.md section 3 table (minutes)")
    ap.add_argument("--cpu-suite-budget", type=float, default=60.0)
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU gradient all-reduce: fused NVLink peer-memory kernel (default) or NCCL")
    for flag in ("mc", "cpu", "small", "variants", "e2e", "workloads", "numpy", "peak"):
        ap.add_argument(f"--skip-{flag}", action="store_true")
    ap.add_argument("--skip-fp32", action="store_true", help="(alias of --skip-variants)")
    args = ap.parse_args()
    if args.act is None:
        args.act = "ReLU" if args.workload == "hjb" else "Sine"
    args.skip_variants = args.skip_variants or args.skip_fp32
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
