"""Times the UNMODIFIED reference (baseline/_ref, see install_ref.py) on the host cores: the CPU baseline of bench.py.

Runs as its own process with CUDA hidden (the reference would otherwise move itself to cuda:0, DeepBSDE.py:143-149):

    python baseline/ref_bench.py train --paths 1000 --steps 10 --warmup 1 [--anomaly]     # DeepBSDE.train()
    python baseline/ref_bench.py suite
    python baseline/ref_bench.py mc --paths 200000

Each mode prints ONE JSON object.  `train` goes through the reference's own public API -- BlackScholesBarenblatt(...)
.train(K, lr) of DeepBSDE.py:265-295 -- on the BSB-100D FC-Sine 4x256 network; the NAIS-Net / HJB rows of `suite`
replay the reference's own iteration body (zero_grad, fetch_minibatch, loss_function, backward, clip, Adam.step:
with_corr_high_dimension_pde.py:412-425) because its train() rewrites N to 2-3 steps or
raises (Q4).
"""
import os

os.environ["CUDA_VISIBLE_DEVICES"] = ""          # before torch: the reference must run on the host cores

import argparse
import contextlib
import importlib.util
import io
import json
import sys
import time
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
D, N, H = 100, 50, 256


def available() -> bool:
    return os.path.exists(os.path.join(REF, "DeepBSDE.py"))


def load(fname):
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.gridspec", "seaborn", "mpl_toolkits",
              "mpl_toolkits.mplot3d"]:
        sys.modules.setdefault(m, MagicMock())
    name = "baseline_ref_" + os.path.basename(fname).replace(".", "_")
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, fname))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


def host_info():
    model = ""
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cpu_model": model}


def bsb_train(paths, steps, warmup, anomaly):
    """seconds per iteration of DeepBSDE.BlackScholesBarenblatt.train() (median over `steps` single-iteration calls
    would restart Adam each time; the reference API is one train(K) call, so time train(K) and divide)."""
    mod = load("DeepBSDE.py")
    torch.autograd.set_detect_anomaly(bool(anomaly))   # DeepBSDE.py:11 switches it on at import ("as shipped")
    torch.manual_seed(1234)
    np.random.seed(1234)
    Xi = np.array([1.0, 0.5] * (D // 2))[None, :]
    model = mod.BlackScholesBarenblatt(Xi, 1.0, paths, N, D, [D + 1] + 4 * [H] + [1], "FC", "Sine")
    assert model.device.type == "cpu"
    with contextlib.redirect_stdout(io.StringIO()):
        if warmup > 0:
            model.train(warmup, 1e-3)
        t0 = time.perf_counter()
        model.train(steps, 1e-3)
        sec = (time.perf_counter() - t0) / steps
    return sec


def body_rate(fname, cls, ctor, iters, warmup, clip=1.0):
    """seconds per iteration of the reference's own iteration body for the classes whose train() cannot be used."""
    import torch.optim as optim
    mod = load(fname)
    torch.autograd.set_detect_anomaly(False)
    torch.manual_seed(1234)
    np.random.seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ctor(getattr(mod, cls))
        model.N = N                      # fixed N = 50 for throughput (BASELINE.json configs), not the scheduled 2-3
        opt = optim.Adam(model.model.parameters(), lr=1e-3)
        times = []
        for i in range(warmup + iters):
            t0 = time.perf_counter()
            opt.zero_grad()
            tb, Wb = model.fetch_minibatch()
            loss = model.loss_function(tb, Wb, model.Xi)[0]
            loss.backward()
            if clip:
                torch.nn.utils.clip_grad_norm_(model.model.parameters(), max_norm=clip)
            opt.step()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.median(times))


def basket_ctor(dim, act, corr="no_correlation"):
    layers = [dim + 1] + 4 * [H] + [1]
    return lambda cls: cls(np.ones((1, dim)), 1.0, 100, N, dim, None, layers, "Naisnet", act, corr)


def hjb_ctor(dim, paths):
    layers = [dim + 1] + 4 * [H] + [1]
    return lambda cls: cls(np.zeros((1, dim)), 1.0, paths, N, dim, layers, "Naisnet", "ReLU")


def mc_rate(paths):
    mod = load("numerics/multidimensional_mc_pricer.py")
    np.random.seed(0)
    model = mod.BlackScholesModel(0.05, 0.2, D, True)
    option = mod.BasketOption(np.ones(D) / D, 1.0)
    pricer = mod.MonteCarloPricer(model, option, 1.0, N, paths)
    t0 = time.perf_counter()
    price = pricer.price(np.ones(D))
    sec = time.perf_counter() - t0
    return paths / sec, float(price), sec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["train", "suite", "mc"])
    ap.add_argument("--paths", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--anomaly", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget", type=float, default=60.0, help="suite: rough wall-clock budget in seconds")
    args = ap.parse_args()
    if args.threads > 0:
        torch.set_num_threads(args.threads)
    out = {"available": available(), **host_info()}
    if not available():
        print(json.dumps(out))
        return
    if args.mode == "train":
        sec = bsb_train(args.paths, args.steps, args.warmup, args.anomaly)
        out.update({"paths": args.paths, "steps": args.steps, "warmup": args.warmup, "anomaly": bool(args.anomaly),
                    "sec_per_iter": sec, "iters_per_s": 1.0 / sec, "paths_per_s": args.paths / sec,
                    "api": "DeepBSDE.BlackScholesBarenblatt.train(K, 1e-3), unmodified, CPU"})
    elif args.mode == "mc":
        pps, price, sec = mc_rate(args.paths)
        out.update({"paths": args.paths, "paths_per_s": pps, "price": price, "sec": sec,
                    "api": "numerics/multidimensional_mc_pricer.MonteCarloPricer.price, unmodified (single NumPy thread + BLAS)"})
    else:
        t_start = time.perf_counter()
        rows = {}
        for paths, steps in ((100, 10), (1000, 4), (4096, 2)):
            for anomaly in (False, True):
                if time.perf_counter() - t_start > args.budget and rows:
                    continue
                try:
                    sec = bsb_train(paths, steps if not anomaly else max(2, steps // 2), 1, anomaly)
                    rows[f"bsb_fc_sine_M{paths}_anomaly_{'on' if anomaly else 'off'}"] = {
                        "iters_per_s": 1.0 / sec, "sec_per_iter": sec, "paths": paths}
                except Exception as e:   # noqa: BLE001 -- e.g. out of host memory at M = 4096
                    rows[f"bsb_fc_sine_M{paths}_anomaly_{'on' if anomaly else 'off'}"] = {"error": repr(e)[:200]}
        for dim in (5, 10, 50, 100):
            for act in ("Sine", "ReLU"):
                if time.perf_counter() - t_start > 2 * args.budget:
                    continue
                sec = body_rate("with_corr_high_dimension_pde.py", "CallOption", basket_ctor(dim, act), 3, 1)
                rows[f"basket_nais_{act.lower()}_D{dim}_M100"] = {"iters_per_s": 1.0 / sec, "sec_per_iter": sec}
        if time.perf_counter() - t_start <= 3 * args.budget:
            sec = body_rate("with_corr_high_dimension_pde.py", "CallOption",
                            basket_ctor(100, "Sine", "random_correlation"), 3, 1)
            rows["corr100_nais_sine_M100"] = {"iters_per_s": 1.0 / sec, "sec_per_iter": sec}
            sec = body_rate("hjb_implement.py", "HamiltonJacobiBellman", hjb_ctor(100, 16), 3, 1)
            rows["hjb100_nais_relu_M16"] = {"iters_per_s": 1.0 / sec, "sec_per_iter": sec}
        out["rows"] = rows
        out["note"] = ("median seconds per iteration; BSB rows through DeepBSDE.train(); NAIS / HJB rows replay the "
                       "reference's iteration body at fixed N = 50 (its train() rewrites N or raises)")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
