"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.npz

Inputs are reproduced from seeds (torch.manual_seed -> network init; np.random.seed -> correlation matrix and
Brownian increments), so a fixture stores only the seeds, checksums of the regenerated inputs (so a drift of
either RNG stream is detected loudly) and the reference's outputs.  The reference's `train()` is not used
for the K-step traces because it silently rewrites N or crashes (Q4); the trace replays the
reference's own iteration body -- zero_grad / fetch_minibatch / loss_function / backward / [clip] / Adam.step --
with the reference's own methods (DeepBSDE.py:274-280, with_corr_high_dimension_pde.py:412-425).
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import json
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference(fname: str):
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.gridspec", "seaborn",
              "mpl_toolkits", "mpl_toolkits.mplot3d"]:
        sys.modules.setdefault(m, MagicMock())
    name = "ref_" + os.path.basename(fname).replace(".", "_")
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, fname))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    torch.autograd.set_detect_anomaly(False)   # DeepBSDE.py:11 switches it on at import; numerics are unaffected
    return mod


CASES = [
    # name, file, class, ctor-arity, problem key (oracle), D, M, N, H, n_layers_total, mode, act, Xi, corr, K, lr, clip
    dict(name="bsb100_fc_sine", file="DeepBSDE.py", cls="BlackScholesBarenblatt", arity="short", problem="bsb",
         D=100, M=100, N=50, layers=[101, 256, 256, 256, 256, 1], mode="FC", act="Sine", xi="bsb",
         corr="no_correlation", K=5, lr=1e-3, clip=None),
    dict(name="bsb10_fc_relu", file="DeepBSDE.py", cls="BlackScholesBarenblatt", arity="short", problem="bsb",
         D=10, M=24, N=12, layers=[11, 64, 64, 64, 1], mode="FC", act="ReLU", xi="bsb",
         corr="no_correlation", K=4, lr=1e-3, clip=None),
    dict(name="call1d_fc_sine_m1", file="1d_BSPDE_case.py", cls="CallOption", arity="mm", problem="call1d",
         D=1, M=1, N=50, layers=[2, 256, 256, 256, 256, 1], mode="FC", act="Sine", xi="ones",
         corr=None, K=4, lr=1e-3, clip=1.0),
    dict(name="call1d_fc_sine_m100_quirk", file="1d_BSPDE_case.py", cls="CallOption", arity="mm", problem="call1d",
         D=1, M=100, N=50, layers=[2, 256, 256, 256, 256, 1], mode="FC", act="Sine", xi="ones",
         corr=None, K=2, lr=1e-3, clip=1.0),
    dict(name="callnd100_fc_tanh", file="nd_BSPDE_case.py", cls="CallOption", arity="mm", problem="callnd",
         D=100, M=32, N=20, layers=[101, 128, 128, 128, 1], mode="FC", act="Tanh", xi="ones",
         corr=None, K=3, lr=1e-3, clip=1.0),
    dict(name="basket5_nais_sine", file="with_corr_high_dimension_pde.py", cls="CallOption", arity="corr",
         problem="basket", D=5, M=100, N=50, layers=[6, 256, 256, 256, 256, 1], mode="Naisnet", act="Sine",
         xi="ones", corr="no_correlation", K=4, lr=1e-3, clip=1.0),
    dict(name="basket100_nais_relu_corr", file="with_corr_high_dimension_pde.py", cls="CallOption", arity="corr",
         problem="basket", D=100, M=64, N=50, layers=[101, 256, 256, 256, 256, 1], mode="Naisnet", act="ReLU",
         xi="ones", corr="random_correlation", K=4, lr=1e-3, clip=1.0),
    dict(name="basket10_nais_sine_5l", file="with_corr_high_dimension_pde.py", cls="CallOption", arity="corr",
         problem="basket", D=10, M=32, N=16, layers=[11, 64, 64, 64, 1], mode="Naisnet", act="Sine",
         xi="ones", corr="restricted_random_correlation", K=3, lr=1e-3, clip=1.0),
    dict(name="bsptest50_fc_sine", file="with_corr_high_dimension_pde.py", cls="BSPDETestCase", arity="corr",
         problem="bsptest", D=50, M=32, N=25, layers=[51, 128, 128, 1], mode="FC", act="Sine",
         xi="ones", corr="no_correlation", K=3, lr=1e-3, clip=1.0),
    dict(name="hjb100_nais_relu", file="hjb_implement.py", cls="HamiltonJacobiBellman", arity="short",
         problem="hjb", D=100, M=16, N=50, layers=[101, 256, 256, 256, 256, 1], mode="Naisnet", act="ReLU",
         xi="zeros", corr=None, K=4, lr=1e-3, clip=1.0),
    dict(name="hjb20_nais_tanh_4l", file="hjb_implement.py", cls="HamiltonJacobiBellman", arity="short",
         problem="hjb", D=20, M=16, N=10, layers=[21, 64, 64, 1], mode="Naisnet", act="Tanh",
         xi="zeros", corr=None, K=3, lr=1e-3, clip=1.0),
    # Heston 2-factor FBSNN (heston_dnnpde.py:519-659): D = 1 Brownian driver, states (S, v), clip 1.0
    dict(name="heston_fc_sine", file="heston_dnnpde.py", cls="HestonFBSNN", arity="heston", problem="heston",
         D=1, M=48, N=20, layers=[2, 64, 64, 64, 1], mode="FC", act="Sine", xi="ones", corr=None, K=4, lr=1e-3,
         clip=1.0, payoff="discontinuous"),
    dict(name="heston_nais_tanh_smooth", file="heston_dnnpde.py", cls="HestonFBSNN", arity="heston", problem="heston",
         D=1, M=16, N=12, layers=[2, 32, 32, 32, 1], mode="Naisnet", act="Tanh", xi="ones", corr=None, K=3, lr=1e-3,
         clip=1.0, payoff="continuous"),
]

TORCH_SEED = 1234
NUMPY_SEED = 4321


def make_xi(kind, D):
    if kind == "bsb":
        base = np.array([1.0, 0.5] * (D // 2) + [1.0] * (D % 2))
        return base[None, :]
    if kind == "ones":
        return np.ones((1, D))
    if kind == "zeros":
        return np.zeros((1, D))
    raise ValueError(kind)


def param_checksums(model):
    sums, asums = [], []
    for _, p in model.named_parameters():
        d = p.detach().double()
        sums.append(float(d.sum()))
        asums.append(float(d.abs().sum()))
    return np.array(sums), np.array(asums)


def run_case(c):
    mod = load_reference(c["file"])
    cls = getattr(mod, c["cls"])
    Xi = make_xi(c["xi"], c["D"])
    torch.manual_seed(TORCH_SEED)
    np.random.seed(NUMPY_SEED)
    T = 1.0
    with contextlib.redirect_stdout(io.StringIO()):
        if c["arity"] == "short":
            model = cls(Xi, T, c["M"], c["N"], c["D"], c["layers"], c["mode"], c["act"])
        elif c["arity"] == "mm":
            model = cls(Xi, T, c["M"], c["N"], c["D"], None, c["layers"], c["mode"], c["act"])
        elif c["arity"] == "heston":
            model = cls(Xi, T, c["M"], c["N"], c["D"], int(c["N"] ** (1 / 5)), c["layers"], c["mode"], c["act"],
                        payoff_type=c["payoff"])
        else:
            model = cls(Xi, T, c["M"], c["N"], c["D"], None, c["layers"], c["mode"], c["act"], c["corr"])
    out = {}
    names = [k for k, _ in model.model.named_parameters()]
    out["param_names"] = np.array(names)
    out["param_sum"], out["param_abssum"] = param_checksums(model.model)
    corr = getattr(model, "correlation_matrix", None)
    if corr is not None and c["corr"] not in (None, "no_correlation"):
        out["corr_matrix"] = np.asarray(corr)

    # ---- single evaluation: loss, trajectories, parameter gradients ------------------------------------
    captured = []
    orig_net_u = model.net_u

    def spy(t, X):
        res = orig_net_u(t, X)
        if len(res) == 3:                        # Heston: (u, dU/dS, dU/dv)
            captured.append(torch.cat([res[1], res[2]], dim=1).detach().clone())
        else:
            captured.append(res[1].detach().clone())
        return res

    with contextlib.redirect_stdout(io.StringIO()):
        t_b, W_b = model.fetch_minibatch()
        out["W_sum"] = np.array([float(W_b.double().sum()), float(W_b.double().abs().sum())])
        model.net_u = spy
        loss, X, Y, Y0 = model.loss_function(t_b, W_b, model.Xi)
        model.net_u = orig_net_u
        model.model.zero_grad()
        loss.backward()
    Z = torch.stack(captured, dim=1)
    out["loss"] = np.float32(loss.detach().numpy())
    out["Y0"] = np.float32(float(Y0))
    out["Y"] = Y.detach().numpy()[:, :, 0]
    keep = min(4, c["M"])
    out["X_head"] = X.detach().numpy()[:keep]
    out["Z_head"] = Z.numpy()[:keep]
    out["X_absmean"] = np.float64(X.detach().double().abs().mean())
    out["Z_norm"] = np.float64(Z.double().norm())
    gn, gs = [], []
    for k, p in model.model.named_parameters():
        g = p.grad.detach().double()
        gn.append(float(g.norm()))
        gs.append(float(g.sum()))
        if g.numel() <= 30000:
            out["grad::" + k] = p.grad.detach().numpy().copy()
    out["grad_norm"] = np.array(gn)
    out["grad_sum"] = np.array(gs)

    # ---- K Adam iterations on fresh batches (continuing the NumPy stream) --------------------------------
    opt = torch.optim.Adam(model.model.parameters(), lr=c["lr"])
    tl, ty = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(c["K"]):
            opt.zero_grad()
            t_b, W_b = model.fetch_minibatch()
            loss, X, Y, Y0 = model.loss_function(t_b, W_b, model.Xi)
            loss.backward()
            if c["clip"] is not None:
                torch.nn.utils.clip_grad_norm_(model.model.parameters(), max_norm=c["clip"])
            opt.step()
            tl.append(float(loss.detach()))
            ty.append(float(Y0))
    out["trace_loss"] = np.array(tl, dtype=np.float64)
    out["trace_Y0"] = np.array(ty, dtype=np.float64)
    out["final_param_sum"], out["final_param_abssum"] = param_checksums(model.model)
    last = names[-2]
    out["final::" + last] = dict(model.model.named_parameters())[last].detach().numpy().copy()
    meta = dict(c)
    meta.update(torch_seed=TORCH_SEED, numpy_seed=NUMPY_SEED, T=T, torch=torch.__version__, numpy=np.__version__)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **out)
    print(f"{c['name']:32s} loss={float(out['loss']):.6e} Y0={float(out['Y0']):.6f} trace={tl}")


def run_train_api_case():
    """DeepBSDE.FBSNN.train() itself (no schedule in that file): returned graph after 3 iterations."""
    mod = load_reference("DeepBSDE.py")
    D, M, N = 100, 100, 50
    torch.manual_seed(TORCH_SEED)
    np.random.seed(NUMPY_SEED)
    model = mod.BlackScholesBarenblatt(make_xi("bsb", D), 1.0, M, N, D, [D + 1] + 4 * [256] + [1], "FC", "Sine")
    with contextlib.redirect_stdout(io.StringIO()):
        graph = model.train(3, 1e-3)
        np.random.seed(42)
        t_test, W_test = model.fetch_minibatch()
        X_pred, Y_pred = model.predict(make_xi("bsb", D), t_test, W_test)
    np.savez_compressed(os.path.join(OUT, "bsb100_train_api.npz"), graph=graph,
                        Y_pred=Y_pred.detach().numpy()[:, :, 0], X_pred_head=X_pred.detach().numpy()[:2],
                        meta=np.array(json.dumps(dict(torch_seed=TORCH_SEED, numpy_seed=NUMPY_SEED, D=D, M=M, N=N))))
    print("bsb100_train_api graph", graph.tolist())


def run_mc_cases():
    mod = load_reference("numerics/multidimensional_mc_pricer.py")
    out = {}
    for tag, D, n, corr in [("d5", 5, 20000, True), ("d100", 100, 4000, True), ("d8_nocorr", 8, 5000, False)]:
        np.random.seed(0)
        model = mod.BlackScholesModel(0.05, 0.20, D, corr)
        option = mod.BasketOption(np.ones(D) / D, 1.0)
        pricer = mod.MonteCarloPricer(model, option, 1.0, 50, n)
        price = pricer.price(np.ones(D))
        out[f"{tag}_price"] = np.float64(price)
        out[f"{tag}_corr"] = model.correlation
        out[f"{tag}_analytic"] = np.float64(mod.AnalyticalBlackScholes(0.05, 0.20, D).price(np.ones(D), 1.0, 1.0))
        out[f"{tag}_cfg"] = np.array(json.dumps(dict(D=D, n=n, corr=corr, rate=0.05, sigma=0.2, T=1.0, N=50, seed=0)))
        print(f"mc {tag}: price={price:.6f}")
    np.savez_compressed(os.path.join(OUT, "mc_pricer.npz"), **out)


def run_basket_pricer_case():
    """basket_pricer.py: MonteCarloSimulator.simulate + BasketOptionPricer.price under a fixed NumPy seed."""
    sys.modules.setdefault("sklearn", MagicMock()), sys.modules.setdefault("sklearn.decomposition", MagicMock())
    mod = load_reference("basket_pricer.py")
    S0 = np.linspace(0.9, 1.1, 5)
    corr = np.full((5, 5), 0.25) + 0.75 * np.eye(5)
    np.random.seed(12)
    sim = mod.MonteCarloSimulator(S0, 0.05, 0.2, 1.0, 0.1, corr.copy())
    paths = sim.simulate(3000)
    price = mod.BasketOptionPricer(1.0, 1.0, corr.copy()).price(paths, 0.05)
    np.savez_compressed(os.path.join(OUT, "basket_pricer.npz"), S0=S0, corr=corr, price=np.float64(price),
                        paths_shape=np.array(paths.shape), paths_sum=np.float64(paths.sum()),
                        paths_terminal_head=paths[:, -1, :8],
                        cfg=np.array(json.dumps(dict(seed=12, r=0.05, sigma=0.2, T=1.0, dt=0.1, n=3000, strike=1.0))))
    print(f"basket_pricer: price={price:.6f} shape={paths.shape}")


def run_schedule_case():
    """nd_BSPDE_case.CallOption.train() ITSELF with its N-schedule crossing iteration 4000: the
    iteration counter is started at 3996 through the class's own resume mechanism (`previous_it = self.iteration[-1]`,
    nd_BSPDE_case.py:326-328), so 8 iterations run at N = ceil(Mm) = 3, 3, 3, 3 and then N = ceil(Mm^2) = 5, 5, 5, 5.
    Per-iteration losses are recorded by wrapping the instance's loss_function (the source is not edited)."""
    mod = load_reference("nd_BSPDE_case.py")
    D, M, N, Mm = 8, 12, 50, 2.2
    layers = [D + 1, 64, 64, 1]
    torch.manual_seed(TORCH_SEED)
    np.random.seed(NUMPY_SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        model = mod.CallOption(make_xi("ones", D), 1.0, M, N, D, Mm, layers, "FC", "Sine")
        names = [k for k, _ in model.model.named_parameters()]
        psum, pabs = param_checksums(model.model)
        model.iteration, model.training_loss = [3996], [0.0]
        trace, Ns = [], []
        orig = model.loss_function

        def spy(t, W, Xi):
            res = orig(t, W, Xi)
            trace.append(float(res[0].detach()))
            Ns.append(int(W.shape[1] - 1))
            return res

        model.loss_function = spy
        graph, min_loss, min_state = model.train(8, 1e-3)
    last = names[-2]
    np.savez_compressed(
        os.path.join(OUT, "nd_schedule_trace.npz"), graph=np.asarray(graph, dtype=np.float64), trace_loss=np.array(trace),
        trace_N=np.array(Ns), min_loss=np.float64(min_loss), min_X=min_state[0].numpy(), min_Y=min_state[1].numpy(),
        param_names=np.array(names), param_sum=psum, param_abssum=pabs,
        **{"final::" + last: dict(model.model.named_parameters())[last].detach().numpy().copy()},
        meta=np.array(json.dumps(dict(torch_seed=TORCH_SEED, numpy_seed=NUMPY_SEED, D=D, M=M, N=N, Mm=Mm, layers=layers,
                                      start_it=3996, K=8, lr=1e-3))))
    print("nd_schedule_trace N per iteration", Ns, "losses", [f"{x:.5e}" for x in trace], "min", min_loss)


def run_k30_case():
    """BSB-100D, M = 100 (the shipped configuration): loss / Y0 trajectories over K = 30 Adam iterations, the horizon
    SURVEY section 8(c) states its tolerance bars for."""
    mod = load_reference("DeepBSDE.py")
    D, M, N, K = 100, 100, 50, 30
    torch.manual_seed(TORCH_SEED)
    np.random.seed(NUMPY_SEED)
    model = mod.BlackScholesBarenblatt(make_xi("bsb", D), 1.0, M, N, D, [D + 1] + 4 * [256] + [1], "FC", "Sine")
    names = [k for k, _ in model.model.named_parameters()]
    psum, pabs = param_checksums(model.model)
    opt = torch.optim.Adam(model.model.parameters(), lr=1e-3)
    tl, ty = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(K):
            opt.zero_grad()
            t_b, W_b = model.fetch_minibatch()
            loss, X, Y, Y0 = model.loss_function(t_b, W_b, model.Xi)
            loss.backward()
            opt.step()
            tl.append(float(loss.detach()))
            ty.append(float(Y0))
    last = names[-2]
    np.savez_compressed(
        os.path.join(OUT, "bsb100_k30.npz"), trace_loss=np.array(tl), trace_Y0=np.array(ty), param_names=np.array(names),
        param_sum=psum, param_abssum=pabs,
        **{"final::" + last: dict(model.model.named_parameters())[last].detach().numpy().copy()},
        meta=np.array(json.dumps(dict(torch_seed=TORCH_SEED, numpy_seed=NUMPY_SEED, D=D, M=M, N=N, K=K, lr=1e-3))))
    print("bsb100_k30 loss[0], loss[-1], Y0[-1]:", tl[0], tl[-1], ty[-1])


def run_comparator_case():
    """The two closed-form comparators of the drivers on a fixed random prediction tensor."""
    nd = load_reference("nd_BSPDE_case.py")
    wc = load_reference("with_corr_high_dimension_pde.py")
    rng = np.random.RandomState(7)
    B, NT, A = 6, 11, 9
    S = (0.6 + 0.8 * rng.rand(B, NT, A)).astype(np.float32)
    t = np.tile(np.linspace(0.0, 1.0, NT, dtype=np.float32)[None, :, None], (B, 1, 1))
    t[:, -1, :] = 0.95                                      # keep tau > 0 (tau = 0 with S == K is 0/0 upstream)
    price, delta = nd.BasketOptionPriceCalculator.calculate_option_prices(torch.from_numpy(S), torch.from_numpy(t), 1.0,
                                                                          0.05, 0.2, 1.0)
    Xavg = 0.7 + 0.6 * rng.rand(5, NT)
    Xavg[0, 3] = 1.0
    times = np.linspace(0.0, 1.0, NT)
    p2, d2 = wc.BasicOptionPriceCalculator().calculate_call_option_prices(Xavg, times, 1.0, 0.05, 0.2, 1.0, 25)
    np.savez_compressed(os.path.join(OUT, "comparators.npz"), S=S, t=t, nd_price=price.numpy(), nd_delta=delta.numpy(),
                        Xavg=Xavg, times=times, basic_price=p2, basic_delta=d2,
                        cfg=np.array(json.dumps(dict(K=1.0, r=0.05, sigma=0.2, T=1.0, dims=25))))
    print("comparators: nd price mean", float(price.mean()), "basic price mean", float(p2.mean()))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("make_golden.py needs /root/reference (build container only)")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for c in CASES:
        if not only or c["name"] in only:
            run_case(c)
    if not only or "train_api" in only:
        run_train_api_case()
    if not only or "mc" in only:
        run_mc_cases()
    if not only or "basket_pricer" in only:
        run_basket_pricer_case()
    if not only or "schedule" in only:
        run_schedule_case()
    if not only or "k30" in only:
        run_k30_case()
    if not only or "comparators" in only:
        run_comparator_case()
