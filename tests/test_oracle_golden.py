"""Pin the CPU oracle (oracle/*.py) against outputs of the unmodified reference (tests/golden/*.npz)."""
import json

import numpy as np
import pytest
import torch

from oracle import fbsnn_oracle as orc
from oracle import mc_oracle as mco
from tests import golden_util as gu


@pytest.mark.parametrize("name", gu.solver_cases())
def test_solver_oracle_reproduces_reference(name):
    g, meta = gu.load(name)
    torch.set_num_threads(8)
    sol = gu.rebuild_inputs(meta, g, squeeze_quirk=True)
    t, W = sol.fetch_minibatch()
    wsum = np.array([float(W.double().sum()), float(W.double().abs().sum())])
    assert np.allclose(wsum, g["W_sum"], rtol=1e-12), "NumPy RNG stream drifted (Brownian increments)"
    loss, X, Y, Z, grads = sol.grads(t, W)
    # same op sequence as the reference => agreement to fp32 round-off of the BLAS reduction order
    assert abs(float(loss) - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    assert np.allclose(Y[:, :, 0].numpy(), g["Y"], rtol=1e-5, atol=1e-6)
    k = g["X_head"].shape[0]
    assert np.allclose(X[:k].numpy(), g["X_head"], rtol=1e-6, atol=1e-7)
    assert np.allclose(Z[:k].numpy(), g["Z_head"], rtol=1e-4, atol=1e-6)
    gn = np.array([float(v.double().norm()) for v in grads.values()])
    assert np.allclose(gn, g["grad_norm"], rtol=1e-4, atol=1e-7)
    for kname, v in grads.items():
        key = "grad::" + kname
        if key in g.files:
            ref = g[key]
            assert np.abs(v.numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-7, kname
    # K optimiser iterations, continuing the NumPy stream exactly as the fixture generator did
    sol.make_optimizer(meta["lr"])
    tl, ty = [], []
    for _ in range(meta["K"]):
        t, W = sol.fetch_minibatch()
        l, y0 = sol.train_step(t, W, clip=meta["clip"])
        tl.append(l), ty.append(y0)
    assert np.allclose(tl, g["trace_loss"], rtol=2e-4), (tl, g["trace_loss"])
    assert np.allclose(ty, g["trace_Y0"], rtol=1e-3, atol=1e-4), (ty, g["trace_Y0"])


def test_squeeze_quirk_only_matters_for_d1():
    """the un-dimmed squeeze changes the loss only for D == 1 with M > 1."""
    g, meta = gu.load("call1d_fc_sine_m100_quirk")
    a = gu.rebuild_inputs(meta, g, squeeze_quirk=True)
    t, W = a.fetch_minibatch()
    la = float(a.loss_function(t, W)[0])
    a.squeeze_quirk = False
    lb = float(a.loss_function(t, W)[0])
    assert abs(la - float(g["loss"])) <= 2e-6 * abs(la)
    assert abs(la - lb) > 1e-2 * abs(la)
    g2, meta2 = gu.load("bsb10_fc_relu")
    b = gu.rebuild_inputs(meta2, g2, squeeze_quirk=True)
    t, W = b.fetch_minibatch()
    l1 = float(b.loss_function(t, W)[0])
    b.squeeze_quirk = False
    assert float(b.loss_function(t, W)[0]) == l1


def test_bsb_closed_form_terminal():
    """Known answer endorsed by the reference (DeepBSDE.py:345-349): u(T, x) = sum x^2, u(0, Xi) = 77.105."""
    Xi = gu.make_xi("bsb", 100)
    assert abs(float(orc.bsb_exact(0.0, Xi, 1.0)[0, 0]) - 77.1049) < 1e-3
    assert float(orc.bsb_exact(1.0, Xi, 1.0)[0, 0]) == 62.5


@pytest.mark.parametrize("tag", ["d5", "d100", "d8_nocorr"])
def test_mc_oracle_reproduces_reference(tag):
    g, _ = gu.load("mc_pricer")
    cfg = json.loads(str(g[f"{tag}_cfg"]))
    np.random.seed(cfg["seed"])
    D = cfg["D"]
    corr = mco.random_correlation(D, cfg["corr"])
    assert np.array_equal(corr, g[f"{tag}_corr"])
    price = mco.mc_price(np.ones(D), cfg["rate"], cfg["sigma"], corr, cfg["corr"], np.ones(D) / D, 1.0,
                         cfg["T"], cfg["N"], cfg["n"])
    assert price == float(g[f"{tag}_price"])          # same RNG stream, same arithmetic: bit-exact
    assert abs(mco.analytic_single_asset(np.ones(D), 1.0, cfg["rate"], cfg["sigma"], D, cfg["T"])
               - float(g[f"{tag}_analytic"])) < 1e-15


def test_basket_pricer_oracle_reproduces_reference():
    """oracle restatement of basket_pricer.py:41-66 vs the reference run under the same NumPy seed: bit-exact."""
    g, _ = gu.load("basket_pricer")
    cfg = json.loads(str(g["cfg"]))
    np.random.seed(cfg["seed"])
    paths = mco.simulate_asset_paths(g["S0"], cfg["r"], cfg["sigma"], cfg["T"], cfg["dt"], g["corr"], cfg["n"])
    assert tuple(paths.shape) == tuple(g["paths_shape"])
    assert np.array_equal(paths[:, -1, :8], g["paths_terminal_head"]) and paths.sum() == float(g["paths_sum"])
    assert mco.mean_basket_price(paths, cfg["r"], cfg["strike"], cfg["T"]) == float(g["price"])
    # pathwise deltas: consistent with a common-random-number central bump of the same oracle (noise-free check)
    d, se = mco.pathwise_deltas(paths, g["S0"], cfg["r"], cfg["strike"], cfg["T"])
    h = 1e-3
    for i in (0, 4):
        pr = []
        for sgn in (+1, -1):
            S = g["S0"].copy()
            S[i] += sgn * h
            np.random.seed(cfg["seed"])
            pr.append(mco.mean_basket_price(mco.simulate_asset_paths(S, cfg["r"], cfg["sigma"], cfg["T"], cfg["dt"],
                                                                     g["corr"], cfg["n"]), cfg["r"], cfg["strike"], cfg["T"]))
        assert abs((pr[0] - pr[1]) / (2 * h) - d[i]) < 0.02 * abs(d[i]) + 1e-4


def test_hjb_exact_oracle_terminal_and_jensen():
    """hjb_u_exact restatement (hjb_implement.py:1085-1094; an inline closure upstream, so it cannot be imported):
    u(T, x) = g(x) exactly, and -ln E[exp(-g)] <= E[g] (Jensen)."""
    rng = np.random.RandomState(2)
    X = rng.normal(size=(3, 5))
    t = np.array([[0.0], [0.5], [1.0]])
    np.random.seed(0)
    u = mco.hjb_u_exact(t, X, 1.0, 5, MC=20000)
    assert u.shape == (3, 1)
    assert abs(u[2, 0] - np.log(0.5 + 0.5 * np.sum(X[2] ** 2))) < 1e-12
    np.random.seed(0)
    W = np.random.normal(size=(20000, 3, 5))
    eg = np.mean(np.log(0.5 + 0.5 * np.sum((X + np.sqrt(2.0 * np.abs(1.0 - t)) * W) ** 2, axis=2)), axis=0)
    assert np.all(u[:, 0] <= eg + 1e-12)


def test_mc_moments_formula():
    np.random.seed(3)
    D = 6
    corr = mco.random_correlation(D)
    w = np.ones(D) / D
    mean, var = mco.terminal_moments(np.ones(D), 0.05, 0.2, corr, w, 1.0)
    paths = mco.gbm_paths(np.ones(D), 0.05, 0.2, corr, True, 1.0, 4, 200000)
    b = (paths[:, -1, :] * w).sum(1)
    assert abs(b.mean() - mean) < 5 * np.sqrt(var / b.size)
    assert abs(b.var() - var) < 0.03 * var
