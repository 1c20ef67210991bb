"""No-GPU checks: the C-ABI library loads and exports every symbol of include/fbsnn_b200.h, the ctypes structs
match the C layout, the Python surface mirrors the reference's (constructor arities, state_dict keys, NumPy
Brownian stream, error behaviour), and compute entry points fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import dnnpde_b200 as pde
from oracle import fbsnn_oracle as orc
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = pde._lib.load()
    header = open(os.path.join(ROOT, "include", "fbsnn_b200.h")).read()
    declared = set(re.findall(r"\b((?:fbsnn|mc)_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/fbsnn_b200.h but not exported"
    assert set(pde._lib.EXPORTS) == declared
    assert lib.fbsnn_version() == pde.spec.ABI_VERSION


def test_peer_buffer_layout_without_gpu():
    """Symmetric [gradient | loss | flags] buffer of the fused NVLink all-reduce (host-side layout arithmetic)."""
    lib = pde._lib.load()
    fo, tot = ctypes.c_int64(), ctypes.c_int64()
    assert lib.fbsnn_peer_buffer_floats(223744, ctypes.byref(fo), ctypes.byref(tot)) == 0
    assert fo.value >= 223744 + 4 and fo.value % 64 == 0 and tot.value == fo.value + 32   # ready[16] + done[16]
    assert lib.fbsnn_peer_buffer_floats(223745, ctypes.byref(fo), ctypes.byref(tot)) == -1     # not a multiple of 4
    assert lib.fbsnn_peer_allreduce_adam(None, None, None, 2, 0, None, None, None, 8, None, None) == -1
    assert lib.fbsnn_peer_wait(None, 8, 2, None, None) == -1


def test_workspace_and_validation_without_gpu():
    lib = pde._lib.load()
    sol = pde.BlackScholesBarenblatt(gu.make_xi("bsb", 100), 1.0, 100, 50, 100, [101] + 4 * [256] + [1], "FC", "Sine")
    sp = sol._spec()
    assert ctypes.sizeof(sp) == 3 * 4 + 8 * 4 + 6 * 4 + 5 * 4 + 4 + 4 + 4 * 10 * 8 + 8 + 3 * 4 + 5 * 4   # matches the C struct (v101)
    need = ctypes.c_size_t(0)
    assert lib.fbsnn_workspace_bytes(ctypes.byref(sp), 100, 1, ctypes.byref(need)) == 0
    rows = 100 * 51
    assert need.value >= rows * 256 * 4 * 20          # 5 row arrays per hidden layer
    fwd = ctypes.c_size_t(0)
    assert lib.fbsnn_workspace_bytes(ctypes.byref(sp), 100, 0, ctypes.byref(fwd)) == 0 and fwd.value < need.value
    sp.width[0] = 250                                  # not a multiple of 4
    assert lib.fbsnn_workspace_bytes(ctypes.byref(sp), 100, 1, ctypes.byref(need)) == -2
    assert b"multiple of 4" in lib.fbsnn_last_error()
    sp = sol._spec()
    sp.phi_kind = 9
    assert lib.fbsnn_workspace_bytes(ctypes.byref(sp), 100, 1, ctypes.byref(need)) == -2


def test_state_dict_keys_and_flat_layout_match_reference_naming():
    torch.manual_seed(0)
    fc = pde.BlackScholesBarenblatt(np.ones((1, 10)), 1.0, 4, 5, 10, [11, 32, 32, 32, 32, 1], "FC", "Sine")
    assert list(fc.model.state_dict()) == ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias",
                                           "6.weight", "6.bias", "8.weight", "8.bias"]
    na = pde.BasketCallOption(np.ones((1, 5)), 1.0, 4, 5, 5, None, [6, 32, 32, 32, 32, 1], "Naisnet", "ReLU")
    assert list(na.model.state_dict()) == [
        "layer1.weight", "layer1.bias", "layer2.weight", "layer2.bias", "layer2_input.weight", "layer2_input.bias",
        "layer3.weight", "layer3.bias", "layer3_input.weight", "layer3_input.bias", "layer4.weight", "layer4.bias",
        "layer4_input.weight", "layer4_input.bias", "layer5.weight", "layer5.bias"]
    for s in (fc, na):
        assert s._fp.is_intact()
        for name, p in s.model.named_parameters():
            o = s._fp.offsets[name]
            assert o % 32 == 0 and torch.equal(s._fp.flat[o:o + p.numel()].view(p.shape), p.data)
    sp = na._spec()
    assert sp.n_hidden == 4 and sp.net_kind == 1 and sp.off_Win[2] >= 0 and sp.off_Win[1] == -1
    # loading a reference checkpoint keeps the flat aliasing
    ref = orc.build_model([6, 32, 32, 32, 32, 1], "Naisnet", "ReLU")
    na.model.load_state_dict(ref.state_dict())
    assert na._fp.is_intact()
    assert torch.equal(na._fp.flat[na._fp.offsets["layer3.weight"]:][:1024].view(32, 32), ref.layer3.weight.data)


def test_same_seed_gives_reference_initial_weights_on_cpu():
    """Construction consumes the torch RNG exactly like the reference (default Linear init, then xavier)."""
    for mode, layers in (("FC", [11, 16, 16, 1]), ("Naisnet", [11, 16, 16, 16, 1])):
        torch.manual_seed(5)
        ours = pde.BSPDETestCase(np.ones((1, 10)), 1.0, 4, 5, 10, None, layers, mode, "Tanh", device="cpu")
        torch.manual_seed(5)
        ref = orc.build_model(layers, mode, "Tanh")
        for (k1, v1), (k2, v2) in zip(ours.model.state_dict().items(), ref.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)


@pytest.mark.parametrize("name", ["heston_fc_sine", "heston_nais_tanh_smooth"])
def test_heston_constructor_reproduces_reference_initial_weights(name):
    """HestonFBSNN builds its network like heston_dnnpde.py:519-585 (base net with layers[0] inputs, input layer(s)
    swapped for 3 inputs, xavier gain 0.5, zero biases): the same torch seed gives the REFERENCE's initial weights
    (parameter checksums stored in the fixture), and the same NumPy seed its first Brownian minibatch."""
    g, meta = gu.load(name)
    torch.manual_seed(meta["torch_seed"])
    np.random.seed(meta["numpy_seed"])
    sol = pde.HestonFBSNN(gu.make_xi("ones", 1), meta["T"], meta["M"], meta["N"], 1, int(meta["N"] ** (1 / 5)),
                          meta["layers"], meta["mode"], meta["act"], payoff_type=meta["payoff"], device="cpu")
    names = [k for k, _ in sol.model.named_parameters()]
    assert names == [str(x) for x in g["param_names"]]
    sums = np.array([float(p.detach().double().sum()) for _, p in sol.model.named_parameters()])
    asums = np.array([float(p.detach().double().abs().sum()) for _, p in sol.model.named_parameters()])
    assert np.allclose(sums, g["param_sum"], rtol=0, atol=1e-12) and np.allclose(asums, g["param_abssum"], rtol=0, atol=1e-12)
    t, W = sol.fetch_minibatch()
    assert W.shape == (meta["M"], meta["N"] + 1, 1)
    wsum = np.array([float(W.double().sum()), float(W.double().abs().sum())])
    assert np.allclose(wsum, g["W_sum"], rtol=1e-12)
    sp = sol._spec()
    assert (sp.D, sp.noise_dim, sp.clamp_u, sp.zt_dims) == (2, 1, 1, 1) and sol.D == 1 and sol._fp.is_intact()
    with pytest.raises(RuntimeError):
        sol.net_u(np.zeros((2, 1)), np.ones((2, 2)))          # no CPU fallback


def test_fetch_minibatch_is_the_reference_numpy_stream():
    g, meta = gu.load("basket10_nais_sine_5l")
    oracle = gu.rebuild_inputs(meta, g)
    np.random.seed(meta["numpy_seed"])
    sol = pde.BasketCallOption(gu.make_xi("ones", 10), 1.0, meta["M"], meta["N"], 10, None, meta["layers"],
                               "Naisnet", "Sine", meta["corr"], device="cpu")
    assert np.array_equal(sol.correlation_matrix, g["corr_matrix"])
    t, W = sol.fetch_minibatch()
    assert t.shape == (meta["M"], meta["N"] + 1, 1) and W.shape == (meta["M"], meta["N"] + 1, 10)
    assert t.dtype == torch.float32 and float(t[0, 0, 0]) == 0.0 and torch.all(W[:, 0, :] == 0)
    wsum = np.array([float(W.double().sum()), float(W.double().abs().sum())])
    assert np.allclose(wsum, g["W_sum"], rtol=1e-12)


def test_constructor_arities_attributes_and_errors():
    short = pde.HamiltonJacobiBellman(np.zeros((1, 4)), 1.0, 8, 6, 4, [5, 16, 16, 1], "Naisnet", "ReLU")
    assert short.Mm is None and short.strike == 1.0 and short.correlation_type == "no_correlation"
    long_ = pde.CallOption1D(np.ones((1, 1)), 1.0, 8, 6, 1, 2.0, [2, 16, 16, 1], "FC", "Sine")
    assert long_.Mm == 2.0 and long_.strike == 1.0 and long_.D == 1
    nd = pde.CallOptionND(np.ones((1, 7)), 1.0, 8, 6, 7, None, [8, 16, 16, 1], "FC", "Sine")
    assert nd.strike == 7.0
    for attr in ("device", "Xi", "T", "M", "N", "D", "mode", "activation", "model", "training_loss", "iteration",
                 "optimizer", "correlation_matrix"):
        assert hasattr(short, attr)
    assert short.Xi.requires_grad and short.Xi.dtype == torch.float32
    with pytest.raises(ValueError):
        pde.BasketCallOption(np.ones((1, 4)), 1.0, 8, 6, 4, None, [5, 16, 16, 1], "FC", "Sine", "bogus_correlation")
    with pytest.raises(ValueError):
        pde.BlackScholesBarenblatt(np.ones((1, 4)), 1.0, 8, 6, 4, [5, 16, 16, 1], "FC", "Softplus")
    with pytest.raises(NotImplementedError):
        pde.BlackScholesBarenblatt(np.ones((1, 4)), 1.0, 8, 6, 4, [5, 16, 16, 1], "SDEnet", "Sine")
    with pytest.raises(ValueError):
        short.train(1, 1e-3, optimizer_type="Nadam")
    # N-schedule of the reference (opt-in): Mm = 50 ** (1/5) gives N = 3 for it < 4000, then 5, 11, 23, 51
    sch = pde.CallOption1D(np.ones((1, 1)), 1.0, 8, 50, 1, 50 ** (1 / 5), [2, 16, 16, 1], "FC", "Sine",
                           n_schedule="reference")
    assert [sch._scheduled_N(i) for i in (0, 3999, 4000, 8000, 12000, 16000)] == [3, 3, 5, 11, 23, 51]


def test_compute_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sol = pde.BlackScholesBarenblatt(np.ones((1, 4)), 1.0, 8, 6, 4, [5, 16, 16, 1], "FC", "Sine")
    t, W = sol.fetch_minibatch()
    for call in (lambda: sol.loss_function(t, W, sol.Xi), lambda: sol.train(1, 1e-3),
                 lambda: sol.predict(np.ones((1, 4)), t, W), lambda: sol.net_u(t[:, 0], W[:, 0])):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    model = pde.BlackScholesModel(0.05, 0.2, 5, True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pde.MonteCarloPricer(model, pde.BasketOption(np.ones(5) / 5, 1.0), 1.0, 50, 1000).price(np.ones(5))
    pr = pde.MonteCarloPricer(model, pde.BasketOption(np.ones(5) / 5, 1.0), 1.0, 50, 1000)
    for call in (lambda: pr.price_and_delta(np.ones(5)), lambda: pde.hjb_u_exact(np.zeros((3, 1)), np.ones((3, 4))),
                 lambda: pde.basket_pricer.MonteCarloSimulator(np.ones(3), 0.05, 0.2, 1.0, 0.1).simulate(10)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_basket_pricer_host_surface():
    """basket_pricer.py:7-39 host-side pieces: step count, Cholesky factor, positive-definite repair."""
    bp = pde.basket_pricer
    sim = bp.MonteCarloSimulator(np.ones(3), 0.05, 0.2, 1.0, 0.25, np.eye(3))
    assert sim.num_assets == 3 and sim.num_steps == 4 and np.allclose(sim.L, np.eye(3))
    bad = np.array([[1.0, 1.0], [1.0, 1.0]])                       # singular: repaired by adding eps * I
    fixed = bp.MonteCarloSimulator(np.ones(2), 0.05, 0.2, 1.0, 0.5, bad)
    assert np.all(np.linalg.eigvalsh(fixed.correlation_matrix) > 0) and np.array_equal(bad, [[1.0, 1.0], [1.0, 1.0]])
    assert bp.MonteCarloSimulator(np.ones(2), 0.05, 0.2, 1.0, 0.5).correlation_matrix is None
    paths = np.ones((2, 3, 5))
    paths[:, -1, :] = [[1.5] * 5, [0.9] * 5]
    assert abs(bp.BasketOptionPricer(1.0, 1.0).price(paths, 0.05) - np.exp(-0.05) * 0.2) < 1e-12


def test_mc_host_objects_match_reference():
    g, _ = gu.load("mc_pricer")
    np.random.seed(0)
    m = pde.BlackScholesModel(0.05, 0.20, 5, True)
    assert np.array_equal(m.correlation, g["d5_corr"])
    assert abs(pde.AnalyticalBlackScholes(0.05, 0.2, 5).price(np.ones(5), 1.0, 1.0) - float(g["d5_analytic"])) < 1e-12
    opt = pde.BasketOption(np.ones(5) / 5, 1.0)
    S = np.array([[1.2, 1.0, 0.9, 1.1, 1.3], [0.5, 0.6, 0.7, 0.8, 0.9]])
    assert np.allclose(opt.payoff(S), [0.1, 0.0])
