"""Size-independent properties of the CUDA path at sizes the CPU oracle cannot reach, edge cases, and the
Monte-Carlo pricer (statistical parity with the reference algorithm, known answers, sharding invariance)."""
import json
import math

import numpy as np
import pytest
import torch

from tests import golden_util as gu

pytestmark = pytest.mark.gpu

D, N = 100, 50
LAYERS = [D + 1] + 4 * [256] + [1]


def _bsb(M, precision="fp32", **kw):
    import dnnpde_b200 as pde
    torch.manual_seed(11)
    return pde.BlackScholesBarenblatt(gu.make_xi("bsb", D), 1.0, M, N, D, LAYERS, "FC", "Sine", precision=precision,
                                      seed=5, **kw)


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_loss_and_gradient_are_additive_over_path_shards(precision):
    """loss and d(loss)/d(theta) are sums over paths: full batch == sum of shards (what multi-GPU sharding uses)."""
    M = 8192
    sol = _bsb(M, precision)
    t, W = sol.fetch_minibatch_device(iteration=3)
    loss, X, Y, _, g = sol.loss_grad_flat(t, W)
    loss, g, Y = float(loss), g.clone(), Y.clone()
    acc_l, acc_g = 0.0, torch.zeros_like(g)
    for lo, hi in ((0, 3000), (3000, 8192)):
        sol.M = hi - lo
        l, _, Ys, _, gs = sol.loss_grad_flat(t[lo:hi].contiguous(), W[lo:hi].contiguous())
        acc_l += float(l)
        acc_g += gs
        assert torch.equal(Ys, Y[lo:hi])                       # per-row results do not depend on the batch
    assert abs(acc_l - loss) <= 2e-6 * abs(loss)
    assert float((acc_g - g).abs().max()) <= 2e-5 * float(g.abs().max())
    # all paths start at Xi => one Y0, one Z0
    assert float(Y[:, 0, 0].max() - Y[:, 0, 0].min()) == 0.0


def test_path_advance_matches_fp64_recursion_at_full_size():
    M = 65536
    sol = _bsb(M)
    t, W = sol.fetch_minibatch_device(iteration=1)
    X, Y = sol.predict(gu.make_xi("bsb", D), t, W)
    dW = (W[:, 1:] - W[:, :-1]).double()
    ref = torch.cumprod(torch.cat((torch.ones(M, 1, D, device=W.device, dtype=torch.float64), 1 + 0.4 * dW), 1), 1)
    ref = ref * torch.as_tensor(gu.make_xi("bsb", D), device=W.device)
    assert float((X.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert torch.isfinite(Y).all() and Y.shape == (M, N + 1, 1)


def test_device_brownian_statistics_and_shard_invariance():
    import dnnpde_b200 as pde
    np.random.seed(2)
    sol = pde.BasketCallOption(np.ones((1, 20)), 1.0, 20000, 10, 20, None, [21, 64, 64, 1], "FC", "Sine",
                               "random_correlation", seed=9)
    t, W = sol.fetch_minibatch_device(iteration=4)
    assert torch.all(W[:, 0] == 0) and torch.allclose(t[0, :, 0], torch.linspace(0, 1, 11, device=W.device))
    inc = (W[:, 1:] - W[:, :-1]).reshape(-1, 20).double()
    cov = (inc.t() @ inc / inc.shape[0]).cpu().numpy()
    target = 0.1 * sol.correlation_matrix
    assert np.abs(cov - target).max() <= 0.02 * np.abs(target).max()
    assert float(inc.mean().abs()) < 5e-3
    # a shard of the global path range reproduces the same rows bit for bit
    t2, W2 = sol.fetch_minibatch_device(iteration=4, path_offset=7000, n_paths=500)
    assert torch.equal(W2, W[7000:7500])
    # a different iteration gives a different stream
    _, W3 = sol.fetch_minibatch_device(iteration=5)
    assert not torch.equal(W3, W)


def test_philox_training_is_deterministic_and_learns():
    losses = []
    for _ in range(2):
        sol = _bsb(512, "fp32", brownian="philox")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            sol.train(40, 1e-3)
        losses.append(sol.last_losses.copy())
    assert np.array_equal(losses[0], losses[1])                # no atomics anywhere: bitwise reproducible
    assert losses[0][-5:].mean() < 0.7 * losses[0][:5].mean()


@pytest.mark.parametrize("M,Nn,Dd", [(1, 50, 100), (3, 1, 4), (5, 7, 1), (130, 3, 33)])
def test_edge_shapes_match_oracle(M, Nn, Dd):
    import dnnpde_b200 as pde
    from oracle import fbsnn_oracle as orc
    torch.manual_seed(1)
    np.random.seed(1)
    layers = [Dd + 1, 32, 32, 1]
    Xi = np.random.uniform(0.5, 1.5, (M, Dd))                  # per-path initial states (xi_rows == M)
    oracle = orc.OracleSolver("bsptest", Xi, 1.0, M, Nn, Dd, layers, "FC", "Tanh", squeeze_quirk=False)
    sol = pde.BSPDETestCase(Xi, 1.0, M, Nn, Dd, None, layers, "FC", "Tanh")
    sol.model.load_state_dict(oracle.model.state_dict())
    t, W = oracle.fetch_minibatch()
    ol, oX, oY, oZ, og = oracle.grads(t, W)
    loss, X, Y, Z, g = sol.loss_grad_flat(t, W, want_Z=True)
    assert abs(float(loss) - float(ol)) <= 2e-5 * abs(float(ol)) + 1e-7
    assert torch.allclose(X.cpu(), oX, rtol=1e-6, atol=1e-7)
    assert torch.allclose(Y.cpu(), oY, rtol=2e-5, atol=2e-6)
    assert torch.allclose(Z.cpu(), oZ, rtol=1e-4, atol=1e-5)
    for name, p in sol.model.named_parameters():
        o = sol._fp.offsets[name]
        got = g[o:o + p.numel()].view(p.shape).cpu()
        assert float((got - og[name]).abs().max()) <= 1e-4 * float(og[name].abs().max()) + 1e-6, name


def test_predict_broadcasts_like_the_reference():
    import dnnpde_b200 as pde
    torch.manual_seed(2)
    np.random.seed(2)
    sol = pde.BasketCallOption(np.ones((1, 6)), 1.0, 9, 5, 6, None, [7, 32, 32, 1], "FC", "Sine")
    t, W = sol.fetch_minibatch()
    X, Y = sol.predict(np.ones((1, 6)), t.cpu().numpy(), W.cpu().numpy())      # NumPy inputs, singleton Xi
    assert X.shape == (9, 6, 6) and Y.shape == (9, 6, 1) and sol.M == 9
    X1, Y1 = sol.predict(np.ones((1, 6)), t[:1], W[:1])                          # batch of one
    assert sol.M == 1 and torch.allclose(Y1[0], Y[0], rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        sol.M = 9
        sol.loss_function(t[:, :4], W[:, :4], sol.Xi)                            # N mismatch


def test_two_phase_trainer_and_prediction_sampler():
    """The reference's executor-level callers of the hot path (with_corr...:619-660, :704-726)."""
    import contextlib, io
    import dnnpde_b200 as pde
    from dnnpde_b200.with_corr_high_dimension_pde import CallOption, PredictionGenerator, TrainingPhases
    torch.manual_seed(4)
    np.random.seed(4)
    D6, M6, N6 = 6, 32, 8
    Xi = np.ones((1, D6))
    model = CallOption(Xi, 1.0, M6, N6, D6, None, [D6 + 1, 64, 64, 64, 1], "Naisnet", "Sine", "random_correlation")
    phases = TrainingPhases(model)
    with contextlib.redirect_stdout(io.StringIO()):
        graph, min_loss, state, logs = phases.train_initial_phase(30, 1e-3, "Adam")
        graph2, min_loss2, state2, _ = phases.fine_tuning_phase(10, 1e-5)
    assert graph.shape[0] == 2 and np.isfinite(min_loss) and state[0].shape == (M6, N6 + 1, D6)
    assert phases.min_loss == min_loss2 and model.iteration[0] == 0
    t_test, W_test, X_pred, Y_pred = PredictionGenerator(model, Xi, 5).generate_predictions()
    assert t_test.shape == (5 * M6, N6 + 1, 1) and X_pred.shape == (5 * M6, N6 + 1, D6) and Y_pred.shape == (5 * M6, N6 + 1, 1)
    assert W_test.shape == (M6, N6 + 1, D6) and model.M == M6
    # the batched forward equals the per-sample loop of the reference on the same NumPy stream
    np.random.seed(42)
    t0, W0 = model.fetch_minibatch()
    X0, Y0 = model.predict(Xi, t0, W0)
    assert np.allclose(Y_pred[:M6], Y0.cpu().numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(X_pred[:M6], X0.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# Monte-Carlo pricer
# ---------------------------------------------------------------------------------------------------------------
def _pricer(Dm, n, corr=True, seed=3, N_steps=50, strike=1.0):
    import dnnpde_b200 as pde
    np.random.seed(0)
    model = pde.BlackScholesModel(0.05, 0.2, Dm, corr)
    return model, pde.MonteCarloPricer(model, pde.BasketOption(np.ones(Dm) / Dm, strike), 1.0, N_steps, n, seed=seed)


@pytest.mark.parametrize("tag", ["d5", "d100", "d8_nocorr"])
def test_mc_price_agrees_with_reference_within_standard_error(tag):
    g, _ = gu.load("mc_pricer")
    cfg = json.loads(str(g[f"{tag}_cfg"]))
    model, pr = _pricer(cfg["D"], 1 << 22, cfg["corr"])
    assert np.array_equal(model.correlation, g[f"{tag}_corr"])          # same NumPy-generated correlation matrix
    price, se = pr.price(np.ones(cfg["D"]), return_stderr=True)
    ref = float(g[f"{tag}_price"])                                       # the reference's own estimate at n = cfg["n"]
    se_ref = 0.08 / math.sqrt(cfg["n"])                                  # payoff std < 0.08 for these baskets
    assert abs(price - ref) <= 4 * math.hypot(se, se_ref), (price, ref, se, se_ref)
    assert se < 1e-4


def test_mc_matches_exact_moments_and_black_scholes():
    from oracle import mc_oracle as mco
    # strike 0 => price = exp(-rT) E[basket] exactly; variance from the closed-form covariance of correlated GBM
    model, pr = _pricer(100, 1 << 22, True, strike=0.0)
    price, se = pr.price(np.ones(100), return_stderr=True)
    mean, var = mco.terminal_moments(np.ones(100), 0.05, 0.2, model.correlation, np.ones(100) / 100, 1.0)
    assert abs(price - math.exp(-0.05) * mean) <= 4 * se
    assert abs(se * math.sqrt(1 << 22) / math.exp(-0.05) - math.sqrt(var)) <= 0.01 * math.sqrt(var)
    # D = 1: the Black-Scholes formula is a known answer (the reference's AnalyticalBlackScholes, exact for D = 1)
    import dnnpde_b200 as pde
    model1, pr1 = _pricer(1, 1 << 23, False)
    p1, se1 = pr1.price(np.ones(1), return_stderr=True)
    bs = pde.AnalyticalBlackScholes(0.05, 0.2, 1).price(np.ones(1), 1.0, 1.0)
    assert abs(p1 - bs) <= 4 * se1, (p1, bs, se1)


def test_mc_paths_and_pricer_share_streams_and_shard_exactly():
    model, pr = _pricer(12, 40000, True, seed=17, N_steps=9)
    S0 = np.linspace(0.8, 1.2, 12)
    paths = model.generate_paths(S0, 1.0, 9, 40000, seed=17)
    assert paths.shape == (40000, 10, 12) and paths.dtype == np.float64 and np.allclose(paths[:, 0], S0)
    pay = math.exp(-0.05) * np.maximum((paths[:, -1] * (np.ones(12) / 12)).sum(1) - 1.0, 0)
    price = pr.price(S0)
    assert abs(price - pay.mean()) <= 2e-6 * pay.mean()        # same Philox keys: path-wise equal up to rounding
    # sharding the global path range over "ranks" reproduces the sums (sum, sum of squares) to double round-off
    full = pr.price_async(S0, 40000, 0, 17).cpu().numpy()
    parts = sum(pr.price_async(S0, hi - lo, lo, 17).cpu().numpy() for lo, hi in ((0, 12345), (12345, 40000)))
    assert np.allclose(parts, full, rtol=1e-12)
    # log-returns are Gaussian with the right per-step variance
    lr = np.log(paths[:, 1:] / paths[:, :-1]).reshape(-1, 12)
    assert abs(lr.std() - 0.2 * math.sqrt(1 / 9)) < 2e-3


def test_mc_pathwise_deltas_match_oracle_and_bump_with_common_random_numbers():
    """mc_basket_price_delta (basket_pricer.py:68-81): pathwise deltas vs the NumPy oracle's
    pathwise estimator (statistical) and vs a central bump of the GPU price under common random numbers (exact
    same Philox paths, so the comparison is noise-free up to O(eps^2) and fp32 rounding)."""
    from oracle import mc_oracle as mco
    import dnnpde_b200 as pde
    Dm, n = 6, 1 << 20
    np.random.seed(3)
    model = pde.BlackScholesModel(0.05, 0.2, Dm, True)
    S0 = np.linspace(0.9, 1.1, Dm)
    pr = pde.MonteCarloPricer(model, pde.BasketOption(np.ones(Dm) / Dm, 1.0), 1.0, 10, n, seed=5)
    price, deltas, se = pr.price_and_delta(S0, return_stderr=True)
    assert abs(price - pr.price(S0)) <= 1e-12 * max(price, 1.0)            # same kernel body, same sums
    np.random.seed(4)
    paths = mco.simulate_asset_paths(S0, 0.05, 0.2, 1.0, 0.1, model.correlation, 200000)
    od, ose = mco.pathwise_deltas(paths, S0, 0.05, 1.0, 1.0)
    assert np.all(np.abs(deltas - od) <= 5 * ose + 1e-4), (deltas, od, ose)
    assert abs(price - mco.mean_basket_price(paths, 0.05, 1.0, 1.0)) <= 5 * (se + 0.1 / math.sqrt(200000))
    for i in (0, Dm - 1):                                                   # central bump, identical Philox streams
        h = 2e-2
        up, dn = S0.copy(), S0.copy()
        up[i] += h
        dn[i] -= h
        fd = (pr.price(up) - pr.price(dn)) / (2 * h)
        assert abs(fd - deltas[i]) <= 2e-3 * abs(deltas[i]) + 1e-5, (i, fd, deltas[i])
    # sum_i S0_i delta_i = E[disc 1{B>K} B] >= price (homogeneity of the payoff in S0)
    assert float(np.dot(S0, deltas)) >= price


def test_basket_pricer_surface_matches_reference_layout():
    """MonteCarloSimulator / BasketOptionPricer of basket_pricer.py:7-86 on the device generator."""
    from oracle import mc_oracle as mco
    import dnnpde_b200 as pde
    bp = pde.basket_pricer
    S0 = np.ones(4)
    corr = np.full((4, 4), 0.3) + 0.7 * np.eye(4)
    sim = bp.MonteCarloSimulator(S0, 0.05, 0.2, 1.0, 0.05, corr, seed=11)
    paths = sim.simulate(100000)
    assert paths.shape == (4, 21, 100000) and paths.dtype == np.float64 and np.allclose(paths[:, 0, :], 1.0)
    pricer = bp.BasketOptionPricer(1.0, 1.0, corr, seed=11)
    price = pricer.price(paths, 0.05)
    np.random.seed(8)
    ref_paths = mco.simulate_asset_paths(S0, 0.05, 0.2, 1.0, 0.05, corr, 100000)
    ref = mco.mean_basket_price(ref_paths, 0.05, 1.0, 1.0)
    assert abs(price - ref) <= 5 * 0.15 / math.sqrt(100000) * math.sqrt(2)
    lr = np.log(paths[:, 1:, :] / paths[:, :-1, :])
    c = np.corrcoef(lr.reshape(4, -1))
    assert np.abs(c - corr).max() < 0.01                                    # increments carry the requested correlation
    p2, deltas = pricer.price_and_delta(S0, paths, 0.05, 0.2, 1.0, 0.05)
    od, ose = mco.pathwise_deltas(ref_paths, S0, 0.05, 1.0, 1.0)
    assert p2 == price and np.all(np.abs(deltas - od) <= 5 * ose + 2e-4)


def test_hjb_cole_hopf_exact_matches_oracle_and_terminal_condition():
    """mc_hjb_exact (hjb_implement.py:1085-1094) vs the NumPy restatement; at t = T the
    formula collapses to g(X) exactly."""
    from oracle import mc_oracle as mco
    import dnnpde_b200 as pde
    D, NC, T = 20, 7, 1.0
    rng = np.random.RandomState(0)
    t = np.linspace(0, T, NC)[:, None]
    X = rng.normal(size=(NC, D)) * 0.7
    got = pde.hjb_u_exact(t, X, T, MC=400000, seed=9)
    np.random.seed(1)
    ref = mco.hjb_u_exact(t, X, T, D, MC=100000)
    assert got.shape == (NC, 1) and got.dtype == np.float64
    assert np.abs(got - ref).max() <= 5e-3, (got.ravel(), ref.ravel())       # MC noise of the 1e5-sample oracle ~1e-3
    g_T = np.log(0.5 + 0.5 * np.sum(X[-1] ** 2))
    assert abs(got[-1, 0] - g_T) <= 1e-5 * max(1.0, abs(g_T))
    again = pde.hjb_u_exact(t, X, T, MC=400000, seed=9)
    assert np.array_equal(got, again)                                         # Philox: reproducible


def test_heston_surface_and_clamp():
    """HestonFBSNN (heston_dnnpde.py:519-699): (u, dU/dS, dU/dv) outputs with u clamped at 0 and
    the derivative masked, predict() -> (S, v, Y), train() -> column_stack, in-kernel Brownian driver (1 column)."""
    from oracle import fbsnn_oracle as orc
    import dnnpde_b200 as pde
    torch.manual_seed(4)
    np.random.seed(4)
    layers = [2, 32, 32, 32, 1]
    Xi = np.array([[1.0]])
    oracle = orc.HestonOracle(Xi, 1.0, 24, 8, layers, "FC", "Sine")
    tq = torch.rand(50, 1)
    Xq = torch.cat([0.5 + torch.rand(50, 1), 0.05 + 0.4 * torch.rand(50, 1)], 1).requires_grad_(True)
    with torch.no_grad():                                                    # centre the raw output on the clamp
        raw = oracle.model(torch.cat((tq, Xq), 1))
        oracle.model[-1].bias -= raw.median()
    sol = pde.HestonFBSNN(Xi, 1.0, 24, 8, 1, 1, layers, "FC", "Sine", precision="fp32")
    sol.model.load_state_dict(oracle.model.state_dict())
    ou, oS, ov = oracle.net_u(tq, Xq)
    u, dS, dv = sol.net_u(tq, Xq.detach())
    assert u.shape == (50, 1) and dS.shape == (50, 1) and dv.shape == (50, 1)
    assert (ou.detach() == 0).any() and (ou.detach() > 0).any(), "test rows must straddle the clamp"
    assert torch.allclose(u.cpu(), ou.detach(), atol=2e-6) and torch.allclose(dS.cpu(), oS.detach(), atol=2e-5)
    assert torch.allclose(dv.cpu(), ov.detach(), atol=2e-5)
    with torch.no_grad():
        raw = oracle.model(torch.cat((tq, Xq), 1))
    assert (dS.cpu()[raw < -1e-5] == 0).all() and (dv.cpu()[raw < -1e-5] == 0).all()   # masked where the clamp is active
    t, W = sol.fetch_minibatch()
    assert W.shape == (24, 9, 1)
    S, v, Y = sol.predict(Xi, t, W)
    ol, oX, oY, _ = oracle.loss_function(t.cpu(), W.cpu())
    assert S.shape == (24, 9, 1) and v.shape == (24, 9, 1) and Y.shape == (24, 9, 1)
    assert torch.allclose(torch.cat([S, v], 2).cpu(), oX.detach(), atol=1e-6)
    assert torch.allclose(Y.cpu(), oY.detach(), atol=5e-6)
    out = sol.train(3, 1e-3)
    assert out.shape == (1, 3) and out[0, 0] == 0                            # (iteration, mean loss, Y0) rows
    philox = pde.HestonFBSNN(Xi, 1.0, 512, 8, 1, 1, layers, "FC", "Sine", precision="fp32", brownian="philox",
                             payoff_type="continuous")
    philox.train(20, 1e-3)
    assert np.isfinite(philox.last_losses).all()
    t2, W2 = philox.fetch_minibatch_device(seed=3)
    inc = (W2[:, 1:] - W2[:, :-1]).reshape(-1)
    assert W2.shape == (512, 9, 1) and abs(float(inc.std()) - math.sqrt(1 / 8)) < 0.02


@pytest.mark.parametrize("problem", ["bsb", "hjb", "basket", "heston", "heston_smooth"])
def test_fused_path_loss_kernel_matches_row_parallel_kernels(problem):
    """From 2048 paths up the residual + seeds run as one per-path kernel (loss_path_kernel); below, as the two
    row-parallel kernels.  The whole batch (fused kernel) must equal the sum over shards of 1024 paths (row-parallel
    kernels): loss, gradients, Y -- for every form of phi / g, the Heston clamp mask and the first-component payoffs."""
    import dnnpde_b200 as pde
    torch.manual_seed(3)
    M, Nn = 4096, 12
    if problem.startswith("heston"):
        sol = pde.HestonFBSNN(np.array([[1.0]]), 1.0, M, Nn, 1, 1, [2, 64, 64, 1], "FC", "Sine", precision="fp32", seed=9,
                              payoff_type="continuous" if problem.endswith("smooth") else "discontinuous")
        with torch.no_grad():
            sol.model[-1].bias += 0.05                          # straddle the clamp
    else:
        Dd = 10
        cls = {"bsb": pde.BlackScholesBarenblatt, "hjb": pde.HamiltonJacobiBellman, "basket": pde.BasketCallOption}[problem]
        layers = [Dd + 1, 64, 64, 64, 1]
        xi = np.zeros((1, Dd)) if problem == "hjb" else np.ones((1, Dd))
        if problem == "basket":
            sol = cls(xi, 1.0, M, Nn, Dd, None, layers, "FC", "Tanh", precision="fp32", seed=9)
        else:
            sol = cls(xi, 1.0, M, Nn, Dd, layers, "FC", "Sine", precision="fp32", seed=9)
    t, W = sol.fetch_minibatch_device(iteration=2)
    loss, X, Y, _, g = sol.loss_grad_flat(t, W)
    loss, g, Y = float(loss), g.clone(), Y.clone()
    acc_l, acc_g = 0.0, torch.zeros_like(g)
    for lo in range(0, M, 1024):
        sol.M = 1024
        l, _, Ys, _, gs = sol.loss_grad_flat(t[lo:lo + 1024].contiguous(), W[lo:lo + 1024].contiguous())
        acc_l += float(l)
        acc_g += gs
        assert torch.equal(Ys, Y[lo:lo + 1024])
    assert np.isfinite(loss) and abs(acc_l - loss) <= 5e-6 * abs(loss), (acc_l, loss)
    assert float((acc_g - g).abs().max()) <= 5e-5 * float(g.abs().max())


def test_tensor_core_variants_agree_with_fp32_at_large_batch():
    """M = 16 384 paths (835 584 rows): the large-shape code paths -- persistent multi-tile sweeps, the 16-epilogue-warp
    F kernel, the CTA-pair weight-gradient kernel with ~11 000-row K chunks, the fused per-path loss kernel -- against
    the fp32 SIMT variant on the same minibatch and weights."""
    M = 16384
    ref = _bsb(M, "fp32")
    t, W = ref.fetch_minibatch_device(iteration=7)
    l0, _, Y0, _, g0 = ref.loss_grad_flat(t, W)
    l0, Y0, g0 = float(l0), Y0.clone(), g0.clone()
    state = {k: v.clone() for k, v in ref.model.state_dict().items()}
    del ref
    torch.cuda.empty_cache()
    for precision, tol_l, tol_g in (("tf32x3", 2e-5, 2e-4), ("tf32", 1e-2, 8e-2)):
        sol = _bsb(M, precision)
        sol.model.load_state_dict(state)
        l1, _, Y1, _, g1 = sol.loss_grad_flat(t, W)
        assert abs(float(l1) - l0) <= tol_l * abs(l0), (precision, float(l1), l0)
        assert float((Y1 - Y0).abs().max()) <= 50 * tol_l * float(Y0.abs().max())
        assert float((g1 - g0).abs().max()) <= tol_g * float(g0.abs().max()), precision
        del sol
        torch.cuda.empty_cache()


@pytest.mark.parametrize("mode", ["FC", "Naisnet"])
@pytest.mark.parametrize("H", [64, 128, 192, 256])
@pytest.mark.parametrize("M", [3, 200, 1500])
def test_tensor_core_kernel_forms_across_shapes(mode, H, M):
    """Every dispatch of the tcgen05 kernels (8- and 16-epilogue-warp sweeps, CTA-pair weight gradients, narrow
    column tiles at few row tiles, fused bias column sums indexed per CTA or per m-tile, SIMT for what does not fit)
    against the fp32 SIMT variant: hidden widths 64..256, FC and NAIS-Net, ragged row counts."""
    import dnnpde_b200 as pde
    Dd, Nn = 31, 7                                     # d_in = 32: the input-width GEMMs are tensor-core eligible too
    layers = [Dd + 1, H, H, H, 1]
    torch.manual_seed(H + M)
    args = (np.ones((1, Dd)), 1.0, M, Nn, Dd, None, layers, mode, "Sine")
    ref = pde.BasketCallOption(*args, precision="fp32", seed=3)
    t, W = ref.fetch_minibatch_device(iteration=1)
    l0, _, Y0, Z0, g0 = ref.loss_grad_flat(t, W, want_Z=True)
    l0, Y0, Z0, g0 = float(l0), Y0.clone(), Z0.clone(), g0.clone()
    for precision, tl, tg in (("tf32x3", 5e-5, 5e-4), ("tf32", 2e-2, 1e-1)):
        sol = pde.BasketCallOption(*args, precision=precision, seed=3)
        sol.model.load_state_dict(ref.model.state_dict())
        l1, _, Y1, Z1, g1 = sol.loss_grad_flat(t, W, want_Z=True)
        assert abs(float(l1) - l0) <= tl * abs(l0) + 1e-7, (precision, float(l1), l0)
        assert float((Y1 - Y0).abs().max()) <= 20 * tl * float(Y0.abs().max()) + 1e-6
        assert float((Z1 - Z0).norm()) <= 20 * tl * float(Z0.norm()) + 1e-6
        assert float((g1 - g0).abs().max()) <= tg * float(g0.abs().max()) + 1e-7, precision
