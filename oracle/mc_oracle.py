"""CPU oracle for the correlated-GBM basket Monte-Carlo pricer -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of /root/reference/numerics/multidimensional_mc_pricer.py:
  * CorrelationMatrix ................ :7-36
  * BlackScholesModel.generate_paths . :49-67
  * BasketOption.payoff .............. :75-77
  * MonteCarloPricer.price ........... :88-93
  * AnalyticalBlackScholes.price ..... :96-108 (closed form used by the reference as a sanity print)

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned by tests/test_oracle_golden.py against outputs of the unmodified reference (same NumPy seed =>
bit-identical price), fixtures written by oracle/make_golden.py.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import norm


def random_correlation(dim: int, with_correlation: bool = True) -> np.ndarray:
    """Unit-diagonal SPD matrix from the NumPy global RNG (uniform entries, symmetrised, + dim*I)."""
    if not with_correlation:
        return np.eye(dim)
    a = np.random.rand(dim, dim)
    a = 0.5 * (a + a.T)
    a += dim * np.eye(dim)
    s = np.diag(1.0 / np.sqrt(np.diag(a)))
    return s @ a @ s


def gbm_paths(S0, rate, sigma, corr, with_correlation, T, N, n_paths) -> np.ndarray:
    """(n_paths, N+1, D) float64 path tensor, one standard_normal((n, D)) draw per step."""
    D = corr.shape[0]
    dt = T / N
    paths = np.zeros((n_paths, N + 1, D))
    paths[:, 0, :] = S0
    L = np.linalg.cholesky(corr) if with_correlation else np.eye(D)
    for k in range(1, N + 1):
        z = np.random.standard_normal((n_paths, D)) @ L.T
        paths[:, k, :] = paths[:, k - 1, :] * np.exp((rate - 0.5 * sigma ** 2) * dt + sigma * np.sqrt(dt) * z)
    return paths


def basket_payoff(S_T, weights, strike):
    return np.maximum(np.sum(S_T * weights, axis=1) - strike, 0.0)


def mc_price(S0, rate, sigma, corr, with_correlation, weights, strike, T, N, n_paths, return_payoffs=False):
    paths = gbm_paths(S0, rate, sigma, corr, with_correlation, T, N, n_paths)
    disc = np.exp(-rate * T) * basket_payoff(paths[:, -1, :], weights, strike)
    if return_payoffs:
        return float(np.mean(disc)), disc
    return float(np.mean(disc))


def analytic_single_asset(S0, strike, rate, sigma, dim, T) -> float:
    """The reference's `AnalyticalBlackScholes.price`: Black-Scholes on mean(S0) with vol sigma/sqrt(dim)
    (exact only for independent assets in the large-dim limit; the reference prints it as a sanity value)."""
    S = float(np.mean(S0))
    sigma = sigma / np.sqrt(dim)
    d1 = (np.log(S / strike) + (rate + 0.5 * sigma ** 2) * T) / (sigma * np.sqrt(T))
    d2 = d1 - sigma * np.sqrt(T)
    return float(S * norm.cdf(d1) - strike * np.exp(-rate * T) * norm.cdf(d2))


def terminal_moments(S0, rate, sigma, corr, weights, T):
    """Exact mean and variance of the terminal basket sum_d w_d S_d(T) under correlated GBM with
    Brownian covariance `corr` (used for size-independent statistical checks of the GPU pricer)."""
    S0 = np.asarray(S0, dtype=np.float64)
    w = np.asarray(weights, dtype=np.float64)
    f = w * S0 * np.exp(rate * T)
    mean = f.sum()
    cov = np.outer(f, f) * (np.exp(sigma ** 2 * T * corr) - 1.0)
    return float(mean), float(cov.sum())


# ------------------------------------------------------------------------------------------------------------
# basket_pricer.py (MonteCarloSimulator :7-53, BasketOptionPricer :56-86) and the HJB Cole-Hopf comparator
# (hjb_implement.py:1085-1094) -- SURVEY.md section 8f rows 3-4
# ------------------------------------------------------------------------------------------------------------
def simulate_asset_paths(S0, r, sigma, T, dt, corr, num_simulations):
    """basket_pricer.py:41-53: (num_assets, num_steps+1, num_simulations) via one np.random.normal draw,
    tensordot with the Cholesky factor and a cumulative product."""
    S0 = np.asarray(S0, dtype=np.float64)
    n_assets, n_steps = len(S0), int(T / dt)
    L = np.linalg.cholesky(corr) if corr is not None else np.eye(n_assets)
    Z = np.random.normal(size=(n_assets, n_steps, num_simulations))
    cz = np.tensordot(L, Z, axes=(1, 0))
    inc = np.exp((r - 0.5 * sigma ** 2) * dt + sigma * np.sqrt(dt) * cz)
    return np.cumprod(np.insert(inc, 0, S0[:, np.newaxis], axis=1), axis=1)


def mean_basket_price(asset_paths, r, strike, T):
    """basket_pricer.py:62-66: equal-weight basket (np.mean over assets), terminal payoff, discounted mean."""
    avg = np.mean(asset_paths, axis=0)
    return float(np.exp(-r * T) * np.mean(np.maximum(avg[-1, :] - strike, 0)))


def pathwise_deltas(asset_paths, S0, r, strike, T):
    """d price / d S0_i = E[e^{-rT} 1{basket > K} (1/D) S_T,i / S0_i]: the limit (epsilon -> 0, common random
    numbers) of the bump-and-revalue loop basket_pricer.py:68-81; upstream re-simulates with FRESH noise per bump,
    whose estimate is dominated by Monte-Carlo noise / epsilon (quirk documented in DESIGN.md)."""
    S_T = asset_paths[:, -1, :]                       # (assets, sims)
    D = S_T.shape[0]
    ind = (S_T.mean(axis=0) > strike).astype(np.float64)
    contrib = np.exp(-r * T) * ind[None, :] * S_T / (D * np.asarray(S0, dtype=np.float64)[:, None])
    return contrib.mean(axis=1), contrib.std(axis=1) / np.sqrt(contrib.shape[1])


def hjb_u_exact(t, X, T, D, MC=10 ** 5):
    """hjb_implement.py:1085-1094 verbatim: t (NC, 1), X (NC, D) -> (NC, 1), NumPy global RNG."""
    def g(Xv):
        return np.log(0.5 + 0.5 * np.sum(Xv ** 2, axis=2, keepdims=True))
    NC = t.shape[0]
    W = np.random.normal(size=(MC, NC, D))
    return -np.log(np.mean(np.exp(-g(X + np.sqrt(2.0 * np.abs(T - t)) * W)), axis=0))
