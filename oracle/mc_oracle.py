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
