"""Surface of the reference's basket_pricer.py (MonteCarloSimulator :7-53, BasketOptionPricer :56-86) on the
fused Philox path generator.

* `MonteCarloSimulator.simulate(n)` returns the reference's layout (num_assets, num_steps+1, n), float64 NumPy
  (or the fp32 device tensor with as_tensor=True), from `mc_generate_paths`.
* `BasketOptionPricer.price(asset_paths, r)` is the reference formula on whatever array it is given.
* `BasketOptionPricer.delta` / `price_and_delta`: upstream bumps each S0_i by 1e-4 and re-simulates with FRESH
  noise, so its estimate is Monte-Carlo noise divided by epsilon.  Here the deltas are the pathwise estimator
  E[e^{-rT} 1{basket > K} S_T,i / (D S0_i)] (the epsilon -> 0 limit under common random numbers), accumulated per
  asset inside the pricing kernel (`mc_basket_price_delta`): one pass instead of D + 1 simulations.
"""
from __future__ import annotations

import numpy as np

from .mc_pricer import BasketOption, BlackScholesModel, MonteCarloPricer


class _FixedCorrelationModel(BlackScholesModel):
    """BlackScholesModel with a caller-supplied correlation matrix (upstream's simulator takes the matrix)."""

    def __init__(self, rate, sigma, correlation):
        self.rate, self.sigma = rate, sigma
        self.dimensions = correlation.shape[0]
        self.with_correlation = True
        self.correlation_matrix = None
        self.correlation = correlation


class MonteCarloSimulator:
    def __init__(self, S0, r, sigma, T, dt, correlation_matrix=None, seed=None):
        self.S0 = np.asarray(S0, dtype=np.float64)
        self.r, self.sigma, self.T, self.dt = r, sigma, T, dt
        self.num_assets = len(self.S0)
        self.num_steps = int(T / dt)
        self.seed = seed
        if correlation_matrix is not None:
            self.correlation_matrix = self._make_positive_definite(np.array(correlation_matrix, dtype=np.float64))
            self.L = np.linalg.cholesky(self.correlation_matrix)
        else:
            self.correlation_matrix = None
            self.L = np.eye(self.num_assets)

    @staticmethod
    def _is_positive_definite(matrix):
        try:
            np.linalg.cholesky(matrix)
            return True
        except np.linalg.LinAlgError:
            return False

    def _make_positive_definite(self, matrix):
        epsilon = 1e-10                                    # basket_pricer.py:32-39
        while not self._is_positive_definite(matrix):
            matrix = matrix + epsilon * np.eye(matrix.shape[0])
            epsilon *= 10
        return matrix

    def _model(self):
        corr = self.correlation_matrix if self.correlation_matrix is not None else np.eye(self.num_assets)
        m = _FixedCorrelationModel(self.r, self.sigma, corr)
        m.with_correlation = self.correlation_matrix is not None
        return m

    def simulate(self, num_simulations, as_tensor=False):
        paths = self._model().generate_paths(self.S0, self.num_steps * self.dt, self.num_steps, num_simulations,
                                             seed=self.seed, as_tensor=True)            # (n, N+1, D) on the device
        out = paths.permute(2, 1, 0)                                                   # (D, N+1, n) like upstream
        return out.contiguous() if as_tensor else out.double().cpu().numpy()


class BasketOptionPricer:
    def __init__(self, strike, T, correlation_matrix=None, seed=None):
        self.strike, self.T, self.correlation_matrix, self.seed = strike, T, correlation_matrix, seed

    def price(self, asset_paths, r):
        avg = asset_paths.mean(axis=0) if isinstance(asset_paths, np.ndarray) else asset_paths.mean(dim=0).cpu().numpy()
        return float(np.exp(-r * self.T) * np.mean(np.maximum(avg[-1, :] - self.strike, 0)))

    def _pricer(self, S0, n, r, sigma, T, dt):
        sim = MonteCarloSimulator(S0, r, sigma, T, dt, self.correlation_matrix)
        D = len(S0)
        return MonteCarloPricer(sim._model(), BasketOption(np.ones(D) / D, self.strike), sim.num_steps * dt,
                                sim.num_steps, n, seed=self.seed)

    def delta(self, S0, asset_paths, r, sigma, T, dt, epsilon=1e-4):
        """Pathwise deltas over asset_paths.shape[2] fresh paths (epsilon is accepted for signature parity)."""
        return self._pricer(S0, asset_paths.shape[2], r, sigma, T, dt).price_and_delta(S0)[1]

    def price_and_delta(self, S0, asset_paths, r, sigma, T, dt):
        return self.price(asset_paths, r), self.delta(S0, asset_paths, r, sigma, T, dt)
