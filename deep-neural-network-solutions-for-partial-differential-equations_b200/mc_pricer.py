"""Correlated-GBM basket Monte-Carlo pricer with the surface of numerics/multidimensional_mc_pricer.py
(CorrelationMatrix :7-36, BlackScholesModel :39-67, BasketOption :70-77, MonteCarloPricer :80-93,
AnalyticalBlackScholes :96-108), computed by the Philox-in-kernel path generator with a fused payoff reduction.

`MonteCarloPricer.price(S0)` never materialises paths (the reference allocates an (n, N+1, D) float64 tensor,
40.8 TB at n = 1e9).  Streams are Philox4x32-10 keyed by (seed, global path id): not NumPy's MT19937 stream, so
prices agree with the reference statistically (within standard-error bands), not bit-wise; the seed is drawn
from the NumPy global RNG unless given, so `np.random.seed(s)` still makes a run reproducible.  With
torch.distributed initialised and `data_parallel=True` the global path range is sharded over ranks and three
scalars are all-reduced; the result does not depend on the number of GPUs.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import parallel
from . import spec as S


class CorrelationMatrix:
    """Random SPD matrix with unit diagonal from the NumPy global RNG (host, once; :12-36)."""

    def __init__(self, dimensions, with_correlation=True):
        self.dimensions = dimensions
        self.with_correlation = with_correlation
        self.matrix = self.generate_correlation_matrix()

    def generate_correlation_matrix(self):
        if self.with_correlation:
            return self.generate_positive_definite_correlation_matrix()
        return np.eye(self.dimensions)

    def generate_positive_definite_correlation_matrix(self):
        d = self.dimensions
        a = np.random.rand(d, d)
        a = 0.5 * (a + a.T)
        a += d * np.eye(d)
        s = np.diag(1.0 / np.sqrt(np.diag(a)))
        return s @ a @ s


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("the Monte-Carlo pricer needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _draw_seed(seed: Optional[int], shared: bool = False, device=None) -> int:
    """Explicit seed, or one drawn from the NumPy global RNG.  `shared`: the pricer shards ONE global path range over
    the ranks, so every rank must key Philox with the same seed -- rank 0's draw is broadcast (each rank still draws,
    so the NumPy streams stay in step)."""
    s = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 62))
    if shared and seed is None and parallel.is_distributed():
        import torch.distributed as dist
        dev = _device(device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([s], dtype=torch.int64, device=dev)
        dist.broadcast(t, src=0)
        s = int(t.item())
    return s


class BlackScholesModel:
    def __init__(self, rate, sigma, dimensions, with_correlation=False):
        self.rate = rate
        self.sigma = sigma
        self.dimensions = dimensions
        self.with_correlation = with_correlation
        self.correlation_matrix = CorrelationMatrix(dimensions, with_correlation)
        self.correlation = self.correlation_matrix.matrix

    def _chol_T(self, dev) -> Optional[torch.Tensor]:
        """Transposed lower Cholesky factor on the device (the kernel reads column d of L^T contiguously)."""
        if not self.with_correlation:
            return None
        L = np.linalg.cholesky(self.correlation)
        return torch.from_numpy(np.ascontiguousarray(L.T)).float().to(dev)

    def _mc_spec(self, T, N, strike=0.0) -> S.McSpec:
        if np.ndim(self.sigma) != 0 and np.size(self.sigma) != 1:
            raise NotImplementedError("the fused pricer takes one scalar volatility for all assets (what every driver of "
                                      "the reference passes); per-asset sigma arrays are not implemented")
        return S.McSpec(int(self.dimensions), int(N), float(self.rate), float(np.asarray(self.sigma).reshape(-1)[0]),
                        float(T), float(strike))

    def generate_paths(self, S0, T, N, num_simulations, seed: Optional[int] = None, as_tensor: bool = False,
                       path_offset: int = 0, device=None):
        """(num_simulations, N+1, D) GBM paths, S_t = S_{t-1} exp((r - sigma^2/2) dt + sigma sqrt(dt) (L z_t)) (:49-67).
        Returns float64 NumPy like the reference (or the fp32 device tensor with as_tensor=True)."""
        dev = _device(device)
        lib = _lib.load()
        sp = self._mc_spec(T, N)
        S0d = torch.as_tensor(np.broadcast_to(np.asarray(S0, dtype=np.float32), (self.dimensions,)).copy()).to(dev)
        cT = self._chol_T(dev)
        out = torch.empty(num_simulations, N + 1, self.dimensions, device=dev)
        with torch.cuda.device(dev):
            rc = lib.mc_generate_paths(ctypes.byref(sp), ctypes.c_void_p(S0d.data_ptr()),
                                       None if cT is None else ctypes.c_void_p(cT.data_ptr()), num_simulations,
                                       _draw_seed(seed), path_offset, ctypes.c_void_p(out.data_ptr()),
                                       ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"mc_generate_paths failed ({rc}); D must be <= 256")
        return out if as_tensor else out.double().cpu().numpy()


class BasketOption:
    def __init__(self, weights, strike):
        self.weights = weights
        self.strike = strike

    def payoff(self, S):
        if isinstance(S, torch.Tensor):
            w = torch.as_tensor(np.asarray(self.weights), dtype=S.dtype, device=S.device)
            return torch.clamp((S * w).sum(dim=1) - self.strike, min=0)
        return np.maximum(np.sum(S * self.weights, axis=1) - self.strike, 0)


class MonteCarloPricer:
    def __init__(self, model, option, T, N, num_simulations, seed: Optional[int] = None,
                 data_parallel: bool = False, device=None):
        self.model = model
        self.option = option
        self.T = T
        self.N = N
        self.num_simulations = num_simulations
        self.seed = seed
        self.data_parallel = data_parallel
        self.device = device
        self.last_stderr = None
        self.last_seed = None

    def price_async(self, S0, num_simulations: Optional[int] = None, path_offset: int = 0, seed=None):
        """Enqueue the fused simulate+payoff+reduce kernels for global paths [offset, offset + n); returns the
        device tensor [sum, sum of squares] (fp64) without synchronising."""
        dev = _device(self.device)
        lib = _lib.load()
        m = self.model
        n = int(self.num_simulations if num_simulations is None else num_simulations)
        sp = m._mc_spec(self.T, self.N, self.option.strike)
        D = m.dimensions
        S0d = torch.as_tensor(np.broadcast_to(np.asarray(S0, dtype=np.float32), (D,)).copy()).to(dev)
        wd = torch.as_tensor(np.broadcast_to(np.asarray(self.option.weights, dtype=np.float32), (D,)).copy()).to(dev)
        cT = m._chol_T(dev)
        scratch = torch.empty(lib.mc_scratch_bytes(), dtype=torch.uint8, device=dev)
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        self._keep = (S0d, wd, cT, scratch)
        with torch.cuda.device(dev):
            rc = lib.mc_basket_price(ctypes.byref(sp), ctypes.c_void_p(S0d.data_ptr()), ctypes.c_void_p(wd.data_ptr()),
                                     None if cT is None else ctypes.c_void_p(cT.data_ptr()), n, int(seed), path_offset,
                                     ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(sums.data_ptr()),
                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"mc_basket_price failed ({rc}); D must be <= 256")
        return sums

    def price_and_delta(self, S0, return_stderr: bool = False):
        """Price and the D pathwise deltas d price / d S0_d from ONE fused simulation pass
        (mc_basket_price_delta): the quantity basket_pricer.py:68-81 estimates by bump-and-revalue."""
        dev = _device(self.device)
        lib = _lib.load()
        m = self.model
        n_total = int(self.num_simulations)
        seed = _draw_seed(self.seed, shared=self.data_parallel, device=self.device)
        self.last_seed = seed
        if self.data_parallel and parallel.is_distributed():
            lo, hi = parallel.shard_range(n_total, parallel.rank(), parallel.world_size())
        else:
            lo, hi = 0, n_total
        D = m.dimensions
        out = torch.zeros(2 + D, dtype=torch.float64, device=dev)
        if hi > lo:
            sp = m._mc_spec(self.T, self.N, self.option.strike)
            S0d = torch.as_tensor(np.broadcast_to(np.asarray(S0, dtype=np.float32), (D,)).copy()).to(dev)
            wd = torch.as_tensor(np.broadcast_to(np.asarray(self.option.weights, dtype=np.float32), (D,)).copy()).to(dev)
            cT = m._chol_T(dev)
            scratch = torch.empty(lib.mc_scratch_bytes(), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                rc = lib.mc_basket_price_delta(ctypes.byref(sp), ctypes.c_void_p(S0d.data_ptr()),
                                               ctypes.c_void_p(wd.data_ptr()),
                                               None if cT is None else ctypes.c_void_p(cT.data_ptr()), hi - lo, seed, lo,
                                               ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                               ctypes.c_void_p(out[2:].data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"mc_basket_price_delta failed ({rc}); D must be <= 256")
        if self.data_parallel:
            parallel.allreduce_sums(out)
        vals = out.cpu().numpy()
        mean = float(vals[0]) / n_total
        var = max(float(vals[1]) / n_total - mean * mean, 0.0)
        self.last_stderr = math.sqrt(var / n_total)
        deltas = vals[2:] / n_total
        return (mean, deltas, self.last_stderr) if return_stderr else (mean, deltas)

    def price(self, S0, return_stderr: bool = False):
        """exp(-rT) * mean payoff (:88-93).  `return_stderr=True` also returns the standard error, which the
        reference does not compute."""
        n_total = int(self.num_simulations)
        seed = _draw_seed(self.seed, shared=self.data_parallel, device=self.device)
        self.last_seed = seed
        if self.data_parallel and parallel.is_distributed():
            lo, hi = parallel.shard_range(n_total, parallel.rank(), parallel.world_size())
        else:
            lo, hi = 0, n_total
        if hi > lo:
            sums = self.price_async(S0, hi - lo, lo, seed)
        else:
            sums = torch.zeros(2, dtype=torch.float64, device=_device(self.device))
        if self.data_parallel:
            parallel.allreduce_sums(sums)
        s, q = (float(v) for v in sums.cpu())
        mean = s / n_total
        var = max(q / n_total - mean * mean, 0.0)
        self.last_stderr = math.sqrt(var / n_total)
        return (mean, self.last_stderr) if return_stderr else mean


class AnalyticalBlackScholes:
    """Host-side sanity value the reference prints next to the MC price (:96-108): Black-Scholes on the mean
    spot with volatility sigma / sqrt(D)."""

    def __init__(self, rate, sigma, dimensions):
        self.rate = rate
        self.sigma = sigma
        self.dimensions = dimensions
        self.sigma_avg = sigma / np.sqrt(dimensions)

    def price(self, S0, strike, T):
        from math import erf, log, sqrt, exp
        cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
        s = float(np.mean(S0))
        d1 = (log(s / strike) + (self.rate + 0.5 * self.sigma_avg ** 2) * T) / (self.sigma_avg * sqrt(T))
        d2 = d1 - self.sigma_avg * sqrt(T)
        return s * cdf(d1) - strike * exp(-self.rate * T) * cdf(d2)
