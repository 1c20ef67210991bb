"""Closed-form comparators the reference's drivers plot the learned solution against, evaluated on the device for a
whole prediction tensor in one launch:

  * `BasketOptionPriceCalculator` -- nd_BSPDE_case.py:621-658 (per-asset Black-Scholes call, equal-weighted mean)
  * `BasicOptionPriceCalculator`  -- with_corr_high_dimension_pde.py:663-700 (Black-Scholes on the basket average with
    volatility sigma / sqrt(D); upstream loops over every (sample, step) in Python)

Same class / method names, argument order and return types as upstream.  No CPU fallback.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .mc_pricer import _device


def _run(mode, S, t, rows, cols, ntimes, K, r, sigma, T, dims, dev):
    lib = _lib.load()
    S = S.to(device=dev, dtype=torch.float64).contiguous()
    t = t.to(device=dev, dtype=torch.float64).contiguous()
    n_out = rows if mode == 0 else rows * cols
    price = torch.empty(n_out, dtype=torch.float64, device=dev)
    delta = torch.empty(n_out, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mc_bs_comparator(mode, ctypes.c_void_p(S.data_ptr()), ctypes.c_void_p(t.data_ptr()), rows, cols, ntimes,
                                  float(K), float(r), float(sigma), float(T), int(dims), ctypes.c_void_p(price.data_ptr()),
                                  ctypes.c_void_p(delta.data_ptr()),
                                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"mc_bs_comparator failed ({rc})")
    return price, delta


class BasketOptionPriceCalculator:
    @staticmethod
    def black_scholes_call(S, K, T, r, sigma, q=0):
        """Element-wise Black-Scholes call and delta (nd_BSPDE_case.py:622-634); small helper kept for drivers."""
        S, K, T, r, sigma, q = [torch.as_tensor(x).to(S.device) for x in [S, K, T, r, sigma, q]]
        d1 = (torch.log(S / K) + (r - q + 0.5 * sigma ** 2) * T) / (sigma * torch.sqrt(T))
        d2 = d1 - sigma * torch.sqrt(T)
        normal = torch.distributions.Normal(0, 1)
        return S * torch.exp(-q * T) * normal.cdf(d1) - K * torch.exp(-r * T) * normal.cdf(d2), normal.cdf(d1)

    @staticmethod
    def calculate_option_prices(S, t, K, r, sigma, T):
        """S (batch, steps, assets), t (batch, steps, 1) -> basket price and delta, each (batch, steps, 1)."""
        S = torch.as_tensor(S)
        t = torch.as_tensor(t)
        batch, steps, assets = S.shape
        dev = S.device if S.is_cuda else _device()
        price, delta = _run(0, S.reshape(-1, assets), t.reshape(-1), batch * steps, assets, 0, K, r, sigma, T, 1, dev)
        shape = (batch, steps, 1)
        return price.reshape(shape).to(S.dtype), delta.reshape(shape).to(S.dtype)


class BasicOptionPriceCalculator:
    @staticmethod
    def black_scholes_call(S, K, T, r, sigma, dimensions, q=0):
        from math import erf, exp, log, sqrt
        cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
        sigma_avg = sigma / np.sqrt(dimensions)
        S_avg = float(np.mean(S))
        d1 = (log(S_avg / K) + (r + 0.5 * sigma_avg ** 2) * T) / (sigma_avg * sqrt(T))
        d2 = d1 - sigma_avg * sqrt(T)
        return S_avg * cdf(d1) - K * exp(-r * T) * cdf(d2), cdf(d1)

    def calculate_call_option_prices(self, X_pred, time_array, K, r, sigma, T, dimensions, q=0):
        """X_pred (rows, cols) basket averages (NumPy or tensor), time_array (n,) -> (prices, deltas) float64 NumPy."""
        if torch.is_tensor(X_pred):
            X = X_pred.detach()
        else:
            X = torch.as_tensor(np.asarray(X_pred, dtype=np.float64))
        rows, cols = X.shape
        tt = torch.as_tensor(np.asarray(time_array, dtype=np.float64)).reshape(-1)
        dev = X.device if X.is_cuda else _device()
        price, delta = _run(1, X, tt, rows, cols, tt.numel(), K, r, sigma, T, dimensions, dev)
        return price.reshape(rows, cols).cpu().numpy(), delta.reshape(rows, cols).cpu().numpy()
