"""Problem definitions: the reference's FBSNN subclasses with their mu / sigma / phi / g callables, each tagged
with the `ProblemSpec` row the kernels evaluate (SURVEY.md section 8a problem table).

Several reference scripts define different classes under the same name `CallOption`; they are distinct classes
here and the per-script alias modules (DeepBSDE, with_corr_high_dimension_pde, nd_BSPDE_case, bspde_1d_case,
hjb_implement) export them under the upstream names.
"""
from __future__ import annotations

import math

import torch

from . import spec as S
from .fbsnn import FBSNN


class BlackScholesBarenblatt(FBSNN):
    """100-D Black-Scholes-Barenblatt (DeepBSDE.py:326-341): short constructor, no clipping, log every 100."""
    problem_spec = S.ProblemSpec(S.MU_ZERO, 0.0, S.SIGMA_PROP, 0.4, S.PHI_BSB, 0.05, S.G_SUMSQ)
    _y0_as_float = True
    _log_every = 100
    _train_returns = "graph"
    _clip_norm = None

    def __init__(self, Xi, T, M, N, D, layers, mode, activation, **kw):
        super().__init__(Xi, T, M, N, D, layers, mode, activation, **kw)

    def phi_tf(self, t, X, Y, Z):
        return 0.05 * (Y - torch.sum(X * Z, dim=1, keepdim=True))

    def g_tf(self, X):
        return torch.sum(X ** 2, 1, keepdim=True)

    def mu_tf(self, t, X, Y, Z):
        return super().mu_tf(t, X, Y, Z)

    def sigma_tf(self, t, X, Y):
        return 0.4 * torch.diag_embed(X)


def u_exact(t, X, T=1.0, r=0.05, sigma_max=0.4):
    """Closed-form BSB solution the reference plots against (DeepBSDE.py:345-349)."""
    import numpy as np
    return np.exp((r + sigma_max ** 2) * (T - t)) * np.sum(X ** 2, 1, keepdims=True)


def hjb_u_exact(t, X, T=1.0, MC=10 ** 5, seed=None, device=None):
    """Cole-Hopf Monte-Carlo "exact" solution the HJB driver plots against (hjb_implement.py:1085-1094):
    t (NC, 1), X (NC, D) -> (NC, 1) float64 NumPy, `MC` Philox draws per time point on the device (mc_hjb_exact).
    The seed comes from the NumPy global RNG unless given, so np.random.seed(s) keeps runs reproducible."""
    import ctypes

    import numpy as np

    from . import _lib
    from .mc_pricer import _device, _draw_seed
    dev = _device(device)
    lib = _lib.load()
    td = torch.as_tensor(np.asarray(t, dtype=np.float32)).reshape(-1).contiguous().to(dev)
    Xd = torch.as_tensor(np.asarray(X, dtype=np.float32)).reshape(td.numel(), -1).contiguous().to(dev)
    out = torch.empty(td.numel(), dtype=torch.float64, device=dev)
    scratch = torch.empty(lib.mc_scratch_bytes(), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mc_hjb_exact(Xd.shape[1], td.numel(), ctypes.c_void_p(td.data_ptr()), ctypes.c_void_p(Xd.data_ptr()),
                              float(T), int(MC), _draw_seed(seed), ctypes.c_void_p(scratch.data_ptr()),
                              ctypes.c_void_p(out.data_ptr()),
                              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"mc_hjb_exact failed ({rc})")
    return out.cpu().numpy().reshape(-1, 1)


class BasketCallOption(FBSNN):
    """Basket-mean call, `CallOption` of with_corr_high_dimension_pde.py:546-596 and hjb_implement.py:543-586.
    phi keeps the upstream form r*Y (the avg_XZ term is computed and discarded there)."""
    problem_spec = S.ProblemSpec(S.MU_LINEAR, 0.05, S.SIGMA_PROP, 0.20, S.PHI_RY, 0.05, S.G_CALL_MEAN)
    _schedule_kind = "recursive"

    def __init__(self, Xi, T, M, N, D, Mm, layers, mode, activation, correlation_type="no_correlation", **kw):
        super().__init__(Xi, T, M, N, D, Mm, layers, mode, activation, correlation_type, **kw)

    def phi_tf(self, t, X, Y, Z):
        return 0.05 * Y

    def g_tf(self, X):
        avg = torch.mean(X, dim=1, keepdim=True)
        return torch.maximum(avg - self.strike, torch.tensor(0.0).to(X.device))

    def mu_tf(self, t, X, Y, Z):
        return 0.05 * X

    def sigma_tf(self, t, X, Y):
        return 0.20 * torch.diag_embed(X)


class BSPDETestCase(FBSNN):
    """with_corr_high_dimension_pde.py:599-616."""
    problem_spec = S.ProblemSpec(S.MU_LINEAR, 0.05, S.SIGMA_PROP, 0.20, S.PHI_BSB, 0.05, S.G_SUMSQ)
    _schedule_kind = "recursive"

    def __init__(self, Xi, T, M, N, D, Mm, layers, mode, activation, correlation_type="no_correlation", **kw):
        super().__init__(Xi, T, M, N, D, Mm, layers, mode, activation, correlation_type, **kw)

    def g_tf(self, X):
        return torch.sum(X ** 2, dim=1, keepdim=True)

    def phi_tf(self, t, X, Y, Z):
        return 0.05 * (Y - torch.sum(X * Z, dim=1, keepdim=True))

    def mu_tf(self, t, X, Y, Z):
        return 0.05 * X

    def sigma_tf(self, t, X, Y):
        return 0.20 * torch.diag_embed(X)


class CallOption1D(FBSNN):
    """`CallOption` of 1d_BSPDE_case.py:510-560: r = 0.01, sigma = 0.25, phi = r*Y, g = max(sum X - K, 0), K = D."""
    problem_spec = S.ProblemSpec(S.MU_LINEAR, 0.01, S.SIGMA_PROP, 0.25, S.PHI_RY, 0.01, S.G_CALL_SUM)
    _strike_per_dim = True
    _train_returns = "triple"
    _log_every = 100

    def __init__(self, Xi, T, M, N, D, Mm, layers, mode, activation, **kw):
        super().__init__(Xi, T, M, N, D, Mm, layers, mode, activation, **kw)

    def phi_tf(self, t, X, Y, Z):
        return 0.01 * Y

    def g_tf(self, X):
        return torch.maximum(torch.sum(X, dim=1, keepdim=True) - self.strike, torch.tensor(0.0).to(X.device))

    def mu_tf(self, t, X, Y, Z):
        return 0.01 * X

    def sigma_tf(self, t, X, Y):
        return 0.25 * torch.diag_embed(X)


class CallOptionND(FBSNN):
    """`CallOption` of nd_BSPDE_case.py:503-540: r = 0.05, sigma = 0.2, phi = r(Y - X.Z), g = max(sum X - K, 0), K = D."""
    problem_spec = S.ProblemSpec(S.MU_LINEAR, 0.05, S.SIGMA_PROP, 0.20, S.PHI_BSB, 0.05, S.G_CALL_SUM)
    _strike_per_dim = True
    _train_returns = "triple"
    _log_every = 100

    def __init__(self, Xi, T, M, N, D, Mm, layers, mode, activation, **kw):
        super().__init__(Xi, T, M, N, D, Mm, layers, mode, activation, **kw)

    def phi_tf(self, t, X, Y, Z):
        return 0.05 * (Y - torch.sum(X * Z, dim=1, keepdim=True))

    def g_tf(self, X):
        return torch.maximum(torch.sum(X, dim=1, keepdim=True) - self.strike, torch.tensor(0.0).to(X.device))

    def mu_tf(self, t, X, Y, Z):
        return 0.05 * X

    def sigma_tf(self, t, X, Y):
        return 0.20 * torch.diag_embed(X)


class HamiltonJacobiBellman(FBSNN):
    """100-D HJB (hjb_implement.py:590-604): sigma = sqrt(2) I, phi = |Z|^2, g = ln(0.5 + 0.5 |X|^2).
    Upstream passes Mm=None and then crashes in train(); here Mm=None means 'no schedule'."""
    problem_spec = S.ProblemSpec(S.MU_ZERO, 0.0, S.SIGMA_CONST, float(torch.sqrt(torch.tensor(2.0))),
                                 S.PHI_ZSQ, 1.0, S.G_LOGQ)

    def __init__(self, Xi, T, M, N, D, layers, mode, activation, **kw):
        super().__init__(Xi, T, M, N, D, None, layers, mode, activation, **kw)

    def phi_tf(self, t, X, Y, Z):
        return torch.sum(Z ** 2, dim=1, keepdim=True)

    def g_tf(self, X):
        return torch.log(0.5 + 0.5 * torch.sum(X ** 2, dim=1, keepdim=True))

    def mu_tf(self, t, X, Y, Z):
        return super().mu_tf(t, X, Y, Z)

    def sigma_tf(self, t, X, Y):
        return math.sqrt(2.0) * super().sigma_tf(t, X, Y)
