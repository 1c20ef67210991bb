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


class HestonFBSNN(FBSNN):
    """Heston 2-factor FBSNN (heston_dnnpde.py:519-699): state (S, v) on ONE Brownian driver (the reference passes
    D = 1 to its base class and lets einsum broadcast dW over both diffusion columns), network input (t, S, v),
    u clamped at 0, outputs (u, dU/dS, dU/dv), terminal call payoff on S ('discontinuous') or its sigmoid-smoothed
    form ('continuous'), phi = r Y.  train() follows that file: N-schedule from Mm, clip 1.0, log every 100,
    returns column_stack((iteration, training_loss, Y0_values))."""
    _check_layers = False
    _skip_nonfinite = True
    _log_every = 100
    _train_returns = "heston"
    _clip_norm = 1.0
    _schedule_kind = "mm"

    def __init__(self, Xi, T, M, N, D, Mm, layers, mode, activation, correlation_type="no_correlation", kappa=2.0,
                 theta=0.2, sigma=0.3, rho=0.8, v0=0.2, payoff_type='discontinuous', **kw):
        import torch.nn as nn

        from .networks import FlatParams
        if payoff_type not in ("discontinuous", "continuous"):
            raise ValueError("Invalid payoff type. Choose 'discontinuous' or 'continuous'.")
        kw.setdefault("n_schedule", None)
        super().__init__(Xi, T, M, N, 1, Mm, layers, mode, activation, correlation_type, **kw)
        self.kappa, self.theta, self.sigma, self.rho, self.v0 = kappa, theta, sigma, rho, v0
        self.payoff_type = payoff_type
        self.y0_values = []
        self.Y0_values = []
        layers = list(layers)
        # swap the input layer(s) for 3 inputs (t, S, v), exactly in the reference's order (RNG stream), :533-544
        if self.mode == "FC":
            self.model[0] = nn.Linear(in_features=3, out_features=layers[1]).to(self.device)
        else:
            self.model.layer1 = nn.Linear(in_features=3, out_features=layers[1]).to(self.device)
            self.model.layer2_input = nn.Linear(in_features=3, out_features=layers[2]).to(self.device)
            if len(layers) >= 5:
                self.model.layer3_input = nn.Linear(in_features=3, out_features=layers[3]).to(self.device)
            if len(layers) == 6:
                self.model.layer4_input = nn.Linear(in_features=3, out_features=layers[4]).to(self.device)
            self.model.layers = [3] + layers[1:]
        self.layers = [3] + layers[1:]
        self.initialize_weights()
        self._fp = FlatParams(self.model, "FC" if self.mode == "FC" else "NAIS", self.device)
        self.problem_spec = S.ProblemSpec(S.MU_HESTON, 0.05, S.SIGMA_HESTON, 0.0, S.PHI_RY, 0.05,
                                          S.G_CALL_FIRST if payoff_type == "discontinuous" else S.G_CALL_FIRST_SMOOTH)

    @property
    def _sdim(self):
        return 2

    def _state_xi(self, Xi):
        """[S0, v0] rows: the caller's Xi carries S0 only (:621-626); predict() may pass (S, v) pairs (:668-671)."""
        Xi = Xi.reshape(-1, Xi.shape[-1] if Xi.dim() > 1 else 1)
        if Xi.shape[1] == 1:
            Xi = torch.cat([Xi, torch.full_like(Xi, self.v0)], dim=1)
        return Xi[:, :2]

    def _fill_problem(self, sp):
        super()._fill_problem(sp)
        sp.noise_dim, sp.clamp_u, sp.zt_dims = 1, 1, 1
        sp.h_kappa, sp.h_theta, sp.h_xi, sp.h_rho, sp.h_v0 = self.kappa, self.theta, self.sigma, self.rho, self.v0

    def initialize_weights(self):
        for param in self.model.parameters():           # heston_dnnpde.py:580-585
            if len(param.shape) > 1:
                torch.nn.init.xavier_uniform_(param, gain=0.5)
            else:
                torch.nn.init.zeros_(param)

    def net_u(self, t, X):
        """(u, dU/dS, dU/dv) for X = (S, v) rows (:560-577)."""
        X = self._as_f32(X, self.device)
        if X.dim() == 1:
            X = X.unsqueeze(0)
        u, du = super().net_u(t, X)
        return u, du[:, 0:1], du[:, 1:2]

    def predict(self, Xi_star, t_star, W_star):
        X, Y = super().predict(Xi_star, t_star, W_star)
        return X[:, :, 0:1], X[:, :, 1:2], Y             # S, v, Y (:683)

    def calculate_greeks(self, S, v, t):
        """(Y, delta, gamma) on a grid (:685-699).  delta is the analytic in-kernel dU/dS; gamma -- a second
        autograd pass upstream -- is a central difference of delta (net_u outputs are not autograd-differentiable)."""
        import numpy as np
        S = np.asarray(S, dtype=np.float32).reshape(-1)
        v = np.asarray(v, dtype=np.float32).reshape(-1)
        t = np.broadcast_to(np.asarray(t, dtype=np.float32).reshape(-1), S.shape).copy()
        h = 1e-3 * np.maximum(np.abs(S), 1.0)
        Y, delta, _ = self.net_u(t[:, None], np.stack([S, v], -1))
        _, dp, _ = self.net_u(t[:, None], np.stack([S + h, v], -1))
        _, dm, _ = self.net_u(t[:, None], np.stack([S - h, v], -1))
        gamma = (dp - dm).cpu().numpy()[:, 0] / (2 * h)
        return Y.cpu().numpy(), delta.cpu().numpy(), gamma[:, None]

    def phi_tf(self, t, X, Y, Z):
        return 0.05 * Y

    def g_tf(self, X):
        Sv = X[:, 0:1] if X.dim() > 1 else X
        if self.payoff_type == 'discontinuous':
            return torch.maximum(Sv - self.strike, torch.tensor(0.0).to(Sv.device))
        return (Sv - self.strike) / (1 + torch.exp(-10.0 * (Sv - self.strike)))

    def mu_tf(self, t, X, Y=None, Z=None):
        Sv, v = X[:, 0:1], X[:, 1:2]
        return torch.cat([0.05 * Sv, self.kappa * (self.theta - v)], dim=1).clamp(-100, 100)

    def sigma_tf(self, t, X, Y=None):
        Sv, v = X[:, 0:1], X[:, 1:2]
        sS = torch.sqrt(torch.clamp(v, min=1e-8)) * Sv
        sv = self.sigma * torch.sqrt(torch.clamp(v, min=1e-8))
        m = torch.zeros((Sv.shape[0], 2, 2), device=X.device)
        m[:, 0, 0], m[:, 1, 1] = sS.squeeze(-1), sv.squeeze(-1)
        m[:, 0, 1], m[:, 1, 0] = self.rho * sv.squeeze(-1), self.rho * sS.squeeze(-1)
        return m.clamp(-100, 100)
