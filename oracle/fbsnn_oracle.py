"""CPU oracle for the FBSNN training step  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a plain-PyTorch (CPU, autograd) restatement of the reference's forward-backward
SDE solver, written only so that the CUDA path can be checked against it.  Nothing under
`deep-neural-network-solutions-for-partial-differential-equations_b200/` imports it; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may.

Parity status: the reference ships no golden vectors or tests of its own, so this
restatement is pinned against outputs of the *unmodified reference run in the build container*
(`oracle/make_golden.py` imports `/root/reference`, writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays them).  Where the reference has a result-changing quirk
the oracle follows the reference and says so.

Reference lines restated here (all relative to /root/reference):
  * Sine activation ................ Functions/Sine.py:6-12
  * FC network construction ........ DeepBSDE.py:166-172
  * NAIS-Net ....................... Functions/naisnet.py:6-95 (inline copy with_corr_high_dimension_pde.py:39-129)
  * xavier init of Linear weights .. DeepBSDE.py:185-187
  * net_u (u and Du by autograd) ... DeepBSDE.py:189-194
  * Dg_tf .......................... DeepBSDE.py:196-200
  * loss_function .................. DeepBSDE.py:202-245, with_corr_high_dimension_pde.py:270-314
  * fetch_minibatch ................ DeepBSDE.py:247-262, with_corr_high_dimension_pde.py:316-353
  * train iteration ................ DeepBSDE.py:274-280, with_corr_high_dimension_pde.py:412-425
  * problem callables .............. DeepBSDE.py:326-341, 1d_BSPDE_case.py:526-560, nd_BSPDE_case.py:517-539,
                                     with_corr_high_dimension_pde.py:561-616, hjb_implement.py:594-604
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------------
class SineAct(nn.Module):
    def forward(self, x):
        return torch.sin(x)


def make_activation(name: str) -> nn.Module:
    if name == "Sine":
        return SineAct()
    if name == "ReLU":
        return nn.ReLU()
    if name == "Tanh":
        return nn.Tanh()
    raise ValueError(f"unknown activation {name!r}")


class StableResNet(nn.Module):
    """NAIS-Net with 1..3 stable blocks (len(layers) in {4,5,6}); attribute names follow the
    reference so that state_dict keys are interchangeable (layer1, layer2, layer2_input, ...)."""

    def __init__(self, layers, activation: nn.Module, stable: bool = True, epsilon: float = 0.01):
        super().__init__()
        if len(layers) not in (4, 5, 6):
            raise ValueError("NAIS-Net needs 4, 5 or 6 layer sizes")
        self.layers = list(layers)
        self.n_blocks = len(layers) - 3
        self.layer1 = nn.Linear(layers[0], layers[1])
        for b in range(self.n_blocks):
            k = b + 2
            setattr(self, f"layer{k}", nn.Linear(layers[k - 1], layers[k]))
            setattr(self, f"layer{k}_input", nn.Linear(layers[0], layers[k]))
        k = self.n_blocks + 2
        setattr(self, f"layer{k}", nn.Linear(layers[k - 1], layers[k]))
        # registration order (layer1, layer2, layer2_input, layer3, [layer3_input, layer4, [layer4_input,
        # layer5]]) equals the reference's, which fixes the init RNG stream.
        self.activation = activation
        self.epsilon = epsilon
        self.stable = stable

    def stable_matrix(self, lin: nn.Linear) -> torch.Tensor:
        delta = 1 - 2 * self.epsilon
        rtr = lin.weight.t() @ lin.weight
        nrm = torch.norm(rtr)
        if nrm > delta:
            rtr = delta ** 0.5 * rtr / (nrm ** 0.5)
        return rtr + torch.eye(rtr.shape[0], dtype=rtr.dtype) * self.epsilon

    def forward(self, x):
        inp = x
        out = self.activation(self.layer1(x))
        for b in range(self.n_blocks):
            k = b + 2
            lin = getattr(self, f"layer{k}")
            keep = out
            if self.stable:
                out = F.linear(out, -self.stable_matrix(lin), lin.bias)
                out = out + getattr(self, f"layer{k}_input")(inp)
            else:
                out = lin(out)
            out = self.activation(out) + keep
        return getattr(self, f"layer{self.n_blocks + 2}")(out)


def build_model(layers, mode: str, activation: str) -> nn.Module:
    act = make_activation(activation)
    if mode == "FC":
        mods = []
        for i in range(len(layers) - 2):
            mods += [nn.Linear(layers[i], layers[i + 1]), act]
        mods.append(nn.Linear(layers[-2], layers[-1]))
        model = nn.Sequential(*mods)
    elif mode in ("Naisnet", "NAIS-Net"):
        model = StableResNet(layers, act, stable=True)
    else:
        raise ValueError(f"unknown mode {mode!r}")

    def init(m):
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)

    model.apply(init)
    return model


# --------------------------------------------------------------------------------------------
# problems
# --------------------------------------------------------------------------------------------
@dataclass
class Problem:
    """mu = mu_c * X (or 0); sigma = sigma_c * diag(X) (prop) or sigma_c * I (const)."""
    name: str
    mu_c: float
    sigma_prop: bool
    sigma_c: float
    phi: str          # 'bsb' : c (Y - sum X Z) | 'ry' : c Y | 'zsq' : sum Z^2
    phi_c: float
    g: str            # 'sumsq' | 'call_sum' | 'call_mean' | 'logq'
    strike_per_dim: bool = False   # strike = 1.0 * D (1d/nd files) instead of 1.0

    def strike(self, D):
        return 1.0 * D if self.strike_per_dim else 1.0


PROBLEMS = {
    # DeepBSDE.py:326-341
    "bsb": Problem("bsb", 0.0, True, 0.4, "bsb", 0.05, "sumsq"),
    # with_corr_high_dimension_pde.py:599-616
    "bsptest": Problem("bsptest", 0.05, True, 0.2, "bsb", 0.05, "sumsq"),
    # 1d_BSPDE_case.py:526-560 (strike = 1.0 * D, :160)
    "call1d": Problem("call1d", 0.01, True, 0.25, "ry", 0.01, "call_sum", True),
    # nd_BSPDE_case.py:517-539
    "callnd": Problem("callnd", 0.05, True, 0.2, "bsb", 0.05, "call_sum", True),
    # with_corr_high_dimension_pde.py:561-596
    "basket": Problem("basket", 0.05, True, 0.2, "ry", 0.05, "call_mean"),
    # hjb_implement.py:594-604
    "hjb": Problem("hjb", 0.0, False, float(torch.sqrt(torch.tensor(2.0))), "zsq", 1.0, "logq"),  # fp32 sqrt, as the reference
}


def mu_fn(p: Problem, X):
    return p.mu_c * X if p.mu_c != 0.0 else torch.zeros_like(X)


def sigma_fn(p: Problem, X):
    """(M,D,D) diffusion matrix, materialised exactly as the reference does."""
    if p.sigma_prop:
        return p.sigma_c * torch.diag_embed(X)
    if p.name == "hjb":
        return torch.sqrt(torch.tensor(2.0)) * torch.diag_embed(torch.ones_like(X))
    return p.sigma_c * torch.diag_embed(torch.ones_like(X))


def phi_fn(p: Problem, X, Y, Z):
    if p.phi == "bsb":
        return p.phi_c * (Y - torch.sum(X * Z, dim=1, keepdim=True))
    if p.phi == "ry":
        return p.phi_c * Y
    if p.phi == "zsq":
        return torch.sum(Z ** 2, dim=1, keepdim=True)
    raise ValueError(p.phi)


def g_fn(p: Problem, X, D):
    if p.g == "sumsq":
        return torch.sum(X ** 2, dim=1, keepdim=True)
    if p.g == "call_sum":
        return torch.maximum(torch.sum(X, dim=1, keepdim=True) - p.strike(D), torch.tensor(0.0, dtype=X.dtype))
    if p.g == "call_mean":
        return torch.maximum(torch.mean(X, dim=1, keepdim=True) - p.strike(D), torch.tensor(0.0, dtype=X.dtype))
    if p.g == "logq":
        return torch.log(0.5 + 0.5 * torch.sum(X ** 2, dim=1, keepdim=True))
    raise ValueError(p.g)


# --------------------------------------------------------------------------------------------
# solver pieces
# --------------------------------------------------------------------------------------------
def net_u(model, t, X):
    u = model(torch.cat((t, X), 1))
    Du = torch.autograd.grad(u, X, grad_outputs=torch.ones_like(u), allow_unused=True,
                             retain_graph=True, create_graph=True)[0]
    return u, Du


def loss_function(model, prob: Problem, t, W, Xi, M, N, D, squeeze_quirk: bool = True,
                  return_Z: bool = False):
    """Sum-of-squares FBSDE residual.  `squeeze_quirk=True` keeps the reference's un-dimmed
    torch.squeeze (DeepBSDE.py:224) which, for D == 1 and M > 1, mixes paths;
    False uses the per-path dot product (what the CUDA path computes).  The two agree for D >= 2."""
    loss = 0
    Xs, Ys, Zs = [], [], []
    t0 = t[:, 0, :]
    W0 = W[:, 0, :]
    if Xi.shape[0] == 1:
        X0 = Xi.view(1, D).repeat(M, 1)
    else:
        X0 = Xi.view(M, D)
    Y0, Z0 = net_u(model, t0, X0)
    Xs.append(X0), Ys.append(Y0), Zs.append(Z0)
    for n in range(N):
        t1 = t[:, n + 1, :]
        W1 = W[:, n + 1, :]
        dW = (W1 - W0).unsqueeze(-1)
        sdw = torch.squeeze(torch.matmul(sigma_fn(prob, X0), dW), dim=-1)
        X1 = X0 + mu_fn(prob, X0) * (t1 - t0) + sdw
        if squeeze_quirk:
            sdw_y = torch.squeeze(torch.matmul(sigma_fn(prob, X0), dW))
        else:
            sdw_y = sdw
        Y1_tilde = Y0 + phi_fn(prob, X0, Y0, Z0) * (t1 - t0) + torch.sum(Z0 * sdw_y, dim=1, keepdim=True)
        Y1, Z1 = net_u(model, t1, X1)
        loss = loss + torch.sum((Y1 - Y1_tilde) ** 2)
        t0, W0, X0, Y0, Z0 = t1, W1, X1, Y1, Z1
        Xs.append(X0), Ys.append(Y0), Zs.append(Z0)
    gT = g_fn(prob, X1, D)
    DgT = torch.autograd.grad(gT, X1, grad_outputs=torch.ones_like(gT), allow_unused=True,
                              retain_graph=True, create_graph=True)[0]
    loss = loss + torch.sum((Y1 - gT) ** 2)
    loss = loss + torch.sum((Z1 - DgT) ** 2)
    X = torch.stack(Xs, dim=1)
    Y = torch.stack(Ys, dim=1)
    if return_Z:
        return loss, X, Y, Y[0, 0, 0], torch.stack(Zs, dim=1)
    return loss, X, Y, Y[0, 0, 0]


def fetch_minibatch(M, N, D, T, chol: Optional[np.ndarray] = None):
    """Host Brownian sampler on the NumPy *global* RNG, cumulative t and W, fp32 tensors."""
    Dt = np.zeros((M, N + 1, 1))
    DW = np.zeros((M, N + 1, D))
    dt = T / N
    Dt[:, 1:, :] = dt
    inc = np.sqrt(dt) * np.random.normal(size=(M, N, D))
    if chol is not None:
        inc = np.einsum('ij,mnj->mni', chol, inc)
    DW[:, 1:, :] = inc
    t = torch.from_numpy(np.cumsum(Dt, axis=1)).float()
    W = torch.from_numpy(np.cumsum(DW, axis=1)).float()
    return t, W


class OracleSolver:
    """Minimal stand-in for a reference FBSNN subclass, CPU only."""

    def __init__(self, problem: str, Xi, T, M, N, D, layers, mode, activation,
                 chol: Optional[np.ndarray] = None, squeeze_quirk: bool = True, dtype=torch.float32):
        self.prob = PROBLEMS[problem]
        self.T, self.M, self.N, self.D = T, M, N, D
        self.dtype = dtype
        self.Xi = torch.from_numpy(np.asarray(Xi)).to(dtype)
        self.Xi.requires_grad = True
        self.model = build_model(layers, mode, activation).to(dtype)
        self.chol = chol
        self.squeeze_quirk = squeeze_quirk
        self.optimizer = None

    def fetch_minibatch(self):
        t, W = fetch_minibatch(self.M, self.N, self.D, self.T, self.chol)
        return t.to(self.dtype), W.to(self.dtype)

    def loss_function(self, t, W, Xi=None, return_Z=False):
        Xi = self.Xi if Xi is None else Xi
        return loss_function(self.model, self.prob, t, W, Xi, t.shape[0], self.N, self.D,
                             self.squeeze_quirk, return_Z)

    def make_optimizer(self, lr):
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)

    def train_step(self, t, W, clip: Optional[float] = None):
        """zero_grad -> loss -> backward -> [clip_grad_norm_] -> Adam.step; returns (loss, Y0)."""
        self.optimizer.zero_grad()
        loss, _, _, y0 = self.loss_function(t, W)
        loss.backward()
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=clip)
        self.optimizer.step()
        return float(loss.detach()), float(y0.detach())

    def grads(self, t, W):
        self.model.zero_grad()
        loss, X, Y, y0, Z = self.loss_function(t, W, return_Z=True)
        loss.backward()
        return (loss.detach(), X.detach(), Y.detach(), Z.detach(),
                {k: p.grad.detach().clone() for k, p in self.model.named_parameters()})


def bsb_exact(t, X, T, r=0.05, sigma=0.4):
    """Closed form endorsed by the reference, DeepBSDE.py:345-349."""
    return np.exp((r + sigma ** 2) * (T - t)) * np.sum(X ** 2, axis=-1, keepdims=True)


# --------------------------------------------------------------------------------------------
# Heston 2-factor FBSNN (heston_dnnpde.py:519-659) -- SURVEY.md section 8f row 4
# --------------------------------------------------------------------------------------------
class HestonOracle:
    """Restatement of HestonFBSNN: the base class is built with D = 1 (one Brownian driver, layers[0] inputs), then
    the input layer(s) are swapped for 3 inputs (t, S, v) and every matrix re-initialised with xavier(gain 0.5),
    biases zeroed (:533-585) -- same module construction order as the reference, hence the same torch RNG stream."""

    def __init__(self, Xi, T, M, N, layers, mode, activation, kappa=2.0, theta=0.2, sigma=0.3, rho=0.8, v0=0.2,
                 payoff_type="discontinuous", dtype=torch.float32):
        self.T, self.M, self.N, self.D = T, M, N, 1
        self.kappa, self.theta, self.sigma, self.rho, self.v0 = kappa, theta, sigma, rho, v0
        self.payoff_type, self.strike, self.dtype = payoff_type, 1.0, dtype
        self.Xi = torch.from_numpy(np.asarray(Xi)).to(dtype)
        self.Xi.requires_grad = True
        self.model = build_model(layers, mode, activation)
        if mode == "FC":
            self.model[0] = nn.Linear(3, layers[1])
        else:
            self.model.layer1 = nn.Linear(3, layers[1])
            self.model.layer2_input = nn.Linear(3, layers[2])
            if len(layers) >= 5:
                self.model.layer3_input = nn.Linear(3, layers[3])
            if len(layers) == 6:
                self.model.layer4_input = nn.Linear(3, layers[4])
        for p in self.model.parameters():
            if len(p.shape) > 1:
                nn.init.xavier_uniform_(p, gain=0.5)
            else:
                nn.init.zeros_(p)
        self.model = self.model.to(dtype)
        self.chol = None
        self.optimizer = None

    def fetch_minibatch(self):
        t, W = fetch_minibatch(self.M, self.N, 1, self.T, np.eye(1))      # :309-343 (Cholesky of the 1x1 identity)
        return t.to(self.dtype), W.to(self.dtype)

    def g_tf(self, S):
        if self.payoff_type == "discontinuous":
            return torch.maximum(S - self.strike, torch.tensor(0.0, dtype=S.dtype))
        return (S - self.strike) / (1 + torch.exp(-10.0 * (S - self.strike)))

    def net_u(self, t, X):
        S, v = X[:, 0:1], X[:, 1:2]
        u = torch.clamp(self.model(torch.cat((t, S, v), dim=1)), min=0.0)
        Du = torch.autograd.grad(u, (S, v), grad_outputs=torch.ones_like(u), create_graph=True, retain_graph=True)
        return u, Du[0], Du[1]

    def mu_tf(self, X):
        S, v = X[:, 0:1], X[:, 1:2]
        return torch.cat([0.05 * S, self.kappa * (self.theta - v)], dim=1).clamp(-100, 100)

    def sigma_tf(self, X):
        S, v = X[:, 0:1], X[:, 1:2]
        sS = torch.sqrt(torch.clamp(v, min=1e-8)) * S
        sv = self.sigma * torch.sqrt(torch.clamp(v, min=1e-8))
        m = torch.zeros((S.shape[0], 2, 2), dtype=X.dtype)
        m[:, 0, 0] = sS.squeeze()
        m[:, 1, 1] = sv.squeeze()
        m[:, 0, 1] = self.rho * sv.squeeze()
        m[:, 1, 0] = self.rho * sS.squeeze()
        return m.clamp(-100, 100)

    def loss_function(self, t, W, Xi=None, return_Z=False):
        Xi = self.Xi if Xi is None else Xi
        M = t.shape[0]
        loss = 0
        Xs, Ys, Zs = [], [], []
        t0, W0 = t[:, 0, :], W[:, 0, :]
        S0 = Xi[:, 0:1].repeat(M, 1) if Xi.shape[0] == 1 else Xi[:, 0:1]
        X0 = torch.cat([S0, torch.full((M, 1), self.v0, dtype=S0.dtype)], dim=1)
        Y0, ZS0, Zv0 = self.net_u(t0, X0)
        Xs.append(X0), Ys.append(Y0), Zs.append(torch.cat([ZS0, Zv0], 1))
        for n in range(self.N):
            t1, W1 = t[:, n + 1, :], W[:, n + 1, :]
            dW = W1 - W0
            sig = self.sigma_tf(X0)
            X1 = X0 + self.mu_tf(X0) * (t1 - t0) + torch.einsum('mij,mj->mi', sig, dW)   # dW (M,1) broadcasts over j
            Y1_tilde = Y0 + 0.05 * Y0 * (t1 - t0) + torch.sum(
                ZS0 * torch.sum(sig[:, 0, :] * dW, dim=1, keepdim=True) +
                Zv0 * torch.sum(sig[:, 1, :] * dW, dim=1, keepdim=True), dim=1, keepdim=True)
            Y1, ZS1, Zv1 = self.net_u(t1, X1)
            loss = loss + torch.sum((Y1 - Y1_tilde) ** 2)
            t0, W0, X0, Y0, ZS0, Zv0 = t1, W1, X1, Y1, ZS1, Zv1
            Xs.append(X0), Ys.append(Y0), Zs.append(torch.cat([ZS0, Zv0], 1))
        S1 = X1[:, 0:1]
        gT = self.g_tf(S1)
        DgT = torch.autograd.grad(gT, S1, grad_outputs=torch.ones_like(gT), allow_unused=True, retain_graph=True,
                                  create_graph=True)[0]
        loss = loss + torch.sum((Y1 - self.g_tf(X1[:, 0:1])) ** 2) + torch.sum((ZS1 - DgT) ** 2)
        X, Y = torch.stack(Xs, dim=1), torch.stack(Ys, dim=1)
        if return_Z:
            return loss, X, Y, Y[0, 0, 0], torch.stack(Zs, dim=1)
        return loss, X, Y, Y[0, 0, 0]

    make_optimizer = OracleSolver.make_optimizer
    train_step = OracleSolver.train_step
    grads = OracleSolver.grads
