"""Host model of the CUDA pass structure (TEST INFRASTRUCTURE).

The CUDA path never calls autograd: it evaluates the network row-wise with four sweeps
(forward F, input-adjoint A, tangent T, backward B) plus weight-gradient contractions G, all
written out analytically.  This file states exactly those passes with dense torch ops (any dtype,
CPU) so that the algebra can be checked against autograd double-backward in the no-GPU test
suite (tests/test_passes_model.py) before it is trusted on the device.  Array names match DESIGN.md.
"""
from __future__ import annotations

import math

import torch


def act_triplet(kind: str, z):
    """activation value, first and second derivative."""
    if kind == "Sine":
        g = torch.sin(z)
        return g, torch.cos(z), -g
    if kind == "ReLU":
        return torch.relu(z), (z > 0).to(z.dtype), torch.zeros_like(z)
    if kind == "Tanh":
        g = torch.tanh(z)
        a = 1 - g * g
        return g, a, -2 * g * a
    raise ValueError(kind)


def nais_matrix(Wl, eps=0.01):
    """B = -(s * W^T W + eps I) and the pieces needed for its backward."""
    delta = 1 - 2 * eps
    R = Wl.t() @ Wl
    n = torch.linalg.norm(R)
    scaled = bool(n > delta)
    s = math.sqrt(delta) / torch.sqrt(n) if scaled else torch.ones((), dtype=Wl.dtype)
    B = -(s * R + eps * torch.eye(R.shape[0], dtype=Wl.dtype))
    return B, (R, n, s, scaled)


def nais_matrix_backward(Wl, ctx, Bbar):
    R, n, s, scaled = ctx
    Abar = -Bbar
    Rbar = s * Abar
    if scaled:
        Rbar = Rbar - 0.5 * s * torch.sum(Abar * R) / (n * n) * R
    return Wl @ (Rbar + Rbar.t())


def unpack(params: dict, mode: str):
    """-> list of per-layer dicts {Wm, Win, bias} for hidden layers 1..L, and (w_out, b_out)."""
    layers = []
    if mode == "FC":
        idx = sorted({int(k.split(".")[0]) for k in params})
        for i in idx[:-1]:
            layers.append(dict(Wm=params[f"{i}.weight"], Win=None, bias=params[f"{i}.bias"], key=str(i)))
        wout, bout = params[f"{idx[-1]}.weight"], params[f"{idx[-1]}.bias"]
        return layers, wout, bout, str(idx[-1])
    nb = sum(1 for k in params if k.endswith("_input.weight"))
    layers.append(dict(Wm=params["layer1.weight"], Win=None, bias=params["layer1.bias"], key="layer1"))
    for k in range(2, nb + 2):
        B, ctx = nais_matrix(params[f"layer{k}.weight"])
        layers.append(dict(Wm=B, Win=params[f"layer{k}_input.weight"],
                           bias=params[f"layer{k}.bias"] + params[f"layer{k}_input.bias"],
                           key=f"layer{k}", ctx=ctx, Wraw=params[f"layer{k}.weight"]))
    last = f"layer{nb + 2}"
    return layers, params[f"{last}.weight"], params[f"{last}.bias"], last


def forward_adjoint(params, mode, act, x):
    """F and A sweeps on rows x (R, d).  Returns u (R,), Dufull (R, d) and the saved arrays."""
    layers, wout, bout, _ = unpack(params, mode)
    res = mode != "FC"
    L = len(layers)
    h_prev = x
    sv = dict(x=x, g=[None] * (L + 1), a=[None] * (L + 1), h=[None] * (L + 1), delta=[None] * (L + 1),
              s=[None] * (L + 1))
    sv["h"][0] = x
    for l in range(1, L + 1):
        ly = layers[l - 1]
        z = h_prev @ ly["Wm"].t() + ly["bias"]
        if ly["Win"] is not None:
            z = z + x @ ly["Win"].t()
        g, a, c = act_triplet(act, z)
        h = g + h_prev if (res and l >= 2) else g
        sv["g"][l], sv["a"][l], sv["h"][l] = g, a, h
        sv["c%d" % l] = c
        h_prev = h
    u = h_prev @ wout.t().reshape(-1) + bout.reshape(())
    ht = wout.reshape(1, -1).expand(x.shape[0], -1)
    du = torch.zeros_like(x)
    for l in range(L, 0, -1):
        ly = layers[l - 1]
        delta = ht * sv["a"][l]
        sv["delta"][l] = delta
        sv["s"][l] = ht * sv["c%d" % l]
        if l > 1:
            nxt = delta @ ly["Wm"]
            if res:
                nxt = nxt + ht
                du = du + delta @ ly["Win"]
            ht = nxt
        else:
            du = du + delta @ ly["Wm"]
    return u, du, sv


def tangent_backward_wgrad(params, mode, act, sv, ybar, V):
    """T and B sweeps and the G contractions.  ybar (R,), V (R, d) with V[:, 0] = 0.  -> dict of grads."""
    layers, wout, bout, lastkey = unpack(params, mode)
    res = mode != "FC"
    L = len(layers)
    x = sv["x"]
    hd = [None] * (L + 1)
    zz = [None] * (L + 1)
    for l in range(1, L + 1):
        ly = layers[l - 1]
        if l == 1:
            dbar = V @ ly["Wm"].t()
        else:
            dbar = hd[l - 1] @ ly["Wm"].t()
            if ly["Win"] is not None:
                dbar = dbar + V @ ly["Win"].t()
        gd = dbar * sv["a"][l]
        zz[l] = dbar * sv["s"][l]
        hd[l] = gd + hd[l - 1] if (res and l >= 2) else gd
    grads = {}
    grads[f"{lastkey}.weight"] = (hd[L].sum(0) + (ybar[:, None] * sv["h"][L]).sum(0)).reshape(wout.shape)
    grads[f"{lastkey}.bias"] = ybar.sum().reshape(bout.shape)
    hb = ybar[:, None] * wout.reshape(1, -1)
    for l in range(L, 0, -1):
        ly = layers[l - 1]
        zbar = hb * sv["a"][l] + zz[l]
        key = ly["key"]
        if l == 1:
            grads[f"{key}.weight"] = zbar.t() @ x + sv["delta"][1].t() @ V
            grads[f"{key}.bias"] = zbar.sum(0)
        else:
            Wm_bar = zbar.t() @ sv["h"][l - 1] + sv["delta"][l].t() @ hd[l - 1]
            if ly["Win"] is not None:
                grads[f"{key}.weight"] = nais_matrix_backward(ly["Wraw"], ly["ctx"], Wm_bar)
                grads[f"{key}_input.weight"] = zbar.t() @ x + sv["delta"][l].t() @ V
                grads[f"{key}.bias"] = zbar.sum(0)
                grads[f"{key}_input.bias"] = zbar.sum(0)
            else:
                grads[f"{key}.weight"] = Wm_bar
                grads[f"{key}.bias"] = zbar.sum(0)
            nxt = zbar @ ly["Wm"]
            hb = nxt + hb if res else nxt
    return grads


# ---------------------------------------------------------------------------------------------------------
# problem-side pieces (closed enumeration, mirrors the device loss kernel)
# ---------------------------------------------------------------------------------------------------------
def advance_paths(prob, t, W, Xi):
    """Euler-Maruyama recursion; returns X (M, N+1, D) and sdw (M, N, D) = sigma(X_n) dW_n."""
    M, N1, D = W.shape
    X = torch.empty(M, N1, D, dtype=W.dtype)
    sdw = torch.empty(M, N1 - 1, D, dtype=W.dtype)
    X[:, 0] = Xi.expand(M, D)
    for n in range(N1 - 1):
        dt = t[:, n + 1] - t[:, n]
        dW = W[:, n + 1] - W[:, n]
        sig = prob.sigma_c * X[:, n] if prob.sigma_prop else prob.sigma_c * torch.ones_like(X[:, n])
        sdw[:, n] = sig * dW
        X[:, n + 1] = X[:, n] + (prob.mu_c * X[:, n]) * dt + sdw[:, n]
    return X, sdw


def loss_and_seeds(prob, t, X, sdw, Y, Z, strike):
    """loss (sum of squares) and its derivatives ybar = dL/dY (M,N+1), zbar = dL/dZ (M,N+1,D)."""
    M, N1, D = X.shape
    N = N1 - 1
    dt = (t[:, 1:, 0] - t[:, :-1, 0])
    Xn, Yn, Zn = X[:, :-1], Y[:, :-1], Z[:, :-1]
    if prob.phi == "bsb":
        phi = prob.phi_c * (Yn - (Xn * Zn).sum(-1))
        phi_y = torch.full_like(Yn, prob.phi_c)
        phi_z = -prob.phi_c * Xn
    elif prob.phi == "ry":
        phi = prob.phi_c * Yn
        phi_y = torch.full_like(Yn, prob.phi_c)
        phi_z = torch.zeros_like(Zn)
    else:
        phi = (Zn * Zn).sum(-1)
        phi_y = torch.zeros_like(Yn)
        phi_z = 2 * Zn
    e = Y[:, 1:] - (Yn + phi * dt + (Zn * sdw).sum(-1))
    XT, YT, ZT = X[:, -1], Y[:, -1], Z[:, -1]
    if prob.g == "sumsq":
        g, dg = (XT * XT).sum(-1), 2 * XT
    elif prob.g in ("call_sum", "call_mean"):
        base = XT.sum(-1) if prob.g == "call_sum" else XT.mean(-1)
        scale = 1.0 if prob.g == "call_sum" else 1.0 / D
        g = torch.clamp(base - strike, min=0)
        ind = (base > strike).to(X.dtype) + 0.5 * (base == strike).to(X.dtype)
        dg = (ind * scale)[:, None].expand(M, D)
    else:
        q = 0.5 + 0.5 * (XT * XT).sum(-1)
        g, dg = torch.log(q), XT / q[:, None]
    loss = (e * e).sum() + ((YT - g) ** 2).sum() + ((ZT - dg) ** 2).sum()
    ybar = torch.zeros_like(Y)
    zbar = torch.zeros_like(Z)
    ybar[:, 1:] += 2 * e
    ybar[:, :-1] += -2 * e * (1 + phi_y * dt)
    zbar[:, :-1] += (-2 * e)[..., None] * (phi_z * dt[..., None] + sdw)
    ybar[:, -1] += 2 * (YT - g)
    zbar[:, -1] += 2 * (ZT - dg)
    return loss, ybar, zbar


def full_step(params, mode, act, prob, t, W, Xi, strike):
    """Everything the fused train step computes before the optimizer: loss, X, Y, Z, grads."""
    M, N1, D = W.shape
    X, sdw = advance_paths(prob, t, W, Xi)
    x = torch.cat((t, X), dim=2).reshape(M * N1, D + 1)
    u, du, sv = forward_adjoint(params, mode, act, x)
    Y = u.reshape(M, N1)
    Z = du[:, 1:].reshape(M, N1, D)
    loss, ybar, zbar = loss_and_seeds(prob, t, X, sdw, Y, Z, strike)
    V = torch.cat((torch.zeros(M, N1, 1, dtype=W.dtype), zbar), dim=2).reshape(M * N1, D + 1)
    grads = tangent_backward_wgrad(params, mode, act, sv, ybar.reshape(-1), V)
    return loss, X, Y, Z, grads
