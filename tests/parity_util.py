"""Builds the CUDA solver for a golden fixture with the reference's initial weights and NumPy stream position,
and measures its deviation from the reference outputs stored in the fixture (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import numpy as np
import torch

import dnnpde_b200 as pde
from tests import golden_util as gu

CLASS_OF = {
    "bsb": pde.BlackScholesBarenblatt,
    "bsptest": pde.BSPDETestCase,
    "call1d": pde.CallOption1D,
    "callnd": pde.CallOptionND,
    "basket": pde.BasketCallOption,
    "hjb": pde.HamiltonJacobiBellman,
    "heston": pde.HestonFBSNN,
}


def build_cuda_solver(meta, g, precision="fp32"):
    """-> (cuda solver with the fixture's initial weights, oracle solver in the same state).  On return the NumPy
    global RNG sits exactly where the reference's was before its first fetch_minibatch()."""
    oracle = gu.rebuild_inputs(meta, g, squeeze_quirk=False)
    state = {k: v.detach().clone() for k, v in oracle.model.state_dict().items()}
    np.random.seed(meta["numpy_seed"])
    cls = CLASS_OF[meta["problem"]]
    D = meta["D"]
    Xi = gu.make_xi(meta["xi"], D)
    args = (Xi, meta["T"], meta["M"], meta["N"], D)
    kw = dict(precision=precision)
    if cls is pde.HestonFBSNN:
        sol = cls(*args, int(meta["N"] ** (1 / 5)), meta["layers"], meta["mode"], meta["act"],
                  payoff_type=meta["payoff"], **kw)
    elif cls in (pde.BlackScholesBarenblatt, pde.HamiltonJacobiBellman):
        sol = cls(*args, meta["layers"], meta["mode"], meta["act"], **kw)
    elif cls in (pde.CallOption1D, pde.CallOptionND):
        sol = cls(*args, None, meta["layers"], meta["mode"], meta["act"], **kw)
    else:
        sol = cls(*args, None, meta["layers"], meta["mode"], meta["act"], meta["corr"] or "no_correlation", **kw)
    if "corr_matrix" in g.files:
        assert np.array_equal(sol.correlation_matrix, g["corr_matrix"])
    sol.model.load_state_dict(state)
    assert sol._fp.is_intact()
    return sol, oracle


def has_squeeze_quirk(meta) -> bool:
    """D == 1 with M > 1 makes the reference's un-dimmed squeeze mix paths; the Heston file's
    loss_function has no such squeeze."""
    return meta["D"] == 1 and meta["M"] > 1 and meta["problem"] != "heston"


def unflatten_grads(sol):
    return {n: sol._fp.grad[o:o + p.numel()].view(p.shape).detach().cpu().numpy().copy()
            for (n, p), o in ((np_, sol._fp.offsets[np_[0]]) for np_ in sol.model.named_parameters())}


def single_eval_errors(sol, oracle, g, meta):
    """Relative deviations of one loss/gradient evaluation from the fixture (reference) -- or from the oracle
    with the per-path product for the D == 1, M > 1 quirk case."""
    t, W = sol.fetch_minibatch()
    loss, X, Y, Z, _ = sol.loss_grad_flat(t, W, want_Z=True)
    torch.cuda.synchronize()
    grads = unflatten_grads(sol)
    quirk = has_squeeze_quirk(meta)
    if quirk:
        ol, oX, oY, oZ, og = oracle.grads(t.cpu(), W.cpu())
        ref = dict(loss=float(ol), Y=oY[:, :, 0].numpy(), X_head=oX[:4].numpy(), Z_head=oZ[:4].numpy())
        ref_g = {k: v.numpy() for k, v in og.items()}
    else:
        ref = dict(loss=float(g["loss"]), Y=g["Y"], X_head=g["X_head"], Z_head=g["Z_head"])
        ref_g = {k[6:]: g[k] for k in g.files if k.startswith("grad::")}
    k = ref["X_head"].shape[0]
    out = {}
    out["loss_rel"] = abs(float(loss) - ref["loss"]) / abs(ref["loss"])
    Yc = Y[:, :, 0].cpu().numpy()
    out["Y_rel"] = float(np.abs(Yc - ref["Y"]).max() / (np.abs(ref["Y"]).max() + 1e-30))
    out["X_abs"] = float(np.abs(X[:k].cpu().numpy() - ref["X_head"]).max())
    Zc = Z[:k].cpu().numpy()
    out["Z_rel_l2"] = float(np.linalg.norm(Zc - ref["Z_head"]) / (np.linalg.norm(ref["Z_head"]) + 1e-30))
    worst, worst_name = 0.0, ""
    for name, rg in ref_g.items():
        e = float(np.abs(grads[name] - rg).max() / (np.abs(rg).max() + 1e-30))
        if e > worst:
            worst, worst_name = e, name
    out["grad_rel_max"] = worst
    out["grad_worst"] = worst_name
    if not quirk:
        gn = np.array([np.linalg.norm(grads[str(n)].astype(np.float64)) for n in g["param_names"]])
        out["gradnorm_rel"] = float(np.abs(gn - g["grad_norm"]).max() / (np.abs(g["grad_norm"]).max() + 1e-30))
    return out


def train_trace_errors(sol, g, meta):
    """K optimiser iterations through the public train() (which consumes the NumPy stream like the reference)."""
    sol.train(meta["K"], meta["lr"])
    tl, ty = sol.last_losses.astype(np.float64), sol.last_Y0.astype(np.float64)
    out = {}
    out["trace_loss_rel"] = float(np.max(np.abs(tl - g["trace_loss"]) / np.abs(g["trace_loss"])))
    out["trace_Y0_abs"] = float(np.max(np.abs(ty - g["trace_Y0"])))
    last = [k for k in g.files if k.startswith("final::")][0]
    w = dict(sol.model.named_parameters())[last[7:]].detach().cpu().numpy()
    out["final_w_rel"] = float(np.abs(w - g[last]).max() / np.abs(g[last]).max())
    return out
