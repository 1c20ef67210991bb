"""Shared helpers for replaying tests/golden/*.npz (fixtures written by oracle/make_golden.py from the
unmodified reference).  Inputs are regenerated from the recorded seeds and verified by checksum."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from oracle import fbsnn_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def solver_cases():
    names = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n not in ("mc_pricer", "bsb100_train_api", "basket_pricer", "nd_schedule_trace", "bsb100_k30", "comparators"):
            names.append(n)
    return names


def path(name):
    return os.path.join(GOLDEN_DIR, name + ".npz")


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(g["meta"])) if "meta" in g.files else None
    return g, meta


def make_xi(kind, D):
    if kind == "bsb":
        return np.array([1.0, 0.5] * (D // 2) + [1.0] * (D % 2))[None, :]
    if kind == "ones":
        return np.ones((1, D))
    if kind == "zeros":
        return np.zeros((1, D))
    raise ValueError(kind)


def ref_correlation(kind, D):
    """The reference FBSNN's correlation generator, drawn from the NumPy global RNG in the constructor
    (with_corr_high_dimension_pde.py:187-212).  Restated for fixture replay only."""
    if kind in (None, "no_correlation"):
        return None
    a = np.random.randn(D, D)
    if kind == "restricted_random_correlation":
        a = np.abs(a)
    c = a @ a.T
    np.fill_diagonal(c, 1)
    d = np.sqrt(np.diag(c))
    c = c / np.outer(d, d)
    eps = 1e-6
    while not np.all(np.linalg.eigvals(c) > 0):
        c += eps * np.eye(D)
        eps *= 2
    return c


def rebuild_inputs(meta, g, squeeze_quirk=True):
    """Recreate (oracle solver with the reference's initial weights, corr matrix) under the fixture's seeds and
    verify the regenerated parameters against the fixture checksums."""
    torch.manual_seed(meta["torch_seed"])
    np.random.seed(meta["numpy_seed"])
    D = meta["D"]
    Xi = make_xi(meta["xi"], D)
    if meta["problem"] == "heston":
        sol = orc.HestonOracle(Xi, meta["T"], meta["M"], meta["N"], meta["layers"], meta["mode"], meta["act"],
                               payoff_type=meta["payoff"])
    else:
        sol = orc.OracleSolver(meta["problem"], Xi, meta["T"], meta["M"], meta["N"], D, meta["layers"],
                               meta["mode"], meta["act"], squeeze_quirk=squeeze_quirk)
    # the reference with_corr / hjb constructors draw the correlation matrix *after* building the network
    needs_corr = meta["file"] in ("with_corr_high_dimension_pde.py", "hjb_implement.py")
    corr = ref_correlation(meta["corr"], D) if needs_corr else None
    if corr is not None:
        assert np.allclose(corr, g["corr_matrix"], rtol=0, atol=0), "NumPy RNG stream drifted (corr matrix)"
        sol.chol = np.linalg.cholesky(corr)
    names = [k for k, _ in sol.model.named_parameters()]
    assert names == [str(s) for s in g["param_names"]], (names, g["param_names"])
    sums = np.array([float(p.detach().double().sum()) for _, p in sol.model.named_parameters()])
    assert np.allclose(sums, g["param_sum"], rtol=0, atol=1e-12), "torch init RNG stream drifted"
    return sol
