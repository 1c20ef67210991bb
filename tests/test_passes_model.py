"""The analytic F/A/T/B/G pass algebra (tests/passes_model.py, which the CUDA kernels mirror) against
autograd double-backward of the oracle, in float64 so that any algebra slip shows up at 1e-9."""
import numpy as np
import pytest
import torch

from oracle import fbsnn_oracle as orc
from tests import passes_model as pm

CASES = [
    ("bsb", "FC", "Sine", 6, [7, 16, 16, 16, 16, 1]),
    ("bsb", "FC", "Tanh", 6, [7, 12, 20, 8, 1]),          # unequal hidden widths are legal for FC
    ("bsptest", "FC", "ReLU", 5, [6, 16, 16, 1]),
    ("call1d", "FC", "Sine", 1, [2, 16, 16, 1]),
    ("callnd", "FC", "Tanh", 4, [5, 16, 16, 16, 1]),
    ("basket", "Naisnet", "Sine", 5, [6, 16, 16, 16, 16, 1]),
    ("basket", "Naisnet", "ReLU", 4, [5, 16, 16, 16, 1]),
    ("hjb", "Naisnet", "Tanh", 6, [7, 16, 16, 1]),
    ("hjb", "Naisnet", "Sine", 3, [4, 8, 8, 8, 8, 1]),
]


@pytest.mark.parametrize("problem,mode,act,D,layers", CASES)
def test_pass_algebra_matches_autograd(problem, mode, act, D, layers):
    torch.manual_seed(7)
    np.random.seed(11)
    M, N, T = 5, 6, 1.0
    Xi = np.random.uniform(0.6, 1.4, size=(1, D))
    sol = orc.OracleSolver(problem, Xi, T, M, N, D, layers, mode, act, squeeze_quirk=False, dtype=torch.float64)
    if mode == "Naisnet":
        # exercise both branches of the norm test in the NAIS projection
        with torch.no_grad():
            sol.model.layer2.weight.mul_(0.05)
    t, W = sol.fetch_minibatch()
    loss, X, Y, Z, grads = sol.grads(t, W)
    params = {k: p.detach() for k, p in sol.model.named_parameters()}
    prob = orc.PROBLEMS[problem]
    loss2, X2, Y2, Z2, grads2 = pm.full_step(params, mode, act, prob, t, W, sol.Xi.detach(), prob.strike(D))
    assert torch.allclose(X2, X, rtol=1e-12, atol=1e-12)
    assert torch.allclose(Y2, Y[:, :, 0], rtol=1e-10, atol=1e-12)
    assert torch.allclose(Z2, Z, rtol=1e-10, atol=1e-12)
    assert abs(float(loss2 - loss)) <= 1e-10 * abs(float(loss))
    assert set(grads2) == set(grads)
    for k in grads:
        scale = float(grads[k].abs().max()) + 1e-30
        err = float((grads2[k] - grads[k]).abs().max())
        assert err <= 1e-8 * scale + 1e-12, (k, err, scale)


def test_nais_projection_backward_both_branches():
    torch.manual_seed(0)
    for scale in (1.0, 0.02):
        W = (torch.randn(12, 12, dtype=torch.float64) * scale).requires_grad_(True)
        B, ctx = pm.nais_matrix(W)
        G = torch.randn_like(B)
        (B * G).sum().backward()
        assert ctx[3] == (scale == 1.0)
        got = pm.nais_matrix_backward(W.detach(), tuple(c.detach() if torch.is_tensor(c) else c for c in ctx), G)
        assert torch.allclose(got, W.grad, rtol=1e-10, atol=1e-12)
