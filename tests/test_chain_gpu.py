"""Layer-chained sweep kernels (csrc/gemm_chain.cuh) against (i) the per-layer launches on identical inputs, array by
array, and (ii) the reference fixtures, with the chained dispatch forced on (`fbsnn_set_option("chain", 2)`).

The chained sweeps do the same arithmetic per element as the per-layer epilogues (same k order inside an MMA
accumulator, same activation code); what differs is the order of the column sums / head dot products, so the two
dispatches agree to fp32 rounding (3xTF32) -- tolerances below.
"""
import ctypes

import numpy as np
import pytest
import torch

from tests import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tmem", "smem", "pair"])
def chain_forced(request):
    """Chained dispatch forced on.  "tmem" = chaint_kernel (the default: operand of the next MMA in tensor memory), "smem" =
    chain_kernel (operand in shared memory), "pair" = the cta_group::2 kernel where the shape allows it (3xTF32, widths
    multiples of 64)."""
    import dnnpde_b200 as pde
    lib = pde._lib.load()
    old = lib.fbsnn_set_option(b"chain", 2)
    old_ta = lib.fbsnn_set_option(b"chain_ta", 2 if request.param == "tmem" else 0)
    old_pair = lib.fbsnn_set_option(b"chain_pair", 1 if request.param == "pair" else 0)
    yield lib
    lib.fbsnn_set_option(b"chain", old)
    lib.fbsnn_set_option(b"chain_ta", old_ta)
    lib.fbsnn_set_option(b"chain_pair", old_pair)


def _grads(sol):
    return {n: sol._fp.grad[sol._fp.offsets[n]:sol._fp.offsets[n] + p.numel()].clone()
            for n, p in sol.model.named_parameters()}


def _both(sol, lib, t, W):
    out = {}
    for mode in (0, 2):
        lib.fbsnn_set_option(b"chain", mode)
        loss, X, Y, Z, _ = sol.loss_grad_flat(t, W, want_Z=True)
        torch.cuda.synchronize()
        out[mode] = dict(loss=float(loss), Y=Y.clone(), Z=Z.clone(), g=_grads(sol))
    return out


SHAPES = [
    # D, M, N, layers, act, problem
    (100, 40, 50, [101, 256, 256, 256, 256, 1], "Sine", "bsb"),      # the benchmarked network, ragged last tile
    (100, 3, 50, [101, 256, 256, 256, 256, 1], "Sine", "bsb"),       # two tiles
    (100, 700, 50, [101, 256, 256, 256, 256, 1], "Sine", "bsb"),     # 279 tiles: several tiles per CTA (pair: 140 pairs of 74)
    (100, 2000, 50, [101, 256, 256, 256, 256, 1], "Sine", "bsb"),    # 797 tiles, 5-6 per CTA on a busy chip: the regime in which a
                                                                     # half-team of chaint_kernel can run three chunks ahead of the other
    (126, 90, 30, [127, 128, 192, 64, 1], "Tanh", "bsb"),            # pair-eligible mixed widths, 22 tiles
    (10, 300, 7, [11, 64, 128, 64, 1], "Tanh", "bsb"),               # ldx = 32, mixed widths
    (20, 77, 12, [21, 96, 96, 1], "ReLU", "hjb"),                    # two layers, odd chunk counts, |Z|^2 driver
    (6, 50, 9, [7, 128, 1], "Sine", "bsb"),                          # one hidden layer: no B sweep
    (100, 64, 20, [101, 192, 256, 64, 224, 160, 1], "Sine", "hjb"),  # five layers
    # widths that are multiples of 64 all the way (the TMEM-operand kernel's domain) with 1, 2 and 5 hidden layers, 2 / 4 / 6 / 8
    # chunks per link (with and without parked accumulator chunks), several tiles per CTA
    (100, 50, 20, [101, 256, 1], "Sine", "bsb"),
    (60, 900, 30, [61, 128, 64, 1], "Tanh", "bsb"),
    (100, 500, 40, [101, 64, 128, 256, 192, 64, 1], "Sine", "hjb"),   # (smooth activation: a ReLU unit at its kink may flip between dispatches)
]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"D{s[0]}_M{s[1]}_L{len(s[3]) - 2}_{s[4]}" for s in SHAPES])
def test_chained_sweeps_match_per_layer(shape, precision, chain_forced):
    import dnnpde_b200 as pde
    lib = chain_forced
    D, M, N, layers, act, problem = shape
    torch.manual_seed(7)
    np.random.seed(7)
    if problem == "hjb":
        sol = pde.HamiltonJacobiBellman(np.zeros((1, D)), 1.0, M, N, D, layers, "FC", act, precision=precision)
    else:
        sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * (D // 2))[None, :], 1.0, M, N, D, layers, "FC", act,
                                         precision=precision)
    t, W = sol.fetch_minibatch()
    r = _both(sol, lib, t, W)
    # 3xTF32: both dispatches are fp32-grade, element-wise agreement to rounding.  Single-pass TF32 rounds every operand to
    # 10 bits (and layers the per-layer dispatch cannot tile run in fp32 there), so a ReLU unit can flip across its kink:
    # the bound is the variant's own accuracy, in the rel-L2 form of TOL_TF32
    a, b = r[0], r[2]
    assert np.isfinite(b["loss"])
    if precision == "tf32x3":
        tol = 2e-5
        assert abs(a["loss"] - b["loss"]) <= tol * abs(a["loss"])
        assert float((a["Y"] - b["Y"]).abs().max()) <= tol * float(a["Y"].abs().max())
        assert float((a["Z"] - b["Z"]).abs().max()) <= tol * float(a["Z"].abs().max()) + 1e-12
        for n in a["g"]:
            ga, gb = a["g"][n], b["g"][n]
            assert float((ga - gb).abs().max()) <= 5 * tol * float(ga.abs().max()) + 1e-12, n
    else:
        rel = lambda x, y: float(torch.linalg.norm((x - y).double()) / (torch.linalg.norm(x.double()) + 1e-30))
        assert abs(a["loss"] - b["loss"]) <= 1e-2 * abs(a["loss"])
        assert rel(a["Y"], b["Y"]) <= 1e-2
        assert rel(a["Z"], b["Z"]) <= 5e-2
        for n in a["g"]:
            assert rel(a["g"][n], b["g"][n]) <= 8e-2, n


def test_chained_dispatch_is_what_runs(chain_forced):
    """With the option forced the 4x256 network must take 4 chained launches + the weight-gradient contractions (one batched
    launch at this batch size)."""
    import dnnpde_b200 as pde
    lib = chain_forced
    torch.manual_seed(0)
    np.random.seed(0)
    D = 100
    sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * 50)[None, :], 1.0, 8, 50, D, [D + 1] + 4 * [256] + [1], "FC",
                                     "Sine", precision="tf32x3")
    t, W = sol.fetch_minibatch()
    lib.fbsnn_dense_timing(1)
    sol.loss_grad_flat(t, W)
    torch.cuda.synchronize()
    tags = []
    out4 = (ctypes.c_double * 4)()
    i = 0
    while True:
        tag = lib.fbsnn_dense_timing_entry(i, out4)
        if tag is None:
            break
        tags.append(tag.decode())
        i += 1
    lib.fbsnn_dense_timing(0)
    # small batch: the four weight-gradient contractions share one launch (gemm_tc2g_batched_kernel)
    assert tags in (["F*", "A*", "T*", "B*", "G"], ["F*", "A*", "T*", "B*", "G", "G", "G", "G"]), tags


FC_CASES = [n for n in gu.solver_cases() if "_fc_" in n]


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("name", FC_CASES)
def test_chained_sweeps_match_reference(name, precision, chain_forced):
    """The reference fixtures through the chained dispatch, at each variant's stated tolerance."""
    from tests import parity_util as pu
    from tests.test_parity_gpu import TOL_TF32, TOL_X3
    tolerances = TOL_X3 if precision == "tf32x3" else TOL_TF32
    g, meta = gu.load(name)
    sol, oracle = pu.build_cuda_solver(meta, g, precision=precision)
    errs = pu.single_eval_errors(sol, oracle, g, meta)
    for k, v in errs.items():
        if k in tolerances:
            assert v <= tolerances[k], (name, k, v, errs)
    if pu.has_squeeze_quirk(meta):
        return
    terr = pu.train_trace_errors(sol, g, meta)
    for k, v in terr.items():
        assert v <= tolerances[k], (name, k, v, terr)


def test_chained_predict_and_net_u(chain_forced):
    """Forward-only plan (no s / seeds): predict() and net_u() through the chained F and A sweeps."""
    import dnnpde_b200 as pde
    lib = chain_forced
    torch.manual_seed(3)
    np.random.seed(3)
    D, M, N = 100, 33, 50
    sol = pde.BlackScholesBarenblatt(np.array([1.0, 0.5] * 50)[None, :], 1.0, M, N, D, [D + 1] + 4 * [256] + [1], "FC",
                                     "Sine", precision="tf32x3")
    t, W = sol.fetch_minibatch()
    res = {}
    for mode in (0, 2):
        lib.fbsnn_set_option(b"chain", mode)
        X, Y = sol.predict(sol.Xi.detach(), t, W)
        u, du = sol.net_u(t[:, 5, :], X[:, 5, :])
        torch.cuda.synchronize()
        res[mode] = (Y.clone(), u.clone(), du.clone())
    for a, b in zip(res[0], res[2]):
        assert float((a - b).abs().max()) <= 2e-5 * float(a.abs().max())
