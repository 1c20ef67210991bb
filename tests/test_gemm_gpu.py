"""Dense-layer GEMM kernels in isolation (through the C-ABI test hook): SIMT fp32 and tcgen05 TF32 against torch
fp32/fp64 matmul for the three operand layouts the sweeps use.  TF32 tolerance: 10-bit mantissas on both operands,
fp32 accumulation => |err| <= ~2^-10 * sum|a||b| per entry; asserted as 2e-3 of the row-wise |A||B| bound."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LAYOUTS = [(1, 1), (1, 0), (0, 0)]


def run_gemm(a_kc, b_kc, use_tc, M, N, K, seed=0):
    import dnnpde_b200 as pde
    lib = pde._lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((M, K) if a_kc else (K, M), device="cuda", generator=g)
    B = torch.randn((N, K) if b_kc else (K, N), device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.fbsnn_debug_gemm(a_kc, b_kc, use_tc, M, N, K, ctypes.c_void_p(A.data_ptr()), A.shape[1],
                              ctypes.c_void_p(B.data_ptr()), B.shape[1], ctypes.c_void_p(C.data_ptr()), N,
                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    pde._lib.check(rc, "fbsnn_debug_gemm")
    torch.cuda.synchronize()
    Am = A.double() if a_kc else A.double().t()
    Bm = B.double().t() if b_kc else B.double()
    ref = Am @ Bm
    bound = Am.abs() @ Bm.abs()
    return C, ref, bound


@pytest.mark.parametrize("a_kc,b_kc", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (5100, 104, 256), (77, 64, 101), (256, 256, 5100)])
def test_simt_gemm(a_kc, b_kc, M, N, K):
    C, ref, bound = run_gemm(a_kc, b_kc, 0, M, N, K)
    assert torch.isfinite(C).all()
    assert ((C.double() - ref).abs() <= 2e-6 * bound + 1e-6).all()


@pytest.mark.parametrize("a_kc,b_kc", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 64, 256), (300, 256, 256), (5100, 128, 256),
                                   (40000, 256, 128)])
def test_tcgen05_gemm(a_kc, b_kc, M, N, K):
    if not a_kc:
        M, K = (256 if M % 128 else M), max(K, 1000 if K == 256 else K)   # weight-gradient shape: K = rows, free
    C, ref, bound = run_gemm(a_kc, b_kc, 1, M, N, K)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs()
    assert (err <= 2e-3 * bound + 1e-5).all(), float((err / (bound + 1e-9)).max())
    # and it is genuinely TF32 arithmetic, not fp32 (guards against a silent SIMT dispatch)
    assert float(err.max()) > 1e-6


@pytest.mark.parametrize("a_kc,b_kc", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (300, 256, 256), (5100, 128, 256), (40000, 256, 128)])
def test_tcgen05_3xtf32_gemm_is_fp32_grade(a_kc, b_kc, M, N, K):
    """operands split hi + lo in shared memory, three MMAs per k-step: error ~2^-21 of sum|a||b|, i.e. fp32-grade."""
    if not a_kc:
        M, K = (256 if M % 128 else M), max(K, 1000 if K == 256 else K)
    C, ref, bound = run_gemm(a_kc, b_kc, 2, M, N, K)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs()
    assert (err <= 2e-6 * bound + 1e-6).all(), float((err / (bound + 1e-9)).max())


@pytest.mark.parametrize("a_kc,b_kc", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (128, 256, 64), (300, 256, 256), (5100, 128, 256), (700, 64, 96),
                                   (40000, 256, 128), (100000, 192, 256)])
def test_tcgen05_cta_pair_gemm_is_fp32_grade(a_kc, b_kc, M, N, K):
    """cta_group::2 form of the 3xTF32 kernel (two CTAs on one 256-row UMMA tile, half of B staged per CTA): same
    fp32-grade bound, ragged row counts (last pair half empty) and every column count the sweeps use."""
    if not a_kc:
        M, K = 256, max(K, 1000 if K == 256 else K) + (7 if K == 128 else 0)   # weight gradients: out = 256, K = rows
    C, ref, bound = run_gemm(a_kc, b_kc, 3, M, N, K)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs()
    assert (err <= 2e-6 * bound + 1e-6).all(), float((err / (bound + 1e-9)).max())


@pytest.mark.parametrize("b_kc", [1, 0])
@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (300, 256, 256), (5100, 128, 128), (700, 64, 96), (40000, 256, 256)])
def test_tcgen05_cta_pair_presplit_weights(b_kc, M, N, K):
    """the sweeps' pair form: W_hi / W_lo twins loaded by TMA into one stage, only A split in-kernel (SPLIT = 3)."""
    C, ref, bound = run_gemm(1, b_kc, 4, M, N, K)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs()
    assert (err <= 2e-6 * bound + 1e-6).all(), float((err / (bound + 1e-9)).max())


@pytest.mark.parametrize("N,K", [(256, 64), (256, 1000), (128, 4100), (64, 300), (192, 20000), (256, 200003)])
def test_tcgen05_cta_pair_tmem_a_weight_gradient(N, K):
    """weight-gradient kernel with the A operand in tensor memory (tcgen05.st by four warps, tcgen05.mma with A from
    TMEM): out = 256, K = rows incl. counts that are not multiples of 32, split-K partials reduced by the hook."""
    C, ref, bound = run_gemm(0, 0, 5, 256, N, K)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs()
    tol = 2e-6 if K <= 50000 else 5e-6          # fp32 accumulation over 25 000-row chunks: ~2^-24 sqrt(K) on top
    assert (err <= tol * bound + 1e-6).all(), float((err / (bound + 1e-9)).max())
    if K > 50000:                              # same accumulation length on the shared-memory-A pair kernel: same error level
        C3, _, _ = run_gemm(0, 0, 3, 256, N, K)
        assert float((C3.double() - ref).abs().max()) >= 0.2 * float(err.max())
