/*
 * fbsnn_b200.h -- C-ABI of the B200 (sm_100a) forward-backward-SDE training step and the correlated-GBM
 * basket Monte-Carlo pricer.
 *
 * The reference (timothykski/Deep-neural-network-solutions-for-partial-differential-equations) has no FFI of
 * its own: its boundary is the Python class surface.  Each entry point below states the
 * reference method it replaces.  All pointers are DEVICE pointers unless the name says `host`; every call
 * enqueues work on `stream` and returns without synchronising (no allocation, no host sync => graph-capturable).
 * Return value: 0 on success, a negative FBSNN_E_* code otherwise; fbsnn_last_error() gives the message.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 */
#ifndef FBSNN_B200_H
#define FBSNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FBSNN_MAX_HIDDEN 8

enum { FBSNN_OK = 0, FBSNN_E_BADARG = -1, FBSNN_E_UNSUPPORTED = -2, FBSNN_E_CUDA = -3, FBSNN_E_WORKSPACE = -4 };

/* network topology: reference `mode` strings "FC" (DeepBSDE.py:166-172) and "Naisnet"/"NAIS-Net"
 * (Functions/naisnet.py:6-95) */
enum { FBSNN_NET_FC = 0, FBSNN_NET_NAIS = 1 };
/* activation: Functions/Sine.py:6-12, nn.ReLU, nn.Tanh (with_corr_high_dimension_pde.py:157-162) */
enum { FBSNN_ACT_SINE = 0, FBSNN_ACT_RELU = 1, FBSNN_ACT_TANH = 2 };
/* closed enumeration of the reference's mu_tf / sigma_tf / phi_tf / g_tf callables */
enum { FBSNN_MU_ZERO = 0, FBSNN_MU_LINEAR = 1,              /* 0 | mu_c * X                                   */
       FBSNN_MU_HESTON = 2 };                               /* [mu_c S, kappa (theta - v)], clamped to +-100  */
enum { FBSNN_SIGMA_CONST = 0, FBSNN_SIGMA_PROP = 1,         /* sigma_c * I | sigma_c * diag(X)                */
       FBSNN_SIGMA_HESTON = 2 };                            /* 2x2 Heston diffusion (heston_dnnpde.py:593-605) */
enum { FBSNN_PHI_BSB = 0, FBSNN_PHI_RY = 1, FBSNN_PHI_ZSQ = 2 }; /* c(Y - X.Z) | c Y | |Z|^2                   */
enum { FBSNN_G_SUMSQ = 0, FBSNN_G_CALL_SUM = 1, FBSNN_G_CALL_MEAN = 2, FBSNN_G_LOGQ = 3,
       FBSNN_G_CALL_FIRST = 4,                              /* max(X_0 - K, 0)            (heston_dnnpde.py:548-550) */
       FBSNN_G_CALL_FIRST_SMOOTH = 5 };                     /* (X_0-K)/(1+exp(-10 (X_0-K))) (heston_dnnpde.py:551-556) */
/* arithmetic variant of the dense layers */
enum { FBSNN_PREC_FP32 = 0,      /* SIMT fp32 FMA: parity-grade (tolerances in tests/test_parity_gpu.py)     */
       FBSNN_PREC_TF32 = 1,      /* tcgen05 kind::tf32, fp32 accumulate in TMEM: large-M throughput variant  */
       FBSNN_PREC_TF32X3 = 2 };  /* tcgen05, operands split hi+lo in smem, 3 MMAs per k-step: fp32-grade     */

typedef struct FbsnnSpec {
  int32_t D;                          /* state dimension                                                      */
  int32_t N;                          /* Euler-Maruyama steps; rows per path = N + 1                          */
  int32_t n_hidden;                   /* hidden layers L (FC: len(layers)-2; NAIS: stable blocks + 1)         */
  int32_t width[FBSNN_MAX_HIDDEN];    /* hidden widths, each a multiple of 4; NAIS: all equal                 */
  int32_t net_kind, act_kind;
  int32_t mu_kind, sigma_kind, phi_kind, g_kind;
  float mu_c, sigma_c, phi_c, strike;
  float nais_eps;                     /* 0.01 in the reference (Functions/naisnet.py:27)                      */
  int32_t precision;
  /* offsets (in floats, multiples of 4) into the flat parameter / gradient / Adam buffers; index l = 1..L is
   * hidden layer l, index L+1 the scalar output layer; -1 = absent.  Matrices are PyTorch (out, in) row-major. */
  int64_t off_W[FBSNN_MAX_HIDDEN + 2];    /* FC: Linear l weight; NAIS: layer{l}.weight                       */
  int64_t off_b[FBSNN_MAX_HIDDEN + 2];
  int64_t off_Win[FBSNN_MAX_HIDDEN + 2];  /* NAIS layer{l}_input.weight (l = 2..L)                            */
  int64_t off_bin[FBSNN_MAX_HIDDEN + 2];
  int64_t n_params;                       /* length of the flat buffers in floats (incl. alignment padding)   */
  /* ---- version 101: Heston 2-factor problem (heston_dnnpde.py:519-659); all zero for the other problems ---- */
  int32_t noise_dim;                      /* columns of W; 0 = D.  Heston: D = 2 states (S, v), ONE Brownian driver */
  int32_t clamp_u;                        /* 1: u = max(net, 0), Du masked accordingly (heston_dnnpde.py:568)  */
  int32_t zt_dims;                        /* terminal |Z - grad g|^2 over the first zt_dims components; 0 = all D */
  float h_kappa, h_theta, h_xi, h_rho, h_v0;   /* mean reversion, long-run variance, vol of vol, correlation, v(0) */
} FbsnnSpec;

/* Adam + clip_grad_norm_ hyper-parameters (torch.optim.Adam defaults, DeepBSDE.py:272;
 * clip: with_corr_high_dimension_pde.py:424).  max_grad_norm <= 0 disables clipping. */
typedef struct FbsnnAdam {
  double lr, beta1, beta2, eps, max_grad_norm;
  double skip_nonfinite;   /* != 0: an iteration whose gradient norm is NaN/inf leaves parameters, moments and the
                              Adam step counter untouched (heston_dnnpde.py:408-410 skips NaN-loss iterations) */
} FbsnnAdam;

const char* fbsnn_last_error(void);
int fbsnn_version(void);
/* Run-time switch of the kernel dispatch (tests, A/B measurements; no reference counterpart).  "chain": 0 = one launch
 * per dense layer, 1 = layer-chained sweep kernels once the row tiles fill the chip (default), 2 = always when the
 * network is eligible (FC, widths multiples of 32 in [64, 256], tensor-core precision).  "chain_ta": which chained kernel:
 * 0 = operand of the next layer's MMA in shared memory (chain_kernel), 1 = in tensor memory where that measured faster
 * (default), 2 = in tensor memory for every eligible sweep (widths multiples of 64).  "chain_pair": 0 | 1 = cta_group::2
 * form of chain_kernel.  Returns the previous value. */
int fbsnn_set_option(const char* name, int value);
/* Measurement hooks used by bench.py: number of kernels this library has launched since it was loaded; and
 * optional CUDA-event timing of every dense-layer launch (enable, run, synchronise, read).
 * out8 = {launches, ms, algorithmic FLOPs, tcgen05 launches, tcgen05 ms, tcgen05 FLOPs, algorithmic HBM bytes,
 * tcgen05 algorithmic HBM bytes} (bytes = both operands once + every row array the fused epilogue touches). */
long long fbsnn_launch_count(void);
void fbsnn_dense_timing(int enable);
int fbsnn_dense_timing_read(double* out8);
const char* fbsnn_dense_timing_entry(int i, double* out4);   /* one recorded launch: {ms, FLOPs, bytes, is tcgen05}; returns its sweep tag */

/* Test hook (tests/test_gemm_gpu.py): one dense GEMM with a plain store on the SIMT (use_tc = 0), tcgen05 TF32
 * (use_tc = 1) or tcgen05 3xTF32 (use_tc = 2) kernel.
 * a_kc: A[m*lda + k] (1) | A[k*lda + m] (0);  b_kc: B[n*ldb + k] (1) | B[k*ldb + n] (0);  C[m*ldc + n]. */
int fbsnn_debug_gemm(int a_kc, int b_kc, int use_tc, int M, int N, int K, const float* A, int lda, const float* B,
                     int ldb, float* C, int ldc, void* stream);

/* Test hook (tests/test_chain_gpu.py): float offset and row width of a named per-row array inside the workspace
 * ("xin","Y","zf","V","ybar", and per hidden layer "g","a","delta","szz","hd"), to compare what two dispatch variants
 * leave behind. */
int fbsnn_debug_ws_offset(const FbsnnSpec* spec, int64_t n_paths, int with_grad, const char* name, int layer,
                          int64_t* offset_out, int* width_out);

/* Bytes of device scratch needed for `n_paths` paths (rows = n_paths * (N+1)).  `with_grad` = 0 sizes for
 * forward/predict only. */
int fbsnn_workspace_bytes(const FbsnnSpec* spec, int64_t n_paths, int with_grad, size_t* bytes_out);

/* Replaces FBSNN.fetch_minibatch (DeepBSDE.py:247-262; correlated: with_corr_high_dimension_pde.py:316-353)
 * on the device: t (M,N+1,1) and cumulative W (M,N+1,D), Philox4x32-10 keyed (seed, iteration, global path id).
 * `chol` = lower Cholesky factor (D,D) row-major or NULL.  Not stream-compatible with NumPy's MT19937. */
int fbsnn_fetch_minibatch(const FbsnnSpec* spec, float T, int64_t n_paths, int64_t path_offset, uint64_t seed,
                          uint64_t iteration, const float* chol, void* workspace, size_t workspace_bytes,
                          float* t_out, float* W_out, void* stream);

/* Replaces FBSNN.net_u (DeepBSDE.py:189-194): u (rows,) and Du (rows, D) for arbitrary (t, X) rows. */
int fbsnn_net_u(const FbsnnSpec* spec, const float* params, const float* t, const float* X, int64_t rows,
                void* workspace, size_t workspace_bytes, float* u_out, float* du_out, void* stream);

/* Replaces FBSNN.loss_function forward / FBSNN.predict (DeepBSDE.py:202-245, 297-302).
 * t (M,N+1,1), W (M,N+1,D) cumulative as the reference passes them; Xi (xi_rows, D), xi_rows in {1, M}.
 * Outputs (any may be NULL): X (M,N+1,D), Y (M,N+1), Z (M,N+1,D), loss (1). */
int fbsnn_forward(const FbsnnSpec* spec, const float* params, const float* t, const float* W, const float* Xi,
                  int64_t xi_rows, int64_t n_paths, void* workspace, size_t workspace_bytes, float* X_out,
                  float* Y_out, float* Z_out, float* loss_out, void* stream);

/* loss_function + loss.backward() (DeepBSDE.py:278-279): as fbsnn_forward, plus the gradient of the summed
 * loss w.r.t. every parameter written (not accumulated) into `grads` (flat, same layout as params).
 * If W == NULL the Brownian increments are drawn in-kernel (Philox, as fbsnn_fetch_minibatch with the same
 * seed/iteration/path_offset, t_n = n T/N) and `t` is ignored. */
int fbsnn_loss_grad(const FbsnnSpec* spec, const float* params, float* grads, const float* t, const float* W,
                    const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                    uint64_t seed, uint64_t iteration, const float* chol, void* workspace,
                    size_t workspace_bytes, float* X_out, float* Y_out, float* Z_out, float* loss_out,
                    void* stream);

/* fbsnn_loss_grad for one rank of a data-parallel training step: the Philox iteration is the persistent device counter
 * of `opt_state` (the one fbsnn_train_step uses; advanced by fbsnn_adam_step / fbsnn_peer_allreduce_adam), so that the
 * sharded step draws exactly the single-GPU stream, consecutive train() calls never replay noise, and the call can
 * be captured in a CUDA graph (no host-side iteration argument).  DeepBSDE.py:274-279 per rank. */
int fbsnn_loss_grad_step(const FbsnnSpec* spec, const float* params, float* grads, const float* t, const float* W,
                         const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                         uint64_t seed, const float* chol, const void* opt_state, void* workspace,
                         size_t workspace_bytes, float* X_out, float* Y_out, float* loss_out, void* stream);

/* min_loss / min_loss_state bookkeeping of the reference's train() (with_corr_high_dimension_pde.py:431-433,
 * 1d_BSPDE_case.py:394-396) on the device: if *loss < state[0] then state[0] = *loss, the iteration index is recorded
 * and (X, Y) are copied to (X_best, Y_best); nothing is copied otherwise and the host is never consulted.
 * `state` = 32 bytes: float best loss (initialise to +inf), int flag, int best call index (-1), int call counter (0),
 * int64 Philox iteration of the best step (taken from `opt_state`, nullable) -- with in-kernel increments X of that step
 * can be re-materialised from it later instead of being copied.  n_x / n_y are element counts (multiples of 4; 0 skips). */
int fbsnn_track_min(const float* loss, float* state, const void* opt_state, const float* X, float* X_best, int64_t n_x,
                    const float* Y, float* Y_best, int64_t n_y, void* stream);

/* clip_grad_norm_ + Adam.step on the flat buffers (with_corr_high_dimension_pde.py:424-425).  `opt_state` is
 * FBSNN_OPT_STATE_BYTES of device memory: int64 step counter at byte 0 (zero-initialised by the caller when the
 * optimizer is created -- the reference builds a fresh Adam per train() call, DeepBSDE.py:272), then float
 * clip_coef @8, step_size @12, sqrt(bias_correction2) @16, grad_norm @20, int32 skip flag @32, int64 Philox iteration counter @24 (keep it
 * across optimisers: fbsnn_train_step adds it to `iteration`), int64 count of optimiser steps ever taken @40 (the epoch of
 * the peer all-reduce's barrier: never reset or rewound), and reduction scratch from byte 64.
 * All counters are advanced on the device, so the call can be replayed from a CUDA graph. */
#define FBSNN_OPT_STATE_BYTES 2048
int fbsnn_adam_step(const FbsnnAdam* host_hp, float* params, const float* grads, float* exp_avg,
                    float* exp_avg_sq, int64_t n_params, void* opt_state, void* stream);

/* Multi-GPU: gradient all-reduce fused with its cross-GPU barrier and the clip norm, then the Adam update, over
 * NVLink peer memory instead of NCCL (no reference counterpart -- upstream is single-device).
 * Each rank owns one symmetric buffer of fbsnn_peer_buffer_floats() floats, zero-initialised once and mapped into
 * every peer: [0, n_params) gradients (pass it as `grads` to fbsnn_loss_grad), [n_params] that rank's loss (pass
 * as `loss_out`), u32 flags from *flag_offset_out.  `peer_ptrs_dev` = device array of `world` pointers to the
 * ranks' buffers as mapped in this process, index = rank.
 * Per iteration:  fbsnn_peer_wait (peers finished reading the previous iteration's gradients)  ->
 * fbsnn_loss_grad  ->  fbsnn_peer_allreduce_adam.  One kernel signals/waits the peers (release/acquire at system
 * scope, bounded spin), sums the W buffers in rank order (=> bit-identical parameters on every rank) into
 * grad_sum (local, n_params + 4 floats; [n_params] = global loss) and accumulates the squared norm; clip + Adam
 * follow on the same stream.  The epoch is the step count @40 of opt_state (never reset or rewound), so all ranks
 * must have taken the same number of optimiser steps; the sequence is CUDA-graph capturable. */
int fbsnn_peer_buffer_floats(int64_t n_params, int64_t* flag_offset_out, int64_t* total_out);
int fbsnn_peer_wait(const float* local_buf, int64_t n_params, int world, const void* opt_state, void* stream);
int fbsnn_peer_allreduce_adam(const FbsnnAdam* host_hp, float* params, const void* peer_ptrs_dev, int world, int rank,
                              float* grad_sum, float* exp_avg, float* exp_avg_sq, int64_t n_params, void* opt_state,
                              void* stream);

/* One fused training iteration on one GPU = fbsnn_loss_grad + fbsnn_adam_step.  Multi-GPU callers run
 * fbsnn_loss_grad, all-reduce [grads | loss] over NCCL, then fbsnn_adam_step. */
int fbsnn_train_step(const FbsnnSpec* spec, const FbsnnAdam* host_hp, float* params, float* grads,
                     float* exp_avg, float* exp_avg_sq, void* opt_state, const float* t, const float* W,
                     const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                     uint64_t seed, uint64_t iteration, const float* chol, void* workspace,
                     size_t workspace_bytes, float* X_out, float* Y_out, float* loss_out, void* stream);

/* ---- Monte-Carlo basket pricer (numerics/multidimensional_mc_pricer.py) --------------------------------- */
typedef struct McSpec {
  int32_t D, N;          /* assets, time steps                                                                */
  float rate, sigma, T, strike;
} McSpec;

/* Replaces MonteCarloPricer.price = generate_paths + payoff + discounted mean (:49-93) without storing paths.
 * Simulates global paths [path_offset, path_offset + n_paths); writes sums_out[0] = sum of discounted payoffs,
 * sums_out[1] = sum of squares (double).  chol_T = TRANSPOSED lower Cholesky factor (chol_T[j*D+d] = L[d][j])
 * or NULL for independent assets.  scratch: mc_scratch_bytes(). */
size_t mc_scratch_bytes(void);
long long mc_launch_count(void);   /* kernels launched by the pricer since load (bench.py's gpu_launches) */
int mc_basket_price(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                    uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                    void* stream);

/* As mc_basket_price, plus the pathwise deltas: delta_sums_out[d] = sum over paths of
 * disc * 1{basket > K} * w_d * S_T,d / S0_d  (double, D entries; divide by the global path count).  This is the
 * quantity basket_pricer.py:68-81 (BasketOptionPricer.delta) estimates by bump-and-revalue with fresh noise per
 * bump; the pathwise form is unbiased and costs one pass. */
int mc_basket_price_delta(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                          uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                          double* delta_sums_out, void* stream);

/* Replaces the Cole-Hopf Monte-Carlo "exact" solution of the HJB driver (hjb_implement.py:1088-1094):
 * u_out[n] = -ln mean_{i < n_mc} 1 / (0.5 + 0.5 |X[n,:] + sqrt(2 |T - t[n]|) W_i|^2), W_i ~ N(0, I_D) from Philox
 * keyed by (seed, i, n).  t (n_times), X (n_times, D) device fp32; u_out (n_times) device double. */
int mc_hjb_exact(int32_t D, int32_t n_times, const float* t, const float* X, float T, uint64_t n_mc, uint64_t seed,
                 void* scratch, double* u_out, void* stream);

/* Measurement hook (bench.py): the pricer's generator alone -- Philox4x32-10 + MUFU Box-Muller,
 * every normal consumed by one add -- drawing about n_normals normals; *normals_out = exact count.  Its rate on the
 * same GPU is the roofline denominator of mc_basket_price.  No reference counterpart. */
int mc_normal_rate_probe(uint64_t n_normals, uint64_t seed, void* scratch, uint64_t* normals_out, void* stream);

/* Closed-form comparators of the reference's drivers for a whole prediction tensor in one launch (fp64).
 * mode 0 replaces BasketOptionPriceCalculator.calculate_option_prices (nd_BSPDE_case.py:621-658): S (rows, cols = assets),
 *        t (rows) -> price_out / delta_out (rows): mean over assets of the per-asset Black-Scholes call and delta.
 * mode 1 replaces BasicOptionPriceCalculator.calculate_call_option_prices (with_corr_high_dimension_pde.py:663-700): S
 *        (rows, cols) basket averages, t = time grid (ntimes) -> (rows, cols): Black-Scholes on the average with volatility
 *        sigma / sqrt(dims), payoff and one-sided delta at maturity. */
int mc_bs_comparator(int32_t mode, const double* S, const double* t, int64_t rows, int32_t cols, int32_t ntimes, double K,
                     double r, double sigma, double T, int32_t dims, double* price_out, double* delta_out, void* stream);

/* Replaces BlackScholesModel.generate_paths (:49-67): paths_out (n_paths, N+1, D) float32. */
int mc_generate_paths(const McSpec* spec, const float* S0, const float* chol_T, uint64_t n_paths,
                      uint64_t seed, uint64_t path_offset, float* paths_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FBSNN_B200_H */
