// C-ABI of the FBSNN step (include/fbsnn_b200.h): workspace plan, sweep orchestration, launches.
//
// Row r = m*(N+1) + n is one (path, step) point.  X does not depend on the network parameters for any of the
// reference's problems, so all M*(N+1) rows are evaluated as one batch with four sweeps
//   F (forward)  A (input adjoint -> Z = Du)  T (tangent in direction dL/dZ)  B (backward)
// followed by the weight-gradient contractions G.  tests/passes_model.py states the same algebra on the host.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "gemm_tc16.cuh"
#include "gemm_tc2g.cuh"
#include "gemm_chain.cuh"
#include "kernels.cuh"

namespace fbsnn {

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU_CHECK(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) return fail(FBSNN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));    \
  } while (0)
static long long g_launches = 0;  // kernels launched by this library since load (bench.py's gpu_launches)
#define LAUNCH_CHECK(name)                                                                        \
  do {                                                                                            \
    ++g_launches;                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                                          \
    if (e_ != cudaSuccess) return fail(FBSNN_E_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

constexpr int kMaxL = FBSNN_MAX_HIDDEN;

struct Plan {
  int D, N, L, ldx, ldi, d_in, Dn;   // Dn = Brownian (noise) dimension, = D except for Heston
  bool tf32;
  int kin, ldw;   // K extent / leading dimension of the input-width weight operands (padded to ldx in TF32 mode)
  size_t W1p, Winp[kMaxL + 2];
  size_t Whi[kMaxL + 2], Wlo[kMaxL + 2], Winhi[kMaxL + 2], Winlo[kMaxL + 2];   // 3xTF32: exact-TF32 weight halves
  bool x3;
  int H[kMaxL + 2];
  bool nais;
  long long rows;
  int loss_blocks, col_blocks, col_rows_per_block, wg_split, wg_chunk, gsq_blocks;
  size_t wg_stride;   // floats between the per-contraction split-K partial buffers
  int wg_slots;       // how many of them the workspace holds
  // offsets in floats
  size_t xin, sdw, Y, zf, V, ev, ybar, inc, zraw, part_loss, part_wg, part_col, part_gsq, umask;
  size_t g[kMaxL + 2], a[kMaxL + 2], delta[kMaxL + 2], szz[kMaxL + 2], hd[kMaxL + 2], h[kMaxL + 2],
      ht[kMaxL + 2], hb[kMaxL + 2];
  size_t Bm[kMaxL + 2], Rm[kMaxL + 2], Bbar[kMaxL + 2], Sm[kMaxL + 2], nstate[kMaxL + 2];
  size_t chain_col;   // per-CTA column-sum partials of the layer-chained T / B sweeps
  size_t total;
};

static int round_up(long long v, int m) { return (int)((v + m - 1) / m * m); }

static int validate(const FbsnnSpec* s) {
  if (!s) return fail(FBSNN_E_BADARG, "spec is null");
  if (s->D < 1 || s->N < 1) return fail(FBSNN_E_BADARG, "need D >= 1 and N >= 1 (got D=%d N=%d)", s->D, s->N);
  if (s->n_hidden < 1 || s->n_hidden > kMaxL) return fail(FBSNN_E_UNSUPPORTED, "n_hidden=%d outside 1..%d", s->n_hidden, kMaxL);
  for (int l = 0; l < s->n_hidden; ++l)
    if (s->width[l] < 4 || s->width[l] % 4) return fail(FBSNN_E_UNSUPPORTED, "hidden width %d is not a positive multiple of 4", s->width[l]);
  if (s->net_kind == FBSNN_NET_NAIS) {
    if (s->n_hidden < 2 || s->n_hidden > 4) return fail(FBSNN_E_UNSUPPORTED, "NAIS-Net needs 1..3 stable blocks");
    for (int l = 1; l < s->n_hidden; ++l)
      if (s->width[l] != s->width[0]) return fail(FBSNN_E_UNSUPPORTED, "NAIS-Net needs equal hidden widths");
  } else if (s->net_kind != FBSNN_NET_FC) {
    return fail(FBSNN_E_UNSUPPORTED, "unknown net_kind %d", s->net_kind);
  }
  if (s->act_kind < 0 || s->act_kind > 2) return fail(FBSNN_E_UNSUPPORTED, "unknown act_kind %d", s->act_kind);
  if (s->mu_kind < 0 || s->mu_kind > 2 || s->sigma_kind < 0 || s->sigma_kind > 2 || s->phi_kind < 0 ||
      s->phi_kind > 2 || s->g_kind < 0 || s->g_kind > 5)
    return fail(FBSNN_E_UNSUPPORTED, "problem callables outside the closed enumeration");
  const bool heston = s->sigma_kind == FBSNN_SIGMA_HESTON || s->mu_kind == FBSNN_MU_HESTON;
  if (heston && (s->sigma_kind != FBSNN_SIGMA_HESTON || s->mu_kind != FBSNN_MU_HESTON || s->D != 2 || s->noise_dim != 1))
    return fail(FBSNN_E_UNSUPPORTED, "the Heston problem needs mu and sigma of kind HESTON, D = 2 states and noise_dim = 1");
  if (!heston && s->noise_dim != 0 && s->noise_dim != s->D)
    return fail(FBSNN_E_UNSUPPORTED, "noise_dim %d != D %d is only defined for the Heston problem", s->noise_dim, s->D);
  if (s->zt_dims < 0 || s->zt_dims > s->D) return fail(FBSNN_E_BADARG, "zt_dims %d outside 0..D", s->zt_dims);

  if (s->precision != FBSNN_PREC_FP32 && s->precision != FBSNN_PREC_TF32 && s->precision != FBSNN_PREC_TF32X3)
    return fail(FBSNN_E_UNSUPPORTED, "unknown precision %d", s->precision);
  return 0;
}

static void make_plan(const FbsnnSpec* s, long long rows, bool with_grad, Plan& p) {
  memset(&p, 0, sizeof(p));
  p.D = s->D, p.N = s->N, p.L = s->n_hidden;
  p.d_in = s->D + 1;
  p.tf32 = s->precision != FBSNN_PREC_FP32;   // both tensor-core variants use the padded (UMMA-tileable) layout
  // TF32 variant: the input width is zero-padded to a multiple of 32 so that every dense layer is a TMA/UMMA tile
  p.ldx = round_up(p.d_in, p.tf32 ? 32 : 4);
  p.kin = p.tf32 ? p.ldx : p.d_in;
  p.ldw = p.kin;
  p.Dn = s->noise_dim > 0 ? s->noise_dim : s->D;
  p.ldi = round_up(p.Dn, 4);
  p.nais = s->net_kind == FBSNN_NET_NAIS;
  p.rows = rows;
  p.H[0] = p.d_in;
  for (int l = 1; l <= p.L; ++l) p.H[l] = s->width[l - 1];
  // loss kernels: 8 warps (= 8 rows at a time) per block, grid-stride; few enough partials for a one-block final sum
  p.loss_blocks = (int)std::min<long long>((rows + 7) / 8, (long long)num_sms() * 16);
  p.col_blocks = (int)std::min<long long>(std::max<long long>((rows + 63) / 64, 1), (long long)num_sms() * 2);
  p.col_rows_per_block = (int)((rows + p.col_blocks - 1) / p.col_blocks);
  p.col_blocks = (int)((rows + p.col_rows_per_block - 1) / p.col_rows_per_block);
  // split-K of the weight-gradient contraction: the SIMT kernel wants many CTAs; the persistent tcgen05 kernel
  // wants (M tiles) x splits ~ one work item per SM, and fewer partial tiles to reduce afterwards
  int split = (int)std::min<long long>(std::max<long long>((rows + 511) / 512, 1), 256);
  if (s->precision != FBSNN_PREC_FP32) split = (int)std::min<long long>(std::max<long long>((rows + 63) / 64, 1), 74);
  // small 3xTF32 batches run the contractions of all layers in one batched launch (sweeps_backward): one work item per CTA
  // pair, i.e. SMs / 2 / contractions splits each -- fewer accumulator drains and a third of the partials to reduce
  // (only when every contraction is one the batched kernel takes: 256 outputs, input widths multiples of 64 -- the others
  // run as launches of their own and want the parallelism)
  bool batchable = s->precision == FBSNN_PREC_TF32X3 && rows <= (long long)num_sms() * 256 && p.ldx % 64 == 0;
  for (int l = 1; l <= p.L; ++l) batchable = batchable && p.H[l] == 256;
  if (batchable) {
    const int jobs = p.nais ? 2 * p.L - 1 : p.L;
    split = std::min(split, std::max(1, num_sms() / 2 / std::max(jobs, 1)));
  }
  p.wg_chunk = round_up((rows + split - 1) / split, 32);   // multiple of the tcgen05 kernel's BLOCK_K
  p.wg_split = (int)((rows + p.wg_chunk - 1) / p.wg_chunk);
  p.gsq_blocks = 256;
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += (n + 63) / 64 * 64;  // 256-byte granules
    return o;
  };
  const size_t R = (size_t)rows;
  p.xin = take(R * p.ldx);
  p.sdw = take(R * p.D);
  p.Y = take(R);
  p.zf = take(R * p.ldx);
  p.part_loss = take(2 * (size_t)p.loss_blocks);
  p.ev = take(R);
  if (s->clamp_u) p.umask = take(R);
  for (int l = 1; l <= p.L; ++l) {
    p.g[l] = take(R * p.H[l]);
    p.a[l] = take(R * p.H[l]);
    p.delta[l] = take(R * p.H[l]);
    p.h[l] = (p.nais && l >= 2) ? take(R * p.H[l]) : p.g[l];
    if (p.nais && l >= 2 && l <= p.L - 1) p.ht[l] = take(R * p.H[l]);
  }
  if (p.nais)
    for (int l = 2; l <= p.L; ++l) {
      const size_t hh = (size_t)p.H[l] * p.H[l];
      p.Bm[l] = take(hh), p.Rm[l] = take(hh), p.nstate[l] = take(4);
    }
  p.x3 = s->precision == FBSNN_PREC_TF32X3;
  if (p.tf32) {
    p.W1p = take((size_t)p.H[1] * p.ldx);
    if (p.nais)
      for (int l = 2; l <= p.L; ++l) p.Winp[l] = take((size_t)p.H[l] * p.ldx);
  }
  if (p.x3) {
    for (int l = 1; l <= p.L; ++l) {
      const size_t n = (size_t)p.H[l] * (l == 1 ? p.ldx : p.H[l - 1]);
      p.Whi[l] = take(n), p.Wlo[l] = take(n);
      if (p.nais && l >= 2) p.Winhi[l] = take((size_t)p.H[l] * p.ldx), p.Winlo[l] = take((size_t)p.H[l] * p.ldx);
    }
  }
  if (with_grad) {
    p.V = take(R * p.ldx);
    p.ybar = take(R);
    p.inc = take(R * p.ldi);
    p.zraw = take(R * p.ldi);
    size_t wg = 0;
    for (int l = 1; l <= p.L; ++l) {
      p.szz[l] = take(R * p.H[l]);
      p.hd[l] = take(R * p.H[l]);
      if (p.nais && l >= 2 && l <= p.L - 1) p.hb[l] = take(R * p.H[l]);
      wg = std::max(wg, (size_t)p.H[l] * (size_t)round_up(p.H[l - 1], 4));
      wg = std::max(wg, (size_t)p.H[l] * (size_t)p.ldx);
    }
    p.wg_stride = (wg * p.wg_split + 63) / 64 * 64;
    // one split-K partial buffer per weight-gradient contraction (FC: L, NAIS-Net: 2 L - 1), reduced together in one launch
    p.wg_slots = p.nais ? (2 * p.L - 1 <= kMaxRedJobs ? 2 * p.L - 1 : 1) : p.L;
    p.part_wg = take(p.wg_stride * p.wg_slots);
    p.part_col = take((size_t)kMaxColJobs * std::max(p.col_blocks, 256) * 1024);   // 256 >= CTAs of the tcgen05 grid
    p.part_gsq = take(p.gsq_blocks);
    if (p.tf32 && !p.nais) p.chain_col = take((size_t)num_sms() * chain::kMaxLinks * 2 * 1024);
    if (p.nais)
      for (int l = 2; l <= p.L; ++l) {
        const size_t hh = (size_t)p.H[l] * p.H[l];
        p.Bbar[l] = take(hh), p.Sm[l] = take(hh);
      }
  }
  p.total = off;
}

static ProblemK problem_k(const FbsnnSpec* s, const Plan& p) {
  ProblemK k;
  k.D = s->D, k.N = s->N, k.ldx = p.ldx;
  k.mu_kind = s->mu_kind, k.sigma_kind = s->sigma_kind, k.phi_kind = s->phi_kind, k.g_kind = s->g_kind;
  k.mu_c = s->mu_c, k.sigma_c = s->sigma_c, k.phi_c = s->phi_c, k.strike = s->strike;
  k.zt_dims = s->zt_dims > 0 ? s->zt_dims : s->D;
  k.h_kappa = s->h_kappa, k.h_theta = s->h_theta, k.h_xi = s->h_xi, k.h_rho = s->h_rho;
  return k;
}

// Optional per-launch timing of the dense layers (bench.py's roofline leg): when enabled, every dense() call is
// bracketed by CUDA events on its stream; fbsnn_dense_timing_read() sums them after a synchronise.
constexpr int kMaxTimed = 4096;
static bool g_timing = false;
static int g_ntimed = 0;
static cudaEvent_t g_ev0[kMaxTimed], g_ev1[kMaxTimed];
static double g_flops[kMaxTimed], g_bytes[kMaxTimed];
static bool g_timed_tc[kMaxTimed];
static const char* g_timed_what[kMaxTimed];

// 3xTF32: weight matrices that have exact-TF32 hi/lo twins in the workspace (filled by split_weights())
struct SplitW {
  const float* src;
  const float* hi;
  const float* lo;
};
static thread_local SplitW g_splitw[2 * (kMaxL + 2)];
static thread_local int g_nsplitw = 0;
static const SplitW* find_split(const float* src) {
  for (int i = 0; i < g_nsplitw; ++i)
    if (g_splitw[i].src == src) return &g_splitw[i];
  return nullptr;
}

template <bool A_KC, bool B_KC>
static bool uses_tc(const FbsnnSpec* s, const GemmArgs& g, int nsplit) {
  return s->precision != FBSNN_PREC_FP32 && tc_eligible<A_KC, B_KC>(g, nsplit);
}

// 3xTF32 weight-gradient contractions go to the CTA-pair (cta_group::2) kernel: measured 4.2 vs 5.2 ms per launch at
// M = 65 536 (they are shared-memory-bandwidth bound on one CTA).  The sweeps stay on the single-CTA kernel: their
// k-blocks carry 2-3x fewer MMA cycles, so the pair's longer stage hand-off (remote barrier arrivals) is exposed and
// they measured 4.4-5.0 vs 3.0-3.9 ms (profiles/r01_launch_table_tf32x3_pair_all.txt); with W_hi / W_lo twins in one
// stage (SPLIT = 3) the pair ties the single-CTA kernel (3.3-3.9 ms, ..._pair_split3_all.txt) -- the sweeps are bound
// by HBM and the epilogue's load latency, not by shared memory.  FBSNN_PAIR=0 disables the pair kernel, FBSNN_PAIR=2
// sends every eligible 3xTF32 launch to it (A/B measurements).
// The F sweep runs on the 16-epilogue-warp form of the tcgen05 kernel (gemm_tc16.cuh): its epilogue is instruction-
// bound on the sine/cosine (3.0 vs 3.6 ms per layer, 2.2 vs 3.5 ms for the first layer, measured on one box).  The
// A/T/B sweeps measured SLOWER on it (4.4 / 4.25 / 3.3 vs 3.6 / 3.8 / 2.8 ms: 64-byte row segments and 80 registers
// hurt their load-bound epilogues) and stay on the 8-warp form.  FBSNN_EPI16=0 disables, =2 sends every sweep to it.
template <class Epi, class = void>
struct wants16 : std::false_type {};
template <class Epi>
struct wants16<Epi, std::void_t<decltype(Epi::kManyEpilogueWarps)>> : std::true_type {};
static int epi16_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FBSNN_EPI16");
    mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return mode;
}
template <class Epi>
static bool epi16_enabled(const Epi& epi, bool x3) {
  const int mode = epi16_mode();
  if (mode == 2) return true;
  if constexpr (wants16<Epi>::value) {
    // the last hidden layer's F epilogue also writes delta_L and s_L (four row arrays): store-bound; same box, 16- vs
    // 8-warp form: 4.44 vs 4.58 ms (3xTF32, keep 16), 4.4 vs 3.5 ms (TF32, use 8)
    return mode == 1 && (x3 || epi.wout == nullptr);
  }
  return false;
}
// weight gradients on the pair with the A operand in tensor memory (gemm_tc2g.cuh); FBSNN_GTMEM=0: A in shared memory
static bool gtmem_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FBSNN_GTMEM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
static int narrow3_mode() {   // FBSNN_NARROW3=0: Du through the two-segment (SPLIT = 1) form; =2: every 8-warp sweep on SPLIT = 3
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FBSNN_NARROW3");
    mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return mode;
}
static int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FBSNN_PAIR");
    mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return mode;
}
template <bool A_KC, bool B_KC>
static bool uses_pair(const FbsnnSpec* s, const GemmArgs& g, int nsplit) {
  const int mode = pair_mode();
  return s->precision == FBSNN_PREC_TF32X3 && (mode == 2 || (mode == 1 && !A_KC)) && tc2_eligible<A_KC, B_KC>(g, nsplit);
}
// CTAs of the tcgen05 launch dense() will make for (g, nsplit): the fused bias-gradient partials are per CTA
template <bool A_KC, bool B_KC>
static int tc_launch_grid(const FbsnnSpec* s, const GemmArgs& g, int nsplit) {
  if (uses_pair<A_KC, B_KC>(s, g, nsplit)) return tc2_grid(g, nsplit, num_sms());
  if (s->precision == FBSNN_PREC_TF32X3) return tc_colpart_rows<A_KC, 1>(g, nsplit, num_sms());
  return tc_colpart_rows<A_KC, 0>(g, nsplit, num_sms());
}

// dense layer dispatch: SIMT fp32, or tcgen05 TF32 when the variant is selected and the shape qualifies
template <bool A_KC, bool B_KC, class Epi>
static int dense(const FbsnnSpec* s, const GemmArgs& g, const Epi& epi, int nsplit, cudaStream_t st, const char* what,
                 bool allow_tc = true) {
  const bool tc = allow_tc && s->precision != FBSNN_PREC_FP32 && tc_eligible<A_KC, B_KC>(g, nsplit);
  int slot = -1;
  if (g_timing && g_ntimed < kMaxTimed) {
    slot = g_ntimed++;
    if (!g_ev0[slot]) cudaEventCreate(&g_ev0[slot]), cudaEventCreate(&g_ev1[slot]);
    double k = 0;
    for (int i = 0; i < g.nseg; ++i) k += g.seg[i].K;
    g_flops[slot] = 2.0 * (double)g.M * (double)g.N * k;   // algorithmic FLOPs of this launch (padding included in N)
    // algorithmic HBM bytes: both operands once + every row array the fused epilogue reads or writes
    g_bytes[slot] = 4.0 * (k * ((double)g.M + (double)g.N) + (double)g.M * (double)g.N * epi.io_arrays() * nsplit);
    g_timed_tc[slot] = tc;
    g_timed_what[slot] = what;
    cudaEventRecord(g_ev0[slot], st);
  }
  ++g_launches;
  cudaError_t e;
  if (!tc) {
    e = launch_gemm<A_KC, B_KC>(g, epi, nsplit, num_sms(), st);
  } else if (s->precision == FBSNN_PREC_TF32X3) {
    // sweeps whose weight operands all have pre-split twins: (A, W) -> (A, W_hi) [a_lo*b + a*b] + (A, W_lo) [a*b]
    GemmArgs g2 = g;
    bool presplit = A_KC && 2 * g.nseg <= kMaxSeg;
    if (presplit) {
      g2.nseg = 0;
      for (int i = 0; i < g.nseg && presplit; ++i) {
        const SplitW* w = find_split(g.seg[i].B);
        if (!w) { presplit = false; break; }
        g2.seg[g2.nseg] = g.seg[i], g2.seg[g2.nseg].B = w->hi, g2.mode[g2.nseg++] = 1;
        g2.seg[g2.nseg] = g.seg[i], g2.seg[g2.nseg].B = w->lo, g2.mode[g2.nseg++] = 2;
      }
    }
    const bool pair = uses_pair<A_KC, B_KC>(s, g, nsplit);
    bool done16 = false;
    if constexpr (A_KC && !std::is_same<Epi, EpiPartial>::value) {
      if (presplit && !pair && epi16_enabled(epi, true) && tc16_eligible<B_KC>(g2, nsplit)) {
        e = launch_gemm_tc16<B_KC, 1>(g2, epi, num_sms(), st);
        done16 = true;
      }
    }
    bool done3 = false;
    if constexpr (A_KC && !std::is_same<Epi, EpiPartial>::value) {
      // narrow sweeps (Du: 128 columns): hi / lo weight twins in one 64 KB stage, the A tile fetched once
      if (!done16 && presplit && !pair && g.nseg <= 4 && (narrow3_mode() == 2 || (narrow3_mode() == 1 && g.N <= 128))) {
        GemmArgs g3 = g;
        for (int i = 0; i < g.nseg; ++i) {
          const SplitW* w = find_split(g.seg[i].B);
          g3.seg[i].B = w->hi;
          g3.seg[i + 4] = g.seg[i], g3.seg[i + 4].B = w->lo;
        }
        e = launch_gemm_tc<A_KC, B_KC, 3>(g3, epi, nsplit, num_sms(), st);
        done3 = true;
      }
    }
    if (done16 || done3) {
    } else if (presplit && pair && g.nseg <= 4) {   // pair sweeps: W_hi / W_lo twins in one stage (SPLIT = 3)
      GemmArgs g3 = g;
      for (int i = 0; i < g.nseg; ++i) {
        const SplitW* w = find_split(g.seg[i].B);
        g3.seg[i].B = w->hi;
        g3.seg[i + 4] = g.seg[i], g3.seg[i + 4].B = w->lo;
      }
      e = launch_gemm_tc2<A_KC, B_KC, 3>(g3, epi, nsplit, num_sms(), st);
    } else if (presplit) e = pair ? launch_gemm_tc2<A_KC, B_KC, 1>(g2, epi, nsplit, num_sms(), st)
                           : launch_gemm_tc<A_KC, B_KC, 1>(g2, epi, nsplit, num_sms(), st);
    else {
      bool doneg = false;
      if constexpr (!A_KC && !B_KC && std::is_same<Epi, EpiPartial>::value) {
        if (pair && gtmem_enabled() && tc2g_eligible(g, nsplit)) {
          e = launch_gemm_tc2g(g, epi, nsplit, num_sms(), st);
          doneg = true;
        }
      }
      if (!doneg) e = pair ? launch_gemm_tc2<A_KC, B_KC, 2>(g, epi, nsplit, num_sms(), st)
                           : launch_gemm_tc<A_KC, B_KC, 2>(g, epi, nsplit, num_sms(), st);
    }
  } else {
    bool done16 = false;
    if constexpr (A_KC && !std::is_same<Epi, EpiPartial>::value) {
      if (epi16_enabled(epi, false) && tc16_eligible<B_KC>(g, nsplit)) {
        e = launch_gemm_tc16<B_KC, 0>(g, epi, num_sms(), st);
        done16 = true;
      }
    }
    if (!done16) e = launch_gemm_tc<A_KC, B_KC, 0>(g, epi, nsplit, num_sms(), st);
  }
  if (slot >= 0) cudaEventRecord(g_ev1[slot], st);
  if (e != cudaSuccess) return fail(FBSNN_E_CUDA, "%s gemm %s: %s", tc ? "tcgen05" : "simt", what, cudaGetErrorString(e));
  return 0;
}

struct Net {
  const float* W[kMaxL + 2];    // main matrix of layer l as (H_l x H_{l-1}) row-major: FC W_l, NAIS l>=2: Bm_l
  const float* Wraw[kMaxL + 2];
  const float* Win[kMaxL + 2];
  const float* b[kMaxL + 2];
  const float* bin[kMaxL + 2];
  const float* wout;
  const float* bout;
};

static Net bind_net(const FbsnnSpec* s, const Plan& p, const float* params, float* ws) {
  Net n;
  memset(&n, 0, sizeof(n));
  for (int l = 1; l <= p.L; ++l) {
    n.Wraw[l] = params + s->off_W[l];
    n.b[l] = params + s->off_b[l];
    if (p.nais && l >= 2) {
      n.W[l] = ws + p.Bm[l];
      n.Win[l] = params + s->off_Win[l];
      n.bin[l] = params + s->off_bin[l];
    } else {
      n.W[l] = n.Wraw[l];
    }
  }
  n.wout = params + s->off_W[p.L + 1];
  n.bout = params + s->off_b[p.L + 1];
  if (p.tf32) {   // zero-padded copies (H x ldx), refreshed by prepare_weights() every call
    n.W[1] = ws + p.W1p;
    if (p.nais)
      for (int l = 2; l <= p.L; ++l) n.Win[l] = ws + p.Winp[l];
  }
  return n;
}

// TF32 variant: copy the input-width matrices into their zero-padded (H x ldx) homes
static int prepare_weights(const FbsnnSpec* s, const Plan& p, const float* params, float* ws, cudaStream_t st) {
  if (!p.tf32 || !p.nais) return 0;   // FC: done together with the hi/lo split in split_weights()
  for (int l = 1; l <= p.L; ++l) {
    const float* src = params + (l == 1 ? s->off_W[1] : s->off_Win[l]);
    float* dst = ws + (l == 1 ? p.W1p : p.Winp[l]);
    const int n = p.H[l] * p.ldx;
    pad_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, p.H[l], p.d_in, dst, p.ldx);
    LAUNCH_CHECK("pad_copy");
  }
  return 0;
}

// 3xTF32: write exact-TF32 hi / lo twins of every weight matrix the sweeps use as their B operand and register them.
// FC networks: ONE launch does the zero-padded copy of W_1 and all hi / lo twins (prep_weights_kernel).
static int split_weights(const FbsnnSpec* s, const Plan& p, const Net& n, const float* params, float* ws, cudaStream_t st) {
  g_nsplitw = 0;
  if (!p.nais) {
    if (!p.tf32) return 0;
    PrepJobs js{};
    for (int l = 1; l <= p.L; ++l) {
      if (l >= 2 && !p.x3) break;
      PrepJob& j = js.job[js.njobs++];
      const int cols_src = l == 1 ? p.d_in : p.H[l - 1], ld = l == 1 ? p.ldx : p.H[l - 1];
      j.src = params + s->off_W[l];
      j.pad = l == 1 ? ws + p.W1p : nullptr;
      j.hi = p.x3 ? ws + p.Whi[l] : nullptr, j.lo = p.x3 ? ws + p.Wlo[l] : nullptr;
      j.rows = p.H[l], j.cols = cols_src, j.ld = ld;
      if (p.x3) g_splitw[g_nsplitw++] = SplitW{n.W[l], ws + p.Whi[l], ws + p.Wlo[l]};
    }
    prep_weights_kernel<<<dim3(64, js.njobs), 256, 0, st>>>(js);
    LAUNCH_CHECK("prep_weights");
    return 0;
  }
  if (!p.x3) return 0;
  auto one = [&](const float* src, size_t hi, size_t lo, int count) {
    split_hi_lo_kernel<<<(count + 255) / 256, 256, 0, st>>>(src, count, ws + hi, ws + lo);
    g_splitw[g_nsplitw++] = SplitW{src, ws + hi, ws + lo};
  };
  for (int l = 1; l <= p.L; ++l) {
    one(n.W[l], p.Whi[l], p.Wlo[l], p.H[l] * (l == 1 ? p.ldx : p.H[l - 1]));
    LAUNCH_CHECK("split_hi_lo");
    if (l >= 2) {
      one(n.Win[l], p.Winhi[l], p.Winlo[l], p.H[l] * p.ldx);
      LAUNCH_CHECK("split_hi_lo");
    }
  }
  return 0;
}

// NAIS projection, once per call: Bm_l = -(s W_l^T W_l + eps I)
static int nais_prepare(const FbsnnSpec* s, const Plan& p, const Net& n, float* ws, cudaStream_t st) {
  for (int l = 2; l <= p.L; ++l) {
    const int H = p.H[l];
    GemmArgs g{};
    g.nseg = 1, g.M = H, g.N = H, g.Nb = H, g.kchunk = 0;
    g.seg[0] = GemmSeg{n.Wraw[l], n.Wraw[l], H, H, H};
    int rc = dense<false, false>(s, g, EpiStore{ws + p.Rm[l], H}, 1, st, "nais RtR", false);
    if (rc) return rc;
    nais_project_kernel<<<(H * H + 1023) / 1024, 1024, 0, st>>>(ws + p.Rm[l], H, s->nais_eps, ws + p.Bm[l], ws + p.nstate[l]);
    LAUNCH_CHECK("nais_project");
  }
  return 0;
}


// ----------------------------------------------------------------------------------------------------------------
// Layer-chained sweeps (gemm_chain.cuh): FC networks on the tensor-core variants whose widths tile into 32-column
// chunks.  Option "chain": 0 = per-layer launches only, 1 = chained sweeps once the row tiles fill the chip (the small-M
// steps keep the narrow-tile per-layer launches that spread 40 row tiles over all SMs), 2 = always when eligible.
// ----------------------------------------------------------------------------------------------------------------
static int g_opt_chain = -1;
static int chain_mode() {
  if (g_opt_chain < 0) {
    const char* e = getenv("FBSNN_CHAIN");
    g_opt_chain = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return g_opt_chain;
}
static bool chain_use_ta(const Plan& p);
static bool chain_eligible(const FbsnnSpec* s, const Plan& p) {
  if (chain_mode() == 0 || p.nais || !p.tf32) return false;
  if (p.ldx % 32 || p.ldx > 256 || p.L + 1 > chain::kMaxLinks) return false;
  for (int l = 1; l <= p.L; ++l)
    if (p.H[l] % 32 || p.H[l] < 64 || p.H[l] > 256) return false;
  // below one row tile per SM the shared-memory chain kernel loses to the narrow-tile per-layer launches (more CTAs at work);
  // chaint_kernel still wins there: 4 launches instead of 16 (M = 100, 40 tiles: 0.29 vs 0.40 ms of dense launches per step)
  if (chain_mode() == 1 && p.rows < (long long)num_sms() * 128 && !chain_use_ta(p)) return false;
  return true;
}
// CTA-pair form of the chained sweeps (chain2_kernel): 3xTF32, every MMA width a multiple of 64 (each CTA stages half of
// the weight columns in 32-column boxes).  Option "chain_pair": 0 = single-CTA chain kernel only, 1 = pair when eligible.
static int g_opt_chain_pair = -1;
static int chain_pair_mode() {
  if (g_opt_chain_pair < 0) {
    const char* e = getenv("FBSNN_CHAIN_PAIR");
    g_opt_chain_pair = (e && e[0] >= '0' && e[0] <= '1') ? e[0] - '0' : 0;
  }
  return g_opt_chain_pair;
}
// Operand of the next MMA in tensor memory (chaint_kernel).  Option "chain_ta" / FBSNN_CHAIN_TA: 0 = the shared-memory forms
// above, 1 (default) = where measured faster (3xTF32: every sweep, 43.4 vs 50.1 ms for the four sweeps at M = 65 536;
// single-pass TF32: the F and B sweeps, 7.0 / 6.4 vs 7.8-8.3 / 6.7-7.1 ms -- its A / T sweeps are bound by the row-array
// traffic, where the shared-memory kernel's deeper operand ring does better, 10.0 / 10.5 vs 10.8 / 11.3 ms), 2 = every
// eligible sweep.
static int g_opt_chain_ta = -1;
static int chain_ta_mode() {
  if (g_opt_chain_ta < 0) {
    const char* e = getenv("FBSNN_CHAIN_TA");
    g_opt_chain_ta = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  return g_opt_chain_ta;
}
// chaint_kernel deals the chunks of every link alternately to its two half-teams and ties ring stages to them: every width
// has to be an even number of 32-column chunks (else the shared-memory chain kernel runs)
static bool chain_use_ta(const Plan& p) {
  if (!chain_ta_mode() || p.ldx % 64) return false;
  for (int l = 1; l <= p.L; ++l)
    if (p.H[l] % 64) return false;
  return true;
}
static bool chain_use_pair(const Plan& p) {
  if (chain_use_ta(p)) return false;
  if (!p.x3 || chain_pair_mode() == 0 || p.ldx % 64) return false;
  for (int l = 1; l <= p.L; ++l)
    if (p.H[l] % 64) return false;
  return true;
}
static bool row_map(CUtensorMap* m, const float* base, int width, long long rows) {
  return tc::make_map(m, base, width, rows, width, 32, 128, false);
}
// weight operand of the MMA fed by link i: layer `l` used as W^T (b_mn = false: W[n][k], F / T) or as W (A / B)
static bool weight_maps(const Plan& p, const Net& n, float* ws, int l, bool b_mn, chain::Maps& m, int i) {
  const int K = l == 1 ? p.ldx : p.H[l - 1];   // columns of layer l's matrix (H_l x K, row-major, ld = K)
  const int Hl = p.H[l];
  const float* hi = p.x3 ? ws + p.Whi[l] : n.W[l];
  const float* lo = p.x3 ? ws + p.Wlo[l] : n.W[l];
  const int box_n = chain_use_pair(p) ? Hl / 2 : Hl;   // each CTA of a pair loads its half of the N rows of W^T
  bool ok;
  if (!b_mn) {   // out = A[rows x K] * W^T: B[n = H_l][k]
    ok = tc::make_map(&m.whi[i], hi, K, Hl, K, 32, box_n, false) && tc::make_map(&m.wlo[i], lo, K, Hl, K, 32, box_n, false);
  } else {       // out = A[rows x H_l] * W: B[k = H_l][n = K]
    ok = tc::make_map(&m.whi[i], hi, K, Hl, K, 32, 32, true) && tc::make_map(&m.wlo[i], lo, K, Hl, K, 32, 32, true);
  }
  return ok;
}
static void chain_timing_begin(int& slot, const chain::Args& a, const char* what, cudaStream_t st) {
  slot = -1;
  if (!(g_timing && g_ntimed < kMaxTimed)) return;
  slot = g_ntimed++;
  if (!g_ev0[slot]) cudaEventCreate(&g_ev0[slot]), cudaEventCreate(&g_ev1[slot]);
  double flops = 0, arrays = 0;   // arrays in units of (rows x 1 column) floats
  for (int i = 0; i < a.nlinks; ++i) {
    const chain::LinkD& L = a.link[i];
    if (L.feeds) flops += 2.0 * (double)a.rows * L.width * L.n_next;
    arrays += (double)L.width * ((L.in0 ? 1 : 0) + (L.in2 ? 1 : 0) + (L.out0 ? 1 : 0) + (L.out2 ? 1 : 0));
  }
  g_flops[slot] = flops;
  g_bytes[slot] = 4.0 * (double)a.rows * arrays;
  g_timed_tc[slot] = true;
  g_timed_what[slot] = what;
  cudaEventRecord(g_ev0[slot], st);
}
template <int SWEEP>
static int chain_launch(const FbsnnSpec* s, const Plan& p, const chain::Maps& m, chain::Args& a, const char* what,
                        cudaStream_t st) {
  int slot;
  chain_timing_begin(slot, a, what, st);
  ++g_launches;
  cudaError_t e;
  if (chain_use_ta(p) && (p.x3 || SWEEP == chain::SWEEP_F || SWEEP == chain::SWEEP_B || chain_ta_mode() == 2)) {
    e = s->precision == FBSNN_PREC_TF32X3 ? chain::launch_chaint<SWEEP, true>(m, a, num_sms(), st)
                                          : chain::launch_chaint<SWEEP, false>(m, a, num_sms(), st);
  } else if (chain_use_pair(p)) {
    a.ntiles = (int)((p.rows + 255) / 256);   // 256-row tiles, one per CTA pair
    e = chain::launch_chain2<SWEEP>(m, a, num_sms(), st);
  } else {
    e = s->precision == FBSNN_PREC_TF32X3 ? chain::launch_chain<SWEEP, true>(m, a, num_sms(), st)
                                          : chain::launch_chain<SWEEP, false>(m, a, num_sms(), st);
  }
  if (slot >= 0) cudaEventRecord(g_ev1[slot], st);
  if (e != cudaSuccess) return fail(FBSNN_E_CUDA, "chained sweep %s: %s", what, cudaGetErrorString(e));
  static int debug = -1;
  if (debug < 0) debug = getenv("FBSNN_CHAIN_DEBUG") ? 1 : 0;
  if (debug) {   // bring-up aid: localise a faulting sweep (never set while capturing a graph)
    const cudaError_t e2 = cudaStreamSynchronize(st);
    fprintf(stderr, "[fbsnn] chained sweep %s (%d links, %d tiles): %s\n", what, a.nlinks, a.ntiles, cudaGetErrorString(e2));
    if (e2 != cudaSuccess) return fail(FBSNN_E_CUDA, "chained sweep %s failed: %s", what, cudaGetErrorString(e2));
  }
  return 0;
}
static int internal_act(const FbsnnSpec* s) {
  int act = s->act_kind;
  if (s->precision == FBSNN_PREC_TF32X3 && act == FBSNN_ACT_SINE) act = kActSineCW;
  if (s->precision == FBSNN_PREC_TF32 && act == FBSNN_ACT_SINE) act = kActSineFast;
  if (s->precision == FBSNN_PREC_TF32 && act == FBSNN_ACT_TANH) act = kActTanhFast;
  return act;
}
static void chain_base(const FbsnnSpec* s, const Plan& p, const Net& n, chain::Args& a) {
  memset(&a, 0, sizeof(a));
  a.rows = (int)p.rows, a.ntiles = (int)((p.rows + 127) / 128), a.act = internal_act(s);
  a.wout = n.wout, a.bout = n.bout;
  static int ablate = -1;
  if (ablate < 0) {
    const char* e = getenv("FBSNN_CHAIN_ABLATE");
    ablate = e ? atoi(e) : 0;
  }
  a.ablate = ablate;
  static int pf = -1, hints = -1;
  if (pf < 0) {
    const char* e = getenv("FBSNN_CHAIN_PF");
    pf = e ? atoi(e) : 7;
    const char* h = getenv("FBSNN_CHAIN_HINT");
    hints = h ? atoi(h) : 7;   // measured (M = 65 536, 3xTF32, sum of the four sweeps): 43.4 ms with 7, 43.7 with 5, 45.1 with 0
  }
  a.pf_dist = pf, a.hints = hints;
}

// F and A sweeps as two launches; leaves g_l, a_l, delta_l (, s_l for l < L), Y and Du_full in the workspace
static int chain_forward(const FbsnnSpec* s, const Plan& p, const Net& n, float* ws, bool with_grad, cudaStream_t st) {
  const long long R = p.rows;
  const int L = p.L;
  {
    chain::Maps m;
    chain::Args a;
    chain_base(s, p, n, a);
    a.nlinks = L + 1, a.Y = ws + p.Y;
    bool ok = true;
    a.link[0] = chain::LinkD{p.ldx, p.H[1], chain::LINK_FIRST, 1, 0, 0, 0, 1, 0, 0};
    ok = ok && row_map(&m.in0[0], ws + p.xin, p.ldx, R) && weight_maps(p, n, ws, 1, false, m, 0);
    for (int l = 1; l <= L; ++l) {
      const bool last = l == L;
      a.link[l] = chain::LinkD{p.H[l], last ? 0 : p.H[l + 1], (unsigned char)(last ? chain::LINK_LAST : chain::LINK_MID),
                               0, 0, 1, 1, (unsigned char)(last ? 0 : 1), 0, 0};
      a.bias[l] = n.b[l];
      ok = ok && row_map(&m.out0[l], ws + p.g[l], p.H[l], R) && row_map(&m.out2[l], ws + p.a[l], p.H[l], R);
      if (!last) ok = ok && weight_maps(p, n, ws, l + 1, false, m, l);
    }
    if (!ok) return fail(FBSNN_E_CUDA, "chained F sweep: tensor map encoding failed");
    int rc = chain_launch<chain::SWEEP_F>(s, p, m, a, "F*", st);
    if (rc) return rc;
  }
  {
    chain::Maps m;
    chain::Args a;
    chain_base(s, p, n, a);
    a.nlinks = L + 1, a.with_s = with_grad ? 1 : 0;
    bool ok = true;
    // link 0: delta_L = wout * a_L
    a.link[0] = chain::LinkD{p.H[L], L >= 2 ? p.H[L - 1] : p.ldx, chain::LINK_FIRST, 1, 0, 1, 0, 1, 1, 0};
    ok = ok && row_map(&m.in0[0], ws + p.a[L], p.H[L], R) && row_map(&m.out0[0], ws + p.delta[L], p.H[L], R) &&
         weight_maps(p, n, ws, L, true, m, 0);
    for (int k = 1; k <= L - 1; ++k) {
      const int l = L - k;
      a.link[k] = chain::LinkD{p.H[l], l >= 2 ? p.H[l - 1] : p.ldx, chain::LINK_MID, 1, (unsigned char)(with_grad ? 1 : 0), 1,
                               (unsigned char)(with_grad ? 1 : 0), 1, 1, 0};
      ok = ok && row_map(&m.in0[k], ws + p.a[l], p.H[l], R) && row_map(&m.out0[k], ws + p.delta[l], p.H[l], R) &&
           weight_maps(p, n, ws, l, true, m, k);
      if (with_grad) ok = ok && row_map(&m.in2[k], ws + p.g[l], p.H[l], R) && row_map(&m.out2[k], ws + p.szz[l], p.H[l], R);
    }
    a.link[L] = chain::LinkD{p.ldx, 0, chain::LINK_LAST, 0, 0, 1, 0, 0, 0, 0};
    ok = ok && row_map(&m.out0[L], ws + p.zf, p.ldx, R);
    if (!ok) return fail(FBSNN_E_CUDA, "chained A sweep: tensor map encoding failed");
    int rc = chain_launch<chain::SWEEP_A>(s, p, m, a, "A*", st);
    if (rc) return rc;
  }
  return 0;
}

// T and B sweeps as two launches + the column-sum finish (bias and output-weight gradients)
static int chain_backward(const FbsnnSpec* s, const Plan& p, const Net& n, float* ws, float* grads, cudaStream_t st) {
  const long long R = p.rows;
  const int L = p.L;
  chain::ColFinJobs fin{};
  float* colacc = ws + p.chain_col;
  {
    chain::Maps m;
    chain::Args a;
    chain_base(s, p, n, a);
    a.nlinks = L + 1, a.ybar = ws + p.ybar, a.colacc = colacc;
    bool ok = true;
    a.link[0] = chain::LinkD{p.ldx, p.H[1], chain::LINK_FIRST, 1, 0, 0, 0, 1, 0, 0};
    ok = ok && row_map(&m.in0[0], ws + p.V, p.ldx, R) && weight_maps(p, n, ws, 1, false, m, 0);
    for (int l = 1; l <= L - 1; ++l) {
      a.link[l] = chain::LinkD{p.H[l], p.H[l + 1], chain::LINK_MID, 1, 1, 1, 1, 1, 0, 0};
      ok = ok && row_map(&m.in0[l], ws + p.a[l], p.H[l], R) && row_map(&m.in2[l], ws + p.szz[l], p.H[l], R) &&
           row_map(&m.out0[l], ws + p.hd[l], p.H[l], R) && row_map(&m.out2[l], ws + p.szz[l], p.H[l], R) &&
           weight_maps(p, n, ws, l + 1, false, m, l);
    }
    // last hidden layer: zbar_L (stored, bias_L gradient) and dbar a + ybar g (output-weight gradient, column sums only)
    a.link[L] = chain::LinkD{p.H[L], 0, chain::LINK_LAST, 1, 1, 1, 0, 0, 0, 3};
    ok = ok && row_map(&m.in0[L], ws + p.a[L], p.H[L], R) && row_map(&m.in2[L], ws + p.g[L], p.H[L], R) &&
         row_map(&m.out0[L], ws + p.szz[L], p.H[L], R);
    if (!ok) return fail(FBSNN_E_CUDA, "chained T sweep: tensor map encoding failed");
    int rc = chain_launch<chain::SWEEP_T>(s, p, m, a, "T*", st);
    if (rc) return rc;
    fin.job[fin.njobs++] = chain::ColFinJob{L, 0, p.H[L], grads + s->off_b[L]};
    fin.job[fin.njobs++] = chain::ColFinJob{L, 1, p.H[L], grads + s->off_W[L + 1]};
    fin.nblk = chain_use_pair(p) ? chain::chain2_grid((int)((p.rows + 255) / 256), num_sms())
                                 : chain::chain_grid((int)((p.rows + 127) / 128), num_sms());
  }
  if (L >= 2) {
    chain::Maps m;
    chain::Args a;
    chain_base(s, p, n, a);
    a.nlinks = L, a.colacc = colacc;
    bool ok = true;
    a.link[0] = chain::LinkD{p.H[L], p.H[L - 1], chain::LINK_FIRST, 1, 0, 0, 0, 1, 1, 0};
    ok = ok && row_map(&m.in0[0], ws + p.szz[L], p.H[L], R) && weight_maps(p, n, ws, L, true, m, 0);
    for (int k = 1; k <= L - 1; ++k) {
      const int l = L - k;
      const bool last = l == 1;
      a.link[k] = chain::LinkD{p.H[l], last ? 0 : p.H[l - 1], (unsigned char)(last ? chain::LINK_LAST : chain::LINK_MID), 1, 1, 1,
                               0, (unsigned char)(last ? 0 : 1), 1, 1};
      ok = ok && row_map(&m.in0[k], ws + p.a[l], p.H[l], R) && row_map(&m.in2[k], ws + p.szz[l], p.H[l], R) &&
           row_map(&m.out0[k], ws + p.szz[l], p.H[l], R);
      if (!last) ok = ok && weight_maps(p, n, ws, l, true, m, k);
      fin.job[fin.njobs++] = chain::ColFinJob{k, 0, p.H[l], grads + s->off_b[l]};
    }
    if (!ok) return fail(FBSNN_E_CUDA, "chained B sweep: tensor map encoding failed");
    int rc = chain_launch<chain::SWEEP_B>(s, p, m, a, "B*", st);
    if (rc) return rc;
  }
  chain::chain_colsum_finish_kernel<<<dim3(8, fin.njobs), 256, 0, st>>>(colacc, fin);   // 8 lanes per column, widths <= 256
  LAUNCH_CHECK("chain_colsum_finish");
  return 0;
}

// F and A sweeps over `rows` rows whose inputs are already in ws[xin]; leaves Y in ws[Y], Du_full in ws[zf].
static int sweeps_forward(const FbsnnSpec* s, const Plan& p, const Net& n, float* ws, bool with_grad, cudaStream_t st) {
  const int R = (int)p.rows;
  if (chain_eligible(s, p)) {
    int rc = chain_forward(s, p, n, ws, with_grad, st);
    if (rc) return rc;
    if (s->clamp_u) {
      const long long thr = p.rows * 32;
      clamp_u_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(ws + p.Y, ws + p.zf, p.ldx, p.rows, ws + p.umask);
      LAUNCH_CHECK("clamp_u");
    }
    return 0;
  }
  int act = s->act_kind;
  // 3xTF32 keeps the fp32-grade sine (Cody-Waite + minimax polynomials, ~22 instructions): the F-sweep epilogue is
  // instruction-bound on it (3.3 ms per layer vs 2.8 ms with the MUFU form, measured), the price of fp32-grade results
  if (s->precision == FBSNN_PREC_TF32X3 && act == FBSNN_ACT_SINE) act = kActSineCW;
  if (s->precision == FBSNN_PREC_TF32 && act == FBSNN_ACT_SINE) act = kActSineFast;
  if (s->precision == FBSNN_PREC_TF32 && act == FBSNN_ACT_TANH) act = kActTanhFast;
  for (int l = 1; l <= p.L; ++l) {
    GemmArgs g{};
    g.M = R, g.N = p.H[l], g.Nb = p.H[l], g.kchunk = 0;
    if (l == 1) {
      g.nseg = 1;
      g.seg[0] = GemmSeg{ws + p.xin, n.W[1], p.ldx, p.ldw, p.kin};
    } else {
      g.nseg = 1;
      g.seg[0] = GemmSeg{ws + p.h[l - 1], n.W[l], p.H[l - 1], p.H[l - 1], p.H[l - 1]};
      if (p.nais) {
        g.nseg = 2;
        g.seg[1] = GemmSeg{ws + p.xin, n.Win[l], p.ldx, p.ldw, p.kin};
      }
    }
    EpiFwdT<true> e{};
    e.bias1 = n.b[l], e.bias2 = n.bin[l];
    const bool res = p.nais && l >= 2;
    e.res = res ? ws + p.h[l - 1] : nullptr;
    e.g = ws + p.g[l], e.a = ws + p.a[l];
    e.h = res ? ws + p.h[l] : nullptr;
    if (l == p.L) {
      e.wout = n.wout, e.delta = ws + p.delta[l], e.s = with_grad ? ws + p.szz[l] : nullptr;
    }
    e.ld = p.H[l], e.act = act;
    int rc = p.nais ? dense<true, true>(s, g, e, 1, st, "F") : dense<true, true>(s, g, narrow<EpiFwdT>(e), 1, st, "F");
    if (rc) return rc;
  }
  {
    const long long thr = p.rows * 32;
    head_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(ws + p.h[p.L], p.H[p.L], p.H[p.L], n.wout, n.bout, p.rows, ws + p.Y);
    LAUNCH_CHECK("head");
  }
  for (int l = p.L; l >= 2; --l) {
    GemmArgs g{};
    g.M = R, g.N = p.H[l - 1], g.Nb = p.H[l - 1], g.kchunk = 0, g.nseg = 1;
    g.seg[0] = GemmSeg{ws + p.delta[l], n.W[l], p.H[l], p.H[l - 1], p.H[l]};
    EpiAdjT<true> e{};
    e.a = ws + p.a[l - 1], e.g = ws + p.g[l - 1];
    if (p.nais) {
      if (l == p.L) e.res_head = n.wout; else e.res = ws + p.ht[l];
      e.ht_out = (l - 1 >= 2) ? ws + p.ht[l - 1] : nullptr;
    }
    e.delta = ws + p.delta[l - 1];
    e.s = with_grad ? ws + p.szz[l - 1] : nullptr;
    e.ld = p.H[l - 1], e.act = act;
    int rc = p.nais ? dense<true, false>(s, g, e, 1, st, "A") : dense<true, false>(s, g, narrow<EpiAdjT>(e), 1, st, "A");
    if (rc) return rc;
  }
  {
    GemmArgs g{};
    g.M = R, g.N = p.ldx, g.Nb = p.kin, g.kchunk = 0;
    g.nseg = 1;
    g.seg[0] = GemmSeg{ws + p.delta[1], n.W[1], p.H[1], p.ldw, p.H[1]};
    if (p.nais)
      for (int l = 2; l <= p.L; ++l) g.seg[g.nseg++] = GemmSeg{ws + p.delta[l], n.Win[l], p.H[l], p.ldw, p.H[l]};
    int rc = dense<true, false>(s, g, EpiStore{ws + p.zf, p.ldx}, 1, st, "Du");
    if (rc) return rc;
  }
  if (s->clamp_u) {
    const long long thr = p.rows * 32;
    clamp_u_kernel<<<(unsigned)((thr + 255) / 256), 256, 0, st>>>(ws + p.Y, ws + p.zf, p.ldx, p.rows, ws + p.umask);
    LAUNCH_CHECK("clamp_u");
  }
  return 0;
}

static int run_loss(const FbsnnSpec* s, const Plan& p, float* ws, bool with_grad, float* loss_out, float* ybsum_out, cudaStream_t st) {
  const ProblemK k = problem_k(s, p);
  LossArgs a{};
  a.xin = ws + p.xin, a.zf = ws + p.zf, a.sdw = ws + p.sdw, a.Y = ws + p.Y, a.rows = p.rows;
  a.ev = ws + p.ev, a.part = ws + p.part_loss;
  a.ybar = with_grad ? ws + p.ybar : nullptr, a.V = with_grad ? ws + p.V : nullptr;
  a.umask = s->clamp_u ? ws + p.umask : nullptr;
  // fused single pass, one warp per path: wins once there are enough paths to hide the row-to-row latency of a
  // warp's walk (1.5 vs 4.3 ms at M = 65 536); below ~2k paths the row-parallel pair is faster (18 vs 62 us at M = 100)
  if (with_grad && s->D <= 128 && p.rows / (s->N + 1) >= 2048) {
    loss_path_kernel<<<p.loss_blocks, 256, 0, st>>>(k, a, p.rows / (s->N + 1));
    LAUNCH_CHECK("loss_path");
  } else {
    loss_residual_kernel<<<p.loss_blocks, 256, 0, st>>>(k, a);
    LAUNCH_CHECK("loss_residual");
    if (with_grad) {
      loss_seed_kernel<<<p.loss_blocks, 256, 0, st>>>(k, a);
      LAUNCH_CHECK("loss_seed");
    }
  }
  final_sum2_kernel<<<1, 1024, 0, st>>>(ws + p.part_loss, p.loss_blocks, FinalSum2{loss_out, with_grad ? ybsum_out : nullptr});
  LAUNCH_CHECK("final_sum2");
  return 0;
}

// weight-gradient contraction of one layer into split-K partials; `defer` (nullable) collects the second-stage
// reduction for one batched launch, else it is launched right away
static int wgrad(const FbsnnSpec* s, const Plan& p, float* ws, const float* P0, const float* Q0, const float* P1,
                 const float* Q1, int out, int in_pad, int in_valid, int ldq, float* dst, int ld_dst, cudaStream_t st,
                 RedJobs* defer = nullptr, size_t part_off = 0, tc2g::BatchG* batch = nullptr) {
  GemmArgs g{};
  g.M = out, g.N = in_pad, g.Nb = in_pad, g.kchunk = p.wg_chunk, g.nseg = 2;
  g.seg[0] = GemmSeg{P0, Q0, out, ldq, (int)p.rows};
  g.seg[1] = GemmSeg{P1, Q1, out, ldq, (int)p.rows};
  float* part = ws + p.part_wg + part_off;
  // small batches: the contractions of all layers go into ONE launch of the TMEM-A pair kernel (flushed by the caller)
  if (batch && defer && s->precision == FBSNN_PREC_TF32X3 && tc_eligible<false, false>(g, p.wg_split) &&
      uses_pair<false, false>(s, g, p.wg_split) && gtmem_enabled() && tc2g_eligible(g, p.wg_split) &&
      tc2g_batch_add(*batch, g, EpiPartial{part, out, in_pad}, p.wg_split)) {
    defer->job[defer->njobs++] = RedJob{part, dst, out, in_pad, in_valid, ld_dst};
    return 0;
  }
  int rc = dense<false, false>(s, g, EpiPartial{part, out, in_pad}, p.wg_split, st, "G");
  if (rc) return rc;
  if (defer) {
    defer->job[defer->njobs++] = RedJob{part, dst, out, in_pad, in_valid, ld_dst};
    return 0;
  }
  const int n = out * in_pad;
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, p.wg_split, out, in_pad, in_valid, dst, ld_dst);
  LAUNCH_CHECK("reduce_partials");
  return 0;
}

static int sweeps_backward(const FbsnnSpec* s, const Plan& p, const Net& n, float* ws, float* grads, cudaStream_t st) {
  const int R = (int)p.rows;
  const int col_slots = std::max(p.col_blocks, 256);
  auto col_part = [&](int job) { return ws + p.part_col + (size_t)job * col_slots * 1024; };   // job l = bias of layer l
  bool fused_bias[kMaxL + 2] = {};
  int fused_grid[kMaxL + 2] = {};
  const bool chained = chain_eligible(s, p);
  if (chained) {
    int rc = chain_backward(s, p, n, ws, grads, st);
    if (rc) return rc;
  }
  // ---- T sweep -----------------------------------------------------------------------------------------
  for (int l = 1; l <= p.L && !chained; ++l) {
    GemmArgs g{};
    g.M = R, g.N = p.H[l], g.Nb = p.H[l], g.kchunk = 0;
    if (l == 1) {
      g.nseg = 1;
      g.seg[0] = GemmSeg{ws + p.V, n.W[1], p.ldx, p.ldw, p.kin};
    } else {
      g.nseg = 1;
      g.seg[0] = GemmSeg{ws + p.hd[l - 1], n.W[l], p.H[l - 1], p.H[l - 1], p.H[l - 1]};
      if (p.nais) {
        g.nseg = 2;
        g.seg[1] = GemmSeg{ws + p.V, n.Win[l], p.ldx, p.ldw, p.kin};
      }
    }
    EpiTanT<true> e{};
    e.a = ws + p.a[l], e.s_zz = ws + p.szz[l];
    e.res = (p.nais && l >= 2) ? ws + p.hd[l - 1] : nullptr;
    e.hd = ws + p.hd[l];
    if (l == p.L) {
      e.ybar = ws + p.ybar, e.wout = n.wout;
      if (uses_tc<true, true>(s, g, 1))   // bias_L gradient fused
        e.colpart = col_part(l), fused_bias[l] = true, fused_grid[l] = tc_launch_grid<true, true>(s, g, 1);
    }
    e.ld = p.H[l];
    int rc = p.nais ? dense<true, true>(s, g, e, 1, st, "T") : dense<true, true>(s, g, narrow<EpiTanT>(e), 1, st, "T");
    if (rc) return rc;
  }
  // ---- B sweep -----------------------------------------------------------------------------------------
  for (int l = p.L; l >= 2 && !chained; --l) {
    GemmArgs g{};
    g.M = R, g.N = p.H[l - 1], g.Nb = p.H[l - 1], g.kchunk = 0, g.nseg = 1;
    g.seg[0] = GemmSeg{ws + p.szz[l], n.W[l], p.H[l], p.H[l - 1], p.H[l]};
    EpiBwdT<true> e{};
    e.a = ws + p.a[l - 1], e.zz_zbar = ws + p.szz[l - 1];
    if (p.nais) {
      if (l == p.L) e.ybar = ws + p.ybar, e.wout = n.wout; else e.res = ws + p.hb[l];
      e.hb_out = (l - 1 >= 2) ? ws + p.hb[l - 1] : nullptr;
    }
    e.ld = p.H[l - 1];
    if (uses_tc<true, false>(s, g, 1))
      e.colpart = col_part(l - 1), fused_bias[l - 1] = true, fused_grid[l - 1] = tc_launch_grid<true, false>(s, g, 1);
    int rc = p.nais ? dense<true, false>(s, g, e, 1, st, "B") : dense<true, false>(s, g, narrow<EpiBwdT>(e), 1, st, "B");
    if (rc) return rc;
  }
  // ---- G contractions ------------------------------------------------------------------------------------
  RedJobs red{};
  red.nsplit = p.wg_split;
  // the contractions run back to back into their own partial buffers, one reduction launch after (NAIS-Net networks too
  // deep for kMaxRedJobs buffers reduce every contraction on its own)
  RedJobs* defer = (!p.nais || p.wg_slots > 1) ? &red : nullptr;
  const size_t wg_stride = p.wg_stride;
  int slot = 0;
  // below ~2 row tiles per SM a contraction is a few k-blocks per CTA: all of them in one launch (gemm_tc2g_batched_kernel)
  tc2g::BatchG bt;
  bt.njobs = 0, bt.nsplit = 0;
  static int gbatch = -1;
  if (gbatch < 0) {
    const char* e = getenv("FBSNN_GBATCH");
    gbatch = (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : 1;   // M = 100: one launch of 39 us instead of four of 14-17 us
  }
  tc2g::BatchG* batch = (gbatch && defer && p.rows <= (long long)num_sms() * 256) ? &bt : nullptr;
  auto flush_batch = [&]() -> int {
    if (!batch || bt.njobs == 0) return 0;
    int tslot = -1;
    if (g_timing && g_ntimed < kMaxTimed) {
      tslot = g_ntimed++;
      if (!g_ev0[tslot]) cudaEventCreate(&g_ev0[tslot]), cudaEventCreate(&g_ev1[tslot]);
      double fl = 0, by = 0;
      for (int j = 0; j < bt.njobs; ++j) {
        double k = 0;
        for (int i = 0; i < bt.g[j].nseg; ++i) k += bt.g[j].seg[i].K;
        fl += 2.0 * (double)bt.g[j].M * (double)bt.g[j].N * k;
        by += 4.0 * (k * ((double)bt.g[j].M + (double)bt.g[j].N) + (double)bt.g[j].M * (double)bt.g[j].N * bt.nsplit);
      }
      g_flops[tslot] = fl, g_bytes[tslot] = by, g_timed_tc[tslot] = true, g_timed_what[tslot] = "G";
      cudaEventRecord(g_ev0[tslot], st);
    }
    ++g_launches;
    const cudaError_t e = launch_gemm_tc2g_batched(bt, num_sms(), st);
    if (tslot >= 0) cudaEventRecord(g_ev1[tslot], st);
    bt.njobs = 0;
    if (e != cudaSuccess) return fail(FBSNN_E_CUDA, "tcgen05 batched weight gradients: %s", cudaGetErrorString(e));
    return 0;
  };
  auto nais_wbar = [&](int l) -> int {   // gradient through the stability projection: Wbar_l = W_l (Rbar + Rbar^T)
    const int H = p.H[l];
    nais_project_bwd_kernel<<<(H * H + 1023) / 1024, 1024, 0, st>>>(ws + p.Bbar[l], ws + p.Rm[l], H, ws + p.nstate[l], ws + p.Sm[l]);
    LAUNCH_CHECK("nais_project_bwd");
    GemmArgs g{};
    g.M = H, g.N = H, g.Nb = H, g.kchunk = 0, g.nseg = 1;
    g.seg[0] = GemmSeg{n.Wraw[l], ws + p.Sm[l], H, H, H};
    return dense<true, false>(s, g, EpiStore{grads + s->off_W[l], H}, 1, st, "nais Wbar", false);
  };
  for (int l = 1; l <= p.L; ++l) {
    int rc;
    if (l == 1) {
      rc = wgrad(s, p, ws, ws + p.szz[1], ws + p.xin, ws + p.delta[1], ws + p.V, p.H[1], p.ldx, p.d_in, p.ldx,
                 grads + s->off_W[1], p.d_in, st, defer, defer ? (size_t)(slot++) * wg_stride : 0, batch);
      if (rc) return rc;
      continue;
    }
    float* dst = p.nais ? ws + p.Bbar[l] : grads + s->off_W[l];
    rc = wgrad(s, p, ws, ws + p.szz[l], ws + p.h[l - 1], ws + p.delta[l], ws + p.hd[l - 1], p.H[l], p.H[l - 1],
               p.H[l - 1], p.H[l - 1], dst, p.H[l - 1], st, defer, defer ? (size_t)(slot++) * wg_stride : 0, batch);
    if (rc) return rc;
    if (p.nais) {
      rc = wgrad(s, p, ws, ws + p.szz[l], ws + p.xin, ws + p.delta[l], ws + p.V, p.H[l], p.ldx, p.d_in, p.ldx,
                 grads + s->off_Win[l], p.d_in, st, defer, defer ? (size_t)(slot++) * wg_stride : 0, batch);
      if (rc) return rc;
      if (!defer && (rc = nais_wbar(l))) return rc;
    }
  }
  {
    int rc = flush_batch();
    if (rc) return rc;
  }
  if (defer && red.njobs > 0) {
    int maxn = 0;
    for (int i = 0; i < red.njobs; ++i) maxn = std::max(maxn, red.job[i].rows * red.job[i].cols_pad);
    reduce_partials_batched_kernel<<<dim3((maxn + 255) / 256, red.njobs), 256, 0, st>>>(red);
    LAUNCH_CHECK("reduce_partials_batched");
  }
  if (p.nais && defer)
    for (int l = 2; l <= p.L; ++l) {
      int rc = nais_wbar(l);
      if (rc) return rc;
    }
  if (chained) return 0;   // bias and output-weight gradients came out of the chained sweeps' fused column sums
  // ---- bias / output-layer gradients: column sums -----------------------------------------------------------
  ColJobs js{};
  js.rows = p.rows, js.rows_per_block = p.col_rows_per_block, js.max_width = 1024;
  int maxw = 0;
  for (int l = 1; l <= p.L; ++l) {
    ColJob& j = js.job[js.njobs++];
    j.A = fused_bias[l] ? nullptr : ws + p.szz[l];
    j.B = nullptr, j.y = nullptr;
    j.out = grads + s->off_b[l];
    j.out2 = (p.nais && l >= 2) ? grads + s->off_bin[l] : nullptr;
    j.part = col_part(l), j.nblk = fused_bias[l] ? fused_grid[l] : p.col_blocks;
    j.ld = p.H[l], j.width = p.H[l];
    maxw = std::max(maxw, p.H[l]);
  }
  {
    ColJob& j = js.job[js.njobs++];
    j.A = ws + p.hd[p.L], j.B = ws + p.h[p.L], j.y = ws + p.ybar;
    j.out = grads + s->off_W[p.L + 1], j.out2 = nullptr;
    j.part = col_part(0), j.nblk = p.col_blocks;
    j.ld = p.H[p.L], j.width = p.H[p.L];
  }
  if (maxw > 1024) return fail(FBSNN_E_UNSUPPORTED, "hidden width %d > 1024", maxw);
  colsum_stage1_kernel<<<dim3(p.col_blocks, js.njobs), 256, 0, st>>>(js);
  LAUNCH_CHECK("colsum1");
  colsum_stage2_kernel<<<dim3((maxw + 255) / 256, js.njobs), 256, 0, st>>>(js);
  LAUNCH_CHECK("colsum2");
  return 0;
}

static int check_ws(const Plan& p, const void* ws, size_t bytes) {
  if (!ws) return fail(FBSNN_E_WORKSPACE, "workspace is null");
  if (((uintptr_t)ws & 255) != 0) return fail(FBSNN_E_WORKSPACE, "workspace must be 256-byte aligned");
  if (bytes < p.total * sizeof(float))
    return fail(FBSNN_E_WORKSPACE, "workspace too small: %zu < %zu bytes", bytes, p.total * sizeof(float));
  return 0;
}

static int gen_increments(const FbsnnSpec* s, const Plan& p, float* ws, long long M, float T, long long path_offset,
                          uint64_t seed, uint64_t iteration, const long long* iter_dev, const float* chol,
                          cudaStream_t st) {
  const float sqrt_dt = (float)sqrt((double)T / (double)s->N);
  const long long total = M * (long long)(s->N + 1) * ((p.Dn + 3) / 4);
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)num_sms() * 32);
  float* target = chol ? ws + p.zraw : ws + p.inc;
  brownian_increments_kernel<<<blocks, 256, 0, st>>>(target, M, s->N, p.Dn, sqrt_dt, p.ldi, path_offset, seed,
                                                     iteration, iter_dev);
  LAUNCH_CHECK("brownian_increments");
  if (chol) {  // inc[r, i] = sum_j zraw[r, j] * L[i][j]   (np.einsum('ij,mnj->mni'), with_corr...:339-341)
    GemmArgs g{};
    g.M = (int)p.rows, g.N = p.ldi, g.Nb = p.Dn, g.kchunk = 0, g.nseg = 1;
    g.seg[0] = GemmSeg{ws + p.zraw, chol, p.ldi, p.Dn, p.Dn};
    int rc = dense<true, true>(s, g, EpiStore{ws + p.inc, p.ldi}, 1, st, "chol", false);
    if (rc) return rc;
  }
  return 0;
}

static int loss_grad_impl(const FbsnnSpec* s, const float* params, float* grads, const float* t, const float* W,
                          const float* Xi, int64_t xi_rows, int64_t M, float T, int64_t path_offset, uint64_t seed,
                          uint64_t iteration, const long long* iter_dev, const float* chol, void* workspace,
                          size_t wbytes, float* X_out, float* Y_out, float* Z_out, float* loss_out, bool with_grad,
                          cudaStream_t st) {
  int rc = validate(s);
  if (rc) return rc;
  if (!params || !Xi || M < 1 || (xi_rows != 1 && xi_rows != M)) return fail(FBSNN_E_BADARG, "bad params/Xi/M/xi_rows");
  if (with_grad && !grads) return fail(FBSNN_E_BADARG, "grads is null");
  if (W && !t) return fail(FBSNN_E_BADARG, "t is required with host-supplied W");
  if (!W && !with_grad) return fail(FBSNN_E_BADARG, "forward needs (t, W)");
  const long long rows = (long long)M * (s->N + 1);
  if (rows > 0x7fffffffLL / 4) return fail(FBSNN_E_UNSUPPORTED, "too many rows (%lld); shard the paths", rows);
  Plan p;
  make_plan(s, rows, with_grad, p);
  rc = check_ws(p, workspace, wbytes);
  if (rc) return rc;
  float* ws = (float*)workspace;
  const Net n = bind_net(s, p, params, ws);
  if ((rc = prepare_weights(s, p, params, ws, st))) return rc;
  if (p.nais && (rc = nais_prepare(s, p, n, ws, st))) return rc;
  if ((rc = split_weights(s, p, n, params, ws, st))) return rc;
  if (!W && (rc = gen_increments(s, p, ws, M, T, path_offset, seed, iteration, iter_dev, chol, st))) return rc;
  {
    const ProblemK k = problem_k(s, p);
    PathArgs a{};
    a.t = W ? t : nullptr, a.W = W, a.inc = W ? nullptr : ws + p.inc, a.ldi = p.ldi;
    a.Xi = Xi, a.xi_rows = xi_rows, a.M = M, a.T = T;
    a.xin = ws + p.xin, a.sdw = ws + p.sdw, a.X_out = X_out;
    if (s->sigma_kind == FBSNN_SIGMA_HESTON) {
      path_advance_heston_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(k, a);
    } else {
      const long long thr = (long long)M * s->D;
      path_advance_kernel<<<(unsigned)((thr + 127) / 128), 128, 0, st>>>(k, a);
    }
    LAUNCH_CHECK("path_advance");
  }
  if ((rc = sweeps_forward(s, p, n, ws, with_grad, st))) return rc;
  float* ybsum = with_grad ? grads + s->off_b[p.L + 1] : nullptr;
  if (loss_out || with_grad) {
    if ((rc = run_loss(s, p, ws, with_grad, loss_out, ybsum, st))) return rc;
  }
  if (Y_out) CU_CHECK(cudaMemcpyAsync(Y_out, ws + p.Y, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (Z_out) {
    const long long nz = rows * s->D;
    gather_z_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>(ws + p.zf, p.ldx, s->D, rows, Z_out);
    LAUNCH_CHECK("gather_z");
  }
  if (with_grad && (rc = sweeps_backward(s, p, n, ws, grads, st))) return rc;
  return 0;
}

static int adam_impl(const FbsnnAdam* hp, float* params, const float* grads, float* m, float* v, int64_t n,
                     void* opt_state, float* gsq_part, int gsq_blocks, cudaStream_t st, bool have_gradsq = false) {
  if (!hp || !params || !grads || !m || !v || !opt_state || n < 1) return fail(FBSNN_E_BADARG, "adam: null argument");
  if (!have_gradsq) {
    gradsq_kernel<<<gsq_blocks, 256, 0, st>>>(grads, n, gsq_part);
    LAUNCH_CHECK("gradsq");
  }
  opt_prepare_kernel<<<1, 256, 0, st>>>(gsq_part, gsq_blocks, *hp, (OptState*)opt_state);
  LAUNCH_CHECK("opt_prepare");
  const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)num_sms() * 8);
  adam_kernel<<<blocks, 256, 0, st>>>(params, grads, m, v, n, (float)hp->beta1, (float)hp->beta2, (float)hp->eps,
                                      (const OptState*)opt_state);
  LAUNCH_CHECK("adam");
  return 0;
}

}  // namespace fbsnn

using namespace fbsnn;

extern "C" {

const char* fbsnn_last_error(void) { return g_err; }
int fbsnn_version(void) { return 103; }
// Run-time switches (tests and A/B measurements): "chain" = 0 | 1 | 2 (layer-chained sweeps: off / auto / always when
// eligible).  Returns the previous value, or FBSNN_E_BADARG for an unknown name.
int fbsnn_set_option(const char* name, int value) {
  if (name && !strcmp(name, "chain")) {
    const int old = chain_mode();
    if (value < 0 || value > 2) return fail(FBSNN_E_BADARG, "option chain takes 0, 1 or 2");
    g_opt_chain = value;
    return old;
  }
  if (name && !strcmp(name, "chain_pair")) {
    const int old = chain_pair_mode();
    if (value < 0 || value > 1) return fail(FBSNN_E_BADARG, "option chain_pair takes 0 or 1");
    g_opt_chain_pair = value;
    return old;
  }
  if (name && !strcmp(name, "chain_ta")) {
    const int old = chain_ta_mode();
    if (value < 0 || value > 2) return fail(FBSNN_E_BADARG, "option chain_ta takes 0, 1 or 2");
    g_opt_chain_ta = value;
    return old;
  }
  return fail(FBSNN_E_BADARG, "unknown option %s", name ? name : "(null)");
}
#ifdef FBSNN_CHAIN_PROF
// profiling builds only (tools/chain_prof.py): copies (and optionally clears) the 160 x 24 cycle counters of chaint_kernel
extern "C" int fbsnn_debug_chain_prof(unsigned long long* out, int reset) {
  if (out && cudaMemcpyFromSymbol(out, chain::g_chain_prof, sizeof(chain::g_chain_prof)) != cudaSuccess) return FBSNN_E_CUDA;
  if (reset) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, chain::g_chain_prof) != cudaSuccess) return FBSNN_E_CUDA;
    cudaMemset(p, 0, sizeof(chain::g_chain_prof));
  }
  return 0;
}
#endif
#ifdef FBSNN_CHAIN_PROF
extern "C" int fbsnn_debug_chain_trap(unsigned int* out8, int reset) {
  if (out8 && cudaMemcpyFromSymbol(out8, chain::g_chain_trap, sizeof(chain::g_chain_trap)) != cudaSuccess) return FBSNN_E_CUDA;
  if (reset) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, chain::g_chain_trap) != cudaSuccess) return FBSNN_E_CUDA;
    cudaMemset(p, 0, sizeof(chain::g_chain_trap));
  }
  return 0;
}
#endif
long long fbsnn_launch_count(void) { return g_launches; }
void fbsnn_dense_timing(int enable) { g_timing = enable != 0; g_ntimed = 0; }
// Sums the recorded dense-layer launches (caller has synchronised): out = {n_launches, total ms, total FLOPs,
// n tcgen05 launches, tcgen05 ms, tcgen05 FLOPs, total algorithmic bytes, tcgen05 algorithmic bytes}.
int fbsnn_dense_timing_read(double* out8) {
  if (!out8) return FBSNN_E_BADARG;
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  for (int i = 0; i < g_ntimed; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]) != cudaSuccess) return fail(FBSNN_E_CUDA, "event not complete");
    out8[0] += 1, out8[1] += ms, out8[2] += g_flops[i], out8[6] += g_bytes[i];
    if (g_timed_tc[i]) out8[3] += 1, out8[4] += ms, out8[5] += g_flops[i], out8[7] += g_bytes[i];
  }
  return 0;
}

// One recorded dense launch (caller has synchronised): out4 = {ms, algorithmic FLOPs, algorithmic bytes, 1 if tcgen05};
// returns the sweep tag ("F", "A", "Du", "T", "B", "G", ...) or NULL when i is out of range.
const char* fbsnn_dense_timing_entry(int i, double* out4) {
  if (i < 0 || i >= g_ntimed || !out4) return nullptr;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, g_ev0[i], g_ev1[i]) != cudaSuccess) return nullptr;
  out4[0] = ms, out4[1] = g_flops[i], out4[2] = g_bytes[i], out4[3] = g_timed_tc[i] ? 1.0 : 0.0;
  return g_timed_what[i];
}

// Test hook: one dense GEMM C = A * B with a plain store epilogue, on the SIMT or the tcgen05 kernel.
//   a_kc: A[m*lda + k] (1) or A[k*lda + m] (0);  b_kc: B[n*ldb + k] (1) or B[k*ldb + n] (0)
int fbsnn_debug_gemm(int a_kc, int b_kc, int use_tc, int M, int N, int K, const float* A, int lda, const float* B,
                     int ldb, float* C, int ldc, void* stream) {
  GemmArgs g{};
  g.nseg = 1, g.M = M, g.N = N, g.Nb = N, g.kchunk = 0;
  g.seg[0] = GemmSeg{A, B, lda, ldb, K};
  const EpiStore e{C, ldc};
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t err;
  if (use_tc) {
    bool ok = a_kc && b_kc ? tc_eligible<true, true>(g, 1) : (a_kc ? tc_eligible<true, false>(g, 1) : tc_eligible<false, false>(g, 1));
    if (!ok || (!a_kc && b_kc)) return fail(FBSNN_E_UNSUPPORTED, "shape not eligible for the tcgen05 kernel");
    if (use_tc == 5) {   // CTA-pair weight-gradient kernel with A in tensor memory: split-K over K, partials summed here
      if (a_kc || b_kc) return fail(FBSNN_E_UNSUPPORTED, "the TMEM-A kernel takes the weight-gradient layout only");
      GemmArgs gg = g;
      const int nsp = std::max(1, std::min(8, K / 64));
      gg.kchunk = ((K + nsp - 1) / nsp + 31) / 32 * 32;
      const int nsplit = (K + gg.kchunk - 1) / gg.kchunk;
      if (!tc2g_eligible(gg, nsplit)) return fail(FBSNN_E_UNSUPPORTED, "shape not eligible for the TMEM-A kernel");
      float* part = nullptr;
      if (cudaMalloc(&part, (size_t)nsplit * M * N * sizeof(float)) != cudaSuccess) return fail(FBSNN_E_CUDA, "cudaMalloc");
      err = launch_gemm_tc2g(gg, EpiPartial{part, M, N}, nsplit, num_sms(), st);
      if (err == cudaSuccess) {
        reduce_partials_kernel<<<(M * N + 255) / 256, 256, 0, st>>>(part, nsplit, M, N, N, C, ldc);
        err = cudaGetLastError();
      }
      cudaStreamSynchronize(st);
      cudaFree(part);
    } else if (use_tc == 4) {   // CTA-pair kernel, B pre-split into exact-TF32 hi / lo twins (the sweeps' form); test hook only
      if (!a_kc) return fail(FBSNN_E_UNSUPPORTED, "pre-split B is the sweeps' form (A k-contiguous)");
      if (!(b_kc ? tc2_eligible<true, true>(g, 1) : tc2_eligible<true, false>(g, 1)))
        return fail(FBSNN_E_UNSUPPORTED, "shape not eligible for the CTA-pair tcgen05 kernel");
      const int nb = b_kc ? N * ldb : K * ldb;
      float* hl = nullptr;
      if (cudaMalloc(&hl, 2 * (size_t)nb * sizeof(float)) != cudaSuccess) return fail(FBSNN_E_CUDA, "cudaMalloc");
      split_hi_lo_kernel<<<(nb + 255) / 256, 256, 0, st>>>(B, nb, hl, hl + nb);
      GemmArgs g3 = g;
      g3.seg[0].B = hl;
      g3.seg[4] = g.seg[0], g3.seg[4].B = hl + nb;
      err = b_kc ? launch_gemm_tc2<true, true, 3>(g3, e, 1, num_sms(), st) : launch_gemm_tc2<true, false, 3>(g3, e, 1, num_sms(), st);
      cudaStreamSynchronize(st);
      cudaFree(hl);
    } else if (use_tc == 3) {   // CTA-pair kernel, both operands split in-kernel
      bool ok2 = a_kc && b_kc ? tc2_eligible<true, true>(g, 1) : (a_kc ? tc2_eligible<true, false>(g, 1) : tc2_eligible<false, false>(g, 1));
      if (!ok2) return fail(FBSNN_E_UNSUPPORTED, "shape not eligible for the CTA-pair tcgen05 kernel");
      if (a_kc && b_kc) err = launch_gemm_tc2<true, true, 2>(g, e, 1, num_sms(), st);
      else if (a_kc) err = launch_gemm_tc2<true, false, 2>(g, e, 1, num_sms(), st);
      else err = launch_gemm_tc2<false, false, 2>(g, e, 1, num_sms(), st);
    } else if (use_tc == 2) {
      if (a_kc && b_kc) err = launch_gemm_tc<true, true, 2>(g, e, 1, num_sms(), st);
      else if (a_kc) err = launch_gemm_tc<true, false, 2>(g, e, 1, num_sms(), st);
      else err = launch_gemm_tc<false, false, 2>(g, e, 1, num_sms(), st);
    } else {
      if (a_kc && b_kc) err = launch_gemm_tc<true, true, 0>(g, e, 1, num_sms(), st);
      else if (a_kc) err = launch_gemm_tc<true, false, 0>(g, e, 1, num_sms(), st);
      else err = launch_gemm_tc<false, false, 0>(g, e, 1, num_sms(), st);
    }
  } else {
    if (a_kc && b_kc) err = launch_gemm<true, true>(g, e, 1, num_sms(), st);
    else if (a_kc) err = launch_gemm<true, false>(g, e, 1, num_sms(), st);
    else if (!b_kc) err = launch_gemm<false, false>(g, e, 1, num_sms(), st);
    else return fail(FBSNN_E_UNSUPPORTED, "layout");
  }
  if (err != cudaSuccess) return fail(FBSNN_E_CUDA, "debug gemm: %s", cudaGetErrorString(err));
  return 0;
}

// Test hook: offset (in floats) of a named per-row array of the workspace plan, so that a test can compare the arrays
// two dispatch variants leave behind.  name in {"xin","Y","zf","V","ybar","g","a","delta","szz","hd"}; layer 1..L for
// the per-layer arrays.  Returns the row width in *width_out.
int fbsnn_debug_ws_offset(const FbsnnSpec* spec, int64_t n_paths, int with_grad, const char* name, int layer,
                          int64_t* offset_out, int* width_out) {
  int rc = validate(spec);
  if (rc) return rc;
  if (n_paths < 1 || !name || !offset_out || !width_out) return fail(FBSNN_E_BADARG, "debug_ws_offset: bad argument");
  Plan p;
  make_plan(spec, (long long)n_paths * (spec->N + 1), with_grad != 0, p);
  const bool per_layer = !strcmp(name, "g") || !strcmp(name, "a") || !strcmp(name, "delta") || !strcmp(name, "szz") || !strcmp(name, "hd");
  if (per_layer && (layer < 1 || layer > p.L)) return fail(FBSNN_E_BADARG, "debug_ws_offset: layer outside 1..L");
  size_t off;
  int w;
  if (!strcmp(name, "xin")) off = p.xin, w = p.ldx;
  else if (!strcmp(name, "Y")) off = p.Y, w = 1;
  else if (!strcmp(name, "zf")) off = p.zf, w = p.ldx;
  else if (!strcmp(name, "V")) off = p.V, w = p.ldx;
  else if (!strcmp(name, "ybar")) off = p.ybar, w = 1;
  else if (!strcmp(name, "g")) off = p.g[layer], w = p.H[layer];
  else if (!strcmp(name, "a")) off = p.a[layer], w = p.H[layer];
  else if (!strcmp(name, "delta")) off = p.delta[layer], w = p.H[layer];
  else if (!strcmp(name, "szz")) off = p.szz[layer], w = p.H[layer];
  else if (!strcmp(name, "hd")) off = p.hd[layer], w = p.H[layer];
  else return fail(FBSNN_E_BADARG, "debug_ws_offset: unknown array %s", name);
  if (!with_grad && (!strcmp(name, "V") || !strcmp(name, "ybar") || !strcmp(name, "szz") || !strcmp(name, "hd")))
    return fail(FBSNN_E_BADARG, "debug_ws_offset: %s exists only in the training plan", name);
  *offset_out = (int64_t)off, *width_out = w;
  return 0;
}

int fbsnn_workspace_bytes(const FbsnnSpec* spec, int64_t n_paths, int with_grad, size_t* bytes_out) {
  int rc = validate(spec);
  if (rc) return rc;
  if (n_paths < 1 || !bytes_out) return fail(FBSNN_E_BADARG, "n_paths < 1 or bytes_out null");
  Plan p;
  make_plan(spec, (long long)n_paths * (spec->N + 1), with_grad != 0, p);
  *bytes_out = p.total * sizeof(float);
  return 0;
}

int fbsnn_fetch_minibatch(const FbsnnSpec* spec, float T, int64_t n_paths, int64_t path_offset, uint64_t seed,
                          uint64_t iteration, const float* chol, void* workspace, size_t workspace_bytes,
                          float* t_out, float* W_out, void* stream) {
  int rc = validate(spec);
  if (rc) return rc;
  if (n_paths < 1 || !W_out) return fail(FBSNN_E_BADARG, "n_paths < 1 or W_out null");
  Plan p;
  make_plan(spec, (long long)n_paths * (spec->N + 1), true, p);
  if ((rc = check_ws(p, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  if ((rc = gen_increments(spec, p, ws, n_paths, T, path_offset, seed, iteration, nullptr, chol, st))) return rc;
  const long long thr = (long long)n_paths * p.Dn;
  cumsum_paths_kernel<<<(unsigned)((thr + 127) / 128), 128, 0, st>>>(ws + p.inc, p.ldi, W_out, t_out, n_paths, spec->N, p.Dn, T);
  LAUNCH_CHECK("cumsum_paths");
  return 0;
}

int fbsnn_net_u(const FbsnnSpec* spec, const float* params, const float* t, const float* X, int64_t rows,
                void* workspace, size_t workspace_bytes, float* u_out, float* du_out, void* stream) {
  int rc = validate(spec);
  if (rc) return rc;
  if (!params || !t || !X || rows < 1) return fail(FBSNN_E_BADARG, "net_u: null argument");
  Plan p;
  make_plan(spec, rows, false, p);
  if ((rc = check_ws(p, workspace, workspace_bytes))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  const Net n = bind_net(spec, p, params, ws);
  if ((rc = prepare_weights(spec, p, params, ws, st))) return rc;
  if (p.nais && (rc = nais_prepare(spec, p, n, ws, st))) return rc;
  if ((rc = split_weights(spec, p, n, params, ws, st))) return rc;
  const long long tot = rows * p.ldx;
  pack_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(t, X, rows, spec->D, p.ldx, ws + p.xin);
  LAUNCH_CHECK("pack_rows");
  if ((rc = sweeps_forward(spec, p, n, ws, false, st))) return rc;
  if (u_out) CU_CHECK(cudaMemcpyAsync(u_out, ws + p.Y, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (du_out) {
    const long long nz = rows * spec->D;
    gather_z_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>(ws + p.zf, p.ldx, spec->D, rows, du_out);
    LAUNCH_CHECK("gather_z");
  }
  return 0;
}

int fbsnn_forward(const FbsnnSpec* spec, const float* params, const float* t, const float* W, const float* Xi,
                  int64_t xi_rows, int64_t n_paths, void* workspace, size_t workspace_bytes, float* X_out,
                  float* Y_out, float* Z_out, float* loss_out, void* stream) {
  return loss_grad_impl(spec, params, nullptr, t, W, Xi, xi_rows, n_paths, 0.f, 0, 0, 0, nullptr, nullptr, workspace,
                        workspace_bytes, X_out, Y_out, Z_out, loss_out, false, (cudaStream_t)stream);
}

int fbsnn_loss_grad(const FbsnnSpec* spec, const float* params, float* grads, const float* t, const float* W,
                    const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                    uint64_t seed, uint64_t iteration, const float* chol, void* workspace,
                    size_t workspace_bytes, float* X_out, float* Y_out, float* Z_out, float* loss_out,
                    void* stream) {
  return loss_grad_impl(spec, params, grads, t, W, Xi, xi_rows, n_paths, T, path_offset, seed, iteration, nullptr,
                        chol, workspace, workspace_bytes, X_out, Y_out, Z_out, loss_out, true, (cudaStream_t)stream);
}

int fbsnn_loss_grad_step(const FbsnnSpec* spec, const float* params, float* grads, const float* t, const float* W,
                         const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                         uint64_t seed, const float* chol, const void* opt_state, void* workspace,
                         size_t workspace_bytes, float* X_out, float* Y_out, float* loss_out, void* stream) {
  if (!opt_state) return fail(FBSNN_E_BADARG, "opt_state is null");
  return loss_grad_impl(spec, params, grads, t, W, Xi, xi_rows, n_paths, T, path_offset, seed, 0,
                        &((const OptState*)opt_state)->rng_iter, chol, workspace, workspace_bytes, X_out, Y_out, nullptr,
                        loss_out, true, (cudaStream_t)stream);
}

int fbsnn_track_min(const float* loss, float* state, const void* opt_state, const float* X, float* X_best, int64_t n_x,
                    const float* Y, float* Y_best, int64_t n_y, void* stream) {
  if (!loss || !state || n_x < 0 || n_y < 0 || n_x % 4 || n_y % 4 || (n_x && (!X || !X_best)) || (n_y && (!Y || !Y_best)))
    return fail(FBSNN_E_BADARG, "track_min: null pointer or element counts that are not multiples of 4");
  if ((((uintptr_t)X | (uintptr_t)X_best | (uintptr_t)Y | (uintptr_t)Y_best) & 15) != 0)
    return fail(FBSNN_E_BADARG, "track_min: arrays must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  track_min_decide_kernel<<<1, 1, 0, st>>>(loss, state, (const OptState*)opt_state);
  LAUNCH_CHECK("track_min_decide");
  const long long n4 = (n_x + n_y) / 4;
  if (n4 > 0) {
    const unsigned blocks = (unsigned)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
    track_min_copy_kernel<<<blocks, 256, 0, st>>>(state, (const float4*)X, (float4*)X_best, n_x / 4, (const float4*)Y,
                                                  (float4*)Y_best, n_y / 4);
    LAUNCH_CHECK("track_min_copy");
  }
  return 0;
}

int fbsnn_adam_step(const FbsnnAdam* host_hp, float* params, const float* grads, float* exp_avg,
                    float* exp_avg_sq, int64_t n_params, void* opt_state, void* stream) {
  // scratch for the squared-norm partials lives behind the 64-byte state block (FBSNN_OPT_STATE_BYTES)
  return adam_impl(host_hp, params, grads, exp_avg, exp_avg_sq, n_params, opt_state,
                   (float*)((char*)opt_state + 64), 256, (cudaStream_t)stream);
}

int fbsnn_peer_buffer_floats(int64_t n_params, int64_t* flag_offset_out, int64_t* total_out) {
  if (n_params < 1 || n_params % 4 || !flag_offset_out || !total_out) return fail(FBSNN_E_BADARG, "peer buffer: n_params must be a positive multiple of 4");
  const int64_t fo = (n_params + 4 + 63) / 64 * 64;
  *flag_offset_out = fo;
  *total_out = fo + 2 * kPeerMaxWorld;
  return 0;
}

int fbsnn_peer_wait(const float* local_buf, int64_t n_params, int world, const void* opt_state, void* stream) {
  if (!local_buf || !opt_state || world < 1 || world > kPeerMaxWorld) return fail(FBSNN_E_BADARG, "peer wait: bad argument");
  int64_t fo, tot;
  int rc = fbsnn_peer_buffer_floats(n_params, &fo, &tot);
  if (rc) return rc;
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(local_buf, fo, world, (const OptState*)opt_state);
  LAUNCH_CHECK("peer_wait");
  return 0;
}

int fbsnn_peer_allreduce_adam(const FbsnnAdam* host_hp, float* params, const void* peer_ptrs_dev, int world, int rank,
                              float* grad_sum, float* exp_avg, float* exp_avg_sq, int64_t n_params, void* opt_state,
                              void* stream) {
  if (!peer_ptrs_dev || !grad_sum || !opt_state || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world)
    return fail(FBSNN_E_BADARG, "peer allreduce: bad argument");
  int64_t fo, tot;
  int rc = fbsnn_peer_buffer_floats(n_params, &fo, &tot);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)((char*)opt_state + 64);
  PeerArgs a{};
  a.peers = (float* const*)peer_ptrs_dev, a.world = world, a.rank = rank;
  a.n4 = n_params / 4 + 1;       // + the float4 that carries the loss
  a.n_grad = n_params, a.flag_off = fo;
  a.st = (const OptState*)opt_state;
  a.counter = (unsigned*)((char*)opt_state + 1536);
  peer_reduce_kernel<<<256, 256, 0, st>>>(a, grad_sum, part);
  LAUNCH_CHECK("peer_reduce");
  return adam_impl(host_hp, params, grad_sum, exp_avg, exp_avg_sq, n_params, opt_state, part, 256, st, true);
}

int fbsnn_train_step(const FbsnnSpec* spec, const FbsnnAdam* host_hp, float* params, float* grads,
                     float* exp_avg, float* exp_avg_sq, void* opt_state, const float* t, const float* W,
                     const float* Xi, int64_t xi_rows, int64_t n_paths, float T, int64_t path_offset,
                     uint64_t seed, uint64_t iteration, const float* chol, void* workspace,
                     size_t workspace_bytes, float* X_out, float* Y_out, float* loss_out, void* stream) {
  if (!opt_state) return fail(FBSNN_E_BADARG, "opt_state is null");
  int rc = loss_grad_impl(spec, params, grads, t, W, Xi, xi_rows, n_paths, T, path_offset, seed, iteration,
                          &((const OptState*)opt_state)->rng_iter, chol, workspace, workspace_bytes, X_out, Y_out, nullptr,
                          loss_out, true, (cudaStream_t)stream);
  if (rc) return rc;
  return adam_impl(host_hp, params, grads, exp_avg, exp_avg_sq, spec->n_params, opt_state,
                   (float*)((char*)opt_state + 64), 256, (cudaStream_t)stream);
}

}  // extern "C"
