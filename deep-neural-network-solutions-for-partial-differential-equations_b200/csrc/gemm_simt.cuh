// fp32 SIMT GEMM with fused sweep epilogues (the FBSNN_PREC_FP32 arithmetic variant).
//
//   C[M x N] = sum over segments s of  A_s[M x K_s] * B_s[K_s x N]        (up to 4 K-concatenated segments)
//
// Operand layouts (template flags):
//   A_KC: A[m*lda + k] (k contiguous: activation rows)      !A_KC: A[k*lda + m] (weight-gradient operand P^T)
//   B_KC: B[n*ldb + k] (PyTorch (out,in) weight used as W^T) !B_KC: B[k*ldb + n] (weight used as W, or rows x in)
// Split-K (weight gradients): blockIdx.z selects rows [z*kchunk, (z+1)*kchunk) of every segment and the
// epilogue stores a partial tile; a deterministic second pass sums the partials (no atomics anywhere).
// All output widths N and leading dimensions of epilogue arrays are multiples of 4 (float4 epilogue I/O).
#pragma once
#include <type_traits>
#include <cstring>

#include "common.cuh"

namespace fbsnn {

struct GemmSeg {
  const float* A;
  const float* B;
  int lda, ldb, K;
};
constexpr int kMaxSeg = 8;
struct GemmArgs {
  GemmSeg seg[kMaxSeg];
  // tcgen05 3xTF32 only: 0 = one MMA a*b; 1 = B is an exact-TF32 "hi" matrix, A is split in-kernel: a_lo*b + a*b;
  // 2 = B is the matching "lo" matrix: a*b only; 3 = both operands split in-kernel: a_lo*b + a*b_lo + a*b
  unsigned char mode[kMaxSeg];
  int nseg;
  int M, N;
  int Nb;      // valid columns of the B operands (<= N; columns in [Nb, N) read as zero)
  int kchunk;  // 0: no split-K
};

// ----------------------------------------------------------------------------------------------------
// epilogues: operator()(row, col, acc4) with col % 4 == 0 and col + 3 < N
// ----------------------------------------------------------------------------------------------------
// Every epilogue is split in two so that the tcgen05 kernel can issue the loads of a whole 32-column chunk
// before consuming them:  Frag f = prefetch(r, c)  reads the row-array inputs,  finish(r, c, acc, f)  computes
// and stores.  operator() = finish(prefetch) for the SIMT kernel.
// RES = the NAIS-Net residual-stream variant (one more row array in flight per element group); the FC variant
// carries a smaller fragment so that the tcgen05 epilogue's 4-deep prefetch fits its register budget.
template <bool RES>
struct FragT {
  float4 x, y, z;
};
template <>
struct FragT<false> {
  float4 x, y;
};
// Per-column constants (biases, output weights) are fetched once per column group with col_prefetch(c) -- the tcgen05
// epilogue reuses them for the 8 rows a thread handles in a chunk instead of re-loading them per row.
struct ColNone {};
// l2_prefetch(r, c): pull the 128-byte lines prefetch(r, c) will read into L2 (no registers held while in flight); the
// tcgen05 epilogue calls it one 32-column chunk ahead -- its stalls are the HBM latency of exactly these loads
__device__ __forceinline__ void l2_line(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#define FBSNN_EPI_CALL                                                                                   \
  __device__ __forceinline__ void operator()(int r, int c, float4 v) const {                            \
    finish(r, c, v, prefetch(r, c), col_prefetch(c));                                                    \
  }
#define FBSNN_EPI_NO_COLS                                                                                \
  typedef ColNone ColFrag;                                                                               \
  __device__ __forceinline__ ColFrag col_prefetch(int) const { return ColFrag{}; }

// F sweep: z = acc + bias;  g = act(z), a = act'(z);  h = g (+ h_prev);  last layer also seeds the adjoint
template <bool RES>
struct EpiFwdT {
  struct Frag {
    float4 y;   // h_{l-1} (RES only)
  };
  static constexpr bool kColsum = false;
  static constexpr bool kManyEpilogueWarps = true;   // instruction-bound epilogue (sine/cosine): 16-warp kernel form
  const float* bias1;
  const float* bias2;  // nullable (NAIS: layer{l}_input.bias)
  const float* res;    // nullable: h_{l-1} (NAIS residual stream)
  float* g;
  float* a;
  float* h;            // nullable (only with res)
  const float* wout;   // nullable: last hidden layer -> delta = wout * a, s = wout * c
  float* delta;
  float* s;            // nullable (forward-only mode)
  int ld, act;
  // number of (rows x N) fp32 arrays this epilogue reads + writes (algorithmic HBM traffic, bench.py roofline)
  int io_arrays() const { return 2 + (h ? 2 : 0) + (wout ? (s ? 2 : 1) : 0); }
  __device__ __forceinline__ void l2_prefetch(int r, int c) const {
    if (RES && h) l2_line(res + (size_t)r * ld + c);
  }
  __device__ __forceinline__ Frag prefetch(int r, int c) const {
    Frag f;
    if (RES && h) f.y = ld4(res + (size_t)r * ld + c);
    else f.y = make_float4(0.f, 0.f, 0.f, 0.f);
    return f;
  }
  struct ColFrag {
    float4 b, w;   // bias (sum of both), output weights (last hidden layer)
  };
  __device__ __forceinline__ ColFrag col_prefetch(int c) const {
    ColFrag cf;
    cf.b = ld4(bias1 + c);
    if (bias2) {
      const float4 b2 = ld4(bias2 + c);
      cf.b.x += b2.x, cf.b.y += b2.y, cf.b.z += b2.z, cf.b.w += b2.w;
    }
    cf.w = wout ? ld4(wout + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    return cf;
  }
  __device__ __forceinline__ void finish(int r, int c, float4 v, const Frag& f, const ColFrag& cf) const {
    const float4 b = cf.b;
    const float z[4] = {v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w};
    float gv[4], av[4];
    act_ga4(act, z, gv, av);
    const size_t o = (size_t)r * ld + c;
    st4(g + o, make_float4(gv[0], gv[1], gv[2], gv[3]));
    st4(a + o, make_float4(av[0], av[1], av[2], av[3]));
    if (RES && h) st4(h + o, make_float4(gv[0] + f.y.x, gv[1] + f.y.y, gv[2] + f.y.z, gv[3] + f.y.w));
    if (wout) {
      const float4 w = cf.w;
      st4(delta + o, make_float4(w.x * av[0], w.y * av[1], w.z * av[2], w.w * av[3]));
      if (s)
        st4(s + o, make_float4(w.x * act_c(act, gv[0], av[0]), w.y * act_c(act, gv[1], av[1]),
                               w.z * act_c(act, gv[2], av[2]), w.w * act_c(act, gv[3], av[3])));
    }
  }
  FBSNN_EPI_CALL
};

// A sweep (writes layer l-1): ht = acc (+ ht_l | + wout);  delta = ht * a;  s = ht * c
template <bool RES>
struct EpiAdjT {
  typedef FragT<RES> Frag;
  static constexpr bool kColsum = false;
  const float* a;
  const float* g;
  const float* res;       // nullable: ht_l (NAIS)
  const float* res_head;  // nullable: ht_L = wout broadcast over rows (NAIS, l = L)
  float* ht_out;          // nullable
  float* delta;
  float* s;               // nullable (forward-only mode)
  int ld, act;
  int io_arrays() const { return 2 + (s ? 2 : 0) + (res ? 1 : 0) + (ht_out ? 1 : 0); }
  __device__ __forceinline__ void l2_prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    l2_line(a + o);
    if (s) l2_line(g + o);
    if (RES && res) l2_line(res + o);
  }
  __device__ __forceinline__ Frag prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    Frag f;
    f.x = ld4(a + o);
    f.y = s ? ld4(g + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (RES) f.z = res ? ld4(res + o) : (res_head ? ld4(res_head + c) : make_float4(0.f, 0.f, 0.f, 0.f));
    return f;
  }
  FBSNN_EPI_NO_COLS
  __device__ __forceinline__ void finish(int r, int c, float4 v, const Frag& f, const ColFrag& = ColFrag{}) const {
    const size_t o = (size_t)r * ld + c;
    float ht[4] = {v.x, v.y, v.z, v.w};
    if constexpr (RES) ht[0] += f.z.x, ht[1] += f.z.y, ht[2] += f.z.z, ht[3] += f.z.w;
    const float4 av = f.x, gv = f.y;
    st4(delta + o, make_float4(ht[0] * av.x, ht[1] * av.y, ht[2] * av.z, ht[3] * av.w));
    if (s)
      st4(s + o, make_float4(ht[0] * act_c(act, gv.x, av.x), ht[1] * act_c(act, gv.y, av.y),
                             ht[2] * act_c(act, gv.z, av.z), ht[3] * act_c(act, gv.w, av.w)));
    if (RES && ht_out) st4(ht_out + o, make_float4(ht[0], ht[1], ht[2], ht[3]));
  }
  FBSNN_EPI_CALL
};

// T sweep (layer l): dbar = acc;  hd = dbar * a (+ hd_{l-1});  zz = dbar * s;  last layer: zbar = ybar*wout*a + zz
template <bool RES>
struct EpiTanT {
  typedef FragT<RES> Frag;
  static constexpr bool kColsum = true;
  const float* a;
  float* s_zz;        // in: s, out: zz (or zbar for the last layer)
  const float* res;   // nullable: hd_{l-1}
  float* hd;
  const float* ybar;  // nullable: last layer
  const float* wout;
  float* colpart;     // nullable: [grid][1024] per-CTA column sums of zbar (last layer), tcgen05 kernel only
  int ld;
  int io_arrays() const { return 4 + (res ? 1 : 0); }
  __device__ __forceinline__ void l2_prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    l2_line(a + o);
    l2_line(s_zz + o);
    if (RES && res) l2_line(res + o);
  }
  __device__ __forceinline__ Frag prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    Frag f;
    f.x = ld4(a + o);
    f.y = ld4(s_zz + o);
    if constexpr (RES) f.z = res ? ld4(res + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    return f;
  }
  struct ColFrag {
    float4 w;
  };
  __device__ __forceinline__ ColFrag col_prefetch(int c) const {
    ColFrag cf;
    cf.w = wout ? ld4(wout + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    return cf;
  }
  __device__ __forceinline__ float4 finish(int r, int c, float4 v, const Frag& f, const ColFrag& cf) const {
    const size_t o = (size_t)r * ld + c;
    const float4 av = f.x, sv = f.y;
    float hdv[4] = {v.x * av.x, v.y * av.y, v.z * av.z, v.w * av.w};
    if constexpr (RES) hdv[0] += f.z.x, hdv[1] += f.z.y, hdv[2] += f.z.z, hdv[3] += f.z.w;
    float zz[4] = {v.x * sv.x, v.y * sv.y, v.z * sv.z, v.w * sv.w};
    if (wout) {
      const float yb = ybar[r];
      const float4 w = cf.w;
      zz[0] += yb * w.x * av.x, zz[1] += yb * w.y * av.y, zz[2] += yb * w.z * av.z, zz[3] += yb * w.w * av.w;
    }
    st4(hd + o, make_float4(hdv[0], hdv[1], hdv[2], hdv[3]));
    st4(s_zz + o, make_float4(zz[0], zz[1], zz[2], zz[3]));
    return make_float4(zz[0], zz[1], zz[2], zz[3]);
  }
  FBSNN_EPI_CALL
};

// B sweep (writes layer l-1): hb = acc (+ hb_l | + ybar*wout);  zbar = hb * a + zz
template <bool RES>
struct EpiBwdT {
  typedef FragT<RES> Frag;
  static constexpr bool kColsum = true;
  const float* a;
  float* zz_zbar;
  const float* res;   // nullable: hb_l (NAIS, l < L)
  const float* ybar;  // nullable: NAIS l = L, residual is ybar[r] * wout[c]
  const float* wout;
  float* hb_out;      // nullable
  float* colpart;     // nullable: [grid][1024] per-CTA column sums of zbar, tcgen05 kernel only
  int ld;
  int io_arrays() const { return 3 + (res ? 1 : 0) + (hb_out ? 1 : 0); }
  __device__ __forceinline__ void l2_prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    l2_line(a + o);
    l2_line(zz_zbar + o);
    if (RES && res) l2_line(res + o);
  }
  __device__ __forceinline__ Frag prefetch(int r, int c) const {
    const size_t o = (size_t)r * ld + c;
    Frag f;
    f.x = ld4(a + o);
    f.y = ld4(zz_zbar + o);
    if constexpr (RES) {
      if (res) {
        f.z = ld4(res + o);
      } else if (ybar) {
        const float yb = ybar[r];
        const float4 w = ld4(wout + c);
        f.z = make_float4(yb * w.x, yb * w.y, yb * w.z, yb * w.w);
      } else {
        f.z = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return f;
  }
  FBSNN_EPI_NO_COLS
  __device__ __forceinline__ float4 finish(int r, int c, float4 v, const Frag& f, const ColFrag& = ColFrag{}) const {
    const size_t o = (size_t)r * ld + c;
    float hb[4] = {v.x, v.y, v.z, v.w};
    if constexpr (RES) hb[0] += f.z.x, hb[1] += f.z.y, hb[2] += f.z.z, hb[3] += f.z.w;
    const float4 av = f.x, zz = f.y;
    const float4 zb = make_float4(hb[0] * av.x + zz.x, hb[1] * av.y + zz.y, hb[2] * av.z + zz.z, hb[3] * av.w + zz.w);
    st4(zz_zbar + o, zb);
    if (RES && hb_out) st4(hb_out + o, make_float4(hb[0], hb[1], hb[2], hb[3]));
    return zb;
  }
  FBSNN_EPI_CALL
};

// same members in both variants: the FC form is obtained by reinterpreting the filled-in NAIS form
template <template <bool> class E>
inline E<false> narrow(const E<true>& e) {
  static_assert(sizeof(E<false>) == sizeof(E<true>), "epilogue variants must share their layout");
  E<false> r;
  memcpy(&r, &e, sizeof(r));
  return r;
}

struct FragNone {};
struct EpiStore {
  typedef FragNone Frag;
  static constexpr bool kColsum = false;
  float* out;
  int ld;
  int io_arrays() const { return 1; }
  __device__ __forceinline__ Frag prefetch(int, int) const { return Frag{}; }
  __device__ __forceinline__ void l2_prefetch(int, int) const {}
  FBSNN_EPI_NO_COLS
  __device__ __forceinline__ void finish(int r, int c, float4 v, const Frag&, const ColFrag& = ColFrag{}) const {
    st4(out + (size_t)r * ld + c, v);
  }
  FBSNN_EPI_CALL
};

// split-K partial tile: out[z][M][N]; `z` is blockIdx.z in the SIMT kernel, the split index in the tcgen05 kernel
struct EpiPartial {
  typedef FragNone Frag;
  static constexpr bool kColsum = false;
  float* out;
  int M, N;
  int io_arrays() const { return 1; }
  __device__ __forceinline__ Frag prefetch(int, int) const { return Frag{}; }
  __device__ __forceinline__ void l2_prefetch(int, int) const {}
  FBSNN_EPI_NO_COLS
  __device__ __forceinline__ void finish(int r, int c, float4 v, const Frag&, const ColFrag& = ColFrag{}) const {
    st4(out + ((size_t)blockIdx.z * M + r) * N + c, v);
  }
  __device__ __forceinline__ void finish_split(int split, int r, int c, float4 v) const {
    st4(out + ((size_t)split * M + r) * N + c, v);
  }
  FBSNN_EPI_CALL
};

// ----------------------------------------------------------------------------------------------------
// kernel
// ----------------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), (BM >= 128 ? 2 : 3))
gemm_simt_kernel(const GemmArgs g, const Epi epi) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int A_LD = BM * BK / NT, B_LD = BN * BK / NT;
  constexpr int GM = 4 * BM / TM, GN = 4 * BN / TN;  // row / column stride between a thread's 4-wide groups
  static_assert(TM % 4 == 0 && TN % 4 == 0 && (BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile shape");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float ra[A_LD], rb[B_LD];

  auto kbeg = [&](int s) { return g.kchunk ? (int)min((long long)g.seg[s].K, (long long)blockIdx.z * g.kchunk) : 0; };
  auto kend = [&](int s) {
    return g.kchunk ? (int)min((long long)g.seg[s].K, ((long long)blockIdx.z + 1) * g.kchunk) : g.seg[s].K;
  };
  auto gload = [&](int s, int k0, int ke) {
    const float* __restrict__ A = g.seg[s].A;
    const float* __restrict__ B = g.seg[s].B;
    const int lda = g.seg[s].lda, ldb = g.seg[s].ldb;
#pragma unroll
    for (int p = 0; p < A_LD; ++p) {
      const int e = tid + p * NT;
      int m, k;
      if (A_KC) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < g.M && gk < ke) v = A_KC ? __ldg(A + (size_t)gm * lda + gk) : __ldg(A + (size_t)gk * lda + gm);
      ra[p] = v;
    }
#pragma unroll
    for (int p = 0; p < B_LD; ++p) {
      const int e = tid + p * NT;
      int n, k;
      if (B_KC) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.Nb && gk < ke) v = B_KC ? __ldg(B + (size_t)gn * ldb + gk) : __ldg(B + (size_t)gk * ldb + gn);
      rb[p] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int p = 0; p < A_LD; ++p) {
      const int e = tid + p * NT;
      int m, k;
      if (A_KC) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      As[buf][k][m] = ra[p];
    }
#pragma unroll
    for (int p = 0; p < B_LD; ++p) {
      const int e = tid + p * NT;
      int n, k;
      if (B_KC) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      Bs[buf][k][n] = rb[p];
    }
  };

  // first non-empty segment
  int s = 0;
  while (s < g.nseg && kbeg(s) >= kend(s)) ++s;
  if (s < g.nseg) {
    int k0 = kbeg(s), ke = kend(s);
    gload(s, k0, ke);
    sstore(0);
    __syncthreads();
    int buf = 0;
    while (true) {
      int ns = s, nk0 = k0 + BK, nke = ke;
      if (nk0 >= ke) {
        ++ns;
        while (ns < g.nseg && kbeg(ns) >= kend(ns)) ++ns;
        if (ns < g.nseg) { nk0 = kbeg(ns); nke = kend(ns); }
      }
      const bool more = ns < g.nseg;
      if (more) gload(ns, nk0, nke);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i4 = 0; i4 < TM / 4; ++i4) {
          const float4 t = ld4(&As[buf][kk][i4 * GM + ty * 4]);
          av[i4 * 4 + 0] = t.x, av[i4 * 4 + 1] = t.y, av[i4 * 4 + 2] = t.z, av[i4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4) {
          const float4 t = ld4(&Bs[buf][kk][j4 * GN + tx * 4]);
          bv[j4 * 4 + 0] = t.x, bv[j4 * 4 + 1] = t.y, bv[j4 * 4 + 2] = t.z, bv[j4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (!more) break;
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
      s = ns, k0 = nk0, ke = nke;
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = m0 + (i / 4) * GM + ty * 4 + (i % 4);
    if (r >= g.M) continue;
#pragma unroll
    for (int j4 = 0; j4 < TN / 4; ++j4) {
      const int c = n0 + j4 * GN + tx * 4;
      if (c < g.N) epi(r, c, make_float4(acc[i][j4 * 4 + 0], acc[i][j4 * 4 + 1], acc[i][j4 * 4 + 2], acc[i][j4 * 4 + 3]));
    }
  }
}

// Host launcher: picks the 128x128 (8x8 micro-tile) shape when the grid fills the chip at least twice,
// else 64x64 (4x4) so that small-M steps still spread over all 148 SMs.
template <bool A_KC, bool B_KC, class Epi>
inline cudaError_t launch_gemm(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  const long long big_ctas = (long long)((g.M + 127) / 128) * ((g.N + 127) / 128) * nsplit;
  if (big_ctas >= 2LL * num_sms) {
    dim3 grid((g.M + 127) / 128, (g.N + 127) / 128, nsplit);
    gemm_simt_kernel<128, 128, 16, 8, 8, A_KC, B_KC, Epi><<<grid, 256, 0, st>>>(g, epi);
  } else {
    dim3 grid((g.M + 63) / 64, (g.N + 63) / 64, nsplit);
    if constexpr (std::is_same<Epi, EpiStore>::value) {
      // the NAIS-Net projection products (W^T W, W S: 256 x 256 x 256) are 16 CTAs of 64 x 64: a quarter of the chip for
      // 21 us, six times per small-batch step -- 32 x 32 tiles (64 threads, same k order, so the same bits) use 64 CTAs
      if ((long long)grid.x * grid.y * grid.z * 4 <= num_sms) {
        dim3 grid32((g.M + 31) / 32, (g.N + 31) / 32, nsplit);
        gemm_simt_kernel<32, 32, 16, 4, 4, A_KC, B_KC, Epi><<<grid32, 64, 0, st>>>(g, epi);
        return cudaGetLastError();
      }
    }
    gemm_simt_kernel<64, 64, 16, 4, 4, A_KC, B_KC, Epi><<<grid, 256, 0, st>>>(g, epi);
  }
  return cudaGetLastError();
}

}  // namespace fbsnn
