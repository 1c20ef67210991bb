// Layer-chained sweep kernel (FC networks, tcgen05 variants): ONE persistent launch runs a whole sweep
// (F, A, T or B of fbsnn_api.cu) over all of its layers.  A 128-row tile walks through the layers on chip: the
// epilogue of layer l turns the accumulator (TMEM) into the NEXT layer's A operand directly in shared memory, 32
// columns (= one k-block) at a time, so the carried activation never travels through HBM and the next layer's MMAs
// start as soon as the first k-block exists.  Every row array a LATER sweep needs is read / written by TMA from / to
// the same swizzled 16 KB chunk buffers (no register-level global traffic, no transposes):
//
//   link i = "epilogue stage" between MMA i and MMA i+1.  For every 32-column chunk of the link:
//     input producer   TMA-loads up to two row-array chunks (e.g. a_l, s_l) into buf0 / buf2 of an A-ring stage
//     epilogue group   (4 warps = 128 rows, lane = row) reads the accumulator chunk with tcgen05.ld, combines it with
//                      the loaded chunks, writes the results back IN PLACE: buf0 = next A operand (+ buf1 = its
//                      3xTF32 low part), buf2 = second output; canonical K-major SWIZZLE_128B layout, which is at
//                      the same time what tcgen05.mma reads and what a TMA store expects
//     MMA warp         issues the k-block of MMA i+1 (A = buf0 / buf1, B = weight ring fed by the weight producer)
//     store warp       TMA-stores buf0 / buf2 to the row arrays of HBM
//   The two accumulators (2 x 256 TMEM columns) alternate between consecutive MMAs, so the epilogue of link i
//   overlaps MMA i+1 chunk by chunk.
//
// Warps: 0 weight producer | 1 MMA issuer (owns TMEM) | 2 input producer | 3 store | 4.. epilogue groups.
// Shared memory (3xTF32): 2 weight stages x (W_hi 32 KB + W_lo 32 KB) + 2 A stages x 3 x 16 KB = 224 KB;
// (TF32): 4 weight stages x 32 KB + 3 A stages x 2 x 16 KB = 224 KB.
//
// chain2_kernel is the CTA-PAIR form (cluster of 2, tcgen05.mma.cta_group::2, 256-row tiles) for 3xTF32: each CTA
// stages its own 128 rows of every chunk and only HALF of every weight k-block (the pair's MMA reads both halves), which
// (i) halves the weight traffic through L2 -- measured to be what bounds the sweeps: every 128-row tile re-fetches the
// 512 KB hi/lo weight twins of every layer -- and (ii) frees 64 KB of shared memory for a THIRD A stage, so that input
// load, epilogue, MMA and store of consecutive chunks overlap instead of queueing on two buffers (ablation timings in
// profiles/r02_chain_ablation.txt).  Cross-CTA signalling: the peer's TMA completes on the LEADER's weight barrier, one
// elected thread per chunk arrives on the leader's a_ready, the leader's tcgen05.commit multicasts to both CTAs.
#pragma once
#include <cstdio>

#include "gemm_tc2.cuh"

namespace fbsnn {
namespace chain {

using tc::smem_u32;

constexpr int kMaxLinks = FBSNN_MAX_HIDDEN + 1;
enum { SWEEP_F = 0, SWEEP_A = 1, SWEEP_T = 2, SWEEP_B = 3 };
enum { LINK_FIRST = 0, LINK_MID = 1, LINK_LAST = 2 };
constexpr int CHUNK_BYTES = 128 * 32 * 4;   // 128 rows x 32 fp32 columns

struct LinkD {
  int width;    // columns of this link's row arrays (multiple of 32, <= 256)
  int n_next;   // N of the MMA this link feeds (0: none)
  unsigned char kind, in0, in2, out0, out2, feeds, b_mn, colsum;   // colsum: bit 0 = buf0, bit 1 = buf2
};
struct Maps {
  CUtensorMap in0[kMaxLinks], in2[kMaxLinks], out0[kMaxLinks], out2[kMaxLinks], whi[kMaxLinks], wlo[kMaxLinks];
};
struct Args {
  int nlinks, rows, ntiles, act, with_s;
  int pf_dist;  // chaint_kernel: row-array chunks pulled into L2 ahead of their TMA load (FBSNN_CHAIN_PF)
  int hints;    // chaint_kernel / chain_kernel L2 policies (FBSNN_CHAIN_HINT): bit 0 = row-array stores evict_first, bit 1 = row-array loads
                // evict_first, bit 2 = weight k-blocks evict_last (re-read by every tile while the row arrays stream through)
  int ablate;   // measurement only (FBSNN_CHAIN_ABLATE): 1 no TMA stores, 2 no input loads, 4 no epilogue math, 8 no MMAs, 16 no weight loads
  LinkD link[kMaxLinks];
  const float* bias[kMaxLinks];   // F: bias of the layer whose pre-activation link i receives
  const float* wout;
  const float* bout;
  const float* ybar;              // T
  float* Y;                       // F: output head
  float* colacc;                  // T / B: [grid][kMaxLinks][2][4][256] per-CTA column sums (row quarter q kept apart)
};

template <bool X3>
struct Cfg {
  static constexpr int W_STAGE = X3 ? 65536 : 32768;
  static constexpr int W_STAGES = X3 ? 2 : 4;
  static constexpr int A_BUFS = X3 ? 3 : 2;                  // [buf0][buf1 = lo (3xTF32 only)][buf2]
  static constexpr int A_STAGE = A_BUFS * CHUNK_BYTES;
  static constexpr int A_STAGES = X3 ? 2 : 3;
  // ONE epilogue team works on ONE chunk at a time, its 16 warps splitting the chunk's 32 columns (lane quarter q = warp
  // % 4 owns 32 rows, column slice h = warp / 4 owns 8 columns), and the chunks alternate between the A stages: the
  // epilogue of chunk c+1 then overlaps the MMAs and the TMA store of chunk c.  (Two independent 4-warp groups, one per
  // stage, measured 2x slower: a group's next chunk reuses the stage of its previous one and so waits out that chunk's
  // MMA + store + re-load round trip of ~3 us every time -- profiles/r02_chain_ablation.txt.)
  static constexpr int EPI_WARPS = 16;
  static constexpr int CW = 32 / (EPI_WARPS / 4);            // columns per epilogue warp
  static constexpr int EPI_WARP0 = 4;
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
  static constexpr int B2_OFF = (A_BUFS - 1) * CHUNK_BYTES;  // byte offset of buf2 inside a stage
  static constexpr int EXTRA_BYTES = 2048;                   // barriers, TMEM slot, head partials
  static constexpr int SMEM_BYTES = W_STAGES * W_STAGE + A_STAGES * A_STAGE + EXTRA_BYTES + 1024 /*align*/;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* tm, const void* src, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// bounded wait that says WHO starved before trapping (role: 0 weights, 1 MMA, 2 inputs, 3 store, 4 epilogue)
__device__ __noinline__ void chain_timeout(int role, int what, int link, int chunk) {
  printf("fbsnn chain kernel: block %d thread %d role %d wait %d link %d chunk %d timed out\n", (int)blockIdx.x,
         (int)threadIdx.x, role, what, link, chunk);
  asm volatile("trap;");
}
__device__ __forceinline__ void cwait(uint64_t* b, uint32_t parity, int role, int what, int link, int chunk) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(b);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > tc::kSpinLimit) chain_timeout(role, what, link, chunk);
  }
}
template <int CW>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[CW]) {
  if constexpr (CW == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
  } else if constexpr (CW == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
  } else {
    static_assert(CW == 32, "column slice of 8, 16 or 32");
    FBSNN_TMEM_LD32(taddr, v);
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float lo_part(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ----------------------------------------------------------------------------------------------------------------
// shared pieces of the single-CTA and the CTA-pair kernel
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
// position in the chunk sequence (tile, link, chunk) of one CTA; `step` = tiles between consecutive tiles of this CTA
struct ChunkIt {
  int tile, link, j;
  __device__ __forceinline__ bool valid(const Args& a) const { return tile < a.ntiles; }
  __device__ __forceinline__ void next(const Args& a, int step) {
    if (++j >= (a.link[link].width >> 5)) {
      j = 0;
      if (++link >= a.nlinks) link = 0, tile += step;
    }
  }
};

// One 32-column chunk of link `i` for the thread that owns row `rt` of the tile: combines the accumulator fragment v[]
// with the loaded chunks (buf0 / buf2) and writes the results back in place (buf0, its low part buf1, buf2).
// CW = columns of the chunk this thread handles, starting at column u0 * 4 of the chunk (v[] = its accumulator fragment).
// (i0, i2) = where the loaded input chunks are: the output buffers themselves (in-place, single-CTA kernel) or a separate
// input ring (CTA-pair kernel; pass-through links then have to copy their operand into buf0).
// TA = the operand of the next MMA goes to TENSOR MEMORY (chaint_kernel): its values come back in hi_out[] instead of
// buf0 / buf1, and shared memory is written only for what a TMA store or a column sum reads.
template <int SWEEP, bool X3, int CW, bool INPLACE, bool TA = false>
__device__ __forceinline__ void chunk_math(const Args& a, const LinkD& L, int i, int c0, int rt, int u0, const float* i0,
                                           const float* i2, float* b0, float* b1, float* b2, const uint32_t* v,
                                           float yb, float& yacc, uint32_t* hi_out = nullptr) {
  const int rsw = rt & 7;
  const bool want_lo = X3 && L.feeds && !TA;
#pragma unroll
  for (int uu = 0; uu < CW / 4; ++uu) {
    const int u = u0 + uu;
    const int off = rt * 32 + ((u ^ rsw) << 2);
    float x0[4] = {0.f, 0.f, 0.f, 0.f}, x2[4] = {0.f, 0.f, 0.f, 0.f};
    if (L.in0) { const float4 t = ld4(i0 + off); x0[0] = t.x, x0[1] = t.y, x0[2] = t.z, x0[3] = t.w; }
    if (L.in2) { const float4 t = ld4(i2 + off); x2[0] = t.x, x2[1] = t.y, x2[2] = t.z, x2[3] = t.w; }
    const float ac[4] = {__uint_as_float(v[4 * uu]), __uint_as_float(v[4 * uu + 1]), __uint_as_float(v[4 * uu + 2]),
                         __uint_as_float(v[4 * uu + 3])};
    float o0[4], o2[4];
    bool w0 = true, w2 = false;
    if constexpr (SWEEP == SWEEP_F) {
      if (L.kind == LINK_FIRST) {
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = x0[t];
        w0 = !INPLACE;
      } else {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias[i] + c0 + 4 * u));
        const float z[4] = {ac[0] + b.x, ac[1] + b.y, ac[2] + b.z, ac[3] + b.w};
        act_ga4(a.act, z, o0, o2);
        w2 = true;
        if (L.kind == LINK_LAST) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(a.wout + c0 + 4 * u));
          yacc = fmaf(o0[0], w.x, fmaf(o0[1], w.y, fmaf(o0[2], w.z, fmaf(o0[3], w.w, yacc))));
        }
      }
    } else if constexpr (SWEEP == SWEEP_A) {
      if (L.kind == LINK_FIRST) {          // delta_L = wout * a_L
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.wout + c0 + 4 * u));
        o0[0] = w.x * x0[0], o0[1] = w.y * x0[1], o0[2] = w.z * x0[2], o0[3] = w.w * x0[3];
      } else if (L.kind == LINK_MID) {     // ht = acc; delta = ht * a; s = ht * c(g, a)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          o0[t] = ac[t] * x0[t];
          o2[t] = ac[t] * act_c(a.act, x2[t], x0[t]);
        }
        w2 = a.with_s != 0;
      } else {                             // Du
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = ac[t];
      }
    } else if constexpr (SWEEP == SWEEP_T) {
      if (L.kind == LINK_FIRST) {
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = x0[t];
        w0 = !INPLACE;
      } else if (L.kind == LINK_MID) {     // dbar = acc; hd = dbar * a; zz = dbar * s
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = ac[t] * x0[t], o2[t] = ac[t] * x2[t];
        w2 = true;
      } else {   // last hidden layer: zbar = ybar wout a + dbar (wout c);  wg = dbar a + ybar g  (column sums only)
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wout + c0 + 4 * u));
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float av = x0[t], gv = x2[t];
          const float sv = w[t] * act_c(a.act, gv, av);
          const float zz = ac[t] * sv;
          o0[t] = zz + yb * w[t] * av;
          o2[t] = ac[t] * av + yb * gv;
        }
        w2 = true;
      }
    } else {
      if (L.kind == LINK_FIRST) {
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = x0[t];
        w0 = !INPLACE;
      } else {                             // hb = acc; zbar = hb * a + zz
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = ac[t] * x0[t] + x2[t];
      }
    }
    if constexpr (TA) {
#pragma unroll
      for (int t = 0; t < 4; ++t) hi_out[4 * uu + t] = __float_as_uint(o0[t]);
      w0 = L.out0 || (L.colsum & 1);
      w2 = w2 && (L.out2 || (L.colsum & 2));
    }
    if (w0) st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
    if (want_lo) st4(b1 + off, make_float4(lo_part(o0[0]), lo_part(o0[1]), lo_part(o0[2]), lo_part(o0[3])));
    if (w2) st4(b2 + off, make_float4(o2[0], o2[1], o2[2], o2[3]));
  }
}

// column sums over a group's 128 rows of the chunk in buf0 (/ buf2): warp q adds its 32 rows of column (c0 + lane); the
// partials of the four row quarters stay apart -- every address has ONE owner thread for the whole launch (chunk j of a
// link always goes to group j % G), so the accumulation is a plain, ordered read-modify-write in global memory
__device__ __forceinline__ void chunk_colsum(const Args& a, const LinkD& L, int i, int c0, int q, int lane,
                                             const float* b0, const float* b2, bool first_tile) {
  float s0 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int r = 0; r < 32; ++r) {
    const int rr = q * 32 + r;
    const int idx = rr * 32 + ((((lane >> 2) ^ (rr & 7))) << 2) + (lane & 3);
    s0 += b0[idx];
    if (L.colsum & 2) s2 += b2[idx];
  }
  float* dst = a.colacc + ((size_t)(blockIdx.x * kMaxLinks + i) * 2) * 1024 + q * 256 + c0 + lane;
  // the owner's store and its later REDs to the same address are applied in program order: deterministic, and the
  // L2 round trip of a read-modify-write stays off the chunk's critical path
  if (L.colsum & 1) { if (first_tile) dst[0] = s0; else atomicAdd(dst, s0); }
  if (L.colsum & 2) { if (first_tile) dst[1024] = s2; else atomicAdd(dst + 1024, s2); }
}

// ----------------------------------------------------------------------------------------------------------------
// single-CTA kernel
// ----------------------------------------------------------------------------------------------------------------
template <int SWEEP, bool X3>
__global__ void __launch_bounds__(Cfg<X3>::NUM_THREADS, 1)
chain_kernel(const __grid_constant__ Maps tm, const Args a) {
  using C = Cfg<X3>;
  constexpr int AS = C::A_STAGES, WS = C::W_STAGES;
  const int tile_end = a.ntiles;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wring = smem;
  uint8_t* aring = smem + WS * C::W_STAGE;
  uint8_t* extra = aring + AS * C::A_STAGE;
  uint64_t* bars = (uint64_t*)extra;
  uint64_t* w_full = bars;           // [WS]  weight k-block landed
  uint64_t* w_empty = bars + 4;      // [WS]  MMAs that read it completed
  uint64_t* in_full = bars + 8;      // [AS]  input chunks landed (or: stage handed to the epilogue)
  uint64_t* a_ready = bars + 12;     // [AS]  the epilogue team has written the chunk (one arrival per warp)
  uint64_t* a_free = bars + 16;      // [AS]  MMAs that read the chunk completed + TMA stores have read it
  uint64_t* acc_full = bars + 20;    // [2]
  uint64_t* acc_empty = bars + 22;   // [2]   all epilogue warps have drained the accumulator
  uint32_t* tmem_slot = (uint32_t*)(bars + 24);
  float* ypart = (float*)(extra + 256);   // [EPI_WARPS / 4 - 1][128] head partial sums (F sweep)
  constexpr int EW = C::EPI_WARPS, CW = C::CW;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a.nlinks; ++i) {
      const LinkD& L = a.link[i];
      if (L.in0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in0[i]) : "memory");
      if (L.in2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in2[i]) : "memory");
      if (L.out0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out0[i]) : "memory");
      if (L.out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out2[i]) : "memory");
      if (L.feeds) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.whi[i]) : "memory");
        if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.wlo[i]) : "memory");
      }
    }
    for (int i = 0; i < WS; ++i) tc::mbar_init(&w_full[i], 1), tc::mbar_init(&w_empty[i], 1);
    for (int i = 0; i < AS; ++i) tc::mbar_init(&in_full[i], 1), tc::mbar_init(&a_ready[i], EW), tc::mbar_init(&a_free[i], 2);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_empty[i], EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tc::pdl_trigger();
  tc::pdl_wait();

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      uint32_t ws = 0, wph = 0;
      const uint64_t wpol = l2_evict_last_policy();
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          if (!L.feeds) continue;
          const int nch = L.width >> 5, N = L.n_next;
          const uint32_t bytes = (uint32_t)N * 128u * (X3 ? 2u : 1u);
          for (int j = 0; j < nch; ++j) {
            cwait(&w_empty[ws], wph ^ 1, 0, 0, i, j);
            uint8_t* dst = wring + ws * C::W_STAGE;
            if (a.ablate & 16) {
              tc::mbar_arrive(&w_full[ws]);
            } else {
              tc::mbar_expect_tx(&w_full[ws], bytes);
              if (a.hints & 4) {   // weights: L2 evict_last
                if (L.b_mn) {
                  for (int c = 0; c < N / 32; ++c) {
                    tma_load_2d_hint(dst + c * 4096, &tm.whi[i], &w_full[ws], 32 * c, 32 * j, wpol);
                    if (X3) tma_load_2d_hint(dst + 32768 + c * 4096, &tm.wlo[i], &w_full[ws], 32 * c, 32 * j, wpol);
                  }
                } else {
                  tma_load_2d_hint(dst, &tm.whi[i], &w_full[ws], 32 * j, 0, wpol);
                  if (X3) tma_load_2d_hint(dst + 32768, &tm.wlo[i], &w_full[ws], 32 * j, 0, wpol);
                }
              } else if (L.b_mn) {   // W[k][n] (n contiguous): 32 x 32 boxes, 128B swizzle with 32B atoms
                for (int c = 0; c < N / 32; ++c) {
                  tc::tma_load_2d(dst + c * 4096, &tm.whi[i], &w_full[ws], 32 * c, 32 * j);
                  if (X3) tc::tma_load_2d(dst + 32768 + c * 4096, &tm.wlo[i], &w_full[ws], 32 * c, 32 * j);
                }
              } else {        // W[n][k] (k contiguous): one N x 32 box
                tc::tma_load_2d(dst, &tm.whi[i], &w_full[ws], 32 * j, 0);
                if (X3) tc::tma_load_2d(dst + 32768, &tm.wlo[i], &w_full[ws], 32 * j, 0);
              }
            }
            if (++ws == WS) ws = 0, wph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== input producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const uint64_t pol = l2_evict_first_policy();
      // the row arrays of the chunks a few positions ahead are pulled into L2 now, so that the TMA load issued when the
      // stage frees up pays an L2 hit instead of the HBM latency (that latency sits inside the stage's occupancy)
      ChunkIt pf{(int)blockIdx.x, 0, 0};
      auto prefetch_one = [&]() {
        if (!pf.valid(a) || (a.ablate & 2)) return;
        const LinkD& P = a.link[pf.link];
        if (P.in0) tma_prefetch_2d(&tm.in0[pf.link], 32 * pf.j, pf.tile * 128);
        if (P.in2) tma_prefetch_2d(&tm.in2[pf.link], 32 * pf.j, pf.tile * 128);
        pf.next(a, (int)gridDim.x);
      };
      for (int p = 0; p < AS + 2; ++p) prefetch_one();
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int m0 = tile * 128;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            cwait(&a_free[s], ph ^ 1, 2, 0, i, j);
            prefetch_one();
            uint8_t* st = aring + s * C::A_STAGE;
            uint64_t* bar = &in_full[s];
            if ((L.in0 || L.in2) && !(a.ablate & 2)) {
              tc::mbar_expect_tx(bar, (uint32_t)CHUNK_BYTES * (uint32_t)((L.in0 ? 1 : 0) + (L.in2 ? 1 : 0)));
              if (a.hints & 2) {   // row arrays stream through: L2 evict_first
                if (L.in0) tma_load_2d_hint(st, &tm.in0[i], bar, 32 * j, m0, pol);
                if (L.in2) tma_load_2d_hint(st + C::B2_OFF, &tm.in2[i], bar, 32 * j, m0, pol);
              } else {
                if (L.in0) tc::tma_load_2d(st, &tm.in0[i], bar, 32 * j, m0);
                if (L.in2) tc::tma_load_2d(st + C::B2_OFF, &tm.in2[i], bar, 32 * j, m0);
              }
            } else {
              tc::mbar_arrive(bar);
            }
            if (++s == AS) s = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0, ws = 0, wph = 0, mm = 0;
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          const uint32_t acc = mm & 1;
          uint32_t idesc = 0;
          if (L.feeds) {
            cwait(&acc_empty[acc], ((mm >> 1) & 1) ^ 1, 1, 0, i, 0);
            tc::tc_fence_after();
            idesc = tc::make_idesc(L.n_next, false, L.b_mn != 0);
          }
          const uint32_t tmem_d = tmem_base + acc * 256;
          for (int j = 0; j < nch; ++j) {
            cwait(&a_ready[s], ph, 1, 1, i, j);
            if (L.feeds) {
              cwait(&w_full[ws], wph, 1, 2, i, j);
              tc::tc_fence_after();
              if (!(a.ablate & 8)) {
                const uint32_t a0 = smem_u32(aring + s * C::A_STAGE);
                const uint32_t alo = a0 + CHUNK_BYTES;
                const uint32_t b0 = smem_u32(wring + ws * C::W_STAGE);
                const uint32_t blo = b0 + 32768;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = tc::make_desc(a0 + k * 32, 16, 1024, 2);
                  const uint64_t db = L.b_mn ? tc::make_desc(b0 + k * 1024, 4096, 512, 1) : tc::make_desc(b0 + k * 32, 16, 1024, 2);
                  if (X3) {
                    const uint64_t dal = tc::make_desc(alo + k * 32, 16, 1024, 2);
                    const uint64_t dbl = L.b_mn ? tc::make_desc(blo + k * 1024, 4096, 512, 1) : tc::make_desc(blo + k * 32, 16, 1024, 2);
                    tc::umma_tf32(tmem_d, dal, db, idesc, (j | k) ? 1u : 0u);
                    tc::umma_tf32(tmem_d, da, dbl, idesc, 1u);
                    tc::umma_tf32(tmem_d, da, db, idesc, 1u);
                  } else {
                    tc::umma_tf32(tmem_d, da, db, idesc, (j | k) ? 1u : 0u);
                  }
                }
              }
              tc::umma_commit(&w_empty[ws]);
              if (++ws == WS) ws = 0, wph ^= 1;
            }
            tc::umma_commit(&a_free[s]);   // (also for chunks nothing reads: commits complete in order)
            if (++s == AS) s = 0, ph ^= 1;
          }
          if (L.feeds) {
            tc::umma_commit(&acc_full[acc]);
            ++mm;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== store warp =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      const uint64_t pol = l2_evict_first_policy();
      for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int m0 = tile * 128;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            cwait(&a_ready[s], ph, 3, 0, i, j);
            const uint8_t* st = aring + s * C::A_STAGE;
            const bool do_store = (L.out0 || L.out2) && !(a.ablate & 1);
            if (do_store && (a.hints & 1)) {
              if (L.out0) tma_store_2d_hint(&tm.out0[i], st, 32 * j, m0, pol);
              if (L.out2) tma_store_2d_hint(&tm.out2[i], st + C::B2_OFF, 32 * j, m0, pol);
            } else if (do_store) {
              if (L.out0) tma_store_2d(&tm.out0[i], st, 32 * j, m0);
              if (L.out2) tma_store_2d(&tm.out2[i], st + C::B2_OFF, 32 * j, m0);
            }
            if (do_store) {
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            tc::mbar_arrive(&a_free[s]);
            if (++s == AS) s = 0, ph ^= 1;
          }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
    }
  } else {
    // ===================== epilogue team: one chunk at a time, 16 warps = 4 row quarters x 4 column slices ==========
    const int e = warp - C::EPI_WARP0;
    const int q = warp & 3;          // TMEM lane quarter (rows 32q .. 32q+31 of the tile)
    const int h = e >> 2;            // column slice: columns [h * CW, (h + 1) * CW) of every chunk
    const int rt = q * 32 + lane;    // row inside the tile
    uint32_t s = 0, ph = 0, mmr = 0;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x, first_tile = false) {
      const int row = tile * 128 + rt;
      const bool valid = row < a.rows;
      float yb = 0.f, yacc = 0.f;
      if (SWEEP == SWEEP_T) yb = valid ? __ldg(a.ybar + row) : 0.f;
      for (int i = 0; i < a.nlinks; ++i) {
        const LinkD& L = a.link[i];
        const int nch = L.width >> 5;
        const bool has_acc = L.kind != LINK_FIRST;
        const uint32_t acc = mmr & 1;
        if (has_acc) {
          cwait(&acc_full[acc], (mmr >> 1) & 1, 4, 0, i, 0);
          tc::tc_fence_after();
        }
        for (int j = 0; j < nch; ++j) {
          cwait(&in_full[s], ph, 4, 1, i, j);
          float* b0 = (float*)(aring + s * C::A_STAGE);
          float* b1 = b0 + CHUNK_BYTES / 4;
          float* b2 = (float*)(aring + s * C::A_STAGE + C::B2_OFF);
          uint32_t v[CW];
          if (has_acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (uint32_t)(32 * j + h * CW);
            tmem_ld_cols<CW>(taddr, v);
          }
          const int c0 = 32 * j;
          if (!(a.ablate & 4)) chunk_math<SWEEP, X3, CW, true>(a, L, i, c0, rt, h * (CW / 4), b0, b2, b0, b1, b2, v, yb, yacc);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA / TMA store
          if (L.colsum) {
            named_bar(1, 32 * EW);                                       // the whole chunk is written
            if (h == 0) chunk_colsum(a, L, i, c0, q, lane, b0, b2, first_tile);
          }
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&a_ready[s]);
          if (++s == AS) s = 0, ph ^= 1;
        }
        if (has_acc) {
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
          ++mmr;
        }
        if (SWEEP == SWEEP_F && L.kind == LINK_LAST) {   // output head: u = h_L . wout + bout
          if (h > 0) ypart[(h - 1) * 128 + rt] = yacc;
          named_bar(1, 32 * EW);
          if (h == 0 && valid) {
            float y = yacc;
#pragma unroll
            for (int hh = 1; hh < EW / 4; ++hh) y += ypart[(hh - 1) * 128 + rt];
            a.Y[row] = y + __ldg(a.bout);
          }
          named_bar(1, 32 * EW);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------------------
// CTA-pair kernel (3xTF32): cluster (2,1,1), tile = 256 rows (128 per CTA), tcgen05.mma.cta_group::2 issued by the
// leader (cluster rank 0).  Staging only half of every weight k-block per CTA frees the shared memory for what the
// single-CTA kernel lacks: an INPUT ring separate from the output ring, so that the row-array chunks of the next chunks
// are already on chip when the epilogue gets to them (in the single-CTA kernel a chunk's inputs land in the very buffers
// the previous use of the stage is still being read from by the MMA and the TMA store, which puts the whole load latency
// on the critical path: removing either the loads or the stores there takes 13.4 -> 8.7 ms off the A sweep).
// Per CTA: weights 2 x (W_hi half 16 KB + W_lo half 16 KB) | out ring 2 x (buf0, lo, buf2 = 48 KB) | in ring 2 x (in0,
// in2 = 32 KB) = 224 KB.  One epilogue team of 16 warps per CTA (4 row quarters x 4 column slices) works on one chunk at a
// time; the chunks alternate between the two out-ring stages, so its work overlaps the MMAs / stores of the chunk before.
// ----------------------------------------------------------------------------------------------------------------
struct Cfg2 {
  static constexpr int W_STAGE = 32768, W_STAGES = 2;
  static constexpr int O_STAGE = 3 * CHUNK_BYTES, O_STAGES = 2;
  static constexpr int I_STAGE = 2 * CHUNK_BYTES, I_STAGES = 2;
  static constexpr int EPI_WARPS = 16, CW = 32 / (EPI_WARPS / 4);
  static constexpr int EPI_WARP0 = 4;
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
  static constexpr int B2_OFF = 2 * CHUNK_BYTES;
  static constexpr int SMEM_BYTES = W_STAGES * W_STAGE + O_STAGES * O_STAGE + I_STAGES * I_STAGE + 2048 + 1024;
};
// TMA load into THIS CTA's shared memory that completes on an mbarrier of EITHER CTA of the pair (.cta_group::2): the
// peer's weight half lands in the peer's shared memory and signals the leader's w_full.  Without the qualifier the barrier
// has to live in the destination CTA.
__device__ __forceinline__ void tma_load_2d_to(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cwait_cluster(uint64_t* b, uint32_t parity, int role, int what, int link, int chunk) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(b);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > tc::kSpinLimit) chain_timeout(role, what, link, chunk);
  }
}

template <int SWEEP>
__global__ void __launch_bounds__(Cfg2::NUM_THREADS, 1)
chain2_kernel(const __grid_constant__ Maps tm, const Args a) {
  using C = Cfg2;
  constexpr bool X3 = true;
  constexpr int OS = C::O_STAGES, IS = C::I_STAGES, WS = C::W_STAGES, EW = C::EPI_WARPS, CW = C::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wring = smem;
  uint8_t* oring = smem + WS * C::W_STAGE;
  uint8_t* iring = oring + OS * C::O_STAGE;
  uint8_t* extra = iring + IS * C::I_STAGE;
  uint64_t* bars = (uint64_t*)extra;
  uint64_t* w_full = bars;           // [WS]  LEADER's copy: both CTAs' weight halves landed (TMA of both completes here)
  uint64_t* w_empty = bars + 2;      // [WS]  local: the pair's MMAs that read the stage completed (multicast commit)
  uint64_t* a_ready = bars + 4;      // [OS]  LEADER's copy: both CTAs' teams have written the chunk (2 arrivals)
  uint64_t* a_loc = bars + 6;        // [OS]  local: this CTA's team has written the chunk -> store warp
  uint64_t* o_free = bars + 8;       // [OS]  local: multicast commit + this CTA's store warp
  uint64_t* in_full = bars + 10;     // [IS]  local: input chunks landed
  uint64_t* in_free = bars + 12;     // [IS]  local: the team has consumed them
  uint64_t* acc_full = bars + 14;    // [2]   local: multicast commit
  uint64_t* acc_empty = bars + 16;   // [2]   LEADER's copy: both CTAs' teams drained the accumulator (2 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 18);
  float* ypart = (float*)(extra + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc2::cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a.nlinks; ++i) {
      const LinkD& L = a.link[i];
      if (L.in0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in0[i]) : "memory");
      if (L.in2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in2[i]) : "memory");
      if (L.out0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out0[i]) : "memory");
      if (L.out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out2[i]) : "memory");
      if (L.feeds) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.whi[i]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.wlo[i]) : "memory");
      }
    }
    for (int i = 0; i < WS; ++i) tc::mbar_init(&w_full[i], 1), tc::mbar_init(&w_empty[i], 1);
    for (int i = 0; i < OS; ++i) tc::mbar_init(&a_ready[i], 2), tc::mbar_init(&a_loc[i], 1), tc::mbar_init(&o_free[i], 2);
    for (int i = 0; i < IS; ++i) tc::mbar_init(&in_full[i], 1), tc::mbar_init(&in_free[i], 1);
    for (int i = 0; i < 2; ++i) tc::mbar_init(&acc_full[i], 1), tc::mbar_init(&acc_empty[i], 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  tc2::cluster_sync_all();           // both CTAs' barriers initialised, TMEM allocated
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tc::pdl_trigger();
  tc::pdl_wait();

  if (warp == 0) {
    // ===================== weight producer: this CTA's half of every k-block, completing on the LEADER's barrier ====
    if (lane == 0) {
      uint32_t ws = 0, wph = 0;
      uint32_t wf_leader[WS];
#pragma unroll
      for (int i = 0; i < WS; ++i)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(wf_leader[i]) : "r"(smem_u32(&w_full[i])), "r"(0u));
      for (int tile = cid; tile < a.ntiles; tile += ncl) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          if (!L.feeds) continue;
          const int nch = L.width >> 5, N = L.n_next, NH = N >> 1;
          for (int j = 0; j < nch; ++j) {
            cwait(&w_empty[ws], wph ^ 1, 0, 0, i, j);
            uint8_t* dst = wring + ws * C::W_STAGE;
            const uint32_t bar = ws == 0 ? wf_leader[0] : wf_leader[1];
            if (leader) tc::mbar_expect_tx(&w_full[ws], (uint32_t)N * 256u);   // both halves, hi + lo
            if (L.b_mn) {
              for (int c = 0; c < NH / 32; ++c) {
                tma_load_2d_to(dst + c * 4096, &tm.whi[i], bar, NH * (int)rank + 32 * c, 32 * j);
                tma_load_2d_to(dst + 16384 + c * 4096, &tm.wlo[i], bar, NH * (int)rank + 32 * c, 32 * j);
              }
            } else {
              tma_load_2d_to(dst, &tm.whi[i], bar, 32 * j, NH * (int)rank);
              tma_load_2d_to(dst + 16384, &tm.wlo[i], bar, 32 * j, NH * (int)rank);
            }
            if (++ws == WS) ws = 0, wph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== input producer (own 128 rows): runs IS chunks ahead of the epilogue =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = cid; tile < a.ntiles; tile += ncl) {
        const int m0 = tile * 256 + 128 * (int)rank;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          if (!(L.in0 || L.in2)) continue;
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            cwait(&in_free[s], ph ^ 1, 2, 0, i, j);
            uint8_t* st = iring + s * C::I_STAGE;
            tc::mbar_expect_tx(&in_full[s], (uint32_t)CHUNK_BYTES * (uint32_t)((L.in0 ? 1 : 0) + (L.in2 ? 1 : 0)));
            if (L.in0) tc::tma_load_2d(st, &tm.in0[i], &in_full[s], 32 * j, m0);
            if (L.in2) tc::tma_load_2d(st + CHUNK_BYTES, &tm.in2[i], &in_full[s], 32 * j, m0);
            if (++s == IS) s = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      uint32_t s = 0, ph = 0, ws = 0, wph = 0, mm = 0;
      for (int tile = cid; tile < a.ntiles; tile += ncl) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          const uint32_t acc = mm & 1;
          uint32_t idesc = 0;
          if (L.feeds) {
            cwait_cluster(&acc_empty[acc], ((mm >> 1) & 1) ^ 1, 1, 0, i, 0);
            tc::tc_fence_after();
            idesc = tc2::make_idesc_pair(L.n_next, false, L.b_mn != 0);
          }
          const uint32_t tmem_d = tmem_base + acc * 256;
          for (int j = 0; j < nch; ++j) {
            cwait_cluster(&a_ready[s], ph, 1, 1, i, j);
            if (L.feeds) {
              cwait_cluster(&w_full[ws], wph, 1, 2, i, j);
              tc::tc_fence_after();
              const uint32_t a0 = smem_u32(oring + s * C::O_STAGE);
              const uint32_t alo = a0 + CHUNK_BYTES;
              const uint32_t b0 = smem_u32(wring + ws * C::W_STAGE);
              const uint32_t blo = b0 + 16384;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = tc::make_desc(a0 + k * 32, 16, 1024, 2);
                const uint64_t dal = tc::make_desc(alo + k * 32, 16, 1024, 2);
                const uint64_t db = L.b_mn ? tc::make_desc(b0 + k * 1024, 4096, 512, 1) : tc::make_desc(b0 + k * 32, 16, 1024, 2);
                const uint64_t dbl = L.b_mn ? tc::make_desc(blo + k * 1024, 4096, 512, 1) : tc::make_desc(blo + k * 32, 16, 1024, 2);
                tc2::umma_tf32_pair(tmem_d, dal, db, idesc, (j | k) ? 1u : 0u);
                tc2::umma_tf32_pair(tmem_d, da, dbl, idesc, 1u);
                tc2::umma_tf32_pair(tmem_d, da, db, idesc, 1u);
              }
              tc2::umma_commit_pair(&w_empty[ws]);
              if (++ws == WS) ws = 0, wph ^= 1;
            }
            tc2::umma_commit_pair(&o_free[s]);
            if (++s == OS) s = 0, ph ^= 1;
          }
          if (L.feeds) {
            tc2::umma_commit_pair(&acc_full[acc]);
            ++mm;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== store warp (own 128 rows) =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = cid; tile < a.ntiles; tile += ncl) {
        const int m0 = tile * 256 + 128 * (int)rank;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            cwait(&a_loc[s], ph, 3, 0, i, j);
            const uint8_t* st = oring + s * C::O_STAGE;
            if (L.out0) tma_store_2d(&tm.out0[i], st, 32 * j, m0);
            if (L.out2) tma_store_2d(&tm.out2[i], st + C::B2_OFF, 32 * j, m0);
            if (L.out0 || L.out2) {
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            tc::mbar_arrive(&o_free[s]);
            if (++s == OS) s = 0, ph ^= 1;
          }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    // ===================== epilogue team (own 128 rows, own TMEM lanes) =====================
    const int e = warp - C::EPI_WARP0;
    const int q = warp & 3;
    const int h = e >> 2;
    const int rt = q * 32 + lane;
    const bool elected = e == 0 && lane == 0;
    uint32_t so = 0, pho = 0, si = 0, phi = 0, mmr = 0;
    bool first_tile = true;
    for (int tile = cid; tile < a.ntiles; tile += ncl, first_tile = false) {
      const int row = tile * 256 + 128 * (int)rank + rt;
      const bool valid = row < a.rows;
      float yb = 0.f, yacc = 0.f;
      if (SWEEP == SWEEP_T) yb = valid ? __ldg(a.ybar + row) : 0.f;
      for (int i = 0; i < a.nlinks; ++i) {
        const LinkD& L = a.link[i];
        const int nch = L.width >> 5;
        const bool has_acc = L.kind != LINK_FIRST;
        const bool has_in = L.in0 || L.in2;
        const uint32_t acc = mmr & 1;
        if (has_acc) {
          cwait(&acc_full[acc], (mmr >> 1) & 1, 4, 0, i, 0);
          tc::tc_fence_after();
        }
        for (int j = 0; j < nch; ++j) {
          if (has_in) cwait(&in_full[si], phi, 4, 1, i, j);
          cwait(&o_free[so], pho ^ 1, 4, 2, i, j);      // MMAs + store of the chunk two positions back are done with it
          const float* i0 = (const float*)(iring + si * C::I_STAGE);
          const float* i2 = i0 + CHUNK_BYTES / 4;
          float* b0 = (float*)(oring + so * C::O_STAGE);
          float* b1 = b0 + CHUNK_BYTES / 4;
          float* b2 = (float*)(oring + so * C::O_STAGE + C::B2_OFF);
          uint32_t v[CW];
          if (has_acc) {
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + (uint32_t)(32 * j + h * CW);
            tmem_ld_cols<CW>(taddr, v);
          }
          const int c0 = 32 * j;
          chunk_math<SWEEP, X3, CW, false>(a, L, i, c0, rt, h * (CW / 4), i0, i2, b0, b1, b2, v, yb, yacc);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc::tc_fence_before();
          named_bar(1, 32 * EW);                         // the chunk is written, its inputs are consumed
          if (L.colsum) {
            if (h == 0) chunk_colsum(a, L, i, c0, q, lane, b0, b2, first_tile);
            named_bar(1, 32 * EW);                       // ... and read back: the stage may be handed on
          }
          // ONE thread signals: the input producer and the store warp of this CTA, and the leader's MMA issuer (a
          // cluster-scope release arrive is a gpu-wide fence, so the peer does it once per chunk, not once per warp)
          if (elected) {
            if (has_in) tc::mbar_arrive(&in_free[si]);
            tc::mbar_arrive(&a_loc[so]);
            if (leader) tc::mbar_arrive(&a_ready[so]);
            else tc2::mbar_arrive_cta(&a_ready[so], 0);
          }
          if (has_in && ++si == IS) si = 0, phi ^= 1;
          if (++so == OS) so = 0, pho ^= 1;
        }
        if (has_acc) {
          // every warp's tcgen05.ld of this accumulator completed before the chunk barrier above
          if (elected) {
            if (leader) tc::mbar_arrive(&acc_empty[acc]);
            else tc2::mbar_arrive_cta(&acc_empty[acc], 0);
          }
          ++mmr;
        }
        if (SWEEP == SWEEP_F && L.kind == LINK_LAST) {
          if (h > 0) ypart[(h - 1) * 128 + rt] = yacc;
          named_bar(1, 32 * EW);
          if (h == 0 && valid) {
            float y = yacc;
#pragma unroll
            for (int hh = 1; hh < EW / 4; ++hh) y += ypart[(hh - 1) * 128 + rt];
            a.Y[row] = y + __ldg(a.bout);
          }
          named_bar(1, 32 * EW);
        }
      }
    }
  }
  tc::tc_fence_before();
  tc2::cluster_sync_all();           // no CTA leaves (or frees TMEM) while its peer may still touch it
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}


// ----------------------------------------------------------------------------------------------------------------
// chaint_kernel: the carried operand lives in TENSOR MEMORY.
//
// In chain_kernel the chunk the epilogue writes is at once MMA operand, TMA-store source and the landing zone of the next
// input load, so a stage is only re-loaded after the MMAs AND the store of its previous chunk -- with two 48 KB stages
// next to 128 KB of weight twins the row-array traffic cannot overlap the MMAs (ablation table: A sweep 13.4 ms = 7.1 ms
// without I/O + 6.3 ms of I/O at HBM speed).  Here
//   * the epilogue team first DRAINS the whole accumulator (half into registers, half parked in 128 spare TMEM columns),
//     which frees it for the next MMA at once: one 256-column accumulator suffices;
//   * the last 128 TMEM columns are a ring of A-operand stages (hi in 32 columns, 3xTF32 lo in the next 32): the team
//     writes the next layer's operand with tcgen05.st (lane = row, column = k) and the MMAs take A from TMEM
//     (tcgen05.mma [d], [a], b_desc), so they read only the weight k-block from shared memory;
//   * shared memory beside the weight ring is a pure I/O ring (3-4 stages of in0|in2 -> out0|out2, 32 KB each) that the
//     MMAs never touch: a stage is free again as soon as its TMA store has read it, and loads run 3-4 chunks ahead.
// Warps as chain_kernel: 0 weight producer | 1 MMA issuer | 2 input producer | 3 store | 4..19 epilogue team.
// ----------------------------------------------------------------------------------------------------------------
template <bool X3>
struct CfgT {
  static constexpr int W_STAGE = X3 ? 65536 : 32768;
  static constexpr int W_STAGES = X3 ? 2 : 3;
  static constexpr int IO_STAGE = 2 * CHUNK_BYTES;       // [buf0 = in0 / out0][buf2 = in2 / out2]
  static constexpr int IO_STAGES = X3 ? 3 : 4;
  // TMEM: accumulator [0, 256) | parked accumulator chunks [256, 384) | operand ring [384, 512)
  static constexpr int PARK_COL = 256, TA_COL0 = 384;
  static constexpr int TA_COLS = X3 ? 64 : 32;           // TMEM columns of one operand stage: hi (, lo)
  static constexpr int TA_STAGES = 128 / TA_COLS;
  static constexpr int EPI_WARPS = 16, CW = 16, EPI_WARP0 = 4;   // two half-teams of 8 warps, 16 columns of a chunk per thread
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
  static constexpr int EXTRA_BYTES = 2048;               // barriers (512 B), head partials (1536 B)
  static constexpr int SMEM_BYTES = W_STAGES * W_STAGE + IO_STAGES * IO_STAGE + EXTRA_BYTES + 1024 /*align*/;
};
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// ac[c] = (jq == jj) ? v[c] : ac[c]
__device__ __forceinline__ void sel8(uint32_t (&ac)[8], const uint32_t* v, int jq, int jj) {
  asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %16, %17;\n\t"
      "selp.b32 %0, %8, %0, p;\n\tselp.b32 %1, %9, %1, p;\n\tselp.b32 %2, %10, %2, p;\n\tselp.b32 %3, %11, %3, p;\n\t"
      "selp.b32 %4, %12, %4, p;\n\tselp.b32 %5, %13, %5, p;\n\tselp.b32 %6, %14, %6, p;\n\tselp.b32 %7, %15, %7, p;\n\t}"
      : "+r"(ac[0]), "+r"(ac[1]), "+r"(ac[2]), "+r"(ac[3]), "+r"(ac[4]), "+r"(ac[5]), "+r"(ac[6]), "+r"(ac[7])
      : "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(jq), "r"(jj));
}
// ac[k] = which ? v1[k] : v0[k]
__device__ __forceinline__ void sel16(uint32_t (&ac)[16], const uint32_t (&v0)[16], const uint32_t (&v1)[16], int which) {
#pragma unroll
  for (int k = 0; k < 16; ++k)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.b32 %0, %2, %1, p;\n\t}" : "=r"(ac[k]) : "r"(v0[k]), "r"(v1[k]), "r"(which));
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// bounded wait without a function call: a call anywhere in the kernel (chain_timeout's printf) makes ptxas allocate
// every setmaxnreg region for the smallest limit
#ifdef FBSNN_CHAIN_PROF
// debugging builds: the first wait that times out records who starved {block, role, what, link, chunk, parity} and turns
// every later wait of the launch into a no-op, so that the kernel ends (with garbage) and the host can read the record
__device__ unsigned int g_chain_trap[8];
__device__ __forceinline__ void twait(uint64_t* b, uint32_t parity, int role, int what, int link, int chunk) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(b);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (*(volatile unsigned int*)&g_chain_trap[0]) break;
    if (++spins > (1u << 20)) {
      if (atomicCAS(&g_chain_trap[0], 0u, 1u) == 0u) {
        g_chain_trap[1] = blockIdx.x, g_chain_trap[2] = role, g_chain_trap[3] = what, g_chain_trap[4] = link;
        g_chain_trap[5] = chunk, g_chain_trap[6] = parity, g_chain_trap[7] = threadIdx.x;
        __threadfence();
      }
      break;
    }
  }
}
#else
__device__ __forceinline__ void twait(uint64_t* b, uint32_t parity, int, int, int, int) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(b);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > tc::kSpinLimit) asm volatile("trap;");
  }
}

#endif
// Wait / work cycle accounting of chaint_kernel (builds with -DFBSNN_CHAIN_PROF only: tools/chain_prof.py): every role
// thread attributes the cycles since its previous lap to one slot of g_chain_prof[CTA][slot].
#ifdef FBSNN_CHAIN_PROF
__device__ unsigned long long g_chain_prof[4][160][24];   // [sweep][CTA][slot]
#define CPROF_BEGIN(on) const bool _cpon = (on); long long _cp = clock64();
#define CPROF_LAP(k)                                                                            \
  if (_cpon) {                                                                                  \
    const long long _n = clock64();                                                             \
    atomicAdd(&g_chain_prof[SWEEP][blockIdx.x][k], (unsigned long long)(_n - _cp));                    \
    _cp = _n;                                                                                   \
  }
#else
#define CPROF_BEGIN(on)
#define CPROF_LAP(k)
#endif

// Barriers and rings the epilogue threads of chaint_kernel work with, and a thread's position in them.
struct TeamCtx {
  uint64_t *in_full, *out_ready, *a_ready, *a_free, *acc_full, *acc_empty;
  uint8_t* ioring;
  float* ypart;
  uint32_t lane_base;             // TMEM address of this warp's lane quarter
  int q, h2, grp, rt, lane;
  uint32_t* started;              // [2] per half-team: 1 + index of its latest chunk whose inputs it has seen land
  uint32_t t, tph, s, sph, dr;    // I/O stage + phase, operand stage + phase, accumulators drained so far
  uint32_t gc;                    // chunks of this launch so far (both half-teams count all of them)
};

// 16 columns (this thread's half of chunk j, row rt) of link `i`: straight-line code per (SWEEP, KIND, ACT) -- what the
// generic chunk_math decides per 4-column group at run time (link kind, which buffers exist, activation) costs more
// instructions than the arithmetic itself.  Same arithmetic, in the same order, as chunk_math.
//   ac[]  accumulator slice, b0 / b2 = the I/O stage's buffers (inputs on entry, outputs on exit), hi[] = operand of the
//   next MMA (KIND != LINK_LAST).
template <int SWEEP, int KIND, int ACT>
__device__ __forceinline__ void ta_math16(const Args& a, int i, int c0, int rt, int h2, float* b0, float* b2,
                                          const uint32_t (&ac)[16], float yb, float& yacc, uint32_t (&hi)[16]) {
  const int rsw = rt & 7;
#pragma unroll
  for (int uu = 0; uu < 4; ++uu) {
    const int u = 4 * h2 + uu;
    const int off = rt * 32 + ((u ^ rsw) << 2);
    const int col = c0 + 4 * u;
    const float acc[4] = {__uint_as_float(ac[4 * uu]), __uint_as_float(ac[4 * uu + 1]), __uint_as_float(ac[4 * uu + 2]),
                          __uint_as_float(ac[4 * uu + 3])};
    float o0[4];
    if constexpr (KIND == LINK_FIRST && SWEEP != SWEEP_A) {          // X / V / zbar_L pass through
      const float4 x = ld4(b0 + off);
      o0[0] = x.x, o0[1] = x.y, o0[2] = x.z, o0[3] = x.w;
    } else if constexpr (SWEEP == SWEEP_F) {                         // g = act(z), a = act'(z)
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias[i] + col));
      const float z[4] = {acc[0] + b.x, acc[1] + b.y, acc[2] + b.z, acc[3] + b.w};
      float o2[4];
      act_ga4(ACT, z, o0, o2);
      st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
      st4(b2 + off, make_float4(o2[0], o2[1], o2[2], o2[3]));
      if constexpr (KIND == LINK_LAST) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.wout + col));
        yacc = fmaf(o0[0], w.x, fmaf(o0[1], w.y, fmaf(o0[2], w.z, fmaf(o0[3], w.w, yacc))));
      }
    } else if constexpr (SWEEP == SWEEP_A) {
      if constexpr (KIND == LINK_FIRST) {                            // delta_L = wout * a_L
        const float4 x = ld4(b0 + off);
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.wout + col));
        o0[0] = w.x * x.x, o0[1] = w.y * x.y, o0[2] = w.z * x.z, o0[3] = w.w * x.w;
        st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
      } else if constexpr (KIND == LINK_MID) {                       // ht = acc; delta = ht * a; s = ht * c(g, a)
        const float4 x = ld4(b0 + off);
        const float x0[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = acc[t] * x0[t];
        if (a.with_s) {
          const float4 g4 = ld4(b2 + off);
          const float x2[4] = {g4.x, g4.y, g4.z, g4.w};
          float o2[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) o2[t] = acc[t] * act_c(ACT, x2[t], x0[t]);
          st4(b2 + off, make_float4(o2[0], o2[1], o2[2], o2[3]));
        }
        st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
      } else {                                                       // Du
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = acc[t];
        st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
      }
    } else if constexpr (SWEEP == SWEEP_T) {
      const float4 x = ld4(b0 + off), y = ld4(b2 + off);
      const float x0[4] = {x.x, x.y, x.z, x.w}, x2[4] = {y.x, y.y, y.z, y.w};
      float o2[4];
      if constexpr (KIND == LINK_MID) {                              // dbar = acc; hd = dbar * a; zz = dbar * s
#pragma unroll
        for (int t = 0; t < 4; ++t) o0[t] = acc[t] * x0[t], o2[t] = acc[t] * x2[t];
      } else {   // last hidden layer: zbar = ybar wout a + dbar (wout c);  wg = dbar a + ybar g  (column sums only)
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.wout + col));
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float av = x0[t], gv = x2[t];
          const float sv = w[t] * act_c(ACT, gv, av);
          const float zz = acc[t] * sv;
          o0[t] = zz + yb * w[t] * av;
          o2[t] = acc[t] * av + yb * gv;
        }
      }
      st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
      st4(b2 + off, make_float4(o2[0], o2[1], o2[2], o2[3]));
    } else {                                                         // B: hb = acc; zbar = hb * a + zz
      const float4 x = ld4(b0 + off), y = ld4(b2 + off);
      o0[0] = acc[0] * x.x + y.x, o0[1] = acc[1] * x.y + y.y, o0[2] = acc[2] * x.z + y.z, o0[3] = acc[3] * x.w + y.w;
      st4(b0 + off, make_float4(o0[0], o0[1], o0[2], o0[3]));
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) hi[4 * uu + t] = __float_as_uint(o0[t]);
  }
}

// One link of one tile for an epilogue thread of chaint_kernel: drain the accumulator (KIND != LINK_FIRST), then this
// half-team's chunks.  FEEDS / has-accumulator / column sums follow from (SWEEP, KIND) -- fbsnn_api.cu builds the links so.
template <int SWEEP, int KIND, bool X3, int ACT>
__device__ __forceinline__ void ta_link(const Args& a, const LinkD& L, int i, TeamCtx& c, float yb, float& yacc, bool first_tile) {
  using C = CfgT<X3>;
  constexpr int IOS = C::IO_STAGES, TS = C::TA_STAGES, EW = C::EPI_WARPS, CW = C::CW;
  constexpr bool HAS_ACC = KIND != LINK_FIRST, FEEDS = KIND != LINK_LAST;
  constexpr int COLSUM = (SWEEP == SWEEP_T && KIND == LINK_LAST) ? 3 : (SWEEP == SWEEP_B && KIND != LINK_FIRST) ? 1 : 0;
  const int nch = L.width >> 5;
  const int grp = c.grp, h2 = c.h2;
  // This thread's slice of the accumulator: 16 columns of each of its (up to 4) chunks.  The chunks j < 4 stay in registers,
  // the chunks j >= 4 are PARKED in TMEM columns [256, 384) -- 64 registers per thread would not fit next to the working
  // set.  Each warp reads back only what it parked itself (same lanes, same columns).
  // (Unconditional loads: columns past the link's width are stale accumulator columns nobody uses.)
  uint32_t v0[16], v1[16];
  CPROF_BEGIN(c.q == 0 && c.h2 == 0 && c.grp == 0 && c.lane == 0)
  if constexpr (HAS_ACC) {
    twait(&c.acc_full[0], c.dr & 1, 4, 0, i, 0);
    CPROF_LAP(0)
    tc::tc_fence_after();
    if (nch > 4) {
      uint32_t p0[16], p1[16];
      tmem_ld16_nowait(c.lane_base + (uint32_t)(32 * (4 + grp) + h2 * CW), p0);
      tmem_ld16_nowait(c.lane_base + (uint32_t)(32 * (6 + grp) + h2 * CW), p1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tmem_st16(c.lane_base + (uint32_t)(C::PARK_COL + 32 * grp + h2 * CW), p0);
      tmem_st16(c.lane_base + (uint32_t)(C::PARK_COL + 32 * (2 + grp) + h2 * CW), p1);
    }
    tmem_ld16_nowait(c.lane_base + (uint32_t)(32 * grp + h2 * CW), v0);
    tmem_ld16_nowait(c.lane_base + (uint32_t)(32 * (2 + grp) + h2 * CW), v1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc::tc_fence_before();
    __syncwarp();
    if (c.lane == 0) tc::mbar_arrive(&c.acc_empty[0]);
    ++c.dr;
    CPROF_LAP(1)
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) v0[k] = 0u, v1[k] = 0u;
  }
#pragma unroll 1
  for (int j = 0; j < nch; ++j) {
    if ((j & 1) == grp) {
      uint32_t ac[CW];
      if (!HAS_ACC || j < 4) {   // predicated selects (opaque to the compiler, which would otherwise index v[] in local memory)
        sel16(ac, v0, v1, j >> 1);
      } else {
        tmem_ld16_nowait(c.lane_base + (uint32_t)(C::PARK_COL + 32 * (j - 4) + h2 * CW), ac);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      CPROF_LAP(2)
      if constexpr (IOS & 1) {
        // A parity wait is only meaningful once the PREVIOUS phase of the barrier has completed.  With an odd number of
        // I/O stages a stage alternates between the half-teams, and a half-team three chunks ahead of the other would take
        // the still-incomplete previous fill of the stage for its own (whole chunks of stale data): wait until the other
        // half-team has seen that fill (chunk gc - IOS) land.
        if (c.gc >= (uint32_t)IOS) {
          const volatile uint32_t* seen = c.started + (grp ^ 1);
          uint32_t spins = 0;
          while (*seen + (uint32_t)IOS <= c.gc) {
            if (++spins > tc::kSpinLimit) asm volatile("trap;");
          }
          __threadfence_block();
        }
      }
      twait(&c.in_full[c.t], c.tph, 4, 1, i, j);
      if constexpr (IOS & 1) {
        if (c.q == 0 && c.h2 == 0 && c.lane == 0) *(volatile uint32_t*)(c.started + grp) = c.gc + 1u;   // chunks [0, gc] of mine have landed
      }
      CPROF_LAP(3)
      float* b0 = (float*)(c.ioring + c.t * C::IO_STAGE);
      float* b2 = b0 + CHUNK_BYTES / 4;
      uint32_t hi[CW];
      if (!(a.ablate & 4)) {
        ta_math16<SWEEP, KIND, ACT>(a, i, 32 * j, c.rt, h2, b0, b2, ac, yb, yacc, hi);
      } else {
#pragma unroll
        for (int k = 0; k < CW; ++k) hi[k] = 0u;
      }
      CPROF_LAP(4)
      if constexpr (FEEDS) {
        twait(&c.a_free[c.s], c.sph ^ 1, 4, 2, i, j);
        CPROF_LAP(5)
        tc::tc_fence_after();
        const uint32_t taddr = c.lane_base + (uint32_t)(C::TA_COL0 + c.s * C::TA_COLS + h2 * CW);
        tmem_st16(taddr, hi);
        if constexpr (X3) {
#pragma unroll
          for (int k = 0; k < CW; ++k) hi[k] = __float_as_uint(lo_part(__uint_as_float(hi[k])));
          tmem_st16(taddr + 32, hi);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::tc_fence_before();
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> TMA store
      CPROF_LAP(6)
      if constexpr (COLSUM != 0) {
        named_bar(1 + grp, 32 * EW / 2);                             // the half-team has written the whole chunk
        if (h2 == 0) chunk_colsum(a, L, i, 32 * j, c.q, c.lane, b0, b2, first_tile);
      }
      __syncwarp();
      if (c.lane == 0) {
        if constexpr (FEEDS) tc::mbar_arrive(&c.a_ready[c.s]);
        tc::mbar_arrive(&c.out_ready[c.t]);
      }
      CPROF_LAP(7)
    }
    if (FEEDS && ++c.s == TS) c.s = 0, c.sph ^= 1;
    if (++c.t == IOS) c.t = 0, c.tph ^= 1;
    ++c.gc;
  }
}
// dispatch on the activation: the three codes fbsnn_api.cu's internal_act() produces for this precision
template <int SWEEP, int KIND, bool X3>
__device__ __forceinline__ void ta_link_act(const Args& a, const LinkD& L, int i, TeamCtx& c, float yb, float& yacc, bool first_tile) {
  constexpr bool USES_ACT = (SWEEP == SWEEP_F && KIND != LINK_FIRST) || (SWEEP == SWEEP_A && KIND == LINK_MID) ||
                            (SWEEP == SWEEP_T && KIND == LINK_LAST);
  if constexpr (!USES_ACT) {
    ta_link<SWEEP, KIND, X3, FBSNN_ACT_RELU>(a, L, i, c, yb, yacc, first_tile);
  } else {
    constexpr int SINE = X3 ? kActSineCW : kActSineFast, TANH = X3 ? FBSNN_ACT_TANH : kActTanhFast;
    if (a.act == SINE) ta_link<SWEEP, KIND, X3, SINE>(a, L, i, c, yb, yacc, first_tile);
    else if (a.act == FBSNN_ACT_RELU) ta_link<SWEEP, KIND, X3, FBSNN_ACT_RELU>(a, L, i, c, yb, yacc, first_tile);
    else ta_link<SWEEP, KIND, X3, TANH>(a, L, i, c, yb, yacc, first_tile);
  }
}

template <int SWEEP, bool X3>
__global__ void __launch_bounds__(CfgT<X3>::NUM_THREADS, 1)
chaint_kernel(const __grid_constant__ Maps tm, const Args a) {
  using C = CfgT<X3>;
  constexpr int IOS = C::IO_STAGES, WS = C::W_STAGES, TS = C::TA_STAGES, EW = C::EPI_WARPS, CW = C::CW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wring = smem;
  uint8_t* ioring = smem + WS * C::W_STAGE;
  uint8_t* extra = ioring + IOS * C::IO_STAGE;
  uint64_t* bars = (uint64_t*)extra;
  uint64_t* w_full = bars;            // [4]  weight k-block landed
  uint64_t* w_empty = bars + 4;       // [4]  MMAs that read it completed
  uint64_t* in_full = bars + 8;       // [4]  input chunks landed (or: stage handed to the team)
  uint64_t* out_ready = bars + 12;    // [4]  the team has written the stage's outputs (one arrival per warp)
  uint64_t* io_free = bars + 16;      // [4]  the TMA stores have read the stage
  uint64_t* a_ready = bars + 20;      // [8]  operand stage written to TMEM (one arrival per warp)
  uint64_t* a_free = bars + 28;       // [8]  MMAs that read the operand stage completed
  uint64_t* acc_full = bars + 36;     // [1]
  uint64_t* acc_empty = bars + 37;    // [1]  every team warp holds its accumulator slice in registers
  uint32_t* tmem_slot = (uint32_t*)(bars + 38);
  float* ypart = (float*)(extra + 512);   // [3][128] head partial sums (F sweep)
  uint32_t* started = (uint32_t*)(extra + 384);   // [2] progress of the two half-teams (odd IO_STAGES only)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a.nlinks; ++i) {
      const LinkD& L = a.link[i];
      if (L.in0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in0[i]) : "memory");
      if (L.in2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.in2[i]) : "memory");
      if (L.out0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out0[i]) : "memory");
      if (L.out2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.out2[i]) : "memory");
      if (L.feeds) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.whi[i]) : "memory");
        if (X3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.wlo[i]) : "memory");
      }
    }
    started[0] = started[1] = 0u;
    for (int i = 0; i < WS; ++i) tc::mbar_init(&w_full[i], 1), tc::mbar_init(&w_empty[i], 1);
    for (int i = 0; i < IOS; ++i) tc::mbar_init(&in_full[i], 1), tc::mbar_init(&out_ready[i], EW / 2), tc::mbar_init(&io_free[i], 1);
    for (int i = 0; i < TS; ++i) tc::mbar_init(&a_ready[i], EW / 2), tc::mbar_init(&a_free[i], 1);
    tc::mbar_init(&acc_full[0], 1), tc::mbar_init(&acc_empty[0], EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  tc::pdl_trigger();
  tc::pdl_wait();

  // register re-allocation: the four control warps (one warpgroup) keep 24 registers each, the team's warps get 112 --
  // their accumulator slice (32 registers) sits next to the working set of the fused epilogue
  // (issued inside every role branch: after a merge point ptxas allocates for the smallest limit)
  if (warp == 0) {
    // ===================== weight producer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      uint32_t ws = 0, wph = 0;
      const uint64_t wpol = l2_evict_last_policy();
      CPROF_BEGIN(true)
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          if (!L.feeds) continue;
          const int nch = L.width >> 5, N = L.n_next;
          const uint32_t bytes = (uint32_t)N * 128u * (X3 ? 2u : 1u);
          for (int j = 0; j < nch; ++j) {
            twait(&w_empty[ws], wph ^ 1, 0, 0, i, j);
            CPROF_LAP(16)
            uint8_t* dst = wring + ws * C::W_STAGE;
            if (a.ablate & 16) {
              tc::mbar_arrive(&w_full[ws]);
            } else {
              tc::mbar_expect_tx(&w_full[ws], bytes);
              if (a.hints & 4) {   // weights: L2 evict_last (they are re-read by every tile; the row arrays stream through)
                if (L.b_mn) {
                  for (int c = 0; c < N / 32; ++c) {
                    tma_load_2d_hint(dst + c * 4096, &tm.whi[i], &w_full[ws], 32 * c, 32 * j, wpol);
                    if (X3) tma_load_2d_hint(dst + 32768 + c * 4096, &tm.wlo[i], &w_full[ws], 32 * c, 32 * j, wpol);
                  }
                } else {
                  tma_load_2d_hint(dst, &tm.whi[i], &w_full[ws], 32 * j, 0, wpol);
                  if (X3) tma_load_2d_hint(dst + 32768, &tm.wlo[i], &w_full[ws], 32 * j, 0, wpol);
                }
              } else if (L.b_mn) {   // W[k][n] (n contiguous): 32 x 32 boxes, 128B swizzle with 32B atoms
                for (int c = 0; c < N / 32; ++c) {
                  tc::tma_load_2d(dst + c * 4096, &tm.whi[i], &w_full[ws], 32 * c, 32 * j);
                  if (X3) tc::tma_load_2d(dst + 32768 + c * 4096, &tm.wlo[i], &w_full[ws], 32 * c, 32 * j);
                }
              } else {        // W[n][k] (k contiguous): one N x 32 box
                tc::tma_load_2d(dst, &tm.whi[i], &w_full[ws], 32 * j, 0);
                if (X3) tc::tma_load_2d(dst + 32768, &tm.wlo[i], &w_full[ws], 32 * j, 0);
              }
            }
            CPROF_LAP(17)
            if (++ws == WS) ws = 0, wph ^= 1;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== input producer: runs IOS chunks ahead of the team =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      uint32_t t = 0, ph = 0;
      ChunkIt pf{(int)blockIdx.x, 0, 0};
      auto prefetch_one = [&]() {   // row arrays of the chunks a few positions ahead -> L2
        if (!pf.valid(a) || (a.ablate & 2)) return;
        const LinkD& P = a.link[pf.link];
        if (P.in0) tma_prefetch_2d(&tm.in0[pf.link], 32 * pf.j, pf.tile * 128);
        if (P.in2) tma_prefetch_2d(&tm.in2[pf.link], 32 * pf.j, pf.tile * 128);
        pf.next(a, (int)gridDim.x);
      };
      const uint64_t pol = l2_evict_first_policy();
      for (int p = 0; p < a.pf_dist; ++p) prefetch_one();
      CPROF_BEGIN(true)
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int m0 = tile * 128;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            twait(&io_free[t], ph ^ 1, 2, 0, i, j);
            CPROF_LAP(14)
            prefetch_one();
            uint8_t* st = ioring + t * C::IO_STAGE;
            if ((L.in0 || L.in2) && !(a.ablate & 2)) {
              tc::mbar_expect_tx(&in_full[t], (uint32_t)CHUNK_BYTES * (uint32_t)((L.in0 ? 1 : 0) + (L.in2 ? 1 : 0)));
              if (a.hints & 2) {
                if (L.in0) tma_load_2d_hint(st, &tm.in0[i], &in_full[t], 32 * j, m0, pol);
                if (L.in2) tma_load_2d_hint(st + CHUNK_BYTES, &tm.in2[i], &in_full[t], 32 * j, m0, pol);
              } else {
                if (L.in0) tc::tma_load_2d(st, &tm.in0[i], &in_full[t], 32 * j, m0);
                if (L.in2) tc::tma_load_2d(st + CHUNK_BYTES, &tm.in2[i], &in_full[t], 32 * j, m0);
              }
            } else {
              tc::mbar_arrive(&in_full[t]);
            }
            CPROF_LAP(15)
            if (++t == IOS) t = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      uint32_t s = 0, sph = 0, ws = 0, wph = 0, mm = 0;
      const uint32_t tmem_d = tmem_base;
      CPROF_BEGIN(true)
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          if (!L.feeds) continue;
          const int nch = L.width >> 5;
          twait(&acc_empty[0], (mm & 1) ^ 1, 1, 0, i, 0);   // the accumulator of the MMA before this one is in registers
          CPROF_LAP(8)
          tc::tc_fence_after();
          const uint32_t idesc = tc::make_idesc(L.n_next, false, L.b_mn != 0);
          for (int j = 0; j < nch; ++j) {
            twait(&a_ready[s], sph, 1, 1, i, j);
            CPROF_LAP(9)
            twait(&w_full[ws], wph, 1, 2, i, j);
            CPROF_LAP(10)
            tc::tc_fence_after();
            if (!(a.ablate & 8)) {
              const uint32_t ta = tmem_base + C::TA_COL0 + s * C::TA_COLS;
              const uint32_t b0 = smem_u32(wring + ws * C::W_STAGE);
              const uint32_t blo = b0 + 32768;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t db = L.b_mn ? tc::make_desc(b0 + k * 1024, 4096, 512, 1) : tc::make_desc(b0 + k * 32, 16, 1024, 2);
                if (X3) {
                  const uint64_t dbl = L.b_mn ? tc::make_desc(blo + k * 1024, 4096, 512, 1) : tc::make_desc(blo + k * 32, 16, 1024, 2);
                  umma_tf32_ts(tmem_d, ta + 32 + 8 * k, db, idesc, (j | k) ? 1u : 0u);
                  umma_tf32_ts(tmem_d, ta + 8 * k, dbl, idesc, 1u);
                  umma_tf32_ts(tmem_d, ta + 8 * k, db, idesc, 1u);
                } else {
                  umma_tf32_ts(tmem_d, ta + 8 * k, db, idesc, (j | k) ? 1u : 0u);
                }
              }
            }
            tc::umma_commit(&w_empty[ws]);
            tc::umma_commit(&a_free[s]);
            CPROF_LAP(11)
            if (++ws == WS) ws = 0, wph ^= 1;
            if (++s == TS) s = 0, sph ^= 1;
          }
          tc::umma_commit(&acc_full[0]);
          ++mm;
        }
      }
    }
  } else if (warp == 3) {
    // ===================== store warp =====================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (lane == 0) {
      uint32_t t = 0, ph = 0;
      int pending = -1;   // stage whose stores are committed but not yet known to have been read
      const uint64_t pol = l2_evict_first_policy();
      CPROF_BEGIN(true)
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int m0 = tile * 128;
        for (int i = 0; i < a.nlinks; ++i) {
          const LinkD& L = a.link[i];
          const int nch = L.width >> 5;
          for (int j = 0; j < nch; ++j) {
            twait(&out_ready[t], ph, 3, 0, i, j);
            CPROF_LAP(12)
            const uint8_t* st = ioring + t * C::IO_STAGE;
            const bool do_store = (L.out0 || L.out2) && !(a.ablate & 1);
            if (do_store && (a.hints & 1)) {
              if (L.out0) tma_store_2d_hint(&tm.out0[i], st, 32 * j, m0, pol);
              if (L.out2) tma_store_2d_hint(&tm.out2[i], st + CHUNK_BYTES, 32 * j, m0, pol);
            } else if (do_store) {
              if (L.out0) tma_store_2d(&tm.out0[i], st, 32 * j, m0);
              if (L.out2) tma_store_2d(&tm.out2[i], st + CHUNK_BYTES, 32 * j, m0);
            }
            if (a.hints & 8) {
              // two store groups in flight: the stage of the PREVIOUS chunk is released once its stores have read it, while
              // this chunk's stores are already queued behind them (a store takes ~1 200 cycles from issue to "read", about as
              // long as the MMAs of a chunk: one group at a time makes the store warp the pace-maker)
              if (do_store) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (pending >= 0) tc::mbar_arrive(&io_free[pending]);
                pending = (int)t;
              } else {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (pending >= 0) tc::mbar_arrive(&io_free[pending]);
                pending = -1;
                tc::mbar_arrive(&io_free[t]);
              }
            } else {
              if (do_store) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              tc::mbar_arrive(&io_free[t]);
            }
            CPROF_LAP(13)
            if (++t == IOS) t = 0, ph ^= 1;
          }
        }
      }
      if (pending >= 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        tc::mbar_arrive(&io_free[pending]);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
    }
  } else {
    // ===================== epilogue: TWO half-teams of 8 warps (4 row quarters x 2 column halves) ==================
    // Half-team g works on the chunks j = g, g + 2, ... of every link, so two chunks are in flight: one chunk's chain of
    // waits, TMEM / shared-memory round trips and fences overlaps the other's.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int e = warp - C::EPI_WARP0;
    TeamCtx c;
    c.in_full = in_full, c.out_ready = out_ready, c.a_ready = a_ready, c.a_free = a_free, c.acc_full = acc_full, c.acc_empty = acc_empty;
    c.ioring = ioring, c.ypart = ypart;
    c.q = warp & 3;                  // TMEM lane quarter (rows 32q .. 32q+31 of the tile)
    c.h2 = (e >> 2) & 1;             // column half: columns [16 h2, 16 h2 + 16) of the chunk
    c.grp = e >> 3;                  // half-team
    c.rt = c.q * 32 + lane, c.lane = lane;
    c.lane_base = tmem_base + ((uint32_t)(c.q * 32) << 16);
    c.t = c.tph = c.s = c.sph = c.dr = c.gc = 0;
    c.started = started;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, first_tile = false) {
      const int row = tile * 128 + c.rt;
      const bool valid = row < a.rows;
      float yb = 0.f, yacc = 0.f;
      if (SWEEP == SWEEP_T) yb = valid ? __ldg(a.ybar + row) : 0.f;
      for (int i = 0; i < a.nlinks; ++i) {
        const LinkD& L = a.link[i];
        if (L.kind == LINK_FIRST) ta_link_act<SWEEP, LINK_FIRST, X3>(a, L, i, c, yb, yacc, first_tile);
        else if (L.kind == LINK_MID) ta_link_act<SWEEP, LINK_MID, X3>(a, L, i, c, yb, yacc, first_tile);
        else ta_link_act<SWEEP, LINK_LAST, X3>(a, L, i, c, yb, yacc, first_tile);
        if (SWEEP == SWEEP_F && L.kind == LINK_LAST) {   // output head: u = h_L . wout + bout
          const int hh = c.grp * 2 + c.h2;
          if (hh > 0) ypart[(hh - 1) * 128 + c.rt] = yacc;
          named_bar(3, 32 * EW);
          if (hh == 0 && valid) {
            float y = yacc;
#pragma unroll
            for (int k = 1; k < 4; ++k) y += ypart[(k - 1) * 128 + c.rt];
            a.Y[row] = y + __ldg(a.bout);
          }
          named_bar(3, 32 * EW);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// out[c] = sum over CTAs and row quarters of the chain kernels' column-sum partials (fixed order, double accumulation)
struct ColFinJob {
  int slot, which, width;
  float* out;
};
struct ColFinJobs {
  ColFinJob job[2 * kMaxLinks];
  int njobs, nblk;
};
__global__ void chain_colsum_finish_kernel(const float* __restrict__ colacc, const ColFinJobs js) {
  // eight lanes per column: lane k adds the CTAs b = k, k + 8, ... (row quarters in order), then the eight partial sums are
  // combined in a fixed tree -- deterministic, and 20 instead of 160 dependent-latency loads per thread at M = 100
  const ColFinJob& j = js.job[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = t >> 3, k = t & 7;
  double acc = 0.0;
  if (c < j.width) {
    for (int b = k; b < js.nblk; b += 8) {
      const float* p = colacc + ((size_t)(b * kMaxLinks + j.slot) * 2 + j.which) * 1024 + c;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc += (double)p[q * 256];
    }
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off, 8);
  if (c < j.width && k == 0) j.out[c] = (float)acc;
}

inline int chain_grid(int ntiles, int num_sms) { return ntiles < num_sms ? ntiles : num_sms; }
template <int SWEEP, bool X3>
inline cudaError_t launch_chain(const Maps& m, const Args& a, int num_sms, cudaStream_t st) {
  using C = Cfg<X3>;
  auto kern = chain_kernel<SWEEP, X3>;
  static unsigned long long attr_devs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  return tc::launch_pdl(kern, chain_grid(a.ntiles, num_sms), C::NUM_THREADS, C::SMEM_BYTES, st, 1, m, a);
}
template <int SWEEP, bool X3>
inline cudaError_t launch_chaint(const Maps& m, const Args& a, int num_sms, cudaStream_t st) {
  using C = CfgT<X3>;
  auto kern = chaint_kernel<SWEEP, X3>;
  static unsigned long long attr_devs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  return tc::launch_pdl(kern, chain_grid(a.ntiles, num_sms), C::NUM_THREADS, C::SMEM_BYTES, st, 1, m, a);
}
// CTA-pair launch: a.ntiles = number of 256-row tiles; grid = 2 x min(tiles, SMs / 2), cluster (2,1,1)
template <int SWEEP>
inline cudaError_t launch_chain2(const Maps& m, const Args& a, int num_sms, cudaStream_t st) {
  using C = Cfg2;
  auto kern = chain2_kernel<SWEEP>;
  static unsigned long long attr_devs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int pairs = a.ntiles < num_sms / 2 ? a.ntiles : num_sms / 2;
  return tc::launch_pdl(kern, 2 * pairs, C::NUM_THREADS, C::SMEM_BYTES, st, 2, m, a);
}
inline int chain2_grid(int ntiles2, int num_sms) { return 2 * (ntiles2 < num_sms / 2 ? ntiles2 : num_sms / 2); }

}  // namespace chain
}  // namespace fbsnn
