// Sweep form of the tcgen05 dense-layer kernel with SIXTEEN epilogue warps (gemm_tc.cuh has eight).
//
// Why: ncu's source page (profiles/r01_source_level_sweeps_tf32x3_m65536.txt) shows the sweeps waiting in their
// fused epilogues, not on the tensor pipe or on TMA -- the F sweep issues ~40 instructions per element (26 of them the
// fp32-grade sine/cosine) from two epilogue warps per scheduler, the A/T/B sweeps sit on the loads of their epilogue
// operands.  Both are cured by more epilogue warps in flight.  Same producer / MMA / splitter protocol as
// gemm_tc_kernel; what changes is the epilogue geometry:
//   * 16 epilogue warps: lane quarter q = warp % 4 (the TMEM lanes a warp may read) x column quarter part = e / 4
//   * 16-column chunks: tcgen05.ld 32x32b.x16 (16 registers instead of 32), a 2 KB warp-private transpose tile
//     (16 x 2 KB = the 32 KB the eight 4 KB tiles took), after which 4 lanes cover 64 B of one row
//   * 22 warps (18 without splitters) => at most 88 (112) registers per thread, which the slimmer chunk fits
// Only the sweeps' layout (activation rows k-contiguous) and SPLIT in {0, 1} exist in this form; the weight
// gradients have a trivial epilogue and run on the CTA-pair kernel.
#pragma once
#include "gemm_tc.cuh"

namespace fbsnn {
namespace tc16 {

using namespace tc;

constexpr int EW = 16;                           // epilogue warps
constexpr int TILE16_BYTES = 32 * 16 * 4;        // 2 KB transpose tile per epilogue warp

#define FBSNN_TMEM_LD16(taddr, v)                                                                              \
  asm volatile(                                                                                                \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                         \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),        \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])   \
      : "r"(taddr))

template <int SPLIT, int BN>
struct Cfg16 {
  static_assert(SPLIT == 0 || SPLIT == 1, "sweeps only");
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGES = 3;
  static constexpr int SPLIT_WARPS = SPLIT ? 4 : 0;
  static constexpr int EPI_WARP0 = 2 + SPLIT_WARPS;
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EW);
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES + (SPLIT ? A_STAGE_BYTES : 0);   // [A][B]([A_lo])
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EW * TILE16_BYTES + 1024 /*align*/ + 256;
};

template <bool B_MN, int SPLIT, int BN, class Epi>
__global__ void __launch_bounds__(Cfg16<SPLIT, BN>::NUM_THREADS, 1)
gemm_tc16_kernel(const __grid_constant__ TmSet tm, const GemmArgs g, const Epi epi, const int num_mtiles) {
  using C = Cfg16<SPLIT, BN>;
  constexpr int B_STAGE_BYTES = C::B_BYTES;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_tiles = (float*)(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C::STAGE_BYTES + EW * TILE16_BYTES);
  uint64_t* full = bars;               // [STAGES]
  uint64_t* empty = bars + 3;          // [STAGES]
  uint64_t* sdone = bars + 6;          // [STAGES]
  uint64_t* tfull = bars + 9;          // [2]
  uint64_t* tempty = bars + 11;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = g.N < BN ? g.N : BN;
  const int num_ntiles = (g.N + BN - 1) / BN;
  const int num_work = num_mtiles * num_ntiles;         // work id = (nt, mt), mt fastest

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b[s]) : "memory");
    }
    for (int i = 0; i < STAGES; ++i)
      mbar_init(&full[i], 1), mbar_init(&empty[i], 1), mbar_init(&sdone[i], C::SPLIT_WARPS > 0 ? C::SPLIT_WARPS : 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1), mbar_init(&tempty[i], EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = A_STAGE_BYTES + (uint32_t)N * BK * 4;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int mt = w % num_mtiles, nt = w / num_mtiles;
        const int m0 = mt * BM, n0 = nt * BN;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = 0, ke = g.seg[s].K; k0 < ke; k0 += BK) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], bytes);
            uint8_t* a = smem + stage * C::STAGE_BYTES;
            uint8_t* b = a + A_STAGE_BYTES;
            tma_load_2d(a, &tm.a[s], &full[stage], k0, m0);
            if (B_MN) {
              for (int c = 0; c < N / 32; ++c) tma_load_2d(b + c * 4096, &tm.b[s], &full[stage], n0 + 32 * c, k0);
            } else {
              tma_load_2d(b, &tm.b[s], &full[stage], k0, n0);
            }
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(N, false, B_MN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = 0, ke = g.seg[s].K; k0 < ke; k0 += BK) {
            mbar_wait(SPLIT ? &sdone[stage] : &full[stage], phase);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t b = a + A_STAGE_BYTES;
            const uint32_t alo = b + B_STAGE_BYTES;
            const int mode = SPLIT == 1 ? g.mode[s] : 0;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t dbk = B_MN ? make_desc(b + k * 1024, 4096, 512, 1) : make_desc(b + k * 32, 16, 1024, 2);
              if (mode & 1) {   // a_lo * b
                umma_tf32(tmem_d, make_desc(alo + k * 32, 16, 1024, 2), dbk, idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_tf32(tmem_d, make_desc(a + k * 32, 16, 1024, 2), dbk, idesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit(&empty[stage]);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp < C::EPI_WARP0) {
    // ===================== splitter warps (SPLIT == 1): lo = x - trunc_tf32(x) of the A tile =====================
    constexpr int TEAM = 32 * (C::SPLIT_WARPS > 0 ? C::SPLIT_WARPS : 1);
    uint32_t sstage = 0, sphase = 0;
    const int ts = threadIdx.x - 64;
    constexpr int nA4 = A_STAGE_BYTES / 16;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      for (int s = 0; s < g.nseg; ++s) {
        for (int k0 = 0, ke = g.seg[s].K; k0 < ke; k0 += BK) {
          mbar_wait(&full[sstage], sphase);
          const float4* a = (const float4*)(smem + sstage * C::STAGE_BYTES);
          float4* lo = (float4*)(smem + sstage * C::STAGE_BYTES + A_STAGE_BYTES + B_STAGE_BYTES);
          if (g.mode[s] == 1) {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = a[ts + u * TEAM];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float4 l;
              l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u);
              l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u);
              l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u);
              l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u);
              lo[ts + u * TEAM] = l;
            }
            static_assert(SPLIT == 0 || nA4 == 8 * TEAM, "the A tile is 8 float4 per splitter thread");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdone[sstage]);
          if (++sstage == STAGES) sstage = 0, sphase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int e = warp - C::EPI_WARP0;
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int part = e >> 2;               // column quarter
    const int ncol = N >> 2;               // 16, 32, 48 or 64 columns per warp
    float* tile = epi_tiles + e * (TILE16_BYTES / 4);
    const int rsub = lane >> 2, c4 = lane & 3;            // after the transpose: row 8 i + rsub, columns 4 c4 .. 4 c4 + 3
    constexpr bool kColsum = Epi::kColsum;
    float4 csum[4];                        // per-thread column sums of zbar: chunk x 4 columns
#pragma unroll
    for (int i = 0; i < 4; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int mt = w % num_mtiles;
      const int n0 = (w / num_mtiles) * BN;
      const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      const int r0 = mt * BM + q * 32;
#pragma unroll 1
      for (int ch = 0; ch * 16 < ncol; ++ch) {
        const int ct = part * ncol + ch * 16;   // column inside the CTA tile (TMEM column)
        const int c0 = n0 + ct;                 // global output column
        if (c4 == 0) {
          // operand lines of this warp's next 128-byte column group (or of its first group in its next tile) -> L2
          const bool more = (ch + 1) * 16 < ncol;
          const int wn = w + gridDim.x;
          if (more ? (((c0 + 16) & 31) == 0) : (wn < num_work)) {
            const int pr0 = more ? r0 : (wn % num_mtiles) * BM + q * 32;
            const int pc0 = more ? c0 + 16 : (wn / num_mtiles) * BN + part * ncol;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (pr0 + 8 * i + rsub < g.M) epi.l2_prefetch(pr0 + 8 * i + rsub, pc0);
          }
        }
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + ct;
        FBSNN_TMEM_LD16(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st4(tile + lane * 16 + ((j ^ ((lane >> 1) & 3)) << 2),
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3])));
        __syncwarp();
        const int cc = c4 * 4;
        const typename Epi::ColFrag cf = epi.col_prefetch(c0 + cc);
        typename Epi::Frag f[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + rsub;
          if (r0 + rr < g.M) f[i] = epi.prefetch(r0 + rr, c0 + cc);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = 8 * i + rsub;
          const float4 a4 = ld4(tile + rr * 16 + ((c4 ^ ((rr >> 1) & 3)) << 2));
          if (r0 + rr < g.M) {
            if constexpr (kColsum) {
              const float4 zb = epi.finish(r0 + rr, c0 + cc, a4, f[i], cf);
              cs.x += zb.x, cs.y += zb.y, cs.z += zb.z, cs.w += zb.w;
            } else {
              epi.finish(r0 + rr, c0 + cc, a4, f[i], cf);
            }
          }
        }
        __syncwarp();
        if constexpr (kColsum) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ch == k) csum[k].x += cs.x, csum[k].y += cs.y, csum[k].z += cs.z, csum[k].w += cs.w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }

    if constexpr (kColsum) {
      // fused bias gradient: per-CTA column sums of zbar -> colpart[row][col]; fixed reduction order (the 8 row groups
      // of a warp by shuffle, then the 4 lane-quarter warps of a column quarter through shared memory)
      if (epi.colpart) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 t = csum[ch];
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            t.x += __shfl_xor_sync(0xffffffffu, t.x, o), t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
            t.z += __shfl_xor_sync(0xffffffffu, t.z, o), t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
          }
          if (rsub == 0 && ch * 16 < ncol) st4(tile + ch * 16 + c4 * 4, t);   // this warp's 32-row sums of ncol columns
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");                       // the 16 epilogue warps only
        if (q == 0) {
          const int prow = num_work <= (int)gridDim.x ? (int)(blockIdx.x % num_mtiles) : (int)blockIdx.x;
          const int pn0 = num_work <= (int)gridDim.x ? (int)(blockIdx.x / num_mtiles) * BN : 0;
          for (int c = lane; c < ncol; c += 32) {
            float tot = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              // warp with lane quarter qq in this column quarter: e' with (EPI_WARP0 + e') % 4 == qq, e' / 4 == part
              const int e2 = part * 4 + ((qq - C::EPI_WARP0) & 3);
              tot += epi_tiles[e2 * (TILE16_BYTES / 4) + c];
            }
            epi.colpart[(size_t)prow * 1024 + pn0 + part * ncol + c] = tot;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace tc16

// sweeps (activation rows k-contiguous, no split-K) whose width splits into four column quarters of 16-column chunks
template <bool B_KC>
inline bool tc16_eligible(const GemmArgs& g, int nsplit) {
  return tc_eligible<true, B_KC>(g, nsplit) && nsplit == 1 && !g.kchunk && g.N % 64 == 0;
}

template <bool B_KC, int SPLIT, class Epi, int BN>
inline cudaError_t launch_gemm_tc16_bn(const GemmArgs& g, const Epi& epi, int num_sms, cudaStream_t st) {
  constexpr bool B_MN = !B_KC;
  using C = tc16::Cfg16<SPLIT, BN>;
  tc::TmSet tm;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok = tc::make_map(&tm.a[s], sg.A, sg.K, g.M, sg.lda, 32, tc::BM, false);
    if (B_MN) ok = ok && tc::make_map(&tm.b[s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);
    else      ok = ok && tc::make_map(&tm.b[s], sg.B, sg.K, g.Nb, sg.ldb, 32, g.N < BN ? g.N : BN, false);
    if (!ok) return cudaErrorInvalidValue;
  }
  for (int s = g.nseg; s < kMaxSeg; ++s) tm.a[s] = tm.a[0], tm.b[s] = tm.b[0];
  auto kern = tc16::gemm_tc16_kernel<B_MN, SPLIT, BN, Epi>;
  static unsigned long long attr_devs = 0;   // per device: opt in to the large dynamic shared memory once
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int mtiles = (g.M + tc::BM - 1) / tc::BM;
  const int work = mtiles * ((g.N + BN - 1) / BN);
  const int grid = work < num_sms ? work : num_sms;
  return tc::launch_pdl(kern, grid, C::NUM_THREADS, C::SMEM_BYTES, st, 1, tm, g, epi, mtiles);
}

template <bool B_KC, int SPLIT, class Epi>
inline cudaError_t launch_gemm_tc16(const GemmArgs& g, const Epi& epi, int num_sms, cudaStream_t st) {
  const int bn = tc_pick_bn<true, SPLIT>(g, 1, num_sms);
  if (bn == 64) return launch_gemm_tc16_bn<B_KC, SPLIT, Epi, 64>(g, epi, num_sms, st);
  if (bn == 128) return launch_gemm_tc16_bn<B_KC, SPLIT, Epi, 128>(g, epi, num_sms, st);
  return launch_gemm_tc16_bn<B_KC, SPLIT, Epi, 256>(g, epi, num_sms, st);
}

}  // namespace fbsnn
