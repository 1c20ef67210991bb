// tcgen05 dense-layer GEMM (FBSNN_PREC_TF32): kind::tf32 UMMA, fp32 accumulators in TMEM, TMA-staged operands.
//
// Same contract as gemm_simt.cuh (K-concatenated segments, operand-major flags, split-K, fused sweep epilogues)
// for shapes that are genuine dense GEMMs:  N in {64,128,192,256}, every K a multiple of 32 (sweeps), leading
// dimensions multiples of 4 floats.  Everything else stays on the SIMT kernel.
//
// CTA = 10 warps, persistent (one CTA per SM, static round-robin over 128-row output tiles):
//   warp 0     TMA producer: cp.async.bulk.tensor 128B-swizzled boxes into a 4-stage ring (48 KB per stage)
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32, M=128, N<=256, K=8;
//              owns the 512-column TMEM allocation = two accumulator buffers (epilogue of tile i overlaps MMA of i+1)
//   warps 2-9  epilogue: tcgen05.ld 32x32b (lane quarter = warp_id % 4, column half = (warp_id-2)/4), then the same
//              sweep epilogue functors as the SIMT kernel, reading/writing the row arrays with 16-byte accesses.
// Operand layouts in shared memory are the canonical UMMA SWIZZLE_128B ones:
//   K-major  (activation rows, weights used as W^T): rows x 32 tf32 (128 B), SBO = 1024 B, k-step = +32 B
//   MN-major (weights used as W, weight-gradient operands): 32-bit MN-major operands only exist in the
//             "128B swizzle, 32B atom" mode (TMA SWIZZLE_128B_ATOM_32B, descriptor layout type 1): chunks of
//             [32 k-rows x 32 elements], LBO = 4096 B between 32-element chunks, SBO = 512 B between 4-row groups,
//             k-step (8 rows) = +1024 B
#pragma once
#include <cuda.h>

#include <type_traits>

#include "gemm_simt.cuh"

namespace fbsnn {
namespace tc {

constexpr int BM = 128, BK = 32, STAGES = 3, UMMA_K = 8;
constexpr int A_STAGE_BYTES = BM * BK * 4;       // 16 KB
constexpr int B_STAGE_BYTES = 256 * BK * 4;      // 32 KB (N <= 256)
constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_LD = 36;                                      // padded row of the per-warp 32x32 transpose tile
constexpr int EPI_TILE_BYTES = 32 * EPI_LD * 4;                 // 4.5 KB per epilogue warp
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + NUM_EPI_WARPS * EPI_TILE_BYTES +
                           1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t kSpinLimit = 1u << 24;       // bounded waits: a protocol bug traps instead of hanging the GPU

struct TmSet {
  CUtensorMap a[4];
  CUtensorMap b[4];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  const uint32_t addr = smem_u32(b);
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > kSpinLimit) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// shared-memory matrix descriptor, Blackwell version bit set.  layout: 2 = SWIZZLE_128B (K-major operands),
// 1 = SWIZZLE_128B_BASE32B (the only layout for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)layout << 61;
  return d;
}
// instruction descriptor: D = f32, A = B = tf32, M = 128
__host__ __device__ __forceinline__ uint32_t make_idesc(int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 2u << 7;                       // a_format = TF32
  d |= 2u << 10;                      // b_format = TF32
  d |= (a_mn ? 1u : 0u) << 15;        // a_major: 0 = K, 1 = MN
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

#define FBSNN_TMEM_LD32(taddr, v)                                                                              \
  asm volatile(                                                                                                \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),        \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),  \
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
      : "r"(taddr))

// ----------------------------------------------------------------------------------------------------
template <bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TmSet tm, const GemmArgs g, const Epi epi, const int num_mtiles, const int nsplit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  float* epi_tiles = (float*)(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* bars = (uint64_t*)(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + NUM_EPI_WARPS * EPI_TILE_BYTES);
  uint64_t* full = bars;               // [STAGES]
  uint64_t* empty = bars + STAGES;     // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES; // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = g.N;
  const int num_work = num_mtiles * nsplit;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b[s]) : "memory");
    }
    for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1), mbar_init(&tempty[i], NUM_EPI_WARPS);
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

  // k-range of one work item (split-K only for the weight-gradient contraction, where K = rows)
  auto kbeg = [&](int s, int split) { return g.kchunk ? (int)min((long long)g.seg[s].K, (long long)split * g.kchunk) : 0; };
  auto kend = [&](int s, int split) {
    return g.kchunk ? (int)min((long long)g.seg[s].K, ((long long)split + 1) * g.kchunk) : g.seg[s].K;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = A_STAGE_BYTES + (uint32_t)N * BK * 4;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int mt = w % num_mtiles, split = w / num_mtiles;
        const int m0 = mt * BM;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], bytes);
            uint8_t* a = sA + stage * A_STAGE_BYTES;
            uint8_t* b = sB + stage * B_STAGE_BYTES;
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < BM / 32; ++c) tma_load_2d(a + c * 4096, &tm.a[s], &full[stage], m0 + 32 * c, k0);
            } else {
              tma_load_2d(a, &tm.a[s], &full[stage], k0, m0);
            }
            if (B_MN) {
              for (int c = 0; c < N / 32; ++c) tma_load_2d(b + c * 4096, &tm.b[s], &full[stage], 32 * c, k0);
            } else {
              tma_load_2d(b, &tm.b[s], &full[stage], k0, 0);
            }
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(N, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const int split = w / num_mtiles;
        const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a = smem_u32(sA + stage * A_STAGE_BYTES);
            const uint32_t b = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = A_MN ? make_desc(a + k * 1024, 4096, 512, 1) : make_desc(a + k * 32, 16, 1024, 2);
              const uint64_t bdesc = B_MN ? make_desc(b + k * 1024, 4096, 512, 1) : make_desc(b + k * 32, 16, 1024, 2);
              umma_tf32(tmem_d, adesc, bdesc, idesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit(&empty[stage]);   // smem slot free once these MMAs have read it
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);         // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // TMEM hands each lane one accumulator ROW; the row arrays in global memory want a warp on one row's
    // contiguous columns.  Each 32x32 chunk is therefore transposed through a warp-private padded smem tile, after
    // which 8 lanes cover 128 B of one row and a warp-wide access touches 4 full cache lines instead of 32 partial.
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;      // column half
    const int ncol = N >> 1;
    float* tile = epi_tiles + (warp - 2) * (32 * EPI_LD);
    const int sub = lane >> 3, cc = (lane & 7) * 4;
    uint32_t it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int mt = w % num_mtiles;
      const int split = w / num_mtiles;
      (void)split;
      const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      const int r0 = mt * BM + q * 32;
      for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + c0;
        FBSNN_TMEM_LD32(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4(tile + lane * EPI_LD + 4 * j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
        __syncwarp();
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) {
          typename Epi::Frag f[4];
          float4 a4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (ib * 4 + i) * 4 + sub;
            if (r0 + rr < g.M) f[i] = epi.prefetch(r0 + rr, c0 + cc);
            a4[i] = ld4(tile + rr * EPI_LD + cc);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (ib * 4 + i) * 4 + sub;
            if (r0 + rr < g.M) {
              if constexpr (std::is_same<Epi, EpiPartial>::value) epi.finish_split(split, r0 + rr, c0 + cc, a4[i]);
              else epi.finish(r0 + rr, c0 + cc, a4[i], f[i]);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D fp32 tensor map over X[outer x inner] (inner contiguous, leading dimension ld), 128-byte swizzle.
inline bool make_map(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_inner,
                     int box_outer, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

}  // namespace tc

template <bool A_KC, bool B_KC>
inline bool tc_eligible(const GemmArgs& g, int nsplit) {
  if (g.N < 64 || g.N > 256 || g.N % 64 || g.nseg < 1 || g.nseg > 4) return false;
  if (!A_KC && B_KC) return false;   // (MN-major A, K-major B) never occurs in the sweeps
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    if (sg.lda % 4 || sg.ldb % 4 || ((uintptr_t)sg.A & 15) || ((uintptr_t)sg.B & 15)) return false;
    if (A_KC && sg.K % 32) return false;   // sweeps: K is a feature width; the weight-gradient K (= rows) is free
  }
  if (g.kchunk && g.kchunk % 32) return false;
  if (!g.kchunk && nsplit != 1) return false;
  return true;
}

template <bool A_KC, bool B_KC, class Epi>
inline cudaError_t launch_gemm_tc(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  constexpr bool A_MN = !A_KC, B_MN = !B_KC;
  tc::TmSet tm;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok;
    if (A_MN) ok = tc::make_map(&tm.a[s], sg.A, g.M, sg.K, sg.lda, 32, 32, true);       // P[k = rows][m]
    else      ok = tc::make_map(&tm.a[s], sg.A, sg.K, g.M, sg.lda, 32, tc::BM, false);  // X[m = rows][k]
    if (B_MN) ok = ok && tc::make_map(&tm.b[s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);    // W[k][n] / Q[k = rows][n]
    else      ok = ok && tc::make_map(&tm.b[s], sg.B, sg.K, g.Nb, sg.ldb, 32, g.N, false);  // W[n][k]
    if (!ok) return cudaErrorInvalidValue;
  }
  for (int s = g.nseg; s < 4; ++s) tm.a[s] = tm.a[0], tm.b[s] = tm.b[0];
  auto kern = tc::gemm_tc_kernel<A_MN, B_MN, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int mtiles = (g.M + tc::BM - 1) / tc::BM;
  const int work = mtiles * nsplit;
  const int grid = work < num_sms ? work : num_sms;
  kern<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(tm, g, epi, mtiles, nsplit);
  return cudaGetLastError();
}

}  // namespace fbsnn
