// Weight-gradient contraction of the 3xTF32 variant on a CTA pair with the A operand in TENSOR MEMORY.
//
//   C[256 x N] (+)= sum over segments  P_s^T[256 x K] * Q_s[K x N],   K = rows of one split-K chunk,
//   3xTF32:  P_lo^T Q + P^T Q_lo + P^T Q   (hi parts = what the tensor core reads from the fp32 container)
//
// gemm_tc2_kernel<SPLIT = 2> stages A, A_lo, B, B_lo in shared memory; every k-block then costs a CTA 32 KB of TMA
// writes + 64 KB of splitter traffic + 12 MMAs x 8 KB of operand reads = 192 KB against 128 B/clk x 1536 MMA cycles
// = 196 KB: shared memory, not the tensor pipe, sets the pace (54-59 % tensor pipe measured).  Here the A operand
// never becomes an MMA operand in shared memory: four warps read the freshly landed P tile (plain row-major, no
// swizzle -- it is not an UMMA operand), and write P and P_lo straight into tensor memory with tcgen05.st (lane = output
// row m, column = k); the MMAs take A from TMEM (tcgen05.mma [d], [a_tmem], b_desc) and read only the B half from
// shared memory: 32 KB TMA + 16 KB A reads + 32 KB B split + 12 x 4 KB = 128 KB per k-block.
//
// TMEM map (512 columns): accumulator 256 x N fp32 -> columns [0, 256) of both CTAs (128 rows each); A staging:
// stage s -> P in columns [256 + 64 s, +32), P_lo in [256 + 64 s + 32, +32), four stages.
// Protocol as gemm_tc2.cuh: leader-issued cta_group::2 MMAs, multicast commits, remote arrivals on the leader's
// barriers (one per CTA and k-block, after a named barrier of the splitter team); one accumulator buffer (a cluster
// has one or few split-K work items).
#pragma once
#include "gemm_tc2.cuh"

namespace fbsnn {
namespace tc2g {

using namespace tc2;

constexpr int STAGES_G = 4;
constexpr int A_ROWMAJOR_BYTES = BK * 128 * 4;                     // [32 k][128 m] fp32, no swizzle
constexpr int STAGE_G_BYTES = A_ROWMAJOR_BYTES + 2 * B_HALF_BYTES;   // [A 16K][B 16K][B_lo 16K]
constexpr int A_WARPS = 4, B_WARPS = 8;                           // splitter team: warps 2..5 -> A, warps 6..13 -> B
constexpr int NUM_THREADS_G = 32 * (2 + A_WARPS + B_WARPS);
constexpr int SMEM_G_BYTES = STAGES_G * STAGE_G_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES + 1024 + 256;

#define FBSNN_TMEM_ST32(taddr, v)                                                                              \
  asm volatile(                                                                                                \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                          \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                               \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                       \
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),    \
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),          \
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),        \
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])         \
      : "memory")

// D[tmem] (+)= A[tmem] * B[smem], pair form
__device__ __forceinline__ void umma_tf32_pair_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(z)
      : "memory");
}

// body of one CTA of cluster `cid` of `ncl` clusters working on the split-K chunks of ONE contraction (tma / tmb = its
// tensor maps, in kernel-parameter space)
template <class Epi>
__device__ __forceinline__ void tc2g_body(const CUtensorMap* tma, const CUtensorMap* tmb, const GemmArgs& g, const Epi& epi,
                                          const int nsplit, const int cid, const int ncl) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_tiles = (float*)(smem + STAGES_G * STAGE_G_BYTES);
  uint64_t* bars = (uint64_t*)(smem + STAGES_G * STAGE_G_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES);
  uint64_t* full = bars;               // [4] local
  uint64_t* empty = bars + 4;          // [4] local (multicast commit)
  uint64_t* sdone = bars + 8;          // [4] leader's copy used
  uint64_t* tfull = bars + 12;         // [1] local (multicast commit)
  uint64_t* tempty = bars + 13;        // [1] leader's copy used
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int N = g.N, NH = N >> 1;
  const int num_work = nsplit;         // one 256 x N output per split-K chunk

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tma[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmb[s]) : "memory");
    }
    for (int i = 0; i < STAGES_G; ++i)
      mbar_init(&full[i], 1), mbar_init(&empty[i], 1), mbar_init(&sdone[i], 2);   // one arrival per CTA
    mbar_init(&tfull[0], 1), mbar_init(&tempty[0], 2 * NUM_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  auto kbeg = [&](int s, int split) { return (int)min((long long)g.seg[s].K, (long long)split * g.kchunk); };
  auto kend = [&](int s, int split) { return (int)min((long long)g.seg[s].K, ((long long)split + 1) * g.kchunk); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = A_ROWMAJOR_BYTES + (uint32_t)NH * BK * 4;
      const int m0 = 128 * (int)rank, n0 = NH * (int)rank;
      for (int w = cid; w < num_work; w += ncl) {
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, w), ke = kend(s, w); k0 < ke; k0 += BK) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], bytes);
            uint8_t* a = smem + stage * STAGE_G_BYTES;
            uint8_t* b = a + A_ROWMAJOR_BYTES;
            tma_load_2d(a, &tma[s], &full[stage], m0, k0);                          // P[k0.., m0 .. m0+128), row-major
            for (int c = 0; c < NH / 32; ++c) tma_load_2d(b + c * 4096, &tmb[s], &full[stage], n0 + 32 * c, k0);
            if (++stage == STAGES_G) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_pair(N, false, true);   // A from TMEM is k-major by construction; B MN-major
      uint32_t stage = 0, phase = 0, it = 0;
      for (int w = cid; w < num_work; w += ncl, ++it) {
        mbar_wait_cluster(&tempty[0], (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base;
        uint32_t first = 1;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, w), ke = kend(s, w); k0 < ke; k0 += BK) {
            mbar_wait_cluster(&sdone[stage], phase);
            tc_fence_after();
            const uint32_t b = smem_u32(smem + stage * STAGE_G_BYTES) + A_ROWMAJOR_BYTES;
            const uint32_t blo = b + B_HALF_BYTES;
            const uint32_t ta = tmem_base + 256 + 64 * stage, talo = ta + 32;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t db = make_desc(b + k * 1024, 4096, 512, 1), dblo = make_desc(blo + k * 1024, 4096, 512, 1);
              umma_tf32_pair_ts(tmem_d, talo + 8 * k, db, idesc, first ? 0u : 1u);
              first = 0;
              umma_tf32_pair_ts(tmem_d, ta + 8 * k, dblo, idesc, 1u);
              umma_tf32_pair_ts(tmem_d, ta + 8 * k, db, idesc, 1u);
            }
            umma_commit_pair(&empty[stage]);
            if (++stage == STAGES_G) stage = 0, phase ^= 1;
          }
        }
        umma_commit_pair(&tfull[0]);
      }
    }
  } else {
    // ===================== splitter team, then epilogue =====================
    const bool is_epi = warp >= 2 + A_WARPS;     // the eight B-split warps also run the (tiny) epilogue
    const int e = warp - (2 + A_WARPS);
    const int q = warp & 3;
    const int half = e >> 2;
    const int ncol = N >> 1;
    float* tile = epi_tiles + (is_epi ? e : 0) * (EPI_TILE_BYTES / 4);
    const int sub = lane >> 3, c4 = lane & 7;
    uint32_t sstage = 0, sphase = 0;
    uint32_t it = 0;
    for (int w = cid; w < num_work; w += ncl, ++it) {
      for (int s = 0; s < g.nseg; ++s) {
        for (int k0 = kbeg(s, w), ke = kend(s, w); k0 < ke; k0 += BK) {
          mbar_wait(&full[sstage], sphase);
          uint8_t* st = smem + sstage * STAGE_G_BYTES;
          if (!is_epi) {
            // A: output row m = 32 q + lane; its 32 k-values -> TMEM lane m, columns k (P) and 32 + k (P_lo)
            const float* arow = (const float*)st + 32 * q + lane;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float x = arow[k * 128];
              hi[k] = __float_as_uint(x);
              lo[k] = __float_as_uint(x - __uint_as_float(hi[k] & 0xFFFFE000u));
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + 64 * sstage;
            FBSNN_TMEM_ST32(taddr, hi);
            FBSNN_TMEM_ST32(taddr + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
          } else {
            // B half: lo = x - trunc_tf32(x), same (swizzled) layout as the tile itself
            const float4* b = (const float4*)(st + A_ROWMAJOR_BYTES);
            float4* blo = (float4*)(st + A_ROWMAJOR_BYTES + B_HALF_BYTES);
            const int tb = threadIdx.x - 32 * (2 + A_WARPS);
            const int nB4 = NH * BK * 4 / 16;
            for (int i = tb; i < nB4; i += 32 * B_WARPS) {
              const float4 x = b[i];
              float4 l;
              l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
              l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
              l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
              l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
              blo[i] = l;
            }
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          // the team's twelve warps meet on a named barrier and ONE thread signals the leader: a cluster-scope
          // release arrive costs a gpu-wide fence (ncu: 30 % of this kernel's stall samples when every warp did it)
          asm volatile("bar.sync 2, %0;" ::"n"(32 * (A_WARPS + B_WARPS)) : "memory");
          if (threadIdx.x == 64) mbar_arrive_cta(&sdone[sstage], 0);
          if (++sstage == STAGES_G) sstage = 0, sphase ^= 1;
        }
      }
      if (is_epi) {
        // drain this CTA's 128 accumulator rows
        mbar_wait(&tfull[0], it & 1);
        tc_fence_after();
        const int r0 = 128 * (int)rank + q * 32;
#pragma unroll 1
        for (int ch = 0; ch * 32 < ncol; ++ch) {
          const int ct = half * ncol + ch * 32;
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ct;
          FBSNN_TMEM_LD32(taddr, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st4(tile + lane * 32 + ((j ^ (lane & 7)) << 2),
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3])));
          __syncwarp();
          const int cc = c4 * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + sub;
            const float4 a4 = ld4(tile + rr * 32 + ((c4 ^ (rr & 7)) << 2));
            if (r0 + rr < g.M) {
              if constexpr (std::is_same<Epi, EpiPartial>::value) epi.finish_split(w, r0 + rr, ct + cc, a4);
              else epi.finish(r0 + rr, ct + cc, a4, typename Epi::Frag{}, epi.col_prefetch(ct + cc));
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&tempty[0], 0);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <class Epi>
__global__ void __launch_bounds__(NUM_THREADS_G, 1)
gemm_tc2g_kernel(const __grid_constant__ TmSet tm, const GemmArgs g, const Epi epi, const int nsplit) {
  tc2g_body(tm.a, tm.b, g, epi, nsplit, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1));
}

// Several contractions in ONE launch (blockIdx.y = job): the weight gradients of all layers of a small-batch step, each of
// which alone would be a launch of a few k-blocks per CTA (M = 100: 4 launches of 14-17 us, mostly prologue and drain).
constexpr int kMaxJobsG = 8;
struct BatchG {
  CUtensorMap a[kMaxJobsG][2], b[kMaxJobsG][2];   // two segments per contraction
  GemmArgs g[kMaxJobsG];
  EpiPartial epi[kMaxJobsG];
  int njobs, nsplit;
};
__global__ void __launch_bounds__(NUM_THREADS_G, 1)
gemm_tc2g_batched_kernel(const __grid_constant__ BatchG bt) {
  const int j = blockIdx.y;
  tc2g_body(bt.a[j], bt.b[j], bt.g[j], bt.epi[j], bt.nsplit, (int)(blockIdx.x >> 1), (int)(gridDim.x >> 1));
}

// 2-D fp32 tensor map without swizzle (the P tile is read by threads, not by the tensor core)
inline bool make_map_plain(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld,
                           int box_inner, int box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

}  // namespace tc2g

// weight-gradient shape (both operands rows x features, out = 256) with split-K
inline bool tc2g_eligible(const GemmArgs& g, int nsplit) {
  return tc2_eligible<false, false>(g, nsplit) && g.kchunk > 0 && g.M == 256;
}

template <class Epi>
inline cudaError_t launch_gemm_tc2g(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  tc::TmSet tm;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok = tc2g::make_map_plain(&tm.a[s], sg.A, g.M, sg.K, sg.lda, 128, tc::BK);       // P[k = rows][m]
    ok = ok && tc::make_map(&tm.b[s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);            // Q[k = rows][n]
    if (!ok) return cudaErrorInvalidValue;
  }
  for (int s = g.nseg; s < kMaxSeg; ++s) tm.a[s] = tm.a[0], tm.b[s] = tm.b[0];
  auto kern = tc2g::gemm_tc2g_kernel<Epi>;
  static unsigned long long attr_devs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2g::SMEM_G_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int ncl = std::min(nsplit, num_sms / 2);
  return tc::launch_pdl(kern, 2 * ncl, tc2g::NUM_THREADS_G, tc2g::SMEM_G_BYTES, st, 2, tm, g, epi, nsplit);
}

// adds one contraction to a batch; false if it does not fit (caller launches it on its own)
inline bool tc2g_batch_add(tc2g::BatchG& bt, const GemmArgs& g, const EpiPartial& epi, int nsplit) {
  if (bt.njobs >= tc2g::kMaxJobsG || g.nseg > 2 || (bt.njobs > 0 && bt.nsplit != nsplit)) return false;
  const int j = bt.njobs;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok = tc2g::make_map_plain(&bt.a[j][s], sg.A, g.M, sg.K, sg.lda, 128, tc::BK);
    ok = ok && tc::make_map(&bt.b[j][s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);
    if (!ok) return false;
  }
  for (int s = g.nseg; s < 2; ++s) bt.a[j][s] = bt.a[j][0], bt.b[j][s] = bt.b[j][0];
  bt.g[j] = g, bt.epi[j] = epi, bt.nsplit = nsplit;
  ++bt.njobs;
  return true;
}
inline cudaError_t launch_gemm_tc2g_batched(const tc2g::BatchG& bt, int num_sms, cudaStream_t st) {
  auto kern = tc2g::gemm_tc2g_batched_kernel;
  static unsigned long long attr_devs = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2g::SMEM_G_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int ncl = std::max(1, std::min(bt.nsplit, num_sms / 2 / bt.njobs));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ncl, bt.njobs, 1);
  cfg.blockDim = dim3(tc2g::NUM_THREADS_G, 1, 1);
  cfg.dynamicSmemBytes = tc2g::SMEM_G_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  attr[n].id = cudaLaunchAttributeClusterDimension;
  attr[n].val.clusterDim.x = 2, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
  ++n;
  if (tc::pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, bt);
}

}  // namespace fbsnn
