// tcgen05 dense-layer GEMM (FBSNN_PREC_TF32): kind::tf32 UMMA, fp32 accumulators in TMEM, TMA-staged operands.
//
// Same contract as gemm_simt.cuh (K-concatenated segments, operand-major flags, split-K, fused sweep epilogues)
// for shapes that are genuine dense GEMMs:  N in {64,128,192,256}, every K a multiple of 32 (sweeps), leading
// dimensions multiples of 4 floats.  Everything else stays on the SIMT kernel.
//
// CTA = 10 warps (14 with SPLIT3), persistent (one CTA per SM, static round-robin over 128-row output tiles):
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

#include <cstdlib>

#include <type_traits>

#include "gemm_simt.cuh"

namespace fbsnn {
namespace tc {

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int A_STAGE_BYTES = BM * BK * 4;       // 16 KB
constexpr int B_STAGE_BYTES = 256 * BK * 4;      // 32 KB (N <= 256)
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_TILE_BYTES = 32 * 32 * 4;      // per-warp 32x32 fp32 transpose tile (XOR-swizzled, no padding)
constexpr uint32_t kSpinLimit = 1u << 24;       // bounded waits: a protocol bug traps instead of hanging the GPU

struct TmSet {
  CUtensorMap a[kMaxSeg];
  CUtensorMap b[kMaxSeg];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Programmatic dependent launch: the dense-layer kernels are launched with programmatic stream serialisation, so
// their prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the tail of the previous kernel on
// idle SMs -- what bounds the 37-launch M = 100 step.  pdl_wait() (= cudaGridDependencySynchronize) returns once the
// previous kernel has completed and flushed; pdl_trigger() lets the NEXT kernel start launching early (it still
// waits for this one to finish at its own pdl_wait).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("FBSNN_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}
template <class Kern, class... Args>
inline cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(block, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster, attr[n].val.clusterDim.y = 1, attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
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
// SPLIT > 0 = fp32-grade accuracy on the tensor cores ("3xTF32"): an operand x is used as hi = trunc_tf32(x) (what
// the UMMA reads from the fp32 container anyway) and lo = x - hi; the products lo*hi + hi*lo + hi*hi are accumulated
// in TMEM, the dropped lo*lo term and the truncation of lo are ~2^-21 relative.  Four splitter warps write the lo
// tiles into extra shared memory as soon as the TMA lands.  The kernel stays HBM-bound (the tensor pipe was 13-39 %
// busy with one MMA per k-step); what 3xTF32 costs is shared memory, i.e. pipeline depth:
//   SPLIT = 1 (sweeps): the weight operand arrives pre-split from global memory as two exact-TF32 matrices W_hi,
//              W_lo, used as two K-concatenated segments (A, W_hi) [2 MMAs] and (A, W_lo) [1 MMA]; only A_lo needs
//              a tile -> 64 KB per stage, 3 stages.  The A tile is fetched twice (second time from L2).
//   SPLIT = 2 (weight gradients, both operands are row arrays): A_lo and B_lo tiles -> 96 KB per stage, 2 stages.
// BN = widest output tile of a CTA (256, or 128 for the SPLIT == 2 weight-gradient kernel: with 128 columns a
// stage is 64 KB and three fit, which the serial TMA -> split -> MMA chain of that kernel needs; the two column
// halves are separate work items that share their operand rows through L2).
template <int SPLIT, int BN = 256>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 4;
  // SPLIT == 3 (sweeps no wider than 128 columns, e.g. Du): W_hi / W_lo twins (seg[s].B / seg[s + 4].B) land in the
  // B / B_lo slots of ONE stage, only A is split in-kernel and the A tile is fetched once instead of twice
  static constexpr int STAGES = (SPLIT >= 2 && BN == 256) ? 2 : 3;
  static constexpr int SPLIT_WARPS = SPLIT ? 4 : 0;
  static constexpr int EPI_WARP0 = 2 + SPLIT_WARPS;
  // SPLIT == 2 (weight gradients): the epilogue warps are idle during the long K loop of a work item, so they
  // join the splitters (12 warps) and run the (tiny) epilogue afterwards
  static constexpr int SPLIT_TEAM_WARPS = SPLIT == 2 ? SPLIT_WARPS + 8 : SPLIT_WARPS;
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES + (SPLIT >= 1 ? A_STAGE_BYTES : 0) +
                                     (SPLIT >= 2 ? B_BYTES : 0);   // [A][B]([A_lo]([B_lo]))
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES + 1024 /*align*/ + 256;
};

template <bool A_MN, bool B_MN, int SPLIT, int BN, class Epi>
__global__ void __launch_bounds__(Cfg<SPLIT, BN>::NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TmSet tm, const GemmArgs g, const Epi epi, const int num_mtiles, const int nsplit) {
  using C = Cfg<SPLIT, BN>;
  constexpr int B_STAGE_BYTES = C::B_BYTES;
  constexpr int STAGES = C::STAGES;
  constexpr bool SPLIT3 = SPLIT > 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage layout: [A 16K][B 32K]([A_lo 16K]([B_lo 32K]))
  float* epi_tiles = (float*)(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C::STAGE_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES);
  uint64_t* full = bars;               // [STAGES]  TMA bytes landed
  uint64_t* empty = bars + 3;          // [STAGES]  MMAs that read the stage have completed
  uint64_t* sdone = bars + 6;          // [STAGES]  lo tiles written (SPLIT3)
  uint64_t* tfull = bars + 9;          // [2]       accumulator complete
  uint64_t* tempty = bars + 11;        // [2]       accumulator drained by the epilogue
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = g.N < BN ? g.N : BN;                    // columns of one CTA tile
  const int num_ntiles = (g.N + BN - 1) / BN;
  const int num_mn = num_mtiles * num_ntiles;           // work id = (split, nt, mt), mt fastest
  const int num_work = num_mn * nsplit;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b[s]) : "memory");
    }
    for (int i = 0; i < STAGES; ++i)
      mbar_init(&full[i], 1), mbar_init(&empty[i], 1), mbar_init(&sdone[i], C::SPLIT_TEAM_WARPS);
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
  pdl_trigger();
  pdl_wait();   // everything above overlapped the previous kernel; its outputs are read only from here on

  // k-range of one work item (split-K only for the weight-gradient contraction, where K = rows)
  auto kbeg = [&](int s, int split) { return g.kchunk ? (int)min((long long)g.seg[s].K, (long long)split * g.kchunk) : 0; };
  auto kend = [&](int s, int split) {
    return g.kchunk ? (int)min((long long)g.seg[s].K, ((long long)split + 1) * g.kchunk) : g.seg[s].K;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = A_STAGE_BYTES + (uint32_t)N * BK * 4 * (SPLIT == 3 ? 2 : 1);
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int mt = w % num_mtiles, nt = (w % num_mn) / num_mtiles, split = w / num_mn;
        const int m0 = mt * BM, n0 = nt * BN;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], bytes);
            uint8_t* a = smem + stage * C::STAGE_BYTES;
            uint8_t* b = a + A_STAGE_BYTES;
            if (A_MN) {
#pragma unroll
              for (int c = 0; c < BM / 32; ++c) tma_load_2d(a + c * 4096, &tm.a[s], &full[stage], m0 + 32 * c, k0);
            } else {
              tma_load_2d(a, &tm.a[s], &full[stage], k0, m0);
            }
            if (B_MN) {
              for (int c = 0; c < N / 32; ++c) tma_load_2d(b + c * 4096, &tm.b[s], &full[stage], n0 + 32 * c, k0);
            } else {
              tma_load_2d(b, &tm.b[s], &full[stage], k0, n0);
            }
            if (SPLIT == 3) {   // W_lo twin -> the B_lo slot
              uint8_t* bl = b + B_STAGE_BYTES + A_STAGE_BYTES;
              if (B_MN) {
                for (int c = 0; c < N / 32; ++c) tma_load_2d(bl + c * 4096, &tm.b[s + 4], &full[stage], n0 + 32 * c, k0);
              } else {
                tma_load_2d(bl, &tm.b[s + 4], &full[stage], k0, n0);
              }
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
        const int split = w / num_mn;
        const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        uint32_t first = 1;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
            mbar_wait(SPLIT3 ? &sdone[stage] : &full[stage], phase);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t b = a + A_STAGE_BYTES;
            const uint32_t alo = a + A_STAGE_BYTES + B_STAGE_BYTES, blo = alo + A_STAGE_BYTES;
            const int mode = SPLIT >= 2 ? 3 : (SPLIT == 1 ? g.mode[s] : 0);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              auto da = [&](uint32_t base) { return A_MN ? make_desc(base + k * 1024, 4096, 512, 1) : make_desc(base + k * 32, 16, 1024, 2); };
              auto db = [&](uint32_t base) { return B_MN ? make_desc(base + k * 1024, 4096, 512, 1) : make_desc(base + k * 32, 16, 1024, 2); };
              if (mode & 1) {   // modes 1, 3: a_lo * b
                umma_tf32(tmem_d, da(alo), db(b), idesc, first ? 0u : 1u);
                first = 0;
              }
              if (SPLIT >= 2) umma_tf32(tmem_d, da(a), db(blo), idesc, 1u);
              umma_tf32(tmem_d, da(a), db(b), idesc, first ? 0u : 1u);
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
    // ===================== splitter and epilogue warps =====================
    // Epilogue: TMEM hands each lane one accumulator ROW; the row arrays in global memory want a warp on one row's
    // contiguous columns.  Each 32x32 chunk is therefore transposed through a warp-private XOR-swizzled smem tile,
    // after which 8 lanes cover 128 B of one row and a warp-wide access touches 4 full cache lines instead of 32.
    const bool is_epi = warp >= C::EPI_WARP0;
    const int e = warp - C::EPI_WARP0;
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int half = e >> 2;               // column half
    const int ncol = N >> 1;
    float* tile = epi_tiles + (is_epi ? e : 0) * (EPI_TILE_BYTES / 4);
    const int sub = lane >> 3, c4 = lane & 7;
    constexpr bool kColsum = Epi::kColsum;
    float4 csum[4];                        // per-thread column sums of zbar: chunk x 4 columns (N/2 <= 128)
#pragma unroll
    for (int i = 0; i < 4; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // drain the accumulator of work item w (local index it) through the fused epilogue
    auto drain = [&](int w, uint32_t it) {
      const int mt = w % num_mtiles;
      const int n0 = ((w % num_mn) / num_mtiles) * BN;
      const int split = w / num_mn;
      (void)split;
      const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      const int r0 = mt * BM + q * 32;
#pragma unroll 1
      for (int ch = 0; ch * 32 < ncol; ++ch) {
        const int ct = half * ncol + ch * 32;   // column inside the CTA tile (TMEM column)
        const int c0 = n0 + ct;                 // global output column
        if (c4 == 0) {
          // the lines this warp's loads touch in the NEXT chunk (or in the first chunk of its next tile) -> L2 now
          const bool more = (ch + 1) * 32 < ncol;
          const int wn = w + gridDim.x;
          if (more || wn < num_work) {
            const int pr0 = more ? r0 : (wn % num_mtiles) * BM + q * 32;
            const int pc0 = more ? c0 + 32 : ((wn % num_mn) / num_mtiles) * BN + half * ncol;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (pr0 + 4 * i + sub < g.M) epi.l2_prefetch(pr0 + 4 * i + sub, pc0);
          }
        }
        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + ct;
        FBSNN_TMEM_LD32(taddr, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st4(tile + lane * 32 + ((j ^ (lane & 7)) << 2),
              make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                          __uint_as_float(v[4 * j + 3])));
        __syncwarp();
        const int cc = c4 * 4;
        const typename Epi::ColFrag cf = epi.col_prefetch(c0 + cc);   // biases / output weights of these 4 columns
#pragma unroll
        for (int ib = 0; ib < 2; ++ib) {
          typename Epi::Frag f[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (ib * 4 + i) * 4 + sub;
            if (r0 + rr < g.M) f[i] = epi.prefetch(r0 + rr, c0 + cc);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (ib * 4 + i) * 4 + sub;
            const float4 a4 = ld4(tile + rr * 32 + ((c4 ^ (rr & 7)) << 2));
            if (r0 + rr < g.M) {
              if constexpr (std::is_same<Epi, EpiPartial>::value) {
                epi.finish_split(split, r0 + rr, c0 + cc, a4);
              } else if constexpr (kColsum) {
                const float4 zb = epi.finish(r0 + rr, c0 + cc, a4, f[i], cf);
                cs.x += zb.x, cs.y += zb.y, cs.z += zb.z, cs.w += zb.w;
              } else {
                epi.finish(r0 + rr, c0 + cc, a4, f[i], cf);
              }
            }
          }
        }
        __syncwarp();
        if constexpr (kColsum) {   // rolled loop: select the chunk's accumulator with predicated adds (no local memory)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ch == k) csum[k].x += cs.x, csum[k].y += cs.y, csum[k].z += cs.z, csum[k].w += cs.w;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    };

    // lo = x - trunc_tf32(x) for the operand tiles of every k-block of work item w (SPLIT > 0)
    constexpr int TEAM = 32 * (C::SPLIT_TEAM_WARPS > 0 ? C::SPLIT_TEAM_WARPS : 1);
    uint32_t sstage = 0, sphase = 0;
    auto split_item = [&](int w) {
      const int ts = threadIdx.x - 64;
      const int nA4 = A_STAGE_BYTES / 16, nB4 = N * BK * 4 / 16;
      const int split = w / num_mn;
      for (int s = 0; s < g.nseg; ++s) {
        for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
          mbar_wait(&full[sstage], sphase);
          const float4* a = (const float4*)(smem + sstage * C::STAGE_BYTES);
          float4* lo = (float4*)(smem + sstage * C::STAGE_BYTES + A_STAGE_BYTES + B_STAGE_BYTES);
          // A and B tiles are contiguous ([A 16K][B N*128]) and so are their lo twins: one flat loop, 8 loads in
          // flight per thread.  SPLIT == 1: only A (mode 1 segments), nothing for the W_lo segments (mode 2).
          const int n4 = SPLIT == 2 ? nA4 + nB4 : (SPLIT == 3 || g.mode[s] == 1 ? nA4 : 0);
          for (int i0 = ts; i0 < n4; i0 += TEAM * 8) {
            float4 x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (i0 + u * TEAM < n4) x[u] = a[i0 + u * TEAM];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (i0 + u * TEAM < n4) {
                float4 l;
                l.x = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xFFFFE000u);
                l.y = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xFFFFE000u);
                l.z = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xFFFFE000u);
                l.w = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xFFFFE000u);
                lo[i0 + u * TEAM] = l;
              }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> UMMA (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdone[sstage]);
          if (++sstage == STAGES) sstage = 0, sphase ^= 1;
        }
      }
    };

    if (SPLIT == 2) {
      // weight gradients: all 12 warps split the K loop of a work item, then the epilogue warps drain it
      uint32_t it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        split_item(w);
        if (is_epi) drain(w, it);
      }
    } else if ((SPLIT == 1 || SPLIT == 3) && !is_epi) {
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) split_item(w);
    } else {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) drain(w, it);
    }

    if constexpr (kColsum) {
      // fused bias gradient: per-CTA column sums of zbar -> colpart[blockIdx.x][col]; fixed reduction order
      // (4 row groups of a warp by shuffle, then the 4 lane-quarter warps of a column half through smem).
      if (epi.colpart && is_epi) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 t = csum[ch];
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            t.x += __shfl_xor_sync(0xffffffffu, t.x, o), t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
            t.z += __shfl_xor_sync(0xffffffffu, t.z, o), t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
          }
          if (sub == 0 && ch * 32 < ncol) st4(tile + ch * 32 + c4 * 4, t);   // this warp's 32 x ncol/32 sums
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");                       // the 8 epilogue warps only
        if (q == 0) {
          for (int c = lane; c < ncol; c += 32) {
            float tot = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              // warp with lane-quarter qq and this column half: e' such that (EPI_WARP0 + e') & 3 == qq, e' >> 2 == half
              const int e2 = half * 4 + ((qq - C::EPI_WARP0) & 3);
              tot += epi_tiles[e2 * (EPI_TILE_BYTES / 4) + c];
            }
            // one partial row per CTA -- or per m-tile when every CTA has exactly one work item (narrow column
            // tiles at small row counts: the column tiles of an m-tile then fill one row between them)
            const int prow = num_work <= (int)gridDim.x ? (int)(blockIdx.x % num_mtiles) : (int)blockIdx.x;
            const int pn0 = num_work <= (int)gridDim.x ? (int)((blockIdx.x % num_mn) / num_mtiles) * BN : 0;
            epi.colpart[(size_t)prow * 1024 + pn0 + half * ncol + c] = tot;
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
// Encoded maps are cached per (base, shape, box): the row arrays live at fixed workspace addresses, so after the first
// iteration a launch costs a table lookup instead of a driver call per operand (matters for the small-M steps).
struct MapKey {
  const void* base;
  long long inner, outer, ld;
  int box_inner, box_outer, mn;
  bool operator==(const MapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && mn == o.mn;
  }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool used;
};
constexpr int kMapCacheSlots = 1024;   // open addressing; cleared when three quarters full
inline bool make_map(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_inner,
                     int box_outer, bool mn_major) {
  static thread_local MapSlot* cache = nullptr;
  static thread_local int used = 0;
  if (!cache) cache = (MapSlot*)calloc(kMapCacheSlots, sizeof(MapSlot));
  const MapKey key{base, inner, outer, ld, box_inner, box_outer, mn_major ? 1 : 0};
  unsigned long long h = (unsigned long long)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)inner * 0xC2B2AE3D27D4EB4Full + (unsigned long long)outer * 0x165667B19E3779F9ull +
       (unsigned long long)ld * 0x27D4EB2F165667C5ull + (unsigned long long)(box_inner * 131 + box_outer * 7 + key.mn);
  int idx = (int)((h >> 20) % kMapCacheSlots);
  if (cache) {
    for (int probe = 0; probe < kMapCacheSlots; ++probe, idx = (idx + 1) % kMapCacheSlots) {
      if (!cache[idx].used) break;
      if (cache[idx].key == key) {
        *m = cache[idx].map;
        return true;
      }
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  const bool ok = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  if (ok && cache) {
    if (used >= kMapCacheSlots * 3 / 4) {
      memset(cache, 0, kMapCacheSlots * sizeof(MapSlot));
      used = 0;
      idx = (int)((h >> 20) % kMapCacheSlots);
    }
    while (cache[idx].used) idx = (idx + 1) % kMapCacheSlots;
    cache[idx].key = key, cache[idx].map = *m, cache[idx].used = true;
    ++used;
  }
  return ok;
}

}  // namespace tc

template <bool A_KC, bool B_KC>
inline bool tc_eligible(const GemmArgs& g, int nsplit) {
  if (g.N < 64 || g.N > 256 || g.N % 64 || g.nseg < 1 || g.nseg > kMaxSeg) return false;
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

// column-tile width for (g, nsplit): 256, or -- for sweeps over few rows, where 128-row x 256-column tiles would leave
// most SMs idle -- the narrowest of 128 / 64 that still gives every CTA at most one work item
template <bool A_KC, int SPLIT>
inline int tc_pick_bn(const GemmArgs& g, int nsplit, int num_sms) {
  if (!A_KC || SPLIT >= 2 || nsplit != 1 || g.N % 128) return 256;
  const int mtiles = (g.M + tc::BM - 1) / tc::BM;
  if (g.N == 256 && mtiles * 4 <= num_sms) return 64;
  if (mtiles * (g.N / 128) <= num_sms && g.N > 128) return 128;
  return 256;
}
// rows of fused column-sum partials the launch writes (see the kernel's colpart indexing)
template <bool A_KC, int SPLIT>
inline int tc_colpart_rows(const GemmArgs& g, int nsplit, int num_sms) {
  const int bn = tc_pick_bn<A_KC, SPLIT>(g, nsplit, num_sms);
  const int mtiles = (g.M + tc::BM - 1) / tc::BM;
  const long long work = (long long)mtiles * ((g.N + bn - 1) / bn) * nsplit;
  return work <= num_sms ? mtiles : num_sms;
}

template <bool A_KC, bool B_KC, int SPLIT, class Epi, int BN = 256>
inline cudaError_t launch_gemm_tc_bn(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st);

template <bool A_KC, bool B_KC, int SPLIT, class Epi>
inline cudaError_t launch_gemm_tc(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  if constexpr (SPLIT == 3) {   // three 64 KB stages up to 128 columns, two 96 KB stages beyond
    if (g.N <= 128) return launch_gemm_tc_bn<A_KC, B_KC, 3, Epi, 128>(g, epi, nsplit, num_sms, st);
    return launch_gemm_tc_bn<A_KC, B_KC, 3, Epi, 256>(g, epi, nsplit, num_sms, st);
  }
  if constexpr (A_KC && SPLIT != 2) {
    const int bn = tc_pick_bn<A_KC, SPLIT>(g, nsplit, num_sms);
    if (bn == 64) return launch_gemm_tc_bn<A_KC, B_KC, SPLIT, Epi, 64>(g, epi, nsplit, num_sms, st);
    if (bn == 128) return launch_gemm_tc_bn<A_KC, B_KC, SPLIT, Epi, 128>(g, epi, nsplit, num_sms, st);
  }
  return launch_gemm_tc_bn<A_KC, B_KC, SPLIT, Epi, 256>(g, epi, nsplit, num_sms, st);
}

template <bool A_KC, bool B_KC, int SPLIT, class Epi, int BN>
inline cudaError_t launch_gemm_tc_bn(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  constexpr bool A_MN = !A_KC, B_MN = !B_KC;
  using C = tc::Cfg<SPLIT, BN>;
  tc::TmSet tm;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok;
    if (A_MN) ok = tc::make_map(&tm.a[s], sg.A, g.M, sg.K, sg.lda, 32, 32, true);       // P[k = rows][m]
    else      ok = tc::make_map(&tm.a[s], sg.A, sg.K, g.M, sg.lda, 32, tc::BM, false);  // X[m = rows][k]
    if (B_MN) ok = ok && tc::make_map(&tm.b[s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);    // W[k][n] / Q[k = rows][n]
    else      ok = ok && tc::make_map(&tm.b[s], sg.B, sg.K, g.Nb, sg.ldb, 32, g.N < BN ? g.N : BN, false);  // W[n][k]
    if (!ok) return cudaErrorInvalidValue;
  }
  for (int s = g.nseg; s < kMaxSeg; ++s) tm.a[s] = tm.a[0], tm.b[s] = tm.b[0];
  if (SPLIT == 3) {   // lo twins of the weight operands ride in seg[s + 4].B
    if (g.nseg > 4) return cudaErrorInvalidValue;
    for (int s = 0; s < g.nseg; ++s) {
      const GemmSeg& sg = g.seg[s];
      const float* lo = g.seg[s + 4].B;
      const bool ok = B_MN ? tc::make_map(&tm.b[s + 4], lo, g.Nb, sg.K, sg.ldb, 32, 32, true)
                           : tc::make_map(&tm.b[s + 4], lo, sg.K, g.Nb, sg.ldb, 32, g.N < BN ? g.N : BN, false);
      if (!ok || !lo) return cudaErrorInvalidValue;
    }
  }
  auto kern = tc::gemm_tc_kernel<A_MN, B_MN, SPLIT, BN, Epi>;
  static unsigned long long attr_devs = 0;   // per device: opt in to the large dynamic shared memory once
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int mtiles = (g.M + tc::BM - 1) / tc::BM;
  const int work = mtiles * ((g.N + BN - 1) / BN) * nsplit;
  const int grid = work < num_sms ? work : num_sms;
  return tc::launch_pdl(kern, grid, C::NUM_THREADS, C::SMEM_BYTES, st, 1, tm, g, epi, mtiles, nsplit);
}

}  // namespace fbsnn
