// CTA-pair (cta_group::2) form of the 3xTF32 dense-layer kernel in gemm_tc.cuh.
//
// Why: with three MMAs per k-step the single-CTA 3xTF32 kernel is bound by SHARED-MEMORY bandwidth, not by HBM or
// the tensor pipe: per 32-wide k-block pair an SM moves 96 KB of TMA writes + 32 KB of splitter traffic + 12 MMAs x
// 12 KB of operand reads = 272 KB against 128 B/clk x 1536 MMA cycles = 196 KB (weight gradients: 288 KB).  Pairing two
// SMs on one 256-row UMMA tile halves the B-operand share of every CTA: each CTA stages its own 128 A rows and HALF of
// the <=256 B columns, the leader CTA's single thread issues tcgen05.mma.cta_group::2 (M = 256), and every MMA reads
// 4 KB (A) + 4 KB (B half) per CTA instead of 4 + 8.  Per k-block pair: 64 + 32 + 96 = 192 KB, inside the budget.
// The smaller stages also deepen the ring (4 stages for the sweeps, 3 for the weight gradients instead of 3 / 2).
//
// Protocol (both CTAs run the same code; rank = %cluster_ctarank, leader = rank 0):
//   warp 0     TMA producer of its OWN CTA: waits its local empty[stage], loads A (own rows) + B (own column half)
//              onto its local full[stage]
//   warps 2-5  (+ the epilogue warps for SPLIT == 2) splitters of their own CTA: wait local full[stage], write the lo
//              tiles, fence.proxy.async, then arrive on the LEADER's sdone[stage] (count = 2 x team warps; the peer's
//              arrivals are remote: mapa + mbarrier.arrive.release.cluster)
//   warp 1     leader only: waits sdone[stage] (acquire.cluster), issues the cta_group::2 MMAs, then
//              tcgen05.commit.cta_group::2 ... multicast::cluster to empty[stage] of BOTH CTAs; after the last k-block
//              the same to tfull[acc] of both
//   warps 6-13 epilogue of their own CTA (own 128 accumulator rows in their own TMEM); when drained they arrive on
//              the LEADER's tempty[acc] (count = 2 x 8)
// TMEM is allocated with tcgen05.alloc.cta_group::2 by warp 1 of both CTAs; cluster barriers bracket set-up and
// tear-down so that no CTA touches (or leaves) a peer whose barriers are not live.
// SPLIT = 1, 2 as in gemm_tc.cuh; SPLIT = 3 (sweeps on the pair): the weight operand has exact-TF32 hi / lo twins in
// global memory (seg[s].B = W_hi, seg[s + 4].B = W_lo) which TMA drops into the B and B_lo slots of ONE stage, only A is
// split in-kernel, three MMAs per k-step -- the A tile is fetched once (SPLIT = 1 fetches it for the hi and again for
// the lo segment) and a stage carries 1536 MMA cycles, which covers the pair's stage hand-off latency.
// Single-pass TF32 is not shared-memory bound and has no pair form.
#pragma once
#include "gemm_tc.cuh"

namespace fbsnn {
namespace tc2 {

using namespace tc;

constexpr int B_HALF_BYTES = 128 * BK * 4;   // 16 KB: at most 128 of the 256 B columns per CTA

template <int SPLIT>
struct Cfg2 {
  static_assert(SPLIT >= 1 && SPLIT <= 3, "the CTA-pair kernel exists for the 3xTF32 variants only");
  static constexpr int STAGES = SPLIT >= 2 ? 3 : 4;
  static constexpr int SPLIT_WARPS = 4;
  static constexpr int EPI_WARP0 = 2 + SPLIT_WARPS;
  static constexpr int SPLIT_TEAM_WARPS = SPLIT == 2 ? SPLIT_WARPS + NUM_EPI_WARPS : SPLIT_WARPS;
  static constexpr int NUM_THREADS = 32 * (EPI_WARP0 + NUM_EPI_WARPS);
  // [A 16K][B 16K][A_lo 16K]([B_lo 16K])
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_HALF_BYTES + A_STAGE_BYTES + (SPLIT >= 2 ? B_HALF_BYTES : 0);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES + 1024 /*align*/ + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the barrier that sits at `bar`'s offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
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
    if (++spins > kSpinLimit) asm volatile("trap;");
  }
}
// completion of all prior MMAs of this thread -> arrive on `bar` (same offset) in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(z)
      : "memory");
}
// instruction descriptor: D = f32, A = B = tf32, M = 256 (128 rows in each CTA of the pair)
__host__ __device__ __forceinline__ uint32_t make_idesc_pair(int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (a_mn ? 1u : 0u) << 15;
  d |= (b_mn ? 1u : 0u) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(256 >> 4) << 24;
  return d;
}

template <bool A_MN, bool B_MN, int SPLIT, class Epi>
__global__ void __launch_bounds__(Cfg2<SPLIT>::NUM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ TmSet tm, const GemmArgs g, const Epi epi, const int num_mtiles2,
                const int nsplit) {
  using C = Cfg2<SPLIT>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_tiles = (float*)(smem + STAGES * C::STAGE_BYTES);
  uint64_t* bars = (uint64_t*)(smem + STAGES * C::STAGE_BYTES + NUM_EPI_WARPS * EPI_TILE_BYTES);
  uint64_t* full = bars;               // [STAGES]  local: this CTA's TMA bytes landed
  uint64_t* empty = bars + 4;          // [STAGES]  local: the pair's MMAs that read the stage completed (multicast commit)
  uint64_t* sdone = bars + 8;          // [STAGES]  LEADER's copy is used: lo tiles of both CTAs written
  uint64_t* tfull = bars + 12;         // [2]       local: accumulator complete (multicast commit)
  uint64_t* tempty = bars + 14;        // [2]       LEADER's copy is used: both CTAs' epilogues drained the buffer
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int N = g.N;                                    // <= 256: one column tile, half of it staged per CTA
  const int NH = N >> 1;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int num_work = num_mtiles2 * nsplit;            // work id = (split, mt2), mt2 fastest

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nseg; ++s) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.a[s]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm.b[s]) : "memory");
    }
    for (int i = 0; i < STAGES; ++i)
      mbar_init(&full[i], 1), mbar_init(&empty[i], 1), mbar_init(&sdone[i], 2);   // one arrival per CTA (see below)
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1), mbar_init(&tempty[i], 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                   // both CTAs' barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  auto kbeg = [&](int s, int split) { return g.kchunk ? (int)min((long long)g.seg[s].K, (long long)split * g.kchunk) : 0; };
  auto kend = [&](int s, int split) {
    return g.kchunk ? (int)min((long long)g.seg[s].K, ((long long)split + 1) * g.kchunk) : g.seg[s].K;
  };

  if (warp == 0) {
    // ===================== TMA producer (own rows of A, own half of B) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t bytes = A_STAGE_BYTES + (uint32_t)NH * BK * 4 * (SPLIT == 3 ? 2 : 1);
      for (int w = cid; w < num_work; w += ncl) {
        const int mt2 = w % num_mtiles2, split = w / num_mtiles2;
        const int m0 = mt2 * 256 + 128 * (int)rank, n0 = NH * (int)rank;
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
              for (int c = 0; c < NH / 32; ++c) tma_load_2d(b + c * 4096, &tm.b[s], &full[stage], n0 + 32 * c, k0);
            } else {
              tma_load_2d(b, &tm.b[s], &full[stage], k0, n0);
            }
            if (SPLIT == 3) {   // W_lo twin -> the B_lo slot
              uint8_t* bl = b + B_HALF_BYTES + A_STAGE_BYTES;
              if (B_MN) {
                for (int c = 0; c < NH / 32; ++c) tma_load_2d(bl + c * 4096, &tm.b[s + 4], &full[stage], n0 + 32 * c, k0);
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
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_pair(N, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int w = cid; w < num_work; w += ncl, ++it) {
        const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
        mbar_wait_cluster(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        const int split = w / num_mtiles2;
        uint32_t first = 1;
        for (int s = 0; s < g.nseg; ++s) {
          for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
            mbar_wait_cluster(&sdone[stage], phase);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint32_t b = a + A_STAGE_BYTES;
            const uint32_t alo = b + B_HALF_BYTES, blo = alo + A_STAGE_BYTES;
            const int mode = SPLIT >= 2 ? 3 : g.mode[s];
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              auto da = [&](uint32_t base) { return A_MN ? make_desc(base + k * 1024, 4096, 512, 1) : make_desc(base + k * 32, 16, 1024, 2); };
              auto db = [&](uint32_t base) { return B_MN ? make_desc(base + k * 1024, 4096, 512, 1) : make_desc(base + k * 32, 16, 1024, 2); };
              if (mode & 1) {   // modes 1, 3: a_lo * b
                umma_tf32_pair(tmem_d, da(alo), db(b), idesc, first ? 0u : 1u);
                first = 0;
              }
              if (SPLIT >= 2) umma_tf32_pair(tmem_d, da(a), db(blo), idesc, 1u);
              umma_tf32_pair(tmem_d, da(a), db(b), idesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit_pair(&empty[stage]);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        umma_commit_pair(&tfull[acc]);
      }
    }
  } else {
    // ===================== splitter and epilogue warps (own CTA's tiles, own TMEM rows) =====================
    const bool is_epi = warp >= C::EPI_WARP0;
    const int e = warp - C::EPI_WARP0;
    const int q = warp & 3;
    const int half = e >> 2;
    const int ncol = N >> 1;
    float* tile = epi_tiles + (is_epi ? e : 0) * (EPI_TILE_BYTES / 4);
    const int sub = lane >> 3, c4 = lane & 7;
    constexpr bool kColsum = Epi::kColsum;
    float4 csum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) csum[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto drain = [&](int w, uint32_t it) {
      const int mt2 = w % num_mtiles2;
      const int split = w / num_mtiles2;
      (void)split;
      const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      const int r0 = mt2 * 256 + 128 * (int)rank + q * 32;
#pragma unroll 1
      for (int ch = 0; ch * 32 < ncol; ++ch) {
        const int ct = half * ncol + ch * 32;
        const int c0 = ct;
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
        if constexpr (kColsum) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (ch == k) csum[k].x += cs.x, csum[k].y += cs.y, csum[k].z += cs.z, csum[k].w += cs.w;
        }
      }
      // the eight epilogue warps meet on a named barrier and ONE thread signals the leader: a cluster-scope release
      // arrive is a gpu-wide fence (30 % of the stall samples of the weight-gradient kernel when every warp did it)
      tc_fence_before();
      asm volatile("bar.sync 3, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
      if (threadIdx.x == 32 * C::EPI_WARP0) mbar_arrive_cta(&tempty[acc], 0);
    };

    constexpr int TEAM = 32 * C::SPLIT_TEAM_WARPS;
    uint32_t sstage = 0, sphase = 0;
    auto split_item = [&](int w) {
      const int ts = threadIdx.x - 64;
      const int nA4 = A_STAGE_BYTES / 16, nB4 = NH * BK * 4 / 16;
      const int split = w / num_mtiles2;
      for (int s = 0; s < g.nseg; ++s) {
        for (int k0 = kbeg(s, split), ke = kend(s, split); k0 < ke; k0 += BK) {
          mbar_wait(&full[sstage], sphase);
          const float4* a = (const float4*)(smem + sstage * C::STAGE_BYTES);
          float4* lo = (float4*)(smem + sstage * C::STAGE_BYTES + A_STAGE_BYTES + B_HALF_BYTES);
          // [A 16K][B 16K] and their lo twins [A_lo 16K][B_lo 16K] are laid out alike; with fewer than 128 B columns
          // per CTA the B tile is shorter but still starts at the 16 KB mark
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
          asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy stores -> UMMA (async proxy)
          asm volatile("bar.sync 2, %0;" ::"n"(TEAM) : "memory");
          if (threadIdx.x == 64) mbar_arrive_cta(&sdone[sstage], 0);
          if (++sstage == STAGES) sstage = 0, sphase ^= 1;
        }
      }
    };

    if (SPLIT == 2) {
      uint32_t it = 0;
      for (int w = cid; w < num_work; w += ncl, ++it) {
        split_item(w);
        if (is_epi) drain(w, it);
      }
    } else if (!is_epi) {
      for (int w = cid; w < num_work; w += ncl) split_item(w);
    } else {
      uint32_t it = 0;
      for (int w = cid; w < num_work; w += ncl, ++it) drain(w, it);
    }

    if constexpr (kColsum) {
      if (epi.colpart && is_epi) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 t = csum[ch];
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            t.x += __shfl_xor_sync(0xffffffffu, t.x, o), t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
            t.z += __shfl_xor_sync(0xffffffffu, t.z, o), t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
          }
          if (sub == 0 && ch * 32 < ncol) st4(tile + ch * 32 + c4 * 4, t);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (q == 0) {
          for (int c = lane; c < ncol; c += 32) {
            float tot = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const int e2 = half * 4 + ((qq - C::EPI_WARP0) & 3);
              tot += epi_tiles[e2 * (EPI_TILE_BYTES / 4) + c];
            }
            epi.colpart[(size_t)blockIdx.x * 1024 + half * ncol + c] = tot;
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                   // nobody leaves while the peer may still signal or be read
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace tc2

// shapes the CTA-pair kernel takes: as tc_eligible, plus N a multiple of 64 (32-column chunks per CTA half) and,
// for the weight-gradient layout, an output of exactly 256 rows (both CTAs hold 128 valid rows)
template <bool A_KC, bool B_KC>
inline bool tc2_eligible(const GemmArgs& g, int nsplit) {
  if (!tc_eligible<A_KC, B_KC>(g, nsplit)) return false;
  if (g.N % 64) return false;
  if (!A_KC && g.M != 256) return false;
  return true;
}

// number of CTAs launch_gemm_tc2 uses (the fused bias-gradient partials are indexed by CTA)
inline int tc2_grid(const GemmArgs& g, int nsplit, int num_sms) {
  const int mtiles2 = (g.M + 255) / 256;
  const long long work = (long long)mtiles2 * nsplit;
  const int ncl = (int)std::min<long long>(work, num_sms / 2);
  return 2 * ncl;
}

template <bool A_KC, bool B_KC, int SPLIT, class Epi>
inline cudaError_t launch_gemm_tc2(const GemmArgs& g, const Epi& epi, int nsplit, int num_sms, cudaStream_t st) {
  constexpr bool A_MN = !A_KC, B_MN = !B_KC;
  using C = tc2::Cfg2<SPLIT>;
  tc::TmSet tm;
  const int NH = g.N / 2;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    bool ok;
    if (A_MN) ok = tc::make_map(&tm.a[s], sg.A, g.M, sg.K, sg.lda, 32, 32, true);
    else      ok = tc::make_map(&tm.a[s], sg.A, sg.K, g.M, sg.lda, 32, tc::BM, false);
    if (B_MN) ok = ok && tc::make_map(&tm.b[s], sg.B, g.Nb, sg.K, sg.ldb, 32, 32, true);
    else      ok = ok && tc::make_map(&tm.b[s], sg.B, sg.K, g.Nb, sg.ldb, 32, NH, false);
    if (!ok) return cudaErrorInvalidValue;
  }
  for (int s = g.nseg; s < kMaxSeg; ++s) tm.a[s] = tm.a[0], tm.b[s] = tm.b[0];
  if (SPLIT == 3) {   // lo twins of the weight operands ride in seg[s + 4].B
    if (g.nseg > 4) return cudaErrorInvalidValue;
    for (int s = 0; s < g.nseg; ++s) {
      const GemmSeg& sg = g.seg[s];
      const float* lo = g.seg[s + 4].B;
      const bool ok = B_MN ? tc::make_map(&tm.b[s + 4], lo, g.Nb, sg.K, sg.ldb, 32, 32, true)
                           : tc::make_map(&tm.b[s + 4], lo, sg.K, g.Nb, sg.ldb, 32, NH, false);
      if (!ok || !lo) return cudaErrorInvalidValue;
    }
  }
  auto kern = tc2::gemm_tc2_kernel<A_MN, B_MN, SPLIT, Epi>;
  static unsigned long long attr_devs = 0;   // per device: opt in to the large dynamic shared memory once
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return cudaErrorInvalidDevice;
  if (!((attr_devs >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_devs |= 1ull << dev;
  }
  const int mtiles2 = (g.M + 255) / 256;
  return tc::launch_pdl(kern, tc2_grid(g, nsplit, num_sms), C::NUM_THREADS, C::SMEM_BYTES, st, 2, tm, g, epi, mtiles2, nsplit);
}

}  // namespace fbsnn
