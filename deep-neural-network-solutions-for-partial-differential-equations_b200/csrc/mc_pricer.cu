// Correlated-GBM basket Monte-Carlo pricer (replaces numerics/multidimensional_mc_pricer.py:49-93).
//
// A CTA simulates batches of kMcPB = 64 paths.  Every (path, asset, 4 steps) triple is one Philox4x32-10 block
// keyed by the GLOBAL path id, so prices do not depend on how paths are sharded over GPUs or CTAs.  Per batch:
//   1. normals:   the 64 x D (path, asset) pairs are dealt round-robin to the 256 threads (all lanes busy for any
//                 D that is a multiple of 4); each draws its N step normals and keeps only their sum -> zsT[d][p]
//   2. correlate: the Cholesky matvec is hoisted out of the time loop -- the payoff is terminal-only and
//                 sum_t L z_t == L sum_t z_t, so one D x D matvec per path replaces N of them (path-wise identical
//                 up to rounding).  Register-tiled over 4 paths per thread, L^T and zsT in shared
//                 memory; then the weighted terminal price w_d S0_d exp(drift + vol y) -> eT[d][p]
//   3. payoff:    one thread per path sums its D contributions in a fixed order, discounted payoff, fp64 sums
// All N*D normals per path are drawn, as the reference does; nothing is stored per path.  The bound is instruction
// issue: ~20 instructions per normal (Philox4x32-10: 5 IMAD.WIDE + 5 LOP3; Box-Muller: 2 MUFU + ~4 FP32), measured
// 57 % of the issue slots with the XU / FMA / ALU pipes at 40-47 % each (profiles/r01_ncu_details_mc_basket_v2.txt).
// Unroll depth (1-4) and 2-4 CTAs per SM were measured within 3 % of each other; <2 Philox blocks in flight, 4 CTAs> kept.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "philox.cuh"

namespace fbsnn {

constexpr int kMcThreads = 256;
constexpr int kMcPB = 64;                 // paths per CTA batch
constexpr int kMcMaxBlocks = 148 * 8;
constexpr int kMcMaxDJ = 8;               // D <= 256

struct McK {
  int D, N;
  float drift_T;     // (r - sigma^2/2) * T
  float vol_sqrt_dt; // sigma * sqrt(T/N)
  float disc;        // exp(-r T)
  float strike;
};

// sum over the N steps of the standard normals of asset d on global path gp
template <int UNROLL>
__device__ __forceinline__ float step_normal_sum(uint64_t gp, int d, int N, const PhiloxKeys& keys) {
  float acc0 = 0.f, acc1 = 0.f;
  const int nfull = N >> 2;
  Philox4 ctr{(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)d, 0u};
#pragma unroll UNROLL
  for (int q = 0; q < nfull; ++q) {
    float z[4];
    ctr.w = (uint32_t)q;
    normal4(philox4x32_10(ctr, keys), z);
    acc0 += z[0] + z[1];
    acc1 += z[2] + z[3];
  }
  const int rem = N & 3;
  if (rem) {
    float z[4];
    ctr.w = (uint32_t)nfull;
    normal4(philox4x32_10(ctr, keys), z);
    acc0 += z[0];
    if (rem > 1) acc0 += z[1];
    if (rem > 2) acc1 += z[2];
  }
  return acc0 + acc1;
}

// GREEKS: also accumulates the pathwise deltas  d price / d S0_d = E[ disc 1{basket > K} w_d S_T,d ] / S0_d  (the
// estimator basket_pricer.py:68-81 approximates by bump-and-revalue), one fp64 accumulator per asset in thread d.
template <bool GREEKS, int UNROLL, int MINB>
__global__ void __launch_bounds__(kMcThreads, MINB)
mc_basket_kernel(const McK k, const float* __restrict__ S0, const float* __restrict__ wts,
                 const float* __restrict__ cholT, int chol_in_smem, unsigned long long n_paths,
                 unsigned long long path_offset, uint64_t seed, double* __restrict__ part,
                 double* __restrict__ part_delta) {
  extern __shared__ __align__(16) float smem[];
  const int D = k.D;
  float* zsT = smem;                      // [D][kMcPB] summed normals, path-contiguous
  float* eT = smem + D * kMcPB;           // [D][kMcPB] weighted terminal prices
  float* LT = smem + 2 * D * kMcPB;       // transposed Cholesky factor (optional)
  if (cholT && chol_in_smem) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) LT[i] = cholT[i];
  }
  const float* Lp = cholT ? (chol_in_smem ? LT : cholT) : nullptr;
  const PhiloxKeys keys = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
  const int tid = threadIdx.x;

  double sum = 0.0, sumsq = 0.0, dsum = 0.0;
  __shared__ __align__(16) float ind[kMcPB];   // GREEKS: disc * 1{basket > K} per path of the batch
  const unsigned long long nbatch = (n_paths + kMcPB - 1) / kMcPB;
  for (unsigned long long bi = blockIdx.x; bi < nbatch; bi += gridDim.x) {
    const unsigned long long base = bi * kMcPB;
    __syncthreads();                      // previous batch's eT / zsT fully consumed (and LT loaded)
    // ---- 1. summed step normals of every (path, asset) pair of the batch
    for (int i = tid; i < D * kMcPB; i += kMcThreads) {
      const int p = i & (kMcPB - 1), d = i >> 6;
      const unsigned long long lp = base + p;
      zsT[i] = lp < n_paths ? step_normal_sum<UNROLL>(path_offset + lp, d, k.N, keys) : 0.f;
    }
    __syncthreads();
    // ---- 2. y = L z (lower triangular), weighted terminal prices; 4 paths per thread
    for (int j = tid; j < D * (kMcPB / 4); j += kMcThreads) {
      const int pg = j & (kMcPB / 4 - 1), d = j >> 4;
      float4 y;
      if (Lp) {
        y = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c <= d; ++c) {
          const float l = Lp[c * D + d];
          const float4 z = *reinterpret_cast<const float4*>(zsT + c * kMcPB + 4 * pg);
          y.x = fmaf(l, z.x, y.x), y.y = fmaf(l, z.y, y.y), y.z = fmaf(l, z.z, y.z), y.w = fmaf(l, z.w, y.w);
        }
      } else {
        y = *reinterpret_cast<const float4*>(zsT + d * kMcPB + 4 * pg);
      }
      const float ws0 = wts[d] * S0[d];
      float4 e;
      e.x = ws0 * expf(k.drift_T + k.vol_sqrt_dt * y.x), e.y = ws0 * expf(k.drift_T + k.vol_sqrt_dt * y.y);
      e.z = ws0 * expf(k.drift_T + k.vol_sqrt_dt * y.z), e.w = ws0 * expf(k.drift_T + k.vol_sqrt_dt * y.w);
      *reinterpret_cast<float4*>(eT + d * kMcPB + 4 * pg) = e;
    }
    __syncthreads();
    // ---- 3. basket payoff of path p (fixed summation order: independent of batch position and sharding)
    if (tid < kMcPB) {
      float basket = 0.f;
      const bool valid = base + tid < n_paths;
      if (valid) {
        for (int d = 0; d < D; ++d) basket += eT[d * kMcPB + tid];
        const double pay = (double)(k.disc * fmaxf(basket - k.strike, 0.f));
        sum += pay;
        sumsq += pay * pay;
      }
      if (GREEKS) ind[tid] = (valid && basket > k.strike) ? k.disc : 0.f;
    }
    if (GREEKS) {
      __syncthreads();
      if (tid < D) {
        float acc = 0.f;
#pragma unroll 4
        for (int p4 = 0; p4 < kMcPB / 4; ++p4) {
          const float4 e = *reinterpret_cast<const float4*>(eT + tid * kMcPB + 4 * p4);
          const float4 w = *reinterpret_cast<const float4*>(ind + 4 * p4);
          acc = fmaf(e.x, w.x, fmaf(e.y, w.y, fmaf(e.z, w.z, fmaf(e.w, w.w, acc))));
        }
        dsum += (double)acc;
      }
    }
  }
  if (GREEKS && tid < D) part_delta[(size_t)blockIdx.x * D + tid] = dsum / (double)S0[tid];
  __shared__ double red[32];
  const double bs = block_sum(sum, red);
  const double bq = block_sum(sumsq, red);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = bs;
    part[2 * blockIdx.x + 1] = bq;
  }
}

__global__ void mc_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ out,
                                const double* __restrict__ part_delta, int D, double* __restrict__ delta_out) {
  __shared__ double red[32];
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += part[2 * i], q += part[2 * i + 1];
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) out[0] = s, out[1] = q;
  if (delta_out)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      double acc = 0.0;
      for (int b = 0; b < nblk; ++b) acc += part_delta[(size_t)b * D + d];   // fixed order over CTAs
      delta_out[d] = acc;
    }
}

// ----------------------------------------------------------------------------------------------------
// "Exact" solution of the 100-D HJB test problem by the Cole-Hopf formula (hjb_implement.py:1088-1094):
//   u(t_n, X_n) = -ln E[ exp(-g(X_n + sqrt(2 |T - t_n|) W)) ],  g(x) = ln(0.5 + 0.5 |x|^2),  W ~ N(0, I_D)
// i.e. the mean of 1 / (0.5 + 0.5 |X_n + s W|^2) over n_mc draws per time point.  blockIdx.y = time point; a
// thread draws whole samples (D normals from D/4 Philox blocks keyed by (seed, sample id, n)), fp64 partials.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hjb_exact_kernel(int D, const float* __restrict__ t, const float* __restrict__ X, float T, unsigned long long n_mc,
                 uint64_t seed, double* __restrict__ part) {
  extern __shared__ float xs[];
  const int n = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) xs[d] = X[(size_t)n * D + d];
  __syncthreads();
  const float s = sqrtf(2.f * fabsf(T - t[n]));
  const PhiloxKeys keys = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
  const int D4 = (D + 3) >> 2;
  double acc = 0.0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_mc;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float q = 0.f;
    for (int b = 0; b < D4; ++b) {
      float z[4];
      normal4(philox4x32_10(Philox4{(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)n, (uint32_t)b}, keys), z);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * b + j < D) {
          const float v = fmaf(s, z[j], xs[4 * b + j]);
          q = fmaf(v, v, q);
        }
    }
    acc += (double)(1.f / (0.5f + 0.5f * q));
  }
  __shared__ double red[32];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[(size_t)n * gridDim.x + blockIdx.x] = acc;
}
__global__ void hjb_exact_final_kernel(const double* __restrict__ part, int nblk, unsigned long long n_mc,
                                       double* __restrict__ out) {
  __shared__ double red[32];
  const int n = blockIdx.x;
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) acc += part[(size_t)n * nblk + i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[n] = -log(acc / (double)n_mc);
}

// Full path tensor (n, N+1, D), per-step Cholesky matvec exactly as generate_paths (:59-65); same Philox
// keying as the pricer, so paths[:, -1, :] reproduces the pricer's terminal prices up to rounding.
__global__ void __launch_bounds__(kMcThreads)
mc_paths_kernel(const McK k, const float* __restrict__ S0, const float* __restrict__ cholT,
                unsigned long long n_paths, unsigned long long path_offset, uint64_t seed,
                float* __restrict__ paths) {
  extern __shared__ float smem[];
  const int D = k.D, N = k.N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kMcThreads / 32;
  float* zs = smem + warp * D;
  const PhiloxKeys keys = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
  const int DJ = (D + 31) >> 5;
  const float drift_dt = k.drift_T / (float)N;
  const unsigned long long gw = (unsigned long long)blockIdx.x * nwarp + warp;
  const unsigned long long tw = (unsigned long long)gridDim.x * nwarp;
  for (unsigned long long i = gw; i < n_paths; i += tw) {
    const uint64_t gp = path_offset + i;
    float* out = paths + i * (unsigned long long)(N + 1) * D;
    float s[kMcMaxDJ];
#pragma unroll
    for (int j = 0; j < kMcMaxDJ; ++j) {
      const int d = lane + 32 * j;
      s[j] = (j < DJ && d < D) ? S0[d] : 0.f;
      if (j < DJ && d < D) out[d] = s[j];
    }
    for (int q = 0; q < (N + 3) / 4; ++q) {
      float z4[kMcMaxDJ][4];
#pragma unroll
      for (int j = 0; j < kMcMaxDJ; ++j) {
        const int d = lane + 32 * j;
        if (j < DJ && d < D)
          normal4(philox4x32_10(Philox4{(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)d, (uint32_t)q}, keys), z4[j]);
        else
          z4[j][0] = z4[j][1] = z4[j][2] = z4[j][3] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int step = 4 * q + u + 1;
        if (step > N) break;
        if (cholT) {
#pragma unroll
          for (int j = 0; j < kMcMaxDJ; ++j) {
            const int d = lane + 32 * j;
            if (j < DJ && d < D) zs[d] = z4[j][u];
          }
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < kMcMaxDJ; ++j) {
          const int d = lane + 32 * j;
          if (j < DJ && d < D) {
            float y = z4[j][u];
            if (cholT) {
              y = 0.f;
              const int jmax = min(D, 32 * j + 32);
              for (int c = 0; c < jmax; ++c) y = fmaf(__ldg(cholT + c * D + d), zs[c], y);
            }
            s[j] *= expf(drift_dt + k.vol_sqrt_dt * y);
            out[(unsigned long long)step * D + d] = s[j];
          }
        }
        if (cholT) __syncwarp();
      }
    }
  }
}


// ----------------------------------------------------------------------------------------------------
// Closed-form comparators the reference's drivers plot the learned Y against, evaluated for
// a whole prediction tensor in one launch instead of Python loops over (sample, step):
//   mode 0  BasketOptionPriceCalculator.calculate_option_prices (nd_BSPDE_case.py:621-658): Black-Scholes call per
//           (row, asset) with time to maturity T - t[row], then the equal-weighted mean over the assets
//   mode 1  BasicOptionPriceCalculator.calculate_call_option_prices (with_corr_high_dimension_pde.py:663-700): one
//           Black-Scholes call per (row, col) on the basket average with volatility sigma / sqrt(D) and time to maturity
//           T - time[min(col, ntimes - 1)]; at maturity the payoff and its one-sided delta
// Same IEEE behaviour as the reference formulas (tau = 0 in mode 0 gives d1 = +-inf, i.e. the intrinsic value).
// ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
__device__ __forceinline__ void bs_call(double S, double K, double tau, double r, double sigma, double& price, double& delta) {
  const double sq = sigma * sqrt(tau);
  const double d1 = (log(S / K) + (r + 0.5 * sigma * sigma) * tau) / sq;
  const double d2 = d1 - sq;
  delta = norm_cdf(d1);
  price = S * delta - K * exp(-r * tau) * norm_cdf(d2);
}
__global__ void bs_comparator_kernel(int mode, const double* __restrict__ S, const double* __restrict__ t, long long rows,
                                     int cols, int ntimes, double K, double r, double sigma, double T, int dims,
                                     double* __restrict__ price_out, double* __restrict__ delta_out) {
  if (mode == 0) {   // one warp per row, lanes over assets, fixed-order warp reduction
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const double tau = T - t[row];
    double ps = 0.0, ds = 0.0;
    for (int a = lane; a < cols; a += 32) {
      double p, d;
      bs_call(S[row * cols + a], K, tau, r, sigma, p, d);
      ps += p, ds += d;
    }
    ps = warp_sum(ps), ds = warp_sum(ds);
    if (lane == 0) price_out[row] = ps / cols, delta_out[row] = ds / cols;
  } else {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int j = (int)(i % cols);
    const double tau = T - t[j < ntimes ? j : ntimes - 1];
    const double s = S[i];
    double p, d;
    if (tau > 0.0) {
      bs_call(s, K, tau, r, sigma / sqrt((double)dims), p, d);
    } else {
      p = fmax(s - K, 0.0);
      d = s > K ? 1.0 : (s == K ? 0.5 : 0.0);
    }
    price_out[i] = p, delta_out[i] = d;
  }
}

// Roofline denominator of the pricer: nothing but the generator -- Philox4x32-10 with
// register-resident round keys + the MUFU Box-Muller of normal4() -- at full occupancy, every normal consumed by one
// add.  bench.py times this probe on the same GPU and reports the pricer's normals/s as a fraction of it.
__global__ void __launch_bounds__(256, 4)
normal_rate_probe_kernel(unsigned long long blocks_per_thread, uint64_t seed, double* __restrict__ part) {
  const PhiloxKeys keys = philox_keys((uint32_t)seed, (uint32_t)(seed >> 32));
  const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  Philox4 ctr{(uint32_t)t, (uint32_t)(t >> 32), 0u, 0u};
  float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 2
  for (unsigned long long i = 0; i < blocks_per_thread; ++i) {
    float z[4];
    ctr.z = (uint32_t)i, ctr.w = (uint32_t)(i >> 32);
    normal4(philox4x32_10(ctr, keys), z);
    acc0 += z[0] + z[1];
    acc1 += z[2] + z[3];
  }
  __shared__ double red[32];
  const double bs = block_sum((double)(acc0 + acc1), red);
  if (threadIdx.x == 0) part[blockIdx.x] = bs;
}

}  // namespace fbsnn

using namespace fbsnn;

static long long g_mc_launches = 0;

static int mc_make(const McSpec* spec, McK& k) {
  if (!spec || spec->D < 1 || spec->D > 32 * kMcMaxDJ || spec->N < 1) return FBSNN_E_UNSUPPORTED;
  k.D = spec->D, k.N = spec->N;
  k.drift_T = (float)(((double)spec->rate - 0.5 * (double)spec->sigma * (double)spec->sigma) * (double)spec->T);
  k.vol_sqrt_dt = (float)((double)spec->sigma * sqrt((double)spec->T / (double)spec->N));
  k.disc = (float)exp(-(double)spec->rate * (double)spec->T);
  k.strike = spec->strike;
  return 0;
}

template <bool GREEKS, int UNROLL, int MINB>
static int mc_price_impl(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                         uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                         double* delta_out, void* stream) {
  McK k;
  if (mc_make(spec, k) || !S0 || !weights || !scratch || !sums_out || n_paths == 0 || (GREEKS && !delta_out))
    return FBSNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  auto kern = mc_basket_kernel<GREEKS, UNROLL, MINB>;
  size_t smem = (size_t)2 * k.D * kMcPB * sizeof(float);
  int chol_in_smem = 0;
  // four CTAs per SM (64 registers, <= 56 KB each) hide the Philox/MUFU latencies; L^T joins the batch buffers in
  // shared memory only when that still fits, otherwise the 5 %-of-the-time matvec reads it through L1
  if (chol_T && smem + (size_t)k.D * k.D * sizeof(float) <= (size_t)(224 / MINB) * 1024) {
    chol_in_smem = 1;
    smem += (size_t)k.D * k.D * sizeof(float);
  }
  static int blocks_per_sm = 0, sms = 0;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return FBSNN_E_CUDA;
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    smem_set = smem;
    blocks_per_sm = 0;
  }
  if (blocks_per_sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kMcThreads, smem_set) != cudaSuccess ||
        blocks_per_sm < 1)
      return FBSNN_E_CUDA;
  }
  const unsigned long long want = (n_paths + kMcPB - 1) / kMcPB;
  const unsigned long long cap = (unsigned long long)std::min(sms * blocks_per_sm, kMcMaxBlocks);
  const int blocks = (int)(want < cap ? want : cap);
  double* part = (double*)scratch;
  double* part_delta = part + 2 * (size_t)kMcMaxBlocks;
  kern<<<blocks, kMcThreads, smem, st>>>(k, S0, weights, chol_T, chol_in_smem, n_paths, path_offset, seed, part,
                                         part_delta);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  mc_final_kernel<<<1, 256, 0, st>>>(part, blocks, sums_out, part_delta, k.D, GREEKS ? delta_out : nullptr);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  g_mc_launches += 2;
  return 0;
}

extern "C" {

long long mc_launch_count(void) { return g_mc_launches; }

size_t mc_scratch_bytes(void) { return (size_t)kMcMaxBlocks * (2 + 32 * kMcMaxDJ) * sizeof(double); }

int mc_basket_price(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                    uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                    void* stream) {
  return mc_price_impl<false, 2, 4>(spec, S0, weights, chol_T, n_paths, seed, path_offset, scratch, sums_out, nullptr, stream);
}

int mc_basket_price_delta(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                          uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                          double* delta_sums_out, void* stream) {
  return mc_price_impl<true, 2, 4>(spec, S0, weights, chol_T, n_paths, seed, path_offset, scratch, sums_out, delta_sums_out,
                             stream);
}

int mc_hjb_exact(int32_t D, int32_t n_times, const float* t, const float* X, float T, uint64_t n_mc, uint64_t seed,
                 void* scratch, double* u_out, void* stream) {
  if (D < 1 || D > 4096 || n_times < 1 || n_times > 65535 || !t || !X || !scratch || !u_out || n_mc == 0)
    return FBSNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long want = (n_mc + 255) / 256;
  const size_t cap = mc_scratch_bytes() / sizeof(double) / (size_t)n_times;
  int bx = (int)std::min<unsigned long long>(want, std::min<size_t>(cap, 64));
  if (bx < 1) return FBSNN_E_UNSUPPORTED;
  hjb_exact_kernel<<<dim3(bx, n_times), 256, D * sizeof(float), st>>>(D, t, X, T, n_mc, seed, (double*)scratch);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  hjb_exact_final_kernel<<<n_times, 128, 0, st>>>((const double*)scratch, bx, n_mc, u_out);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  g_mc_launches += 2;
  return 0;
}

// Measurement hook: draws ~n_normals standard normals with the pricer's generator and nothing else (see
// normal_rate_probe_kernel); *normals_out = the exact count drawn.  The caller times it with CUDA events.
int mc_normal_rate_probe(uint64_t n_normals, uint64_t seed, void* scratch, uint64_t* normals_out, void* stream) {
  if (!scratch || !normals_out || n_normals == 0) return FBSNN_E_BADARG;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
    return FBSNN_E_CUDA;
  const int blocks = std::min(sms * 8, kMcMaxBlocks);
  const unsigned long long threads = (unsigned long long)blocks * 256;
  const unsigned long long per_thread = std::max<unsigned long long>(1, n_normals / (4 * threads));
  normal_rate_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(per_thread, seed, (double*)scratch);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  *normals_out = 4ull * per_thread * threads;
  g_mc_launches += 1;
  return 0;
}

int mc_bs_comparator(int32_t mode, const double* S, const double* t, int64_t rows, int32_t cols, int32_t ntimes, double K,
                     double r, double sigma, double T, int32_t dims, double* price_out, double* delta_out, void* stream) {
  if ((mode != 0 && mode != 1) || !S || !t || rows < 1 || cols < 1 || !price_out || !delta_out || (mode == 1 && (ntimes < 1 || dims < 1)))
    return FBSNN_E_BADARG;
  const long long threads = mode == 0 ? (long long)rows * 32 : (long long)rows * cols;
  bs_comparator_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mode, S, t, rows, cols, ntimes, K, r,
                                                                                         sigma, T, dims, price_out, delta_out);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  g_mc_launches += 1;
  return 0;
}

int mc_generate_paths(const McSpec* spec, const float* S0, const float* chol_T, uint64_t n_paths,
                      uint64_t seed, uint64_t path_offset, float* paths_out, void* stream) {
  McK k;
  if (mc_make(spec, k) || !S0 || !paths_out || n_paths == 0) return FBSNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nwarp = kMcThreads / 32;
  const size_t smem = (size_t)nwarp * k.D * sizeof(float);
  const unsigned long long want = (n_paths + nwarp - 1) / nwarp;
  const int blocks = (int)(want < (unsigned long long)kMcMaxBlocks ? want : (unsigned long long)kMcMaxBlocks);
  mc_paths_kernel<<<blocks, kMcThreads, smem, st>>>(k, S0, chol_T, n_paths, path_offset, seed, paths_out);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  g_mc_launches += 1;
  return 0;
}

}  // extern "C"
