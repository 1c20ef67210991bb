// Correlated-GBM basket Monte-Carlo pricer (replaces numerics/multidimensional_mc_pricer.py:49-93).
//
// One warp simulates one path at a time; lane `l` owns assets d = l, l+32, ...  Every (path, asset, 4 steps)
// triple is one Philox4x32-10 block keyed by the GLOBAL path id, so prices do not depend on how paths are
// sharded over GPUs or CTAs.  The N per-step normals of an asset are all drawn (N*D normals per path, as the
// reference does) but the Cholesky matvec is hoisted out of the time loop: the payoff is terminal-only and
// sum_t L z_t == L sum_t z_t, so one D x D matvec per path replaces N of them (path-wise identical up to
// rounding; stated in DESIGN.md).  Nothing is stored per path; payoffs are reduced in double.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "philox.cuh"

namespace fbsnn {

constexpr int kMcThreads = 256;           // 8 warps per CTA
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
__device__ __forceinline__ float step_normal_sum(uint64_t gp, int d, int N, uint32_t k0, uint32_t k1) {
  float acc = 0.f;
  const int nq = (N + 3) >> 2;
  for (int q = 0; q < nq; ++q) {
    float z[4];
    normal4(philox4x32_10(Philox4{(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)d, (uint32_t)q}, k0, k1), z);
    const int rem = N - 4 * q;
    acc += z[0];
    if (rem > 1) acc += z[1];
    if (rem > 2) acc += z[2];
    if (rem > 3) acc += z[3];
  }
  return acc;
}

__global__ void __launch_bounds__(kMcThreads)
mc_basket_kernel(const McK k, const float* __restrict__ S0, const float* __restrict__ wts,
                 const float* __restrict__ cholT, int chol_in_smem, unsigned long long n_paths,
                 unsigned long long path_offset, uint64_t seed, double* __restrict__ part) {
  extern __shared__ float smem[];
  const int D = k.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kMcThreads / 32;
  float* zs = smem + warp * D;          // per-warp summed normals
  float* LT = smem + nwarp * D;         // transposed Cholesky factor (optional)
  if (cholT && chol_in_smem) {
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) LT[i] = cholT[i];
  }
  __syncthreads();
  const float* Lp = cholT ? (chol_in_smem ? LT : cholT) : nullptr;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const int DJ = (D + 31) >> 5;

  float s0[kMcMaxDJ], wv[kMcMaxDJ];
#pragma unroll
  for (int j = 0; j < kMcMaxDJ; ++j) {
    const int d = lane + 32 * j;
    s0[j] = (j < DJ && d < D) ? S0[d] : 0.f;
    wv[j] = (j < DJ && d < D) ? wts[d] : 0.f;
  }

  double sum = 0.0, sumsq = 0.0;
  const unsigned long long gw = (unsigned long long)blockIdx.x * nwarp + warp;
  const unsigned long long tw = (unsigned long long)gridDim.x * nwarp;
  for (unsigned long long i = gw; i < n_paths; i += tw) {
    const uint64_t gp = path_offset + i;
    float zl[kMcMaxDJ];
#pragma unroll
    for (int j = 0; j < kMcMaxDJ; ++j) {
      const int d = lane + 32 * j;
      zl[j] = (j < DJ && d < D) ? step_normal_sum(gp, d, k.N, k0, k1) : 0.f;
    }
    float basket = 0.f;
    if (Lp) {
#pragma unroll
      for (int j = 0; j < kMcMaxDJ; ++j) {
        const int d = lane + 32 * j;
        if (j < DJ && d < D) zs[d] = zl[j];
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < kMcMaxDJ; ++j) {
        const int d = lane + 32 * j;
        if (j < DJ && d < D) {
          float y = 0.f;
          const int jmax = min(D, 32 * j + 32);   // L is lower triangular: columns beyond d are zero
          for (int c = 0; c < jmax; ++c) y = fmaf(Lp[c * D + d], zs[c], y);
          basket = fmaf(wv[j], s0[j] * expf(k.drift_T + k.vol_sqrt_dt * y), basket);
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int j = 0; j < kMcMaxDJ; ++j) {
        const int d = lane + 32 * j;
        if (j < DJ && d < D) basket = fmaf(wv[j], s0[j] * expf(k.drift_T + k.vol_sqrt_dt * zl[j]), basket);
      }
    }
    basket = warp_sum(basket);
    const double pay = (double)(k.disc * fmaxf(basket - k.strike, 0.f));
    sum += pay;
    sumsq += pay * pay;
  }
  __shared__ double red[32];
  const double bs = block_sum(lane == 0 ? sum : 0.0, red);
  const double bq = block_sum(lane == 0 ? sumsq : 0.0, red);
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = bs;
    part[2 * blockIdx.x + 1] = bq;
  }
}

__global__ void mc_final_kernel(const double* __restrict__ part, int nblk, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += part[2 * i], q += part[2 * i + 1];
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) out[0] = s, out[1] = q;
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
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
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
          normal4(philox4x32_10(Philox4{(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)d, (uint32_t)q}, k0, k1), z4[j]);
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

}  // namespace fbsnn

using namespace fbsnn;

extern "C" {

static long long g_mc_launches = 0;
long long mc_launch_count(void) { return g_mc_launches; }

size_t mc_scratch_bytes(void) { return (size_t)kMcMaxBlocks * 2 * sizeof(double); }

static int mc_make(const McSpec* spec, McK& k) {
  if (!spec || spec->D < 1 || spec->D > 32 * kMcMaxDJ || spec->N < 1) return FBSNN_E_UNSUPPORTED;
  k.D = spec->D, k.N = spec->N;
  k.drift_T = (float)(((double)spec->rate - 0.5 * (double)spec->sigma * (double)spec->sigma) * (double)spec->T);
  k.vol_sqrt_dt = (float)((double)spec->sigma * sqrt((double)spec->T / (double)spec->N));
  k.disc = (float)exp(-(double)spec->rate * (double)spec->T);
  k.strike = spec->strike;
  return 0;
}

int mc_basket_price(const McSpec* spec, const float* S0, const float* weights, const float* chol_T,
                    uint64_t n_paths, uint64_t seed, uint64_t path_offset, void* scratch, double* sums_out,
                    void* stream) {
  McK k;
  if (mc_make(spec, k) || !S0 || !weights || !scratch || !sums_out || n_paths == 0) return FBSNN_E_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int nwarp = kMcThreads / 32;
  size_t smem = (size_t)nwarp * k.D * sizeof(float);
  int chol_in_smem = 0;
  if (chol_T && smem + (size_t)k.D * k.D * sizeof(float) <= 160 * 1024) {
    chol_in_smem = 1;
    smem += (size_t)k.D * k.D * sizeof(float);
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mc_basket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return FBSNN_E_CUDA;
  }
  const unsigned long long want = (n_paths + nwarp - 1) / nwarp;
  const int blocks = (int)(want < (unsigned long long)kMcMaxBlocks ? want : (unsigned long long)kMcMaxBlocks);
  mc_basket_kernel<<<blocks, kMcThreads, smem, st>>>(k, S0, weights, chol_T, chol_in_smem, n_paths, path_offset,
                                                     seed, (double*)scratch);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  mc_final_kernel<<<1, 256, 0, st>>>((const double*)scratch, blocks, sums_out);
  if (cudaGetLastError() != cudaSuccess) return FBSNN_E_CUDA;
  g_mc_launches += 2;
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
