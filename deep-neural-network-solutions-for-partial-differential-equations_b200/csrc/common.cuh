// Shared device helpers for the FBSNN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fbsnn_b200.h"

namespace fbsnn {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (result valid in thread 0).
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// internal activation codes of the TF32 variant: MUFU-based sin/cos after a two-term Cody-Waite reduction
// (abs. error ~5e-7, far below TF32's 2^-11 operand rounding) and tanh.approx
constexpr int kActSineFast = 3, kActTanhFast = 4;
// 3xTF32 variant: fp32-grade sin/cos in ~20 instructions (three-term Cody-Waite reduction by pi/2 + the Cephes
// minimax polynomials on [-pi/4, pi/4], ~1 ulp for |z| < 1e4) instead of the library sincosf (~50)
constexpr int kActSineCW = 5;

__device__ __forceinline__ void sincos_cw(float x, float& s, float& c) {
  const float k = rintf(x * 0.636619772367581343f);
  const int q = (int)k;
  float r = fmaf(k, -1.5703125f, x);
  r = fmaf(k, -4.837512969970703125e-4f, r);
  r = fmaf(k, -7.54978995489188216e-8f, r);
  const float r2 = r * r;
  float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2,
                  fmaf(-0.5f, r2, 1.0f));
  if (q & 1) { const float t = sp; sp = cp; cp = t; }
  s = (q & 2) ? -sp : sp;
  c = ((q + 1) & 2) ? -cp : cp;
}

// activation value g, first derivative a  (Functions/Sine.py:11-12, nn.ReLU, nn.Tanh)
__device__ __forceinline__ void act_ga(int act, float z, float& g, float& a) {
  if (act == FBSNN_ACT_SINE) {
    sincosf(z, &g, &a);
  } else if (act == kActSineCW) {
    sincos_cw(z, g, a);
  } else if (act == kActSineFast) {
    const float k = rintf(z * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, z);
    r = fmaf(-k, -1.7484555e-7f, r);
    g = __sinf(r);
    a = __cosf(r);
  } else if (act == kActTanhFast) {
    asm("tanh.approx.f32 %0, %1;" : "=f"(g) : "f"(z));
    a = 1.f - g * g;
  } else if (act == FBSNN_ACT_RELU) {
    g = fmaxf(z, 0.f);
    a = z > 0.f ? 1.f : 0.f;
  } else {
    g = tanhf(z);
    a = 1.f - g * g;
  }
}
// four elements at once: ONE dispatch on the (warp-uniform) activation code, then four independent evaluations the
// scheduler can interleave -- the fused epilogues are latency-bound on exactly this chain
__device__ __forceinline__ void act_ga4(int act, const float z[4], float g[4], float a[4]) {
  switch (act) {
    case kActSineCW:
#pragma unroll
      for (int i = 0; i < 4; ++i) sincos_cw(z[i], g[i], a[i]);
      break;
    case kActSineFast:
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float k = rintf(z[i] * 0.15915494309189535f);
        float r = fmaf(-k, 6.2831854820251465f, z[i]);
        r = fmaf(-k, -1.7484555e-7f, r);
        g[i] = __sinf(r);
        a[i] = __cosf(r);
      }
      break;
    case FBSNN_ACT_RELU:
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = fmaxf(z[i], 0.f), a[i] = z[i] > 0.f ? 1.f : 0.f;
      break;
    case kActTanhFast:
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        asm("tanh.approx.f32 %0, %1;" : "=f"(g[i]) : "f"(z[i]));
        a[i] = 1.f - g[i] * g[i];
      }
      break;
    case FBSNN_ACT_SINE:
#pragma unroll
      for (int i = 0; i < 4; ++i) sincosf(z[i], &g[i], &a[i]);
      break;
    default:
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i] = tanhf(z[i]), a[i] = 1.f - g[i] * g[i];
  }
}
// second derivative from (g, a): sine -g, relu 0, tanh -2 g a
__device__ __forceinline__ float act_c(int act, float g, float a) {
  if (act == FBSNN_ACT_SINE || act == kActSineFast || act == kActSineCW) return -g;
  if (act == FBSNN_ACT_RELU) return 0.f;
  return -2.f * g * a;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace fbsnn
