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

// activation value g, first derivative a  (Functions/Sine.py:11-12, nn.ReLU, nn.Tanh)
__device__ __forceinline__ void act_ga(int act, float z, float& g, float& a) {
  if (act == FBSNN_ACT_SINE) {
    sincosf(z, &g, &a);
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
// second derivative from (g, a): sine -g, relu 0, tanh -2 g a
__device__ __forceinline__ float act_c(int act, float g, float a) {
  if (act == FBSNN_ACT_SINE || act == kActSineFast) return -g;
  if (act == FBSNN_ACT_RELU) return 0.f;
  return -2.f * g * a;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

}  // namespace fbsnn
