// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and Box-Muller normals, written out here so
// that the FBSNN Brownian generator and the MC pricer draw streams keyed only by (seed, iteration, GLOBAL
// path id, step, dimension) -- results do not depend on how paths are sharded over GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fbsnn {

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = Philox4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// uint32 -> uniform in (0, 1]  (never 0, so log() is finite)
__host__ __device__ __forceinline__ float u01(uint32_t v) { return ((float)(v >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// four standard normals from one Philox block (two Box-Muller pairs)
__device__ __forceinline__ void normal4(const Philox4& r, float out[4]) {
  const float r0 = sqrtf(-2.0f * logf(u01(r.x)));
  const float r1 = sqrtf(-2.0f * logf(u01(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u01(r.y), &s0, &c0);
  sincospif(2.0f * u01(r.w), &s1, &c1);
  out[0] = r0 * c0;
  out[1] = r0 * s0;
  out[2] = r1 * c1;
  out[3] = r1 * s1;
}

}  // namespace fbsnn
