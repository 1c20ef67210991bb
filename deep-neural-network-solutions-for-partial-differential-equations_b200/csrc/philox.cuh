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

// Round keys of Philox4x32-10 for a fixed (k0, k1): loop-invariant, so callers that draw many blocks keep the
// twenty keys in registers instead of re-deriving them (two IADDs per round) for every block.
struct PhiloxKeys {
  uint32_t a[10], b[10];
};
__host__ __device__ __forceinline__ PhiloxKeys philox_keys(uint32_t k0, uint32_t k1) {
  PhiloxKeys k;
#pragma unroll
  for (int r = 0; r < 10; ++r) k.a[r] = k0 + 0x9E3779B9u * (uint32_t)r, k.b[r] = k1 + 0xBB67AE85u * (uint32_t)r;
  return k;
}
__device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, const PhiloxKeys& k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;   // one IMAD.WIDE each
    c = Philox4{(uint32_t)(p1 >> 32) ^ c.y ^ k.a[r], (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.b[r], (uint32_t)p0};
  }
  return c;
}

// Four standard normals from one Philox block (two Box-Muller pairs) on the special-function unit:
//   radius = sqrt(-2 ln u) = sqrt(-2 ln2 * lg2(u)), u in (0, 1) from 23 bits;  angle uniform on [-pi, pi), 23 bits
// lg2.approx / sqrt.approx / sin.approx / cos.approx are each one MUFU operation (abs. error <= 2^-21 on this range),
// ~7 instructions per normal instead of ~35 for logf + sqrtf + sincospif.  Every device generator (training
// increments, MC pricer, MC path tensor) uses this one function, so their streams stay path-wise identical.
__device__ __forceinline__ void normal4(const Philox4& r, float out[4]) {
#ifdef __CUDA_ARCH__
  // uniforms by mantissa stuffing (no int->float conversion, which would share the special-function unit with the
  // four MUFU operations below): as_float(0x3f800000 | v >> 9) is uniform on [1, 2)
  const float u0 = __uint_as_float(0x3f800000u | (r.x >> 9)) + -0.99999994f;   // (0, 1): f - 1 + 2^-24, exact
  const float u1 = __uint_as_float(0x3f800000u | (r.z >> 9)) + -0.99999994f;
  const float a0 = (__uint_as_float(0x3f800000u | (r.y >> 9)) - 1.5f) * 6.283185307179586f;   // [-pi, pi)
  const float a1 = (__uint_as_float(0x3f800000u | (r.w >> 9)) - 1.5f) * 6.283185307179586f;
  float l0, l1, r0, r1;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(u0));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(l0 * -1.3862943611198906f));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(l1 * -1.3862943611198906f));
  out[0] = r0 * __cosf(a0);
  out[1] = r0 * __sinf(a0);
  out[2] = r1 * __cosf(a1);
  out[3] = r1 * __sinf(a1);
#else
  (void)r, (void)out;
#endif
}

}  // namespace fbsnn
