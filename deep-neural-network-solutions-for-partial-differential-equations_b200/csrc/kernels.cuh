// Element-wise and reduction kernels of the FBSNN step: Brownian increments, Euler-Maruyama path advance,
// output head, residual loss and its seeds, column sums, split-K reduction, NAIS projection, clip + Adam.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace fbsnn {

// lets a following dense-layer kernel (launched with programmatic stream serialisation, gemm_tc.cuh) begin its
// prologue while this kernel is still running; that kernel still waits for this one's completion before reading
__device__ __forceinline__ void pdl_trigger_next() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Problem constants handed to the path / loss kernels (closed enumeration).
struct ProblemK {
  int D, N, ldx;
  int mu_kind, sigma_kind, phi_kind, g_kind;
  float mu_c, sigma_c, phi_c, strike;
  int zt_dims;                              // terminal Z penalty over the first zt_dims components (= D normally)
  float h_kappa, h_theta, h_xi, h_rho;      // Heston only
};

// ----------------------------------------------------------------------------------------------------
// Brownian increments  (replaces np.random.normal + sqrt(dt) scaling of FBSNN.fetch_minibatch,
// DeepBSDE.py:252-255).  One Philox block -> 4 normals for 4 consecutive dimensions of one (path, step).
// Layout: inc[(m*(N+1) + n)*ldi + d] = increment from step n-1 to n  (row 0 is zero, like the reference DW).
// ----------------------------------------------------------------------------------------------------
__global__ void brownian_increments_kernel(float* __restrict__ inc, long long M, int N, int D, float sqrt_dt,
                                           int ldi, long long path_offset, uint64_t seed, uint64_t iteration,
                                           const long long* __restrict__ iter_dev) {
  // iter_dev (nullable): device-resident counter added to `iteration`, so that a captured CUDA graph draws a
  // fresh stream on every replay (the optimiser step counter is used for this).
  if (iter_dev) iteration += (uint64_t)iter_dev[0];
  const int D4 = (D + 3) / 4;
  const long long total = M * (long long)(N + 1) * D4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(i % D4);
    const long long row = i / D4;
    const int n = (int)(row % (N + 1));
    const long long m = row / (N + 1);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (n > 0) {
      const uint64_t gp = (uint64_t)(m + path_offset);
      const Philox4 ctr{(uint32_t)gp, (uint32_t)(gp >> 32), (uint32_t)(n * D4 + d4), (uint32_t)iteration};
      normal4(philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iteration >> 32)), z);
    }
    float* o = inc + row * ldi + d4 * 4;   // ldi % 4 == 0: columns [D, ldi) are zero padding
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d4 * 4 + j >= D) z[j] = 0.f;
    st4(o, make_float4(sqrt_dt * z[0], sqrt_dt * z[1], sqrt_dt * z[2], sqrt_dt * z[3]));
  }
}

// cumulative (t, W) in the reference's layout from increments (np.cumsum, DeepBSDE.py:257-258).
// One thread per (path, dim), sequential prefix sum along n.
__global__ void cumsum_paths_kernel(const float* inc, int ldi, float* W, float* t, long long M, int N, int D,
                                    float T) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= M * D) return;
  const long long m = idx / D;
  const int d = (int)(idx % D);
  float acc = 0.f;
  for (int n = 0; n <= N; ++n) {
    const long long r = m * (N + 1) + n;
    acc += inc[r * ldi + d];
    W[r * D + d] = acc;
    if (t && d == 0) t[r] = (float)((double)n * (double)T / (double)N);
  }
}

// ----------------------------------------------------------------------------------------------------
// Euler-Maruyama path advance (X recursion of FBSNN.loss_function, DeepBSDE.py:218-222).
// One thread per (path, dim), sequential over n.  Writes the network input rows xin = [t, X, 0-pad],
// sdw[row n] = sigma(X_n) dW_n (consumed by the loss kernels) and optionally X in the reference layout.
// Arithmetic order follows the reference exactly: X1 = (X0 + (mu_c*X0)*(t1-t0)) + (sigma*X0)*dW, no FMA
// contraction, dW and dt re-differenced in fp32 from the cumulative inputs.
// ----------------------------------------------------------------------------------------------------
struct PathArgs {
  const float* t;      // (M, N+1) or null (uniform grid n*T/N)
  const float* W;      // cumulative (M, N+1, D) or null
  const float* inc;    // increments (M, N+1, ldi) (row n = step n-1 -> n), used when W is null
  int ldi;
  const float* Xi;
  long long xi_rows, M;
  float T;
  float* xin;
  float* sdw;
  float* X_out;  // nullable
};

__global__ void path_advance_kernel(const ProblemK p, const PathArgs a) {
  pdl_trigger_next();
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= a.M * p.D) return;
  const long long m = idx / p.D;
  const int d = (int)(idx % p.D);
  const int N = p.N, D = p.D, ldx = p.ldx;
  float x = a.Xi[(a.xi_rows == 1 ? 0 : m) * D + d];
  const long long row0 = m * (N + 1);
  float tn = a.t ? a.t[row0] : 0.f;
  float wn = a.W ? a.W[row0 * D + d] : 0.f;
  auto write_row = [&](long long r) {
    float* xr = a.xin + r * ldx;
    xr[1 + d] = x;
    if (d == 0) xr[0] = tn;
    for (int c = D + 1 + d; c < ldx; c += D) xr[c] = 0.f;   // zero padding, spread over the path's D threads
    if (a.X_out) a.X_out[r * D + d] = x;
  };
  // The recursion is sequential in n, its inputs are not: the increments and times of the next PF steps are fetched
  // before the first of them is needed, so a step costs its arithmetic and not a memory round trip (M = 100: 36 -> 19 us).
  // The arithmetic and its order are untouched (X stays bit-identical to the reference).
  constexpr int PF = 8;
  for (int n0 = 0; n0 < N; n0 += PF) {
    float tv[PF], wv[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int n = n0 + k;
      if (n < N) {
        const long long r = row0 + n;
        tv[k] = a.t ? __ldg(a.t + r + 1) : (float)((double)(n + 1) * (double)a.T / (double)N);
        wv[k] = a.W ? __ldg(a.W + (r + 1) * D + d) : __ldg(a.inc + (r + 1) * a.ldi + d);
      }
    }
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int n = n0 + k;
      if (n < N) {
        const long long r = row0 + n;
        write_row(r);
        const float tn1 = tv[k];
        const float dt = __fsub_rn(tn1, tn);
        float dw;
        if (a.W) {
          dw = __fsub_rn(wv[k], wn);
          wn = wv[k];
        } else {
          dw = wv[k];
        }
        const float sig = p.sigma_kind == FBSNN_SIGMA_PROP ? __fmul_rn(p.sigma_c, x) : p.sigma_c;
        const float sd = __fmul_rn(sig, dw);
        a.sdw[r * D + d] = sd;
        const float mu = p.mu_kind == FBSNN_MU_LINEAR ? __fmul_rn(p.mu_c, x) : 0.f;
        x = __fadd_rn(__fadd_rn(x, __fmul_rn(mu, dt)), sd);
        tn = tn1;
      }
    }
  }
  write_row(row0 + N);
}

// Heston 2-factor path advance (heston_dnnpde.py:587-605, 636-641): state (S, v), ONE Brownian driver W (M, N+1, 1)
// that the reference's einsum broadcasts over both diffusion columns, so
//   S' = S + clamp(mu_c S) dt + (s00 + s01) dW,   v' = v + clamp(kappa (theta - v)) dt + (s10 + s11) dW
// with s00 = sqrt(max(v,1e-8)) S, s11 = xi sqrt(max(v,1e-8)), s01 = rho s11, s10 = rho s00, every entry clamped
// to +-100.  One thread per path; sdw[row] = [(s00+s01) dW, (s10+s11) dW] feeds the generic loss kernels.
__device__ __forceinline__ float clamp100(float x) { return fminf(fmaxf(x, -100.f), 100.f); }
__global__ void path_advance_heston_kernel(const ProblemK p, const PathArgs a) {
  pdl_trigger_next();
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= a.M) return;
  const int N = p.N, ldx = p.ldx;
  float S = a.Xi[(a.xi_rows == 1 ? 0 : m) * 2 + 0], v = a.Xi[(a.xi_rows == 1 ? 0 : m) * 2 + 1];
  const long long row0 = m * (N + 1);
  float tn = a.t ? a.t[row0] : 0.f;
  float wn = a.W ? a.W[row0] : 0.f;
  for (int n = 0; n <= N; ++n) {
    const long long r = row0 + n;
    float* xr = a.xin + r * ldx;
    xr[0] = tn, xr[1] = S, xr[2] = v;
    for (int c = 3; c < ldx; ++c) xr[c] = 0.f;
    if (a.X_out) a.X_out[r * 2] = S, a.X_out[r * 2 + 1] = v;
    if (n == N) break;
    const float tn1 = a.t ? a.t[r + 1] : (float)((double)(n + 1) * (double)a.T / (double)N);
    const float dt = __fsub_rn(tn1, tn);
    float dw;
    if (a.W) {
      const float wn1 = a.W[r + 1];
      dw = __fsub_rn(wn1, wn);
      wn = wn1;
    } else {
      dw = a.inc[(r + 1) * a.ldi];
    }
    const float sq = sqrtf(fmaxf(v, 1e-8f));
    const float sS = __fmul_rn(sq, S), sv = __fmul_rn(p.h_xi, sq);
    const float s00 = clamp100(sS), s01 = clamp100(__fmul_rn(p.h_rho, sv));
    const float s10 = clamp100(__fmul_rn(p.h_rho, sS)), s11 = clamp100(sv);
    const float d0 = __fadd_rn(__fmul_rn(s00, dw), __fmul_rn(s01, dw));
    const float d1 = __fadd_rn(__fmul_rn(s10, dw), __fmul_rn(s11, dw));
    a.sdw[r * 2] = d0, a.sdw[r * 2 + 1] = d1;
    const float muS = clamp100(__fmul_rn(p.mu_c, S));
    const float muv = clamp100(__fmul_rn(p.h_kappa, __fsub_rn(p.h_theta, v)));
    S = __fadd_rn(__fadd_rn(S, __fmul_rn(muS, dt)), d0);
    v = __fadd_rn(__fadd_rn(v, __fmul_rn(muv, dt)), d1);
    tn = tn1;
  }
}

// u = max(u, 0) and Du <- Du * 1{u >= 0} (torch.clamp(u, min=0) and its autograd mask, heston_dnnpde.py:568-576).
// mask (nullable) keeps 1{u_raw >= 0} for the seed kernel.
__global__ void clamp_u_kernel(float* __restrict__ Y, float* __restrict__ zf, int ldx, long long rows,
                               float* __restrict__ mask) {
  pdl_trigger_next();
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const bool keep = Y[r] >= 0.f;
  if (!keep)
    for (int c = lane; c < ldx; c += 32) zf[r * ldx + c] = 0.f;
  __syncwarp();
  if (lane == 0) {
    if (!keep) Y[r] = 0.f;
    if (mask) mask[r] = keep ? 1.f : 0.f;
  }
}

// rows of arbitrary (t, X) for net_u: xin = [t, X, 0-pad]
__global__ void pack_rows_kernel(const float* t, const float* X, long long rows, int D, int ldx, float* xin) {
  pdl_trigger_next();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows * ldx) return;
  const long long r = i / ldx;
  const int c = (int)(i % ldx);
  xin[i] = c == 0 ? t[r] : (c <= D ? X[r * D + c - 1] : 0.f);
}

// dst (rows x ldd) = [src (rows x cols) | 0]   (TF32 variant: zero-padded homes of the input-width matrices)
__global__ void pad_copy_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst, int ldd) {
  pdl_trigger_next();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ldd) return;
  const int r = i / ldd, c = i % ldd;
  dst[i] = c < cols ? src[r * cols + c] : 0.f;
}

// 3xTF32 variant: hi = x with the 13 low mantissa bits cleared (exactly what the tensor core reads), lo = x - hi
__global__ void split_hi_lo_kernel(const float* __restrict__ src, int n, float* __restrict__ hi, float* __restrict__ lo) {
  pdl_trigger_next();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = src[i];
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = x - h;
}

// u[r] = h_L[r,:] . wout + bout      (one warp per row)
__global__ void head_kernel(const float* __restrict__ h, int ld, int H, const float* __restrict__ wout,
                            const float* __restrict__ bout, long long rows, float* __restrict__ u) {
  pdl_trigger_next();
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int c = lane; c < H; c += 32) acc = fmaf(h[r * ld + c], __ldg(wout + c), acc);
  acc = warp_sum(acc);
  if (lane == 0) u[r] = acc + bout[0];
}

// copy column block [1, D] of ZF (rows, ldx) / scalar Y into user-visible outputs
__global__ void gather_z_kernel(const float* __restrict__ zf, int ldx, int D, long long rows, float* __restrict__ Z) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows * D) return;
  Z[i] = zf[(i / D) * ldx + 1 + (i % D)];
}

// ----------------------------------------------------------------------------------------------------
// Residual loss (FBSNN.loss_function, DeepBSDE.py:223-240) -- one warp per (path, step) row.
//   n < N : e_n = Y_{n+1} - (Y_n + phi(X_n,Y_n,Z_n) dt + Z_n . sdw_n)
//   n = N : e_N = Y_N - g(X_N)   and   sum_d (Z_N - grad g(X_N))^2
// ev[r] keeps the residual for the seed kernel; per-block partial sums of squares go to loss_part.
// ----------------------------------------------------------------------------------------------------
struct LossArgs {
  const float* xin;
  const float* zf;
  const float* sdw;
  const float* Y;
  long long rows;
  float* ev;
  float* ybar;
  float* V;
  float* part;   // [2][gridDim.x] partial sums: loss, sum(ybar)
  const float* umask;   // nullable: 1{u_raw >= 0} per row (clamp_u): scales the seeds
};

__device__ __forceinline__ void terminal_g(const ProblemK& p, float sumx, float sumx2, float& g, float& dg_scale) {
  // g(X_N) and the scalar s.t. grad g = dg_scale * X (sumsq, logq) or dg_scale * 1 (calls)
  if (p.g_kind == FBSNN_G_SUMSQ) {
    g = sumx2;
    dg_scale = 2.f;
  } else if (p.g_kind == FBSNN_G_LOGQ) {
    const float q = 0.5f + 0.5f * sumx2;
    g = logf(q);
    dg_scale = 1.f / q;
  } else {
    const float base = p.g_kind == FBSNN_G_CALL_SUM ? sumx : sumx / (float)p.D;
    const float sc = p.g_kind == FBSNN_G_CALL_SUM ? 1.f : 1.f / (float)p.D;
    g = fmaxf(base - p.strike, 0.f);
    dg_scale = base > p.strike ? sc : (base == p.strike ? 0.5f * sc : 0.f);  // torch.maximum splits ties
  }
}
// payoffs of the first state component only (Heston: X = (S, v)): g(S) and dg/dS
__device__ __forceinline__ void terminal_g_first(const ProblemK& p, float x0, float& g, float& dg) {
  const float s = x0 - p.strike;
  if (p.g_kind == FBSNN_G_CALL_FIRST) {
    g = fmaxf(s, 0.f);
    dg = s > 0.f ? 1.f : (s == 0.f ? 0.5f : 0.f);
  } else {                                   // (S-K) / (1 + exp(-alpha (S-K))), alpha = 10
    const float q = 1.f / (1.f + expf(-10.f * s));   // sigmoid(10 s)
    g = s * q;
    dg = q + 10.f * s * q * (1.f - q);               // = q + 10 s e / (1+e)^2 without the inf/inf at s << 0
  }
}
__device__ __forceinline__ bool g_is_first(const ProblemK& p) {
  return p.g_kind == FBSNN_G_CALL_FIRST || p.g_kind == FBSNN_G_CALL_FIRST_SMOOTH;
}

__global__ void loss_residual_kernel(const ProblemK p, const LossArgs a) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  float contrib = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; r < a.rows; r += wstride) {
    const int n = (int)(r % (p.N + 1));
    const float* x = a.xin + r * p.ldx + 1;
    const float* z = a.zf + r * p.ldx + 1;
    if (n < p.N) {
      const float* sd = a.sdw + r * p.D;
      float zs = 0.f, xz = 0.f, z2 = 0.f;
      for (int d = lane; d < p.D; d += 32) {
        const float zv = z[d];
        zs = fmaf(zv, sd[d], zs);
        xz = fmaf(x[d], zv, xz);
        z2 = fmaf(zv, zv, z2);
      }
      zs = warp_sum(zs), xz = warp_sum(xz), z2 = warp_sum(z2);
      const float y = a.Y[r];
      const float dt = a.xin[(r + 1) * p.ldx] - a.xin[r * p.ldx];
      float phi;
      if (p.phi_kind == FBSNN_PHI_BSB) phi = p.phi_c * (y - xz);
      else if (p.phi_kind == FBSNN_PHI_RY) phi = p.phi_c * y;
      else phi = z2;
      const float e = a.Y[r + 1] - (y + phi * dt + zs);
      if (lane == 0) { a.ev[r] = e; contrib += e * e; }
    } else {
      float sx = 0.f, sx2 = 0.f;
      for (int d = lane; d < p.D; d += 32) { const float xv = x[d]; sx += xv; sx2 = fmaf(xv, xv, sx2); }
      sx = warp_sum(sx), sx2 = warp_sum(sx2);
      float g, dgs;
      const bool first = g_is_first(p);
      if (first) terminal_g_first(p, x[0], g, dgs);
      else terminal_g(p, sx, sx2, g, dgs);
      const bool mulx = p.g_kind == FBSNN_G_SUMSQ || p.g_kind == FBSNN_G_LOGQ;
      float zt = 0.f;
      for (int d = lane; d < p.zt_dims; d += 32) {
        const float dg = first ? (d == 0 ? dgs : 0.f) : (mulx ? dgs * x[d] : dgs);
        const float diff = z[d] - dg;
        zt = fmaf(diff, diff, zt);
      }
      zt = warp_sum(zt);
      const float e = a.Y[r] - g;
      if (lane == 0) { a.ev[r] = e; contrib += e * e + zt; }
    }
  }
  const float tot = block_sum(contrib, red);
  if (threadIdx.x == 0) a.part[blockIdx.x] = tot;
}

// Seeds of the reverse sweeps: ybar = dL/dY, V = [0, dL/dZ, 0-pad] per row.
__global__ void loss_seed_kernel(const ProblemK p, const LossArgs a) {
  pdl_trigger_next();
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  float ybsum = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; r < a.rows; r += wstride) {
    float yb = 0.f;
    const int n = (int)(r % (p.N + 1));
    const float* x = a.xin + r * p.ldx + 1;
    const float* z = a.zf + r * p.ldx + 1;
    float* v = a.V + r * p.ldx;
    const float e = a.ev[r];
    if (n < p.N) {
      const float* sd = a.sdw + r * p.D;
      const float dt = a.xin[(r + 1) * p.ldx] - a.xin[r * p.ldx];
      const float phi_y = p.phi_kind == FBSNN_PHI_ZSQ ? 0.f : p.phi_c;
      yb = -2.f * e * (1.f + phi_y * dt);
      if (n > 0) yb += 2.f * a.ev[r - 1];
      for (int d = lane; d < p.D; d += 32) {
        float pz;
        if (p.phi_kind == FBSNN_PHI_BSB) pz = -p.phi_c * x[d];
        else if (p.phi_kind == FBSNN_PHI_RY) pz = 0.f;
        else pz = 2.f * z[d];
        v[1 + d] = -2.f * e * (pz * dt + sd[d]);
      }
    } else {
      float sx = 0.f, sx2 = 0.f;
      for (int d = lane; d < p.D; d += 32) { const float xv = x[d]; sx += xv; sx2 = fmaf(xv, xv, sx2); }
      sx = warp_sum(sx), sx2 = warp_sum(sx2);
      float g, dgs;
      const bool first = g_is_first(p);
      if (first) terminal_g_first(p, x[0], g, dgs);
      else terminal_g(p, sx, sx2, g, dgs);
      const bool mulx = p.g_kind == FBSNN_G_SUMSQ || p.g_kind == FBSNN_G_LOGQ;
      yb = 2.f * e + (p.N > 0 ? 2.f * a.ev[r - 1] : 0.f);
      for (int d = lane; d < p.D; d += 32) {
        const float dg = first ? (d == 0 ? dgs : 0.f) : (mulx ? dgs * x[d] : dgs);
        v[1 + d] = d < p.zt_dims ? 2.f * (z[d] - dg) : 0.f;
      }
    }
    if (a.umask) {                            // clamp_u: d max(u,0)/du and d(mask Du)/dDu
      const float mk = a.umask[r];
      yb *= mk;
      __syncwarp();
      for (int d = lane; d < p.D; d += 32) v[1 + d] *= mk;
    }
    if (lane == 0) {
      v[0] = 0.f;
      for (int c = p.D + 1; c < p.ldx; ++c) v[c] = 0.f;
      a.ybar[r] = yb;
      ybsum += yb;
    }
  }
  const float tot = block_sum(lane == 0 ? ybsum : 0.f, red);
  if (threadIdx.x == 0) a.part[gridDim.x + blockIdx.x] = tot;
}

// ----------------------------------------------------------------------------------------------------
// Fused residual + seeds (training, D <= 128): one warp walks ONE PATH through its N+1 rows in order, so the
// residual e_{n-1} that the seed of row n needs is still in a register -- xin, zf and sdw are read once instead
// of twice and the ev[] round trip disappears.  The next row's operands are loaded before the current row's
// warp reductions (two rows in flight per warp).  Same arithmetic as the two kernels above.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_path_kernel(const ProblemK p, const LossArgs a, long long n_paths) {
  pdl_trigger_next();
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const int N = p.N, D = p.D, ldx = p.ldx;
  float contrib = 0.f, ybsum = 0.f;
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  const bool first = g_is_first(p);
  const bool mulx = p.g_kind == FBSNN_G_SUMSQ || p.g_kind == FBSNN_G_LOGQ;
  for (long long m = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; m < n_paths; m += wstride) {
    const long long row0 = m * (N + 1);
    float xc[4], zc[4], sc[4], xn[4], zn[4], sn[4];
    auto load_row = [&](long long r, bool with_sd, float* x, float* z, float* sd) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        const bool ok = d < D;
        x[j] = ok ? a.xin[r * ldx + 1 + d] : 0.f;
        z[j] = ok ? a.zf[r * ldx + 1 + d] : 0.f;
        sd[j] = (ok && with_sd) ? a.sdw[r * D + d] : 0.f;
      }
    };
    load_row(row0, N > 0, xc, zc, sc);
    float t_cur = a.xin[row0 * ldx], y_cur = a.Y[row0], e_prev = 0.f;
    for (int n = 0; n <= N; ++n) {
      const long long r = row0 + n;
      float t_next = 0.f, y_next = 0.f;
      if (n < N) {
        load_row(r + 1, n + 1 < N, xn, zn, sn);
        t_next = a.xin[(r + 1) * ldx], y_next = a.Y[r + 1];
      }
      float e, yb;
      float vrow[4];
      if (n < N) {
        float zs = 0.f, xz = 0.f, z2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) zs = fmaf(zc[j], sc[j], zs), xz = fmaf(xc[j], zc[j], xz), z2 = fmaf(zc[j], zc[j], z2);
        zs = warp_sum(zs), xz = warp_sum(xz), z2 = warp_sum(z2);
        const float dt = t_next - t_cur;
        float phi;
        if (p.phi_kind == FBSNN_PHI_BSB) phi = p.phi_c * (y_cur - xz);
        else if (p.phi_kind == FBSNN_PHI_RY) phi = p.phi_c * y_cur;
        else phi = z2;
        e = y_next - (y_cur + phi * dt + zs);
        contrib += e * e;
        const float phi_y = p.phi_kind == FBSNN_PHI_ZSQ ? 0.f : p.phi_c;
        yb = -2.f * e * (1.f + phi_y * dt);
        if (n > 0) yb += 2.f * e_prev;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float pz;
          if (p.phi_kind == FBSNN_PHI_BSB) pz = -p.phi_c * xc[j];
          else if (p.phi_kind == FBSNN_PHI_RY) pz = 0.f;
          else pz = 2.f * zc[j];
          vrow[j] = -2.f * e * (pz * dt + sc[j]);
        }
      } else {
        float sx = 0.f, sx2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) sx += xc[j], sx2 = fmaf(xc[j], xc[j], sx2);
        sx = warp_sum(sx), sx2 = warp_sum(sx2);
        float g, dgs;
        if (first) terminal_g_first(p, __shfl_sync(0xffffffffu, xc[0], 0), g, dgs);
        else terminal_g(p, sx, sx2, g, dgs);
        float zt = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int d = lane + 32 * j;
          const float dg = first ? (d == 0 ? dgs : 0.f) : (mulx ? dgs * xc[j] : dgs);
          const float diff = zc[j] - dg;
          const bool in = d < p.zt_dims;
          if (in) zt = fmaf(diff, diff, zt);
          vrow[j] = in ? 2.f * diff : 0.f;
        }
        zt = warp_sum(zt);
        e = y_cur - g;
        contrib += e * e + zt;
        yb = 2.f * e + (N > 0 ? 2.f * e_prev : 0.f);
      }
      if (a.umask) {
        const float mk = a.umask[r];
        yb *= mk;
#pragma unroll
        for (int j = 0; j < 4; ++j) vrow[j] *= mk;
      }
      float* v = a.V + r * ldx;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        if (d < D) v[1 + d] = vrow[j];
      }
      if (lane == 0) {
        v[0] = 0.f;
        a.ybar[r] = yb;
        ybsum += yb;
      }
      for (int c = D + 1 + lane; c < ldx; c += 32) v[c] = 0.f;
      e_prev = e;
      t_cur = t_next, y_cur = y_next;
#pragma unroll
      for (int j = 0; j < 4; ++j) xc[j] = xn[j], zc[j] = zn[j], sc[j] = sn[j];
    }
  }
  // contrib is warp-uniform (every lane holds the reduced values): count it once per warp
  const float tot = block_sum(lane == 0 ? contrib : 0.f, red);
  const float tys = block_sum(lane == 0 ? ybsum : 0.f, red);
  if (threadIdx.x == 0) a.part[blockIdx.x] = tot, a.part[gridDim.x + blockIdx.x] = tys;
}

// out_j[0] = (float) sum_i part[j*n + i], j = 0,1   (single block, double accumulation, deterministic)
struct FinalSum2 {
  float* out0;
  float* out1;
};
__global__ void final_sum2_kernel(const float* __restrict__ part, int n, FinalSum2 o) {
  pdl_trigger_next();
  __shared__ double red[32];
  for (int j = 0; j < 2; ++j) {
    float* out = j == 0 ? o.out0 : o.out1;
    if (!out) continue;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)part[(size_t)j * n + i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = (float)acc;
  }
}

// ----------------------------------------------------------------------------------------------------
// Column sums over rows (bias gradients, output-layer weight gradient), two deterministic stages.
//   job j: out[c] = sum_r ( A[r*ld + c] + (B ? y[r] * B[r*ld + c] : 0) )
// ----------------------------------------------------------------------------------------------------
struct ColJob {
  const float* A;   // null: the per-CTA partials were already written by the tcgen05 sweep epilogue (fused)
  const float* B;   // nullable
  const float* y;   // with B
  float* out;       // final destination(s)
  float* out2;      // nullable second destination (NAIS: layer{l}.bias and layer{l}_input.bias share a gradient)
  float* part;      // [nblk][max_width] partial sums of this job
  int nblk;
  int ld, width;
};
constexpr int kMaxColJobs = 12;
struct ColJobs {
  ColJob job[kMaxColJobs];
  int njobs;
  long long rows;
  int rows_per_block;
  int max_width;
};

// stage 1: 256 threads = 4 row groups x 64 column quads; every thread streams float4 rows with four loads in flight,
// the four row groups are combined through shared memory in a fixed order (widths and ld are multiples of 4)
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const ColJobs js) {
  const ColJob& j = js.job[blockIdx.y];
  if (!j.A) return;
  __shared__ float4 red[4][64];
  const long long r0 = (long long)blockIdx.x * js.rows_per_block;
  const long long r1 = min(js.rows, r0 + js.rows_per_block);
  const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
  for (int c0 = 0; c0 < j.width; c0 += 256) {
    const int c = c0 + 4 * cg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < j.width) {
      long long r = r0 + rg;
      for (; r + 12 < r1; r += 16) {
        float4 a[4], b[4];
        float y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = ld4(j.A + (r + 4 * u) * j.ld + c);
          if (j.B) b[u] = ld4(j.B + (r + 4 * u) * j.ld + c), y[u] = j.y[r + 4 * u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j.B) a[u].x = fmaf(y[u], b[u].x, a[u].x), a[u].y = fmaf(y[u], b[u].y, a[u].y),
                   a[u].z = fmaf(y[u], b[u].z, a[u].z), a[u].w = fmaf(y[u], b[u].w, a[u].w);
          acc.x += a[u].x, acc.y += a[u].y, acc.z += a[u].z, acc.w += a[u].w;
        }
      }
      for (; r < r1; r += 4) {
        float4 a = ld4(j.A + r * j.ld + c);
        if (j.B) {
          const float4 b = ld4(j.B + r * j.ld + c);
          const float y = j.y[r];
          a.x = fmaf(y, b.x, a.x), a.y = fmaf(y, b.y, a.y), a.z = fmaf(y, b.z, a.z), a.w = fmaf(y, b.w, a.w);
        }
        acc.x += a.x, acc.y += a.y, acc.z += a.z, acc.w += a.w;
      }
    }
    red[rg][cg] = acc;
    __syncthreads();
    if (rg == 0 && c < j.width) {
      float4 t = red[0][cg];
#pragma unroll
      for (int g = 1; g < 4; ++g) t.x += red[g][cg].x, t.y += red[g][cg].y, t.z += red[g][cg].z, t.w += red[g][cg].w;
      st4(j.part + (size_t)blockIdx.x * js.max_width + c, t);
    }
    __syncthreads();
  }
}
__global__ void colsum_stage2_kernel(const ColJobs js) {
  const ColJob& j = js.job[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= j.width) return;
  float acc = 0.f;
  for (int b = 0; b < j.nblk; ++b) acc += j.part[(size_t)b * js.max_width + c];
  j.out[c] = acc;
  if (j.out2) j.out2[c] = acc;
}

// out[o*ld_out + i] = sum_z part[z][o][i_pad]   (weight-gradient split-K second stage)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int nsplit, int rows, int cols_pad, int cols,
                                       float* __restrict__ out, int ld_out) {
  pdl_trigger_next();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols_pad) return;
  const int o = idx / cols_pad, i = idx % cols_pad;
  if (i >= cols) return;
  float acc = 0.f;
  for (int z = 0; z < nsplit; ++z) acc += part[(size_t)z * rows * cols_pad + idx];
  out[(size_t)o * ld_out + i] = acc;
}

// the same for up to kMaxRedJobs layers in one launch (blockIdx.y = job): the FC network's weight-gradient
// contractions run back to back into per-layer partial buffers and are reduced together
constexpr int kMaxRedJobs = 10;
struct RedJob {
  const float* part;
  float* out;
  int rows, cols_pad, cols, ld_out;
};
struct RedJobs {
  RedJob job[kMaxRedJobs];
  int njobs, nsplit;
};
__global__ void reduce_partials_batched_kernel(const RedJobs js) {
  const RedJob& j = js.job[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= j.rows * j.cols_pad) return;
  const int o = idx / j.cols_pad, i = idx % j.cols_pad;
  if (i >= j.cols) return;
  float acc = 0.f;
  for (int z = 0; z < js.nsplit; ++z) acc += j.part[(size_t)z * j.rows * j.cols_pad + idx];
  j.out[(size_t)o * j.ld_out + i] = acc;
}

// One launch for the per-iteration weight preparation of the tensor-core variants: job = (src rows x cols) ->
// zero-padded copy (rows x ld) [TF32 input-width matrices] and / or its exact-TF32 hi / lo twins [3xTF32]
constexpr int kMaxPrepJobs = 20;
struct PrepJob {
  const float* src;
  float* pad;   // nullable
  float* hi;    // nullable (with lo)
  float* lo;
  int rows, cols, ld;
};
struct PrepJobs {
  PrepJob job[kMaxPrepJobs];
  int njobs;
};
__global__ void prep_weights_kernel(const PrepJobs js) {
  pdl_trigger_next();
  const PrepJob& j = js.job[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.rows * j.ld; i += gridDim.x * blockDim.x) {
    const int r = i / j.ld, c = i % j.ld;
    const float x = c < j.cols ? j.src[(size_t)r * j.cols + c] : 0.f;
    if (j.pad) j.pad[i] = x;
    if (j.hi) {
      const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      j.hi[i] = h;
      j.lo[i] = x - h;
    }
  }
}

// ----------------------------------------------------------------------------------------------------
// NAIS-Net stability projection (Functions/naisnet.py:30-39), once per iteration instead of per net_u call:
//   R = W^T W;  n = ||R||_F;  s = sqrt(delta)/sqrt(n) if n > delta else 1;  Bm = -(s R + eps I)
// state[0] = n, state[1] = s, state[2] = 1 if scaled.
// Grid of ceil(H*H / 1024) blocks of 1024 threads: EVERY block reduces the whole matrix (the same thread -> element mapping
// and the same block_sum in each, so all blocks hold the bit-identical norm, and the value is the one a single block would
// compute), then scales its own 1024 elements -- one launch, no grid-wide barrier, 8 us instead of 45 (H = 256), which is what
// the small-batch NAIS-Net steps were spending a third of their time on.
// ----------------------------------------------------------------------------------------------------
__global__ void nais_project_kernel(const float* __restrict__ Rm, int H, float eps, float* __restrict__ Bm,
                                    float* __restrict__ state) {
  __shared__ double red[32];
  __shared__ float sh_s;
  double acc = 0.0;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) acc += (double)Rm[i] * (double)Rm[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float n = (float)sqrt(acc);
    const float delta = 1.f - 2.f * eps;
    const bool scaled = n > delta;
    const float s = scaled ? sqrtf(delta) / sqrtf(n) : 1.f;
    if (blockIdx.x == 0) state[0] = n, state[1] = s, state[2] = scaled ? 1.f : 0.f;
    sh_s = s;
  }
  __syncthreads();
  const float s = sh_s;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * H) {
    const int r = i / H, c = i % H;
    Bm[i] = -(s * Rm[i] + (r == c ? eps : 0.f));
  }
}
// Backward of the projection: Sm = Rbar + Rbar^T with Rbar = s*(-Bbar) - [scaled] 0.5 s <-Bbar, R>/n^2 R;
// the caller then forms Wbar = W * Sm with the GEMM.  Same launch shape as nais_project_kernel.
__global__ void nais_project_bwd_kernel(const float* __restrict__ Bbar, const float* __restrict__ Rm, int H,
                                        const float* __restrict__ state, float* __restrict__ Sm) {
  __shared__ double red[32];
  __shared__ float sh_k;
  double acc = 0.0;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) acc -= (double)Bbar[i] * (double)Rm[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float n = state[0], s = state[1];
    sh_k = state[2] != 0.f ? (float)(0.5 * (double)s * acc / ((double)n * (double)n)) : 0.f;
  }
  __syncthreads();
  const float s = state[1], k = sh_k;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * H) {
    const int r = i / H, c = i % H, it = c * H + r;
    Sm[i] = (-s * Bbar[i] - k * Rm[i]) + (-s * Bbar[it] - k * Rm[it]);
  }
}

// ----------------------------------------------------------------------------------------------------
// clip_grad_norm_(max_norm) + Adam (torch.optim.Adam defaults; with_corr_high_dimension_pde.py:424-425)
// opt_state (device, 64 B): [0] int64 step | [8] float clip_coef | [12] float step_size | [16] float bc2_sqrt
//                            | [20] float grad_norm | [24] int64 Philox iteration counter (advanced with the step)
// ----------------------------------------------------------------------------------------------------
struct OptState {
  long long step;                                    // Adam step (reset when a new optimiser is created)
  float clip_coef, step_size, bc2_sqrt, grad_norm;
  long long rng_iter;                                // Philox iteration counter of fbsnn_train_step (never reset)
  int skip;                                          // 1: this iteration's gradient was not finite, no update
  long long epoch;                                   // optimiser steps taken so far, NEVER rewound: epoch of the peer all-reduce's
                                                     // cross-GPU barrier (rng_iter is rewound after a graph-capture dry run)
};

__global__ void gradsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fmaf(g[i], g[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// ----------------------------------------------------------------------------------------------------
// Fused gradient all-reduce over NVLink peer memory (multi-GPU training step): ONE kernel does the cross-GPU
// barrier, the reduction over peer buffers and the squared norm for clip_grad_norm_; Adam follows on the same
// stream.  Replaces ncclAllReduce + the norm kernel; the payload (0.9 MB) is latency-bound, so one pass of 16-byte
// P2P loads is the minimum.
//
// Every rank's gradient kernels write into a symmetric buffer that all peers have mapped:
//   floats [0, n_grad) gradient | [n_grad] loss | ... | u32 flags at float offset flag_off: ready[16], done[16]
// Protocol for epoch e = (OptState::epoch + 1), identical on all ranks and never rewound:
//   1. block 0 stores e into ready[rank] of every peer (release.sys: this rank's gradient kernels finished
//      earlier on the stream); every block spins (acquire.sys) until its own ready[0..W) >= e
//   2. all blocks sum the W buffers in rank order 0..W-1 (so every rank gets bit-identical sums) with L1-bypassing
//      loads, store locally, accumulate the squared norm
//   3. the last block to finish stores e into done[rank] of every peer; peer_wait_kernel at the start of the next
//      iteration spins until done[0..W) >= e before any gradient kernel overwrites the buffer
// Spins are bounded by a wall-clock limit and trap instead of hanging the GPU.
// ----------------------------------------------------------------------------------------------------
constexpr int kPeerMaxWorld = 16;
constexpr unsigned long long kPeerTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_spin(const unsigned* flag, unsigned epoch) {
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if (global_ns() - t0 > kPeerTimeoutNs) {
      printf("fbsnn peer all-reduce: timed out waiting for a peer flag (have %u, want %u)\n", ld_acquire_sys(flag), epoch);
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ float4 ld4_peer(const float* p) {   // L1 caches peer lines; always go to the owner
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
struct PeerArgs {
  float* const* peers;       // device array [world]: every rank's symmetric buffer as mapped here
  int world, rank;
  long long n4;              // float4 count to reduce (gradient + the float4 carrying the loss)
  long long n_grad;
  long long flag_off;        // float offset of the flag block
  const OptState* st;
  unsigned* counter;         // local: blocks finished (self-resetting)
};
__global__ void peer_reduce_kernel(PeerArgs a, float* __restrict__ out, float* __restrict__ part) {
  __shared__ float red[32];
  const unsigned epoch = (unsigned)(a.st->epoch + 1);
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(a.peers[threadIdx.x] + a.flag_off) + a.rank, epoch);
  }
  if (threadIdx.x < a.world)
    peer_spin(reinterpret_cast<const unsigned*>(a.peers[a.rank] + a.flag_off) + threadIdx.x, epoch);
  __syncthreads();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = ld4_peer(a.peers[0] + 4 * i);
    for (int r = 1; r < a.world; ++r) {
      const float4 v = ld4_peer(a.peers[r] + 4 * i);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
    if (4 * i < a.n_grad) acc = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, fmaf(s.w, s.w, acc))));
  }
  acc = block_sum(acc, red);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    part[blockIdx.x] = acc;
    __threadfence();
    last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < a.world) {
    if (threadIdx.x == 0) *a.counter = 0;
    st_release_sys(reinterpret_cast<unsigned*>(a.peers[threadIdx.x] + a.flag_off) + kPeerMaxWorld + a.rank, epoch);
  }
}
// start of an iteration: every peer has finished reading this rank's buffer in the previous epoch
__global__ void peer_wait_kernel(const float* local_buf, long long flag_off, int world, const OptState* st) {
  const unsigned epoch = (unsigned)st->epoch;
  if (epoch != 0 && threadIdx.x < world)
    peer_spin(reinterpret_cast<const unsigned*>(local_buf + flag_off) + kPeerMaxWorld + threadIdx.x, epoch);
}

// ----------------------------------------------------------------------------------------------------
// min_loss_state of the reference's train() (with_corr_high_dimension_pde.py:431-433) without a host round trip:
//   if loss < best: best = loss, best_iter = iter, (X, Y) -> (X_best, Y_best)
// track_min_decide (1 thread) sets the flag, track_min_copy copies only when it is set (no traffic otherwise); the
// comparison is false for a NaN loss, like the reference's `loss < min_loss`.
// state (32 bytes): [0] best loss (float, +inf initially) | [1] flag (int) | [2] best call index (int, -1) | [3] call
// counter (int) | [4..5] int64: Philox iteration the best step drew its increments with (opt_state's counter - 1)
// ----------------------------------------------------------------------------------------------------
__global__ void track_min_decide_kernel(const float* __restrict__ loss, float* __restrict__ state,
                                        const OptState* __restrict__ opt) {
  int* st = reinterpret_cast<int*>(state);
  const float l = loss[0];
  const bool better = l < state[0];
  st[1] = better ? 1 : 0;
  if (better) {
    state[0] = l, st[2] = st[3];
    if (opt) *reinterpret_cast<long long*>(state + 4) = opt->rng_iter - 1;   // the optimiser step already advanced it
  }
  st[3] += 1;
}
__global__ void track_min_copy_kernel(const float* __restrict__ state, const float4* __restrict__ X,
                                      float4* __restrict__ Xb, long long nX4, const float4* __restrict__ Y,
                                      float4* __restrict__ Yb, long long nY4) {
  if (reinterpret_cast<const int*>(state)[1] == 0) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nX4 + nY4; i += stride) {
    if (i < nX4) Xb[i] = X[i];
    else Yb[i - nX4] = Y[i - nX4];
  }
}

__global__ void opt_prepare_kernel(const float* __restrict__ part, int npart, FbsnnAdam hp, OptState* st) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < npart; i += blockDim.x) acc += (double)part[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(acc);
    float coef = 1.f;
    if (hp.max_grad_norm > 0.0) coef = fminf((float)hp.max_grad_norm / (norm + 1e-6f), 1.f);
    st->rng_iter += 1;
    st->epoch += 1;
    st->grad_norm = norm;
    st->skip = (hp.skip_nonfinite != 0.0 && !isfinite(norm)) ? 1 : 0;
    if (st->skip) return;
    const long long step = st->step + 1;
    const double bc1 = 1.0 - pow(hp.beta1, (double)step);
    const double bc2 = 1.0 - pow(hp.beta2, (double)step);
    st->step = step;
    st->clip_coef = coef;
    st->step_size = (float)(hp.lr / bc1);
    st->bc2_sqrt = (float)sqrt(bc2);
  }
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float beta1, float beta2, float eps,
                            const OptState* __restrict__ st) {
  if (st->skip) return;
  const float coef = st->clip_coef, step_size = st->step_size, bc2s = st->bc2_sqrt;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);          // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;      // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2s + eps;
    p[i] -= step_size * (mi / denom);
  }
}

}  // namespace fbsnn
