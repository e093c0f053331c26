// i-vector front half of iv_plda (reference model/iv_plda.py:248-293 add_delta, :380-396
// extract_emb; model/_iv_plda/gmm.py:120-171 UBM posteriors and Baum-Welch statistics;
// model/_iv_plda/ivector_extract.py:94-114 i-vector = (I + sum_c N_c T_c' S_c^-1 T_c)^-1 (sum_c T_c' S_c^-1 F_c)).
//
// Building blocks, each with its adjoint, composed by the host class (speakerguard_b200/model/iv_plda.py);
// the dense contractions (UBM log-likelihoods on packed quadratic features, statistics, the L / linear
// assembly) go through the conv-as-GEMM kernel (sg_gemm), so this file only holds the element-wise
// and per-utterance pieces:
//   delta filters, quadratic feature expansion, row softmax, and a batched SPD solve (fp64 Cholesky).
#include <math.h>
#include <stdint.h>

#include "sg_common.cuh"

// ---------------------------------------------------------------------------------------------
// add_delta: out[b,t,i*F+f] = sum_j s_i[j] x[b, clamp(t+j), f], i = 0..2 (window 3, order 2)
// ---------------------------------------------------------------------------------------------
__constant__ float c_delta1[7];    // first-order filter, offsets -3..3
__constant__ float c_delta2[13];   // second-order filter, offsets -6..6

__global__ void delta_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int ldo, int B, int T, int F) {
  const size_t n = (size_t)B * T * F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const size_t bt = i / F;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float* xb = x + (size_t)b * T * ldx + f;
    float d1 = 0.f, d2 = 0.f;
#pragma unroll
    for (int j = -3; j <= 3; ++j) d1 = fmaf(c_delta1[j + 3], xb[(size_t)min(max(t + j, 0), T - 1) * ldx], d1);
#pragma unroll
    for (int j = -6; j <= 6; ++j) d2 = fmaf(c_delta2[j + 6], xb[(size_t)min(max(t + j, 0), T - 1) * ldx], d2);
    float* o = out + ((size_t)b * T + t) * ldo;
    o[f] = xb[(size_t)t * ldx];
    o[F + f] = d1;
    o[2 * F + f] = d2;
  }
}
// adjoint: dx[b,u,f] = sum_i sum_{t,j : clamp(t+j) == u} s_i[j] dout[b,t,i*F+f]
__global__ void delta_bwd_kernel(const float* __restrict__ dout, int ldo, float* __restrict__ dx, int ldx, int B, int T, int F) {
  const size_t n = (size_t)B * T * F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const size_t bt = i / F;
    const int u = (int)(bt % T), b = (int)(bt / T);
    const float* db = dout + (size_t)b * T * ldo;
    float g = db[(size_t)u * ldo + f];
    // a frame t reaches u through offset j = u - t, or through the clamp when u is an edge frame
    for (int t = max(0, u - 6); t <= min(T - 1, u + 6); ++t) {
      const int j = u - t;
      if (j >= -3 && j <= 3) g = fmaf(c_delta1[j + 3], db[(size_t)t * ldo + F + f], g);
      g = fmaf(c_delta2[j + 6], db[(size_t)t * ldo + 2 * F + f], g);
    }
    if (u == 0) {
      for (int t = 0; t < min(T, 6); ++t)
        for (int j = -6; j < -t; ++j) {                            // t + j < 0 clamps to frame 0
          if (j >= -3) g = fmaf(c_delta1[j + 3], db[(size_t)t * ldo + F + f], g);
          g = fmaf(c_delta2[j + 6], db[(size_t)t * ldo + 2 * F + f], g);
        }
    }
    if (u == T - 1) {
      for (int t = max(0, T - 6); t < T; ++t)
        for (int j = T - t; j <= 6; ++j) {                         // t + j > T-1 clamps to frame T-1
          if (j <= 3) g = fmaf(c_delta1[j + 3], db[(size_t)t * ldo + F + f], g);
          g = fmaf(c_delta2[j + 6], db[(size_t)t * ldo + 2 * F + f], g);
        }
    }
    dx[((size_t)b * T + u) * ldx + f] = g;
  }
}

// ---------------------------------------------------------------------------------------------
// quadratic feature expansion for the full-covariance UBM (gmm.py:120-131):
//   ll[t,c] = gconst_c + (S_c^-1 mu_c) . x_t - 1/2 x_t' S_c^-1 x_t  =  q_t . w_c + gconst_c
//   q_t = [x (F), x_i x_j for i <= j (F(F+1)/2), 0-pad];  rows are stacked per utterance with stride Tp,
//   rows t >= T are zero.
// ---------------------------------------------------------------------------------------------
// tf32 split of an fp32 value: hi keeps the 10 mantissa bits the tensor core reads, lo = x - hi is exact in fp32
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// split3 != 0: the row holds three K-segments [lo | hi | hi] of width kseg (3xTF32 operand layout, see sg_api_iv.cu)
// idx (optional, [kseg]): the factor pair of column k as i | j << 16, j = 0xffff for a linear column, i = 0xffff for padding -
// with it the kernel is a table lookup, two shared-memory reads and three coalesced stores per column (1.11 -> see DESIGN)
__global__ void quad_expand_fwd_kernel(const float* __restrict__ xa, int ldx, float* __restrict__ q, int ldq, int F,
                                       int split3, int kseg, const uint32_t* __restrict__ idx) {
  extern __shared__ float xs[];                                   // [F]
  const int row = blockIdx.x;
  float* qr = q + (size_t)row * ldq;
  if (split3 && idx) {
    for (int i = threadIdx.x; i < F; i += blockDim.x) xs[i] = xa[(size_t)row * ldx + i];
    __syncthreads();
    for (int k = threadIdx.x; k < kseg; k += blockDim.x) {
      const uint32_t ij = __ldg(idx + k);
      const uint32_t i = ij & 0xffffu, j = ij >> 16;
      float v = 0.f;
      if (i != 0xffffu) v = (j == 0xffffu) ? xs[i] : xs[i] * xs[j];
      const float hi = tf32_hi(v);
      qr[k] = v - hi; qr[kseg + k] = hi; qr[2 * kseg + k] = hi;
    }
    return;
  }
  if (split3) {
    for (int i = threadIdx.x; i < F; i += blockDim.x) xs[i] = xa[(size_t)row * ldx + i];
    __syncthreads();
    const int P = F * (F + 1) / 2;
    for (int k = threadIdx.x; k < kseg; k += blockDim.x) {
      float v = 0.f;
      if (k < F) v = xs[k];
      else if (k < F + P) {
        const int p = k - F;
        int i = (int)((2.f * F + 1.f - sqrtf((2.f * F + 1.f) * (2.f * F + 1.f) - 8.f * p)) * 0.5f);
        i = min(max(i, 0), F - 1);
        while (i > 0 && i * F - i * (i - 1) / 2 > p) --i;
        while (i + 1 < F && (i + 1) * F - (i + 1) * i / 2 <= p) ++i;
        v = xs[i] * xs[i + (p - (i * F - i * (i - 1) / 2))];
      }
      const float hi = tf32_hi(v);
      qr[k] = v - hi; qr[kseg + k] = hi; qr[2 * kseg + k] = hi;
    }
    return;
  }
  for (int i = threadIdx.x; i < F; i += blockDim.x) xs[i] = xa[(size_t)row * ldx + i];
  __syncthreads();
  const int P = F * (F + 1) / 2;
  for (int i = threadIdx.x; i < F; i += blockDim.x) qr[i] = xs[i];
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    // p -> (i, j), i <= j, row-major upper triangle: p = i*F - i(i-1)/2 + (j - i)
    int i = (int)((2.f * F + 1.f - sqrtf((2.f * F + 1.f) * (2.f * F + 1.f) - 8.f * p)) * 0.5f);
    i = min(max(i, 0), F - 1);
    while (i > 0 && i * F - i * (i - 1) / 2 > p) --i;
    while (i + 1 < F && (i + 1) * F - (i + 1) * i / 2 <= p) ++i;
    const int j = i + (p - (i * F - i * (i - 1) / 2));
    qr[F + p] = xs[i] * xs[j];
  }
  for (int i = F + P + threadIdx.x; i < ldq; i += blockDim.x) qr[i] = 0.f;
}
// dx_i = dq_lin[i] + sum_j dq_sym(i,j) x_j (1 + [i == j]) + add_i   (add: the path through the first-order statistics)
__global__ void quad_expand_bwd_kernel(const float* __restrict__ dq, int ldq, const float* __restrict__ xa, int ldx,
                                       const float* __restrict__ add, float* __restrict__ dx, int lddx, int T, int Tp, int F) {
  extern __shared__ float xs[];
  const int row = blockIdx.x, b = row / Tp, t = row - b * Tp;
  if (t >= T) return;
  const float* dr = dq + (size_t)row * ldq;
  for (int i = threadIdx.x; i < F; i += blockDim.x) xs[i] = xa[(size_t)row * ldx + i];
  __syncthreads();
  float* o = dx + ((size_t)b * T + t) * lddx;
  for (int i = threadIdx.x; i < lddx; i += blockDim.x) {
    if (i >= F) { o[i] = 0.f; continue; }
    float g = dr[i] + add[(size_t)row * ldx + i];
    for (int j = 0; j < F; ++j) {
      const int a = min(i, j), c = max(i, j);
      const float d = dr[F + a * F - a * (a - 1) / 2 + (c - a)];
      g = fmaf(d * (i == j ? 2.f : 1.f), xs[j], g);
    }
    o[i] = g;
  }
}
// Same result (same terms, same order) from the symmetric matrix expanded in shared memory: the packed row is read once,
// coalesced, and scattered to both (i, j) and (j, i) through the column -> factor-pair table of the forward kernel; every
// output is then a float4 dot product of one padded matrix row with x.  The form above spends 22 instructions per term on the
// packed index and the scattered load (635 M warp instructions at 131 072 rows: issue-bound at 0.70 ms; this one: 0.59 ms).
__global__ void __launch_bounds__(128)
quad_expand_bwd_sym_kernel(const float* __restrict__ dq, int ldq, const float* __restrict__ xa, int ldx,
                           const float* __restrict__ add, float* __restrict__ dx, int lddx, int T, int Tp, int F,
                           const uint32_t* __restrict__ idx) {
  extern __shared__ float xs[];                                   // [F4] x, then M [F][LDM]
  const int row = blockIdx.x, b = row / Tp, t = row - b * Tp;
  if (t >= T) return;
  const int F4 = (F + 3) & ~3, LDM = F4 + 4;                      // F = 72: rows 76 floats apart, conflict-free float4 reads
  float* M = xs + F4;
  const float* dr = dq + (size_t)row * ldq;
  for (int i = threadIdx.x; i < F4; i += blockDim.x) xs[i] = i < F ? xa[(size_t)row * ldx + i] : 0.f;
  for (int e = threadIdx.x; e < F * (F4 - F); e += blockDim.x) M[(e / (F4 - F)) * LDM + F + e % (F4 - F)] = 0.f;
  const int P = F * (F + 1) / 2;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const uint32_t ij = __ldg(idx + F + p);
    const int i = (int)(ij & 0xffffu), j = (int)(ij >> 16);
    const float d = dr[F + p];
    M[i * LDM + j] = (i == j) ? d * 2.f : d;
    if (i != j) M[j * LDM + i] = d;
  }
  __syncthreads();
  float* o = dx + ((size_t)b * T + t) * lddx;
  for (int i = threadIdx.x; i < lddx; i += blockDim.x) {
    if (i >= F) { o[i] = 0.f; continue; }
    float g = dr[i] + add[(size_t)row * ldx + i];
    const float4* mr = reinterpret_cast<const float4*>(M + i * LDM);
    const float4* xv = reinterpret_cast<const float4*>(xs);
    for (int q = 0; q < F4 / 4; ++q) {
      const float4 m4 = mr[q], x4 = xv[q];
      g = fmaf(m4.x, x4.x, g); g = fmaf(m4.y, x4.y, g); g = fmaf(m4.z, x4.z, g); g = fmaf(m4.w, x4.w, g);
    }
    o[i] = g;
  }
}
// feat [B,T,ld] -> Xa [B*Tp, Fa] = [x, 1, 0...]; rows t >= T are zero
__global__ void pad_aug_kernel(const float* __restrict__ feat, int ld, float* __restrict__ xa, int Fa, int B, int T, int Tp, int F) {
  const size_t n = (size_t)B * Tp * Fa;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % Fa);
    const size_t row = i / Fa;
    const int t = (int)(row % Tp), b = (int)(row / Tp);
    float v = 0.f;
    if (t < T) v = f < F ? feat[((size_t)b * T + t) * ld + f] : (f == F ? 1.f : 0.f);
    xa[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// row softmax over C components (gmm.py:133-136); rows with t >= T (padding) are written as zero
// one CTA (256 threads) per row
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float iv_block_reduce(float v, float* red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { float w = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, w) : v + w; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) a = is_max ? fmaxf(a, red[i]) : a + red[i];
  return a;
}
// (ll and post may alias: each thread rewrites only the elements it read)
__global__ void softmax_rows_fwd_kernel(const float* ll, float* post, int C, int T, int Tp) {
  __shared__ float red[8];
  const int row = blockIdx.x, t = row % Tp;
  const float* lr = ll + (size_t)row * C;
  float* pr = post + (size_t)row * C;
  if (t >= T) { for (int c = threadIdx.x; c < C; c += blockDim.x) pr[c] = 0.f; return; }
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lr[c]);
  mx = iv_block_reduce(mx, red, true);
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(lr[c] - mx);
  s = iv_block_reduce(s, red, false);
  const float inv = 1.f / s;
  for (int c = threadIdx.x; c < C; c += blockDim.x) pr[c] = expf(lr[c] - mx) * inv;
}
// One warp per row with the row in registers (C = 128 * NV: NV float4 per lane): one pass over memory, one exp per element,
// shuffle reductions instead of block barriers.  The CTA-per-row form above read the row three times through L1 and ran at
// 2.7 x the time of its HBM traffic at C = 2048.
template <int NV>
__global__ void __launch_bounds__(256)
softmax_rows_fwd_warp_kernel(const float* ll, float* post, int rows, int T, int Tp) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int C = 128 * NV;
  const float4* lr = reinterpret_cast<const float4*>(ll + (size_t)row * C);
  float4* pr = reinterpret_cast<float4*>(post + (size_t)row * C);
  if (row % Tp >= T) {
#pragma unroll
    for (int i = 0; i < NV; ++i) pr[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = lr[i * 32 + lane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
#pragma unroll
  for (int i = 0; i < NV; ++i) pr[i * 32 + lane] = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
}
template <int NV>
__global__ void __launch_bounds__(256)
softmax_rows_bwd_warp_kernel(const float* __restrict__ post, const float* dpost, float* dll, int rows, int T, int Tp, int split3) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  constexpr int C = 128 * NV;
  const int nseg = split3 ? 3 : 1;
  float4* o = reinterpret_cast<float4*>(dll + (size_t)row * C * nseg);
  if (row % Tp >= T) {
    for (int i = lane; i < NV * 32 * nseg; i += 32) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* pr = reinterpret_cast<const float4*>(post + (size_t)row * C);
  const float4* dr = reinterpret_cast<const float4*>(dpost + (size_t)row * C);
  float4 p[NV], d[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { p[i] = __ldg(pr + i * 32 + lane); d[i] = dr[i * 32 + lane]; }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { s = fmaf(d[i].x, p[i].x, s); s = fmaf(d[i].y, p[i].y, s); s = fmaf(d[i].z, p[i].z, s); s = fmaf(d[i].w, p[i].w, s); }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 v = make_float4(p[i].x * (d[i].x - s), p[i].y * (d[i].y - s), p[i].z * (d[i].z - s), p[i].w * (d[i].w - s));
    if (split3) {
      const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      o[i * 32 + lane] = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
      o[NV * 32 + i * 32 + lane] = hi;
      o[2 * NV * 32 + i * 32 + lane] = hi;
    } else {
      o[i * 32 + lane] = v;
    }
  }
}
// dll = post * (dpost - sum_c dpost_c post_c)
// split3: dll is written as [lo | hi | hi] segments of width C into a row of 3C (operand of the 3xTF32 contraction)
__global__ void softmax_rows_bwd_kernel(const float* __restrict__ post, const float* dpost, float* dll, int C, int T, int Tp,
                                        int split3) {
  __shared__ float red[8];
  const int row = blockIdx.x, t = row % Tp;
  const float* pr = post + (size_t)row * C;
  const float* dr = dpost + (size_t)row * C;
  float* o = dll + (size_t)row * C * (split3 ? 3 : 1);
  if (t >= T) { for (int c = threadIdx.x; c < C * (split3 ? 3 : 1); c += blockDim.x) o[c] = 0.f; return; }
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s = fmaf(dr[c], pr[c], s);
  s = iv_block_reduce(s, red, false);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = pr[c] * (dr[c] - s);
    if (split3) { const float hi = tf32_hi(v); o[c] = v - hi; o[C + c] = hi; o[2 * C + c] = hi; }
    else o[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// batched SPD solve: Lp [B, D(D+1)/2] packed upper triangle (fp32) -> fp64 Cholesky (kept in `fac`,
// [B, D, D], lower triangle) ; w = L^-1 rhs.  Second entry point solves with the stored factor and forms
// the adjoints of the system: lambda = L^-1 dw, dLp(i<=j) = -(lambda_i w_j + lambda_j w_i) / (1 + [i==j]).
// One CTA per utterance.
// ---------------------------------------------------------------------------------------------
#define CH_NB 32                 // panel width of the blocked factorisation
#define CH_LD (CH_NB + 1)        // padded panel row stride (doubles)

// rows k0..D-1 of the nb panel columns -> shared memory (upper part of the diagonal block zeroed).  One warp per row, one
// 8-byte asynchronous copy per element, every copy of the panel in flight before the first wait: with plain loads staged through
// registers the panel came in at ~4 loads in flight per thread and a quarter of the kernel's samples sat on this line.
__device__ __forceinline__ void chol_load_panel(const double* __restrict__ A, int D, int k0, int nb, double* Pn) {
  const int rows = D - k0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane < nb) {
    const double* src = reinterpret_cast<const double*>(__cvta_generic_to_global(A + (size_t)(k0 + warp) * D + k0 + lane));
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(Pn + warp * CH_LD + lane);
    for (int r = warp; r < rows; r += nw) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
      src += (size_t)nw * D;
      dst += (uint32_t)(nw * CH_LD * sizeof(double));
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (lane < nb)
    for (int r = warp; r < nb; r += nw)
      if (lane > r) Pn[r * CH_LD + lane] = 0.0;
}

// forward substitution L y = v, then back substitution L' w = y, panel by panel (L lower-triangular, row-major)
__device__ void chol_solve_with_factor(const double* __restrict__ A, int D, double* v /* smem [D], in/out */, double* Pn,
                                       double* red /* smem [8 * CH_NB] */) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < D; k0 += CH_NB) {
    const int nb = min(CH_NB, D - k0), rows = D - k0;
    __syncthreads();
    chol_load_panel(A, D, k0, nb, Pn);
    __syncthreads();
    if (warp == 0) {
      double val = lane < nb ? v[k0 + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        if (lane == j) val /= Pn[j * CH_LD + j];
        const double yj = __shfl_sync(0xffffffffu, val, j);
        if (lane > j && lane < nb) val -= Pn[lane * CH_LD + j] * yj;
      }
      if (lane < nb) v[k0 + lane] = val;
    }
    __syncthreads();
    for (int i = nb + tid; i < rows; i += blockDim.x) {
      double sacc = 0.0;
      for (int c = 0; c < nb; ++c) sacc += Pn[i * CH_LD + c] * v[k0 + c];
      v[k0 + i] -= sacc;
    }
  }
  for (int k0 = (D - 1) / CH_NB * CH_NB; k0 >= 0; k0 -= CH_NB) {
    const int nb = min(CH_NB, D - k0), rows = D - k0;
    __syncthreads();
    chol_load_panel(A, D, k0, nb, Pn);
    __syncthreads();
    {  // s_c = sum_{i >= nb} L[k0+i][k0+c] w[k0+i]
      double sacc = 0.0;
      if (lane < nb)
        for (int i = nb + warp; i < rows; i += (blockDim.x >> 5)) sacc += Pn[i * CH_LD + lane] * v[k0 + i];
      red[warp * CH_NB + lane] = sacc;
    }
    __syncthreads();
    if (warp == 0) {
      double val = 0.0;
      if (lane < nb) {
        val = v[k0 + lane];
        for (int g = 0; g < (int)(blockDim.x >> 5); ++g) val -= red[g * CH_NB + lane];
      }
      for (int j = nb - 1; j >= 0; --j) {
        if (lane == j) val /= Pn[j * CH_LD + j];
        const double wj = __shfl_sync(0xffffffffu, val, j);
        if (lane < j) val -= Pn[j * CH_LD + lane] * wj;
      }
      if (lane < nb) v[k0 + lane] = val;
    }
  }
  __syncthreads();
}

// blocked right-looking Cholesky, in place in the lower triangle of A (row-major [D, D], global memory).  Per 32-column panel:
// the 32 x 32 diagonal block is factorised by the whole CTA, the rows below it are then independent triangular solves
// x L11' = a (one thread per row, the row in registers; the same sequence of fused multiply-adds and divisions per element as the
// column-by-column right-looking form, without its 3 block barriers per column over the full panel height), and the trailing
// matrix gets one rank-32 update.
__device__ void chol_factor_blocked(double* __restrict__ A, int D, double* Pn) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
  for (int k0 = 0; k0 < D; k0 += CH_NB) {
    const int nb = min(CH_NB, D - k0), rows = D - k0;
    __syncthreads();
    chol_load_panel(A, D, k0, nb, Pn);
    __syncthreads();
    // diagonal block: lane -> column, warp -> row
    for (int j = 0; j < nb; ++j) {
      const double d = sqrt(Pn[j * CH_LD + j]);
      __syncthreads();
      if (tid < nb - j) { const int r = j + tid; Pn[r * CH_LD + j] = (r == j) ? d : Pn[r * CH_LD + j] / d; }
      __syncthreads();
      const int c = j + 1 + lane;
      if (c < nb)
        for (int r = j + 1 + warp; r < nb; r += nw)
          if (r >= c) Pn[r * CH_LD + c] -= Pn[r * CH_LD + j] * Pn[c * CH_LD + j];
      __syncthreads();
    }
    // rows below the block (nb == CH_NB whenever there are any)
    for (int r = nb + tid; r < rows; r += nthr) {
      double x[CH_NB];
      double* pr = Pn + r * CH_LD;
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) x[c] = pr[c];
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) {
        const double* lc = Pn + c * CH_LD;           // uniform address: broadcast reads
        double acc = x[c];
#pragma unroll
        for (int k = 0; k < c; ++k) acc = fma(-x[k], lc[k], acc);
        x[c] = acc / lc[c];
      }
#pragma unroll
      for (int c = 0; c < CH_NB; ++c) pr[c] = x[c];
    }
    __syncthreads();
    for (int e = tid; e < rows * nb; e += nthr) {
      const int r = e / nb, c = e - r * nb;
      if (r >= c) A[(size_t)(k0 + r) * D + k0 + c] = Pn[r * CH_LD + c];
    }
    // trailing update in 4x4 register tiles over the lower triangle of the (n2 x n2) block.  A warp takes a patch of 4 x 8
    // tiles: its row operands are 4 distinct panel rows (one shared-memory wavefront), its column operands 8 (two) - with 32
    // consecutive column tiles per warp the rows were 4 * 33 doubles apart on every lane, an 8-way bank conflict that made up
    // 60 % of the kernel's shared-memory wavefronts.
    const int n2 = rows - nb, nt = (n2 + 3) / 4;
    int cnt = 0;
    for (int si = 0; 4 * si < nt; ++si) {
      const int sjmax = min((4 * si + 3) >> 3, (nt - 1) >> 3);
      for (int sj = 0; sj <= sjmax; ++sj, ++cnt) {
        if (cnt % nw != warp) continue;
        const int ti = 4 * si + (lane >> 3), tj = 8 * sj + (lane & 7);
        if (ti >= nt || tj > ti) continue;
        double acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int w = 0; w < 4; ++w) acc[u][w] = 0.0;
        int ri[4], rj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { ri[u] = min(nb + ti * 4 + u, rows - 1) * CH_LD; rj[u] = min(nb + tj * 4 + u, rows - 1) * CH_LD; }
        for (int c = 0; c < nb; ++c) {
          double a[4], bq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = Pn[ri[u] + c]; bq[u] = Pn[rj[u] + c]; }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) acc[u][w] = fma(a[u], bq[w], acc[u][w]);
        }
        const int gj0 = nb + tj * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int gi = nb + ti * 4 + u;
          if (gi >= rows) continue;
          double* arow = A + (size_t)(k0 + gi) * D + k0 + gj0;
          if ((D & 1) == 0 && gj0 + 3 <= gi) {             // the tile's four columns of this row: two 16-byte updates
            double2* a2 = reinterpret_cast<double2*>(arow);
            double2 v0 = a2[0], v1 = a2[1];
            v0.x -= acc[u][0]; v0.y -= acc[u][1]; v1.x -= acc[u][2]; v1.y -= acc[u][3];
            a2[0] = v0; a2[1] = v1;
          } else {
#pragma unroll
            for (int w = 0; w < 4; ++w)
              if (gj0 + w <= gi) arow[w] -= acc[u][w];
          }
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 2)
chol_factor_solve_kernel(const float* __restrict__ Lp, int ldp, const float* __restrict__ rhs, int ldr, float offset,
                         const float* __restrict__ emb_mean, double* __restrict__ fac, float* __restrict__ wfull,
                         float* __restrict__ iv, int D) {
  extern __shared__ double vs[];                                  // [D] right-hand side, [8*32] reduction, [D*33] panel
  double* red = vs + D;
  double* Pn = red + 8 * CH_NB;
  const int b = blockIdx.x;
  double* A = fac + (size_t)b * D * D;
  const float* lp = Lp + (size_t)b * ldp;
  // unpack (upper-triangle packing: p = i*D - i(i-1)/2 + (j-i), i <= j) into the lower triangle; L = I + sum_c N_c U_c
  // through 32 x 32 tiles (one per warp at a time, in the panel buffer) so that both the packed rows and A's rows are read and
  // written along their contiguous direction (writing A[j][i] with j across the lanes was 9 % of the kernel's samples)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, nbk = (D + 31) / 32;
    float* tile = reinterpret_cast<float*>(Pn) + warp * (32 * 33);
    int cnt = 0;
    for (int bi = 0; bi < nbk; ++bi)
      for (int bj = bi; bj < nbk; ++bj, ++cnt) {
        if (cnt % nw != warp) continue;
        for (int ii = 0; ii < 32; ++ii) {
          const int i = 32 * bi + ii, j = 32 * bj + lane;
          if (i < D && j < D && j >= i) tile[ii * 33 + lane] = lp[(size_t)i * D - (size_t)i * (i - 1) / 2 + (j - i)];
        }
        __syncwarp();
        for (int jj = 0; jj < 32; ++jj) {
          const int j = 32 * bj + jj, i = 32 * bi + lane;
          if (j < D && i <= j) A[(size_t)j * D + i] = (double)tile[lane * 33 + jj] + (i == j ? 1.0 : 0.0);
        }
        __syncwarp();
      }
  }
  chol_factor_blocked(A, D, Pn);
  // linear[0] += prior_offset; ivector[0] -= prior_offset (ivector_extract.py:108-113); then the global mean is removed
  for (int i = threadIdx.x; i < D; i += blockDim.x) vs[i] = (double)rhs[(size_t)b * ldr + i] + (i == 0 ? (double)offset : 0.0);
  chol_solve_with_factor(A, D, vs, Pn, red);
  for (int i = threadIdx.x; i < ldr; i += blockDim.x) {
    wfull[(size_t)b * ldr + i] = i < D ? (float)vs[i] : 0.f;
    iv[(size_t)b * ldr + i] = i < D ? (float)(vs[i] - (i == 0 ? (double)offset : 0.0)) - emb_mean[i] : 0.f;
  }
}

__global__ void __launch_bounds__(256)
chol_solve_bwd_kernel(const double* __restrict__ fac, const float* __restrict__ w, const float* __restrict__ dw, int ldr,
                      float* __restrict__ drhs, float* __restrict__ dLp, int ldp, int D) {
  extern __shared__ double vs[];
  double* red = vs + D;
  double* Pn = red + 8 * CH_NB;
  const int b = blockIdx.x;
  const size_t P = (size_t)D * (D + 1) / 2;
  const double* A = fac + (size_t)b * D * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) vs[i] = (double)dw[(size_t)b * ldr + i];
  chol_solve_with_factor(A, D, vs, Pn, red);
  for (int i = threadIdx.x; i < ldr; i += blockDim.x) drhs[(size_t)b * ldr + i] = i < D ? (float)vs[i] : 0.f;
  if (!dLp) return;                 // the caller contracts dL = -lambda w' implicitly (sg_api_iv.cu)
  const float* wb = w + (size_t)b * ldr;
  float* o = dLp + (size_t)b * ldp;
  for (int i = 0; i < D; ++i) {
    const double li = vs[i];
    const double wi = (double)wb[i];
    for (int j = i + threadIdx.x; j < D; j += blockDim.x) {
      // the packed entry (i,j), i < j, stands for both L_ij and L_ji
      const double g = (i == j) ? -li * wi : -(li * (double)wb[j] + vs[j] * wi);
      o[(size_t)i * D - (size_t)i * (i - 1) / 2 + (j - i)] = (float)g;
    }
  }
  for (size_t p = P + threadIdx.x; p < (size_t)ldp; p += blockDim.x) o[p] = 0.f;
}

// out[r, n] = sum_s part[s][r][n]   (split-K partial sums of the skinny contractions)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int rows, int N, float* __restrict__ out, int ldo) {
  const size_t n = (size_t)rows * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += part[(size_t)sp * n + i];
    out[(i / N) * ldo + (i % N)] = acc;
  }
}

// [rows, C] (row stride ld) -> [rows, 3C] = [lo | hi | hi]  (activation-side operand of a 3xTF32 contraction)
__global__ void split3_rows_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int rows, int C) {
  const size_t n = (size_t)rows * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / C;
    const int c = (int)(i - r * C);
    const float v = in[r * ld + c], hi = tf32_hi(v);
    float* o = out + r * 3 * C;
    o[c] = v - hi; o[C + c] = hi; o[2 * C + c] = hi;
  }
}
// same with the output row padded to ldo >= 3C floats (zeros): K of the contraction a multiple of the tensor core's k-block
__global__ void split3_rows_ld_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int ldo, int rows, int C) {
  const size_t n = (size_t)rows * ldo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / ldo;
    const int k = (int)(i - r * ldo), seg = k / C, c = k - seg * C;
    float v = 0.f;
    if (seg < 3) { const float x = in[r * ld + c], hi = tf32_hi(x); v = seg == 0 ? x - hi : hi; }
    out[i] = v;
  }
}
// weight-side operand of a 3xTF32 contraction, K-major: W3[n, :] = [hi | lo | hi | 0](w_n), w_n[d] = src[n * sn + d * sd], d < K
__global__ void build_w3_kernel(const float* __restrict__ src, size_t sn, size_t sd, float* __restrict__ W3, int N, int K, int K3p) {
  const size_t n_el = (size_t)N * K3p;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / K3p;
    const int k = (int)(i - n * K3p), seg = k / K, d = k - seg * K;
    float v = 0.f;
    if (seg < 3) { const float x = src[n * sn + (size_t)d * sd], hi = tf32_hi(x); v = seg == 1 ? x - hi : hi; }
    W3[i] = v;
  }
}
// U [C, Pp] -> UT3 [Pp, 3C] = [hi | lo | hi](U^T)  (weight-side operand, K-major)
__global__ void build_ut3_kernel(const float* __restrict__ U, float* __restrict__ UT3, int C, int Pp) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = (c0 + i < C) ? U[(size_t)(c0 + i) * Pp + p0 + threadIdx.x] : 0.f;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const float v = tile[threadIdx.x][i], hi = tf32_hi(v);
    if (c0 + threadIdx.x >= C) continue;
    float* o = UT3 + (size_t)(p0 + i) * 3 * C + c0 + threadIdx.x;
    o[0] = hi; o[C] = v - hi; o[2 * C] = hi;
  }
}
// dFsT[b, F, c] = -sum_f dFsT[b, f, c] * a[b, f, c]   (zeroth-order statistics gradient from the first-order one)
__global__ void dn_from_df_kernel(float* __restrict__ dFsT, const float* __restrict__ a, int B, int F, int Fa, int C) {
  const size_t n = (size_t)B * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / C;
    const int c = (int)(i - b * C);
    const float* d = dFsT + b * Fa * C + c;
    const float* q = a + b * Fa * C + c;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) acc = fmaf(d[(size_t)f * C], q[(size_t)f * C], acc);
    dFsT[b * Fa * C + (size_t)F * C + c] = -acc;
  }
}

// row-major [R, C] -> [C, R] per batch item (layout helper: X^T for the statistics GEMM)
__global__ void transpose_batched_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C, int ld_in,
                                         int ld_out, size_t stride_in, size_t stride_out) {
  __shared__ float tile[32][33];
  const float* ib = in + blockIdx.z * stride_in;
  float* ob = out + blockIdx.z * stride_out;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? ib[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) ob[(size_t)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static std::atomic<unsigned long long> g_delta_init{0};   // per device: __constant__ memory belongs to a device's context
static int delta_init() {
  if (!sg_first_on_device(&g_delta_init)) return SG_OK;
  // get_scales(window 3, order 2): first filter j/28 for j=-3..3, second = first convolved with itself
  float d1[7], d2[13];
  for (int j = -3; j <= 3; ++j) d1[j + 3] = (float)j / 28.f;
  for (int i = 0; i < 13; ++i) d2[i] = 0.f;
  for (int j = -3; j <= 3; ++j)
    for (int k = -3; k <= 3; ++k) d2[j + k + 6] += (float)j * d1[k + 3];
  for (int i = 0; i < 13; ++i) d2[i] *= 1.f / 28.f;
  SG_CUDA_CHECK(cudaMemcpyToSymbol(c_delta1, d1, sizeof(d1)));
  SG_CUDA_CHECK(cudaMemcpyToSymbol(c_delta2, d2, sizeof(d2)));
  return SG_OK;
}
static int iv_blocks(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

int sg_delta_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int F, int backward, cudaStream_t st) {
  int r = delta_init();
  if (r != SG_OK) return r;
  if (!backward) delta_fwd_kernel<<<iv_blocks((size_t)B * T * F), 256, 0, st>>>(in, ld_in, out, ld_out, B, T, F);
  else delta_bwd_kernel<<<iv_blocks((size_t)B * T * F), 256, 0, st>>>(in, ld_in, out, ld_out, B, T, F);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_pad_aug_launch(const float* feat, int ld, float* xa, int Fa, int B, int T, int Tp, int F, cudaStream_t st) {
  pad_aug_kernel<<<iv_blocks((size_t)B * Tp * Fa), 256, 0, st>>>(feat, ld, xa, Fa, B, T, Tp, F);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_quad_expand_launch(const float* xa, int ldx, float* q, int ldq, int rows, int F, int split3, int kseg, const uint32_t* idx,
                          cudaStream_t st) {
  quad_expand_fwd_kernel<<<rows, 256, F * sizeof(float), st>>>(xa, ldx, q, ldq, F, split3, kseg, idx);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_quad_expand_bwd_launch(const float* dq, int ldq, const float* xa, int ldx, const float* add, float* dx, int lddx,
                              int B, int T, int Tp, int F, const uint32_t* idx, cudaStream_t st) {
  const int F4 = (F + 3) & ~3;
  const size_t sym = ((size_t)F4 + (size_t)F * (F4 + 4)) * sizeof(float);
  if (idx && sym <= 48 * 1024) {
    quad_expand_bwd_sym_kernel<<<B * Tp, 128, sym, st>>>(dq, ldq, xa, ldx, add, dx, lddx, T, Tp, F, idx);
    SG_LAUNCH_CHECK();
    return SG_OK;
  }
  // (measured without gain at 131 072 rows, 0.70 ms: a warp-per-row form that reads each packed row once - warp sum for dx_i,
  // per-lane accumulators for dx_j: 1.16 ms, 72 dependent rounds of load + shuffle reduction; carried packed indices instead
  // of min / max / multiply per term: 0.72 ms; the packed row staged in shared memory by float4 loads: 0.73 ms)
  quad_expand_bwd_kernel<<<B * Tp, 96, F * sizeof(float), st>>>(dq, ldq, xa, ldx, add, dx, lddx, T, Tp, F);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_softmax_rows_launch(const float* a, const float* b, float* out, int rows, int C, int T, int Tp, int backward,
                           int split3, cudaStream_t st) {
  const bool al = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0;
  if (C == 2048 && al) {
    if (!backward) softmax_rows_fwd_warp_kernel<16><<<(rows + 7) / 8, 256, 0, st>>>(a, out, rows, T, Tp);
    else softmax_rows_bwd_warp_kernel<16><<<(rows + 7) / 8, 256, 0, st>>>(a, b, out, rows, T, Tp, split3);
  } else if (C == 1024 && al) {
    if (!backward) softmax_rows_fwd_warp_kernel<8><<<(rows + 7) / 8, 256, 0, st>>>(a, out, rows, T, Tp);
    else softmax_rows_bwd_warp_kernel<8><<<(rows + 7) / 8, 256, 0, st>>>(a, b, out, rows, T, Tp, split3);
  } else if (C == 256 && al) {
    if (!backward) softmax_rows_fwd_warp_kernel<2><<<(rows + 7) / 8, 256, 0, st>>>(a, out, rows, T, Tp);
    else softmax_rows_bwd_warp_kernel<2><<<(rows + 7) / 8, 256, 0, st>>>(a, b, out, rows, T, Tp, split3);
  } else
  if (!backward) softmax_rows_fwd_kernel<<<rows, 256, 0, st>>>(a, out, C, T, Tp);
  else softmax_rows_bwd_kernel<<<rows, 256, 0, st>>>(a, b, out, C, T, Tp, split3);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
// [D] right-hand side, [8 * 32] reduction scratch, panel [D * 33] - which also holds the eight 32 x 33 float tiles of the unpack
static size_t chol_smem(int D) {
  const size_t panel = (size_t)D * CH_LD, tiles = 8 * 32 * 33 / 2;
  return ((size_t)D + 8 * CH_NB + (panel > tiles ? panel : tiles)) * sizeof(double);
}
static int chol_init(int D) {
  static int configured[64] = {0};                         // per device (function attributes are per context)
  const int need = (int)chol_smem(D);
  if (need > 227 * 1024) { sg_set_error("i-vector dimension %d too large for the shared-memory panel (max 800)", D); return SG_EINVAL; }
  int dev = 0;
  SG_CUDA_CHECK(cudaGetDevice(&dev));
  if (need > configured[dev & 63]) {
    SG_CUDA_CHECK(cudaFuncSetAttribute(chol_factor_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
    SG_CUDA_CHECK(cudaFuncSetAttribute(chol_solve_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
    configured[dev & 63] = need;
  }
  return SG_OK;
}
int sg_chol_solve_launch(const float* Lp, int ldp, const float* rhs, int ldr, float offset, const float* emb_mean, double* fac,
                         float* wfull, float* iv, int B, int D, cudaStream_t st) {
  int r = chol_init(D);
  if (r != SG_OK) return r;
  chol_factor_solve_kernel<<<B, 256, chol_smem(D), st>>>(Lp, ldp, rhs, ldr, offset, emb_mean, fac, wfull, iv, D);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_chol_solve_bwd_launch(const double* fac, const float* w, const float* dw, int ldr, float* drhs, float* dLp, int ldp,
                             int B, int D, cudaStream_t st) {
  int r = chol_init(D);
  if (r != SG_OK) return r;
  chol_solve_bwd_kernel<<<B, 256, chol_smem(D), st>>>(fac, w, dw, ldr, drhs, dLp, ldp, D);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_split3_rows_launch(const float* in, int ld, float* out, int rows, int C, cudaStream_t st) {
  split3_rows_kernel<<<iv_blocks((size_t)rows * C), 256, 0, st>>>(in, ld, out, rows, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
// float4 form of the two kernels above for contiguous sources (sd == 1) with K, the strides and the padded width multiples of 4:
// one thread per output float4, 32-bit index arithmetic (the scalar forms with a 64-bit division per element wrote at ~1.2 TB/s)
// act != 0: activation-side layout [lo | hi | hi | 0]; else weight-side [hi | lo | hi | 0]
__global__ void split3_vec_kernel(const float* __restrict__ src, int sn, float4* __restrict__ out, int n4, int K, int K3p, int act) {
  const int kq = K3p >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int n = i / kq, k = (i - n * kq) << 2, seg = k / K, d = k - seg * K;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (seg < 3) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(src + (size_t)n * sn + d));
      const float4 hi = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
      v = (seg == (act ? 0 : 1)) ? make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w) : hi;
    }
    out[i] = v;
  }
}
static bool split3_vec_ok(const void* src, size_t sn, const void* out, size_t n, int K, int K3p) {
  return K % 4 == 0 && K3p % 4 == 0 && sn % 4 == 0 && (((uintptr_t)src | (uintptr_t)out) & 15) == 0 && n * (size_t)(K3p / 4) < (1u << 30);
}
int sg_split3_rows_ld_launch(const float* in, int ld, float* out, int ldo, int rows, int C, cudaStream_t st) {
  if (split3_vec_ok(in, (size_t)ld, out, (size_t)rows, C, ldo)) {
    const int n4 = rows * (ldo / 4);
    split3_vec_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(in, ld, reinterpret_cast<float4*>(out), n4, C, ldo, 1);
    SG_LAUNCH_CHECK();
    return SG_OK;
  }
  split3_rows_ld_kernel<<<iv_blocks((size_t)rows * ldo), 256, 0, st>>>(in, ld, out, ldo, rows, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_build_w3_launch(const float* src, size_t sn, size_t sd, float* W3, int N, int K, int K3p, cudaStream_t st) {
  if (sd == 1 && split3_vec_ok(src, sn, W3, (size_t)N, K, K3p)) {
    const int n4 = N * (K3p / 4);
    split3_vec_kernel<<<(n4 + 255) / 256, 256, 0, st>>>(src, (int)sn, reinterpret_cast<float4*>(W3), n4, K, K3p, 0);
    SG_LAUNCH_CHECK();
    return SG_OK;
  }
  build_w3_kernel<<<iv_blocks((size_t)N * K3p), 256, 0, st>>>(src, sn, sd, W3, N, K, K3p);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_build_ut3_launch(const float* U, float* UT3, int C, int Pp, cudaStream_t st) {
  build_ut3_kernel<<<dim3(Pp / 32, (C + 31) / 32), dim3(32, 8), 0, st>>>(U, UT3, C, Pp);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_dn_from_df_launch(float* dFsT, const float* a, int B, int F, int Fa, int C, cudaStream_t st) {
  dn_from_df_kernel<<<iv_blocks((size_t)B * C), 256, 0, st>>>(dFsT, a, B, F, Fa, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_splitk_reduce_launch(const float* part, int splits, int rows, int N, float* out, int ldo, cudaStream_t st) {
  splitk_reduce_kernel<<<iv_blocks((size_t)rows * N), 256, 0, st>>>(part, splits, rows, N, out, ldo);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_transpose_batched_launch(const float* in, float* out, int R, int C, int ld_in, int ld_out, size_t stride_in,
                                size_t stride_out, int nbatch, cudaStream_t st) {
  transpose_batched_kernel<<<dim3((C + 31) / 32, (R + 31) / 32, nbatch), dim3(32, 8), 0, st>>>(in, out, R, C, ld_in, ld_out,
                                                                                               stride_in, stride_out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
