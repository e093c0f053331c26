// H1: statistics pooling, embedding post-processing, PLDA scoring, decision and attack losses,
// with their adjoints.  Everything here is per-utterance vector work (<= 3072 floats), fp32.
//
// Reference: stats pooling xvecTDNN.py:62; process_emb model/iv_plda.py:411-443 (length-norm
// with detached norm: xvector_extract.py:31-38; PLDA transform plda.py:73-97); scoring
// plda.py:140-190; decision model/defended_model.py:167-170; losses attack/utils.py:7-102.
#include <cuda_bf16.h>
#include <math.h>

#include <stdlib.h>

#include "sg_common.cuh"
#include "sg_head.cuh"

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// ---------------------------------------------------------------------------------------------
// statistics pooling over the valid frames of the last TDNN layer (post-ReLU, BN folded here)
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void pool_fwd_kernel(const AT* __restrict__ r5, int T, int Tv, const float* __restrict__ bn_mean,
                                const float* __restrict__ bn_istd, float* __restrict__ stats,
                                float* __restrict__ save_mean, float* __restrict__ save_std) {
  // grid (C5P/128, B), block (32, 8): each thread owns 4 consecutive channels, 8 row lanes per channel group
  __shared__ float4 part[8][32];
  __shared__ float4 bc[32];
  const int c0 = blockIdx.x * 128 + threadIdx.x * 4, r = threadIdx.y, b = blockIdx.y;
  const AT* base = r5 + (size_t)b * T * SG_C5P + c0;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = r; t < Tv; t += 8) { const float4 v = ld4(base + (size_t)t * SG_C5P); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
  part[r][threadIdx.x] = s;
  __syncthreads();
  if (r == 0) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 v = part[i][threadIdx.x]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    const float inv = 1.f / (float)Tv;
    bc[threadIdx.x] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
  __syncthreads();
  const float4 mean = bc[threadIdx.x];
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = r; t < Tv; t += 8) {
    const float4 v = ld4(base + (size_t)t * SG_C5P);
    const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
    q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
  }
  __syncthreads();
  part[r][threadIdx.x] = q;
  __syncthreads();
  if (r == 0) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float4 v = part[i][threadIdx.x]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    const float inv = 1.f / (float)(Tv - 1);                       // unbiased (torch.std default)
    const float sd[4] = {sqrtf(a.x * inv), sqrtf(a.y * inv), sqrtf(a.z * inv), sqrtf(a.w * inv)};
    const float mu4[4] = {mean.x, mean.y, mean.z, mean.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + i;
      const bool real = c < SG_C5;
      const float mu = real ? bn_mean[c] : 0.f, is = real ? bn_istd[c] : 0.f;
      stats[(size_t)b * SG_STATS + c] = real ? (mu4[i] - mu) * is : 0.f;
      stats[(size_t)b * SG_STATS + SG_C5P + c] = real ? sd[i] * is : 0.f;
      save_mean[(size_t)b * SG_C5P + c] = mu4[i];
      save_std[(size_t)b * SG_C5P + c] = sd[i];
    }
  }
}

// bf16 mode: single pass over r5 (the two-pass form above re-reads every CTA's 70 KB slab through L2, which costs as much
// as the first read).  Shifted sums with the first frame as pivot: d = x - x[0], mean = x[0] + S1/n,
// var = (S2 - S1^2/n)/(n-1) clamped at 0; exact (var = 0) for constant / dead channels.  The fp32 accumulation error is
// far below the bf16 rounding of the activations themselves; fp32 / tf32 modes keep the reference's two-pass form.
// grid (C5P/256, B), block (32, 8): each thread owns 8 consecutive channels (16-byte loads), 8 row lanes.
__device__ __forceinline__ void unpack8(const uint4 u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) { f[2 * k] = __uint_as_float(w[k] << 16); f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
}
__global__ void __launch_bounds__(256)
pool_fwd_onepass_kernel(const __nv_bfloat16* __restrict__ r5, int T, int Tv, const float* __restrict__ bn_mean,
                        const float* __restrict__ bn_istd, float* __restrict__ stats,
                        float* __restrict__ save_mean, float* __restrict__ save_std) {
  __shared__ float p1[8][32][9], p2[8][32][9];                     // [row lane][channel group][8 channels + pad]
  const int c0 = blockIdx.x * 256 + threadIdx.x * 8, r = threadIdx.y, b = blockIdx.y;
  const __nv_bfloat16* base = r5 + (size_t)b * T * SG_C5P + c0;
  float K[8], s1[8], s2[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base)), K);
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  int t = r;
  for (; t + 24 < Tv; t += 32) {                                   // 4 independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldcs(reinterpret_cast<const uint4*>(base + (size_t)(t + 8 * k) * SG_C5P));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8(u[k], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = f[i] - K[i]; s1[i] += d; s2[i] = fmaf(d, d, s2[i]); }
    }
  }
  for (; t < Tv; t += 8) {
    float f[8];
    unpack8(__ldcs(reinterpret_cast<const uint4*>(base + (size_t)t * SG_C5P)), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = f[i] - K[i]; s1[i] += d; s2[i] = fmaf(d, d, s2[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { p1[r][threadIdx.x][i] = s1[i]; p2[r][threadIdx.x][i] = s2[i]; }
  __syncthreads();
  // 256 threads -> 256 channels: thread (x, y) finishes channel y * 32 + x of this block
  const int cl = r * 32 + threadIdx.x, g = cl >> 3, i = cl & 7;
  float a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { a1 += p1[k][g][i]; a2 += p2[k][g][i]; }
  const int c = blockIdx.x * 256 + cl;
  const float n = (float)Tv;
  const float pivot = __bfloat162float(r5[(size_t)b * T * SG_C5P + c]);
  const float mean = pivot + a1 / n;
  const float var = fmaxf(a2 - a1 * a1 / n, 0.f) / (float)(Tv - 1);   // unbiased (torch.std default)
  const float sd = sqrtf(var);
  const bool real = c < SG_C5;
  const float mu = real ? bn_mean[c] : 0.f, is = real ? bn_istd[c] : 0.f;
  stats[(size_t)b * SG_STATS + c] = real ? (mean - mu) * is : 0.f;
  stats[(size_t)b * SG_STATS + SG_C5P + c] = real ? sd * is : 0.f;
  save_mean[(size_t)b * SG_C5P + c] = mean;
  save_std[(size_t)b * SG_C5P + c] = sd;
}

// d(stats) -> d(pre-ReLU layer-5 activation), ReLU mask and row validity applied
// grid (C5P/128, B, tsplit), block (32, 8); each thread owns 4 consecutive channels (float4 traffic)
template <typename AT>
__global__ void pool_bwd_kernel(const AT* __restrict__ r5, int T, int Tv, const float* __restrict__ bn_istd,
                                const float* __restrict__ dstats, const float* __restrict__ save_mean,
                                const float* __restrict__ save_std, AT* __restrict__ dA5) {
  const int c0 = blockIdx.x * 128 + threadIdx.x * 4, b = blockIdx.y;
  float alpha[4], beta[4], mean[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + i;
    const bool real = c < SG_C5;
    const float is = real ? bn_istd[c] : 0.f;
    const float sd = save_std[(size_t)b * SG_C5P + c];
    mean[i] = save_mean[(size_t)b * SG_C5P + c];
    alpha[i] = real ? dstats[(size_t)b * SG_STATS + c] * is / (float)Tv : 0.f;
    beta[i] = (real && sd > 0.f) ? dstats[(size_t)b * SG_STATS + SG_C5P + c] * is / ((float)(Tv - 1) * sd) : 0.f;
  }
  const size_t off = (size_t)b * T * SG_C5P + c0;
  for (int t = blockIdx.z * 8 + threadIdx.y; t < T; t += 8 * gridDim.z) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < Tv) {
      const float4 r = ld4(r5 + off + (size_t)t * SG_C5P);
      o.x = r.x > 0.f ? fmaf(beta[0], r.x - mean[0], alpha[0]) : 0.f;
      o.y = r.y > 0.f ? fmaf(beta[1], r.y - mean[1], alpha[1]) : 0.f;
      o.z = r.z > 0.f ? fmaf(beta[2], r.z - mean[2], alpha[2]) : 0.f;
      o.w = r.w > 0.f ? fmaf(beta[3], r.w - mean[3], alpha[3]) : 0.f;
    }
    st4(dA5 + off + (size_t)t * SG_C5P, o);
  }
}

// Per-(utterance, channel pair) coefficients of the pooling adjoint for the fused layer-5 dgrad (sg_conv_tc.cu, XFORM):
// dA5[t, c] = (t < Tv && r5 > 0) ? alpha' + beta * r5,  alpha' = alpha - beta * mean  (same alpha / beta as pool_bwd_kernel),
// stored as bf16x2 {alpha'(c), alpha'(c+1)}, {beta(c), beta(c+1)} for the packed arithmetic of the tile transform.
__global__ void pool_bwd_params_kernel(int Tv, int npair, const float* __restrict__ bn_istd, const float* __restrict__ dstats,
                                       const float* __restrict__ save_mean, const float* __restrict__ save_std,
                                       uint2* __restrict__ ab) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npair) return;
  float al[2], be[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = 2 * p + k;
    const int b = idx / SG_C5P, c = idx - b * SG_C5P;
    const bool real = c < SG_C5;
    const float is = real ? bn_istd[c] : 0.f;
    const float sd = save_std[idx], mean = save_mean[idx];
    const float alpha = real ? dstats[(size_t)b * SG_STATS + c] * is / (float)Tv : 0.f;
    be[k] = (real && sd > 0.f) ? dstats[(size_t)b * SG_STATS + SG_C5P + c] * is / ((float)(Tv - 1) * sd) : 0.f;
    al[k] = fmaf(-be[k], mean, alpha);
  }
  const __nv_bfloat162 a2 = __floats2bfloat162_rn(al[0], al[1]), b2 = __floats2bfloat162_rn(be[0], be[1]);
  ab[p] = make_uint2(*reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&b2));
}

// Shifted sum over the taps of the per-tap layer-1 dgrad: out[r, f] = sum_k G[r - k * dil, k * 32 + f] (rows before the
// first one contribute nothing, exactly like the zero-filled TMA rows of the K = taps * 512 form; rows of the previous
// utterance hold zeros there because the layer-2 dgrad epilogue zeroes its invalid frames).  One float4 per thread.
__global__ void tap_gather_kernel(const float* __restrict__ G, int ldg, float* __restrict__ out, int ldo, size_t rows, int taps, int dil) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t r = i >> 3;
  const int j = (int)(i & 7);
  if (r >= rows) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < taps; ++k) {
    const long long rr = (long long)r - (long long)k * dil;
    if (rr < 0) break;
    const float4 v = __ldcs(reinterpret_cast<const float4*>(G + (size_t)rr * ldg + k * 32) + j);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *(reinterpret_cast<float4*>(out + r * ldo) + j) = acc;
}

// ---------------------------------------------------------------------------------------------
// block helpers (blockDim.x == 256)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a += red[i];
  return a;
}

// ---------------------------------------------------------------------------------------------
// head forward: e2 (after LDA) -> length-norm -> PLDA transform -> normalised embedding q
// one CTA (256 threads) per utterance
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_fwd_kernel(SgHeadConst H, const float* __restrict__ e2, float* __restrict__ tsave,
                float* __restrict__ scal, float* __restrict__ emb) {
  extern __shared__ float sm[];
  float* v = sm;                 // [Lp]
  float* red = sm + H.Lp;        // [8]
  const int b = blockIdx.x, tid = threadIdx.x, L = H.L;
  float x = 0.f;
  for (int j = tid; j < L; j += 256) { float e = e2[(size_t)b * H.Lp + j]; x = fmaf(e, e, x); }
  const float norm = sqrtf(block_sum(x, red));
  const float ratio = sqrtf((float)L) / norm;                      // xvector_extract.py:31-38
  for (int j = tid; j < L; j += 256) v[j] = e2[(size_t)b * H.Lp + j] * ratio - H.plda_mean[j];
  __syncthreads();
  float tloc[2] = {0.f, 0.f};                                      // L <= 512
  float s = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = tid + 256 * h;
    if (i < L) {
      float a = 0.f;
      for (int j = 0; j < L; ++j) a = fmaf(H.plda_Tt[(size_t)j * L + i], v[j], a);   // plda.py:75
      tloc[h] = a;
      s = fmaf(a * a, H.inv_psi1[i], s);
    }
  }
  s = block_sum(s, red);
  const float factor = sqrtf((float)L / s);                        // plda.py:92-97
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = tid + 256 * h;
    if (i < L) {
      tsave[(size_t)b * H.Lp + i] = tloc[h];
      emb[(size_t)b * L + i] = tloc[h] * factor;
    }
  }
  if (tid == 0) { scal[b * 4 + 0] = ratio; scal[b * 4 + 1] = factor; scal[b * 4 + 2] = s; scal[b * 4 + 3] = norm; }
}

// head backward: dq -> de2 (norm treated as a constant: reference quirk Q2)
__global__ void __launch_bounds__(256)
head_bwd_kernel(SgHeadConst H, const float* __restrict__ dq, const float* __restrict__ tsave,
                const float* __restrict__ scal, float* __restrict__ de2) {
  extern __shared__ float sm[];
  float* dt = sm;                // [Lp]
  float* red = sm + H.Lp;
  const int b = blockIdx.x, tid = threadIdx.x, L = H.L;
  const float ratio = scal[b * 4 + 0], factor = scal[b * 4 + 1], s = scal[b * 4 + 2];
  float dot = 0.f;
  for (int i = tid; i < L; i += 256) dot = fmaf(dq[(size_t)b * L + i], tsave[(size_t)b * H.Lp + i], dot);
  dot = block_sum(dot, red);
  const float k = factor / s * dot;
  for (int i = tid; i < L; i += 256)
    dt[i] = factor * dq[(size_t)b * L + i] - k * H.inv_psi1[i] * tsave[(size_t)b * H.Lp + i];
  __syncthreads();
  for (int j = tid; j < H.Lp; j += 256) {
    float a = 0.f;
    if (j < L)
      for (int i = 0; i < L; ++i) a = fmaf(H.plda_T[(size_t)i * L + j], dt[i], a);
    de2[(size_t)b * H.Lp + j] = a * ratio;
  }
}

// ---------------------------------------------------------------------------------------------
// The same two stages with HEAD_U utterances per CTA and the L x L transform staged in shared memory.  One utterance per
// CTA streams the 160 KB matrix (L = 200) from L2 once per utterance: 164 MB of L2 reads per launch at B = 1024 (33 us),
// and at small B a chain of dependent L2 round trips (the head was 18 % of the B = 128 step).  Here the matrix is read once
// per 8 utterances and every weight fetched from shared memory feeds 8 accumulators.  Per utterance the arithmetic and its
// order are those of head_fwd_kernel / head_bwd_kernel (the fallbacks for L*L beyond the shared-memory budget): same bits.
// ---------------------------------------------------------------------------------------------
// U block sums with two barriers in total; per value the reduction tree is block_sum's (warp butterfly, then the eight warp
// partials in order), so the results are bit-identical to U separate block_sum calls.  red: [U][8]
template <int U>
__device__ __forceinline__ void block_sum_multi(float (&v)[U], float* red) {
#pragma unroll
  for (int u = 0; u < U; ++u) v[u] = warp_sum(v[u]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int u = 0; u < U; ++u) red[u * 8 + (threadIdx.x >> 5)] = v[u];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += red[u * 8 + i];
    v[u] = a;
  }
}

// The matrix arrives by bulk asynchronous copies (cp.async.bulk, one thread issues, an mbarrier counts the bytes) while the
// CTA computes the length norms; U = 8 utterances per CTA for large batches, 2 for small ones (more CTAs, shorter chains).
__device__ __forceinline__ void head_stage_matrix(float* Ts, const float* __restrict__ src, int nfloat, uint64_t* bar) {
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)nfloat * 4u;                     // multiple of 16 (checked by the host)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = min(32768u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(Ts) + off), "l"(reinterpret_cast<const char*>(src) + off), "r"(n), "r"(bar_a) : "memory");
    }
  }
}
__device__ __forceinline__ void head_wait_matrix(uint64_t* bar) {
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_a) : "memory");
}

template <int U>
__global__ void __launch_bounds__(256)
head_fwd_multi_kernel(SgHeadConst H, const float* __restrict__ e2, int B, float* __restrict__ tsave,
                      float* __restrict__ scal, float* __restrict__ emb) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) uint64_t mbar;
  const int L = H.L, Lp = H.Lp, tid = threadIdx.x;
  float* Ts = sm;                              // [L][L]  plda_Tt: row j, column i
  float* v = Ts + (size_t)L * L;               // [U][Lp]
  float* red = v + U * Lp;                     // [U][8]
  const int b0 = blockIdx.x * U, nu = min(U, B - b0);
  head_stage_matrix(Ts, H.plda_Tt, L * L, &mbar);
  float ratio[U], norm[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {                                      // all loads in flight together, one multi-value reduction
    float x = 0.f;
    if (u < nu)
      for (int j = tid; j < L; j += 256) { float e = e2[(size_t)(b0 + u) * Lp + j]; x = fmaf(e, e, x); }
    norm[u] = x;
  }
  block_sum_multi<U>(norm, red);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    ratio[u] = 0.f;
    if (u < nu) {                                                    // CTA-uniform
      norm[u] = sqrtf(norm[u]);
      ratio[u] = sqrtf((float)L) / norm[u];                          // xvector_extract.py:31-38
      for (int j = tid; j < Lp; j += 256) v[u * Lp + j] = j < L ? e2[(size_t)(b0 + u) * Lp + j] * ratio[u] - H.plda_mean[j] : 0.f;
    } else {
      for (int j = tid; j < Lp; j += 256) v[u * Lp + j] = 0.f;
    }
  }
  __syncthreads();                                                   // v[] complete (and the mbarrier initialised long ago)
  head_wait_matrix(&mbar);
  float tloc[2][U];                                                  // L <= 512
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = tid + 256 * h;
#pragma unroll
    for (int u = 0; u < U; ++u) tloc[h][u] = 0.f;
    if (i < L) {
      int j = 0;
      for (; j + 4 <= L; j += 4) {                                   // plda.py:75; per utterance: the same j order as head_fwd_kernel
        const float w0 = Ts[(size_t)j * L + i], w1 = Ts[(size_t)(j + 1) * L + i], w2 = Ts[(size_t)(j + 2) * L + i], w3 = Ts[(size_t)(j + 3) * L + i];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float4 x = *reinterpret_cast<const float4*>(v + u * Lp + j);   // broadcast read
          float a = tloc[h][u];
          a = fmaf(w0, x.x, a); a = fmaf(w1, x.y, a); a = fmaf(w2, x.z, a); a = fmaf(w3, x.w, a);
          tloc[h][u] = a;
        }
      }
      for (; j < L; ++j) {
        const float w = Ts[(size_t)j * L + i];
#pragma unroll
        for (int u = 0; u < U; ++u) tloc[h][u] = fmaf(w, v[u * Lp + j], tloc[h][u]);
      }
    }
  }
  float ssum[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = tid + 256 * h;
      if (i < L) s = fmaf(tloc[h][u] * tloc[h][u], H.inv_psi1[i], s);
    }
    ssum[u] = s;
  }
  block_sum_multi<U>(ssum, red);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < nu) {
      const int b = b0 + u;
      const float s = ssum[u];
      const float factor = sqrtf((float)L / s);                      // plda.py:92-97
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i = tid + 256 * h;
        if (i < L) {
          tsave[(size_t)b * Lp + i] = tloc[h][u];
          emb[(size_t)b * L + i] = tloc[h][u] * factor;
        }
      }
      if (tid == 0) { scal[b * 4 + 0] = ratio[u]; scal[b * 4 + 1] = factor; scal[b * 4 + 2] = s; scal[b * 4 + 3] = norm[u]; }
    }
  }
}

template <int U>
__global__ void __launch_bounds__(256)
head_bwd_multi_kernel(SgHeadConst H, const float* __restrict__ dq, int B, const float* __restrict__ tsave,
                      const float* __restrict__ scal, float* __restrict__ de2) {
  extern __shared__ __align__(16) float sm[];
  __shared__ __align__(8) uint64_t mbar;
  const int L = H.L, Lp = H.Lp, tid = threadIdx.x;
  float* Ts = sm;                              // [L][L]  plda_T: row i, column j
  float* dt = Ts + (size_t)L * L;              // [U][Lp]
  float* red = dt + U * Lp;                    // [U][8]
  const int b0 = blockIdx.x * U, nu = min(U, B - b0);
  head_stage_matrix(Ts, H.plda_T, L * L, &mbar);
  float ratio[U], dots[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float dot = 0.f;
    if (u < nu)
      for (int i = tid; i < L; i += 256) dot = fmaf(dq[(size_t)(b0 + u) * L + i], tsave[(size_t)(b0 + u) * Lp + i], dot);
    dots[u] = dot;
  }
  block_sum_multi<U>(dots, red);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    ratio[u] = 0.f;
    if (u < nu) {
      const int b = b0 + u;
      ratio[u] = scal[b * 4 + 0];
      const float factor = scal[b * 4 + 1], s = scal[b * 4 + 2];
      const float k = factor / s * dots[u];
      for (int i = tid; i < Lp; i += 256)
        dt[u * Lp + i] = i < L ? factor * dq[(size_t)b * L + i] - k * H.inv_psi1[i] * tsave[(size_t)b * Lp + i] : 0.f;
    } else {
      for (int i = tid; i < Lp; i += 256) dt[u * Lp + i] = 0.f;
    }
  }
  __syncthreads();
  head_wait_matrix(&mbar);
  for (int j = tid; j < Lp; j += 256) {
    float a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = 0.f;
    if (j < L) {
      int i = 0;
      for (; i + 4 <= L; i += 4) {
        const float w0 = Ts[(size_t)i * L + j], w1 = Ts[(size_t)(i + 1) * L + j], w2 = Ts[(size_t)(i + 2) * L + j], w3 = Ts[(size_t)(i + 3) * L + j];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float4 x = *reinterpret_cast<const float4*>(dt + u * Lp + i);
          float c = a[u];
          c = fmaf(w0, x.x, c); c = fmaf(w1, x.y, c); c = fmaf(w2, x.z, c); c = fmaf(w3, x.w, c);
          a[u] = c;
        }
      }
      for (; i < L; ++i) {
        const float w = Ts[(size_t)i * L + j];
#pragma unroll
        for (int u = 0; u < U; ++u) a[u] = fmaf(w, dt[u * Lp + i], a[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (u < nu) de2[(size_t)(b0 + u) * Lp + j] = a[u] * ratio[u];
  }
}

// ---------------------------------------------------------------------------------------------
// PLDA log-likelihood-ratio scoring + decision (plda.py:140-190, defended_model.py:167-170)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_fwd_kernel(SgHeadConst H, const float* __restrict__ emb, const float* __restrict__ enroll, int S,
                 float threshold, float* __restrict__ scores, long long* __restrict__ decisions) {
  extern __shared__ float sm[];
  float* q = sm;                 // [L]
  float* sc = sm + H.Lp;         // [S]
  float* red = sc + S;           // [8]
  const int b = blockIdx.x, tid = threadIdx.x, L = H.L, lane = tid & 31, warp = tid >> 5;
  float s2 = 0.f;
  for (int i = tid; i < L; i += 256) {
    float x = emb[(size_t)b * L + i];
    q[i] = x;
    s2 = fmaf(x * x, H.inv_psi1[i], s2);
  }
  s2 = block_sum(s2, red);                                         // also orders the q[] writes
  const float without = -0.5f * (H.logdet_without + H.log2pi_L + s2);
  for (int n = warp; n < S; n += 8) {
    float a = 0.f;
    for (int i = lane; i < L; i += 32) {
      float d = q[i] - H.psi_ratio[i] * enroll[(size_t)n * L + i];
      a = fmaf(d * d, H.inv_var_given[i], a);
    }
    a = warp_sum(a);
    if (lane == 0) {
      float given = -0.5f * (H.logdet_given + H.log2pi_L + a);
      float v = given - without;
      sc[n] = v;
      scores[(size_t)b * S + n] = v;
    }
  }
  __syncthreads();
  if (decisions != nullptr && warp == 0) {
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int n = lane; n < S; n += 32)
      if (sc[n] > best) { best = sc[n]; bi = n; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) decisions[b] = (best > threshold) ? (long long)bi : -1LL;
  }
}

// dscores -> demb
__global__ void __launch_bounds__(256)
score_bwd_kernel(SgHeadConst H, const float* __restrict__ emb, const float* __restrict__ dscores,
                 const float* __restrict__ enroll, int S, float* __restrict__ demb) {
  extern __shared__ float sm[];
  float* ds = sm;                // [S]
  const int b = blockIdx.x, tid = threadIdx.x, L = H.L;
  for (int n = tid; n < S; n += 256) ds[n] = dscores[(size_t)b * S + n];
  __syncthreads();
  for (int i = tid; i < L; i += 256) {
    const float qi = emb[(size_t)b * L + i], pr = H.psi_ratio[i], iv = H.inv_var_given[i], w = H.inv_psi1[i];
    float a = 0.f;
    for (int n = 0; n < S; ++n) a = fmaf(ds[n], -(qi - pr * enroll[(size_t)n * L + i]) * iv + qi * w, a);
    demb[(size_t)b * L + i] = a;
  }
}

// ---------------------------------------------------------------------------------------------
// losses (attack/utils.py:7-102) and d(sum loss)/d(scores).  One warp per utterance.
// ---------------------------------------------------------------------------------------------
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax warp_argmax(float v, int i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  ArgMax r; r.v = v; r.i = i; return r;
}

__global__ void loss_kernel(const float* __restrict__ scores, const long long* __restrict__ y, int B, int S,
                            sg_loss_params lp, float* __restrict__ loss, float* __restrict__ dscores) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* sc = scores + (size_t)b * S;
  float* ds = dscores ? dscores + (size_t)b * S : nullptr;
  const int lab = (int)y[b];
  if (ds) for (int n = lane; n < S; n += 32) ds[n] = 0.f;
  __syncwarp();
  // Labels the reference rejects (an index error for label >= S, the SV assert at attack/utils.py:50 for labels outside
  // {0, -1}) cannot raise from a kernel: they give a NaN loss and a zero gradient instead of an out-of-bounds read.
  if (lab < -1 || lab >= S || (lp.task == SG_TASK_SV && lab > 0)) {
    if (lane == 0) loss[b] = __int_as_float(0x7fc00000);
    return;
  }
  float L = 0.f;
  if (lp.loss == SG_LOSS_CE && lp.task == SG_TASK_CSI) {
    if (lab >= 0) {
      float mx = -INFINITY;
      for (int n = lane; n < S; n += 32) mx = fmaxf(mx, sc[n]);
      mx = warp_max(mx);
      float se = 0.f;
      for (int n = lane; n < S; n += 32) se += expf(sc[n] - mx);
      se = warp_sum(se);
      L = logf(se) + mx - sc[lab];
      if (ds) for (int n = lane; n < S; n += 32) ds[n] = expf(sc[n] - mx) / se - (n == lab ? 1.f : 0.f);
    }
    if (lane == 0) loss[b] = L;
    return;
  }
  // ---- margin family ----
  // every branch yields loss = sum_k coef_k * score[idx_k] + const with at most 2 active indices
  int i0 = -1, i1 = -1; float c0 = 0.f, c1 = 0.f;
  const float thr = lp.threshold, conf = lp.confidence;
  if (lp.task == SG_TASK_SV) {
    const float s = sc[0];
    const bool plus = (lab == 0) ? !lp.targeted : (lp.targeted != 0);   // loss = s + conf - thr
    L = plus ? s + conf - thr : thr + conf - s;
    i0 = 0; c0 = plus ? 1.f : -1.f;
  } else if (lab >= 0) {
    ArgMax oth; { float bv = -INFINITY; int bi = 0x7fffffff;
      for (int n = lane; n < S; n += 32) { float v = (n == lab) ? -10000.f : sc[n]; if (v > bv) { bv = v; bi = n; } }
      oth = warp_argmax(bv, bi); }
    const float real = sc[lab];
    const float gother = (oth.i != lab) ? 1.f : 0.f;               // (1-onehot)*scores: no grad through the label slot
    if (lp.targeted) {
      if (lp.task == SG_TASK_CSI) { L = oth.v + conf - real; i0 = oth.i; c0 = gother; }
      else { L = fmaxf(oth.v, thr) + conf - real; i0 = oth.i; c0 = (oth.v >= thr) ? gother : 0.f; }
      i1 = lab; c1 = -1.f;
    } else if (lp.task == SG_TASK_CSI) {
      L = real + conf - oth.v; i0 = lab; c0 = 1.f; i1 = oth.i; c1 = -gother;
    } else {
      ArgMax mx; { float bv = -INFINITY; int bi = 0x7fffffff;
        for (int n = lane; n < S; n += 32) if (sc[n] > bv) { bv = sc[n]; bi = n; }
        mx = warp_argmax(bv, bi); }
      const float f_rej = mx.v + conf - thr;
      const float f_mis = fmaxf(real, thr) + conf - oth.v;
      L = fminf(f_rej, f_mis);
      // torch.minimum: ties split the gradient evenly
      const float wr = (f_rej < f_mis) ? 1.f : ((f_rej == f_mis) ? 0.5f : 0.f), wm = 1.f - wr;
      if (ds && lane == 0) {
        atomicAdd(&ds[mx.i], wr);
        if (real >= thr) atomicAdd(&ds[lab], wm);
        atomicAdd(&ds[oth.i], -wm * gother);
      }
      i0 = -2;                                                     // handled
    }
  } else if (lp.task == SG_TASK_OSI) {                              // imposter, OSI
    ArgMax mx; { float bv = -INFINITY; int bi = 0x7fffffff;
      for (int n = lane; n < S; n += 32) if (sc[n] > bv) { bv = sc[n]; bi = n; }
      mx = warp_argmax(bv, bi); }
    if (lp.targeted) { L = mx.v + conf - thr; i0 = mx.i; c0 = 1.f; }
    else { L = thr + conf - mx.v; i0 = mx.i; c0 = -1.f; }
  }  // CSI imposter: 0 * sum(scores): zero loss, zero gradient (quirk Q7)
  float g = 1.f;
  if (lp.clip_max) {                                               // torch.max(0, loss): tie -> half gradient
    g = (L > 0.f) ? 1.f : ((L == 0.f) ? 0.5f : 0.f);
    L = fmaxf(L, 0.f);
  }
  __syncwarp();
  if (ds && lane == 0) {
    if (i0 == -2) { for (int n = 0; n < S; ++n) ds[n] *= g; }
    else {
      if (i0 >= 0) ds[i0] += g * c0;
      if (i1 >= 0) ds[i1] += g * c1;
    }
  }
  if (lane == 0) loss[b] = L;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int sg_pool_fwd_launch(const void* r5, int bf16, int B, int T, int Tv, const float* bn_mean, const float* bn_istd,
                       float* stats, float* save_mean, float* save_std, cudaStream_t st) {
  if (bf16 && Tv >= 2) pool_fwd_onepass_kernel<<<dim3(SG_C5P / 256, B), dim3(32, 8), 0, st>>>((const __nv_bfloat16*)r5, T, Tv, bn_mean, bn_istd, stats, save_mean, save_std);
  else if (bf16) pool_fwd_kernel<__nv_bfloat16><<<dim3(SG_C5P / 128, B), dim3(32, 8), 0, st>>>((const __nv_bfloat16*)r5, T, Tv, bn_mean, bn_istd, stats, save_mean, save_std);
  else pool_fwd_kernel<float><<<dim3(SG_C5P / 128, B), dim3(32, 8), 0, st>>>((const float*)r5, T, Tv, bn_mean, bn_istd, stats, save_mean, save_std);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_pool_bwd_launch(const void* r5, int bf16, int B, int T, int Tv, const float* bn_istd, const float* dstats,
                       const float* save_mean, const float* save_std, void* dA5, cudaStream_t st) {
  int tsplit = (B >= 64) ? 1 : 4;
  if (bf16) pool_bwd_kernel<__nv_bfloat16><<<dim3(SG_C5P / 128, B, tsplit), dim3(32, 8), 0, st>>>((const __nv_bfloat16*)r5, T, Tv, bn_istd, dstats, save_mean, save_std, (__nv_bfloat16*)dA5);
  else pool_bwd_kernel<float><<<dim3(SG_C5P / 128, B, tsplit), dim3(32, 8), 0, st>>>((const float*)r5, T, Tv, bn_istd, dstats, save_mean, save_std, (float*)dA5);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_pool_bwd_params_launch(int B, int Tv, const float* bn_istd, const float* dstats, const float* save_mean,
                              const float* save_std, float* ab, cudaStream_t st) {
  const int n = B * SG_C5P / 2;
  pool_bwd_params_kernel<<<(n + 255) / 256, 256, 0, st>>>(Tv, n, bn_istd, dstats, save_mean, save_std, reinterpret_cast<uint2*>(ab));
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_tap_gather_launch(const float* G, int ldg, float* out, int ldo, size_t rows, int taps, int dil, cudaStream_t st) {
  const size_t n = rows * 8;
  tap_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(G, ldg, out, ldo, rows, taps, dil);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
static size_t head_multi_smem(const SgHeadConst& H, int U) { return ((size_t)H.L * H.L + (size_t)U * H.Lp + 8 * U) * sizeof(float); }
static bool head_multi_ok(const SgHeadConst& H) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("SGB200_HEAD_MULTI"); enabled = e ? (atoi(e) != 0) : 1; }
  if (!enabled || head_multi_smem(H, 8) > 200 * 1024 || (H.L * H.L) % 4 != 0 || H.Lp % 4 != 0) return false;
  static std::atomic<unsigned long long> configured{0};
  if (sg_first_on_device(&configured)) {
    if (cudaFuncSetAttribute(head_fwd_multi_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(head_bwd_multi_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(head_fwd_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(head_bwd_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
  }
  return true;
}
int sg_head_fwd_launch(const SgHeadConst& H, const float* e2, int B, float* tsave, float* scal, float* emb, cudaStream_t st) {
  if (head_multi_ok(H)) {
    if (B >= 512) head_fwd_multi_kernel<8><<<(B + 7) / 8, 256, head_multi_smem(H, 8), st>>>(H, e2, B, tsave, scal, emb);
    else head_fwd_multi_kernel<2><<<(B + 1) / 2, 256, head_multi_smem(H, 2), st>>>(H, e2, B, tsave, scal, emb);
  } else head_fwd_kernel<<<B, 256, (H.Lp + 8) * sizeof(float), st>>>(H, e2, tsave, scal, emb);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_head_bwd_launch(const SgHeadConst& H, const float* dq, int B, const float* tsave, const float* scal, float* de2, cudaStream_t st) {
  if (head_multi_ok(H)) {
    if (B >= 512) head_bwd_multi_kernel<8><<<(B + 7) / 8, 256, head_multi_smem(H, 8), st>>>(H, dq, B, tsave, scal, de2);
    else head_bwd_multi_kernel<2><<<(B + 1) / 2, 256, head_multi_smem(H, 2), st>>>(H, dq, B, tsave, scal, de2);
  } else head_bwd_kernel<<<B, 256, (H.Lp + 8) * sizeof(float), st>>>(H, dq, tsave, scal, de2);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_score_fwd_launch(const SgHeadConst& H, const float* emb, int B, const float* enroll, int S, float threshold,
                        float* scores, long long* decisions, cudaStream_t st) {
  score_fwd_kernel<<<B, 256, (H.Lp + S + 8) * sizeof(float), st>>>(H, emb, enroll, S, threshold, scores, decisions);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_score_bwd_launch(const SgHeadConst& H, const float* emb, const float* dscores, int B, const float* enroll, int S,
                        float* demb, cudaStream_t st) {
  score_bwd_kernel<<<B, 256, (S + 8) * sizeof(float), st>>>(H, emb, dscores, enroll, S, demb);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_loss_launch(const float* scores, const long long* y, int B, int S, const sg_loss_params& lp, float* loss,
                   float* dscores, cudaStream_t st) {
  loss_kernel<<<(B + 7) / 8, 256, 0, st>>>(scores, y, B, S, lp, loss, dscores);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
