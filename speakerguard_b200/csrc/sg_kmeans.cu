// C1: FeCo feature compression (reference defense/feature_level.py:18-50, :168-217), replacing the
// libKMCUDA dependency: batched per-utterance Lloyd k-means (k-means++ seeding, L2 metric, stop when
// the fraction of re-assigned frames drops to `tol`, libKMCUDA's default 0.01) plus the differentiable
// segment means the reference builds from the cluster ids (:202-217).
//
// One CTA per utterance: frames and centroids live in shared memory for the whole run (n x dim and
// k x dim floats), so HBM traffic is one read of the features and one write of ids / means.
#include <math.h>

#include "sg_common.cuh"

#define KM_THREADS 256

__device__ __forceinline__ uint32_t km_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  // small counter-based generator (xorshift-multiply mix); quality is irrelevant beyond seeding
  uint32_t x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ (d * 0x27D4EB2Fu);
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  return x;
}

__device__ __forceinline__ float km_block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < KM_THREADS / 32; ++i) a += red[i];
  return a;
}

// feat [B, n, ld] -> ids [B, n] (int32).  smem: X[n*dim], Cn[k*dim], mind[n], ids[n], cnt[k], red[40]
__global__ void __launch_bounds__(KM_THREADS)
feco_kmeans_kernel(const float* __restrict__ feat, int ld, int n, int dim, int k, uint32_t seed_lo, uint32_t seed_hi,
                   int max_iter, float tol, int* __restrict__ ids_out) {
  extern __shared__ float sm[];
  float* X = sm;
  float* Cn = X + (size_t)n * dim;
  float* mind = Cn + (size_t)k * dim;
  int* ids = reinterpret_cast<int*>(mind + n);
  int* cnt = ids + n;
  float* red = reinterpret_cast<float*>(cnt + k);
  __shared__ int s_pick;
  __shared__ int s_changed;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* fb = feat + (size_t)b * n * ld;
  for (int i = tid; i < n * dim; i += KM_THREADS) X[i] = fb[(size_t)(i / dim) * ld + (i % dim)];
  __syncthreads();

  // ---- k-means++ seeding ----
  if (tid == 0) s_pick = (int)(km_hash(seed_lo, seed_hi, (uint32_t)b, 0u) % (uint32_t)n);
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    const int pick = s_pick;
    for (int d = tid; d < dim; d += KM_THREADS) Cn[c * dim + d] = X[pick * dim + d];
    __syncthreads();
    float local = 0.f;
    for (int j = tid; j < n; j += KM_THREADS) {
      float s = 0.f;
      for (int d = 0; d < dim; ++d) { float t = X[j * dim + d] - Cn[c * dim + d]; s = fmaf(t, t, s); }
      const float m = (c == 0) ? s : fminf(mind[j], s);
      mind[j] = m;
      local += m;
    }
    const float total = km_block_sum(local, red);
    if (c + 1 < k) {
      if (tid == 0) {
        // sample j with probability mind[j] / total (serial scan over n <= ~1000 values: negligible)
        const float u = (km_hash(seed_lo, seed_hi, (uint32_t)b, (uint32_t)(c + 1)) >> 8) * (1.0f / 16777216.0f);
        const float r = u * total;
        float acc = 0.f;
        int sel = n - 1;
        for (int j = 0; j < n; ++j) { acc += mind[j]; if (acc > r) { sel = j; break; } }
        s_pick = sel;
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < n; j += KM_THREADS) ids[j] = -1;
  __syncthreads();

  // ---- Lloyd iterations ----
  for (int it = 0; it < max_iter; ++it) {
    if (tid == 0) s_changed = 0;
    __syncthreads();
    int changed = 0;
    for (int j = tid; j < n; j += KM_THREADS) {
      float best = INFINITY; int bi = 0;
      for (int c = 0; c < k; ++c) {
        float s = 0.f;
        for (int d = 0; d < dim; ++d) { float t = X[j * dim + d] - Cn[c * dim + d]; s = fmaf(t, t, s); }
        if (s < best) { best = s; bi = c; }
      }
      if (ids[j] != bi) { ids[j] = bi; ++changed; }
    }
    if (changed) atomicAdd(&s_changed, changed);
    __syncthreads();
    // centroid update, deterministic: thread c scans all frames in order
    for (int c = tid; c < k; c += KM_THREADS) {
      int m = 0;
      for (int j = 0; j < n; ++j) m += (ids[j] == c);
      cnt[c] = m;
      if (m > 0) {
        for (int d = 0; d < dim; ++d) {
          float s = 0.f;
          for (int j = 0; j < n; ++j) s += (ids[j] == c) ? X[j * dim + d] : 0.f;
          Cn[c * dim + d] = s / (float)m;
        }
      }                                                            // empty cluster: keep the old centroid
    }
    __syncthreads();
    if ((float)s_changed <= tol * (float)n) break;                  // uniform: s_changed is shared
    __syncthreads();
  }
  for (int j = tid; j < n; j += KM_THREADS) ids_out[(size_t)b * n + j] = ids[j];
}

// out[b,i,:] = mean of feat[b, ids==i, :]; empty cluster -> feat[b,i,:] when force (feature_level.py:205-216)
__global__ void feco_means_fwd_kernel(const float* __restrict__ feat, int ld_in, const int* __restrict__ ids, int n, int dim,
                                      int k, int force, float* __restrict__ out, int ld_out, int* __restrict__ counts) {
  const int b = blockIdx.y, i = blockIdx.x, d = threadIdx.x;
  const int* ib = ids + (size_t)b * n;
  const float* fb = feat + (size_t)b * n * ld_in;
  float s = 0.f;
  int m = 0;
  for (int j = 0; j < n; ++j)
    if (ib[j] == i) { ++m; if (d < dim) s += fb[(size_t)j * ld_in + d]; }
  float v = 0.f;
  if (d < dim) v = m > 0 ? s / (float)m : ((force && i < n) ? fb[(size_t)i * ld_in + d] : 0.f);
  if (d < ld_out) out[((size_t)b * k + i) * ld_out + d] = v;
  if (d == 0) counts[(size_t)b * k + i] = m;
}
// dfeat[b,j,:] = dout[b, ids[j], :] / count[ids[j]]  (+ dout[b,j,:] if cluster j is empty and force)
__global__ void feco_means_bwd_kernel(const float* __restrict__ dout, int ld_out, const int* __restrict__ ids,
                                      const int* __restrict__ counts, int n, int dim, int k, int force,
                                      float* __restrict__ dfeat, int ld_in) {
  const int b = blockIdx.y, j = blockIdx.x, d = threadIdx.x;
  if (d >= ld_in) return;
  const int id = ids[(size_t)b * n + j];
  float g = 0.f;
  if (d < dim) {
    if (id >= 0 && id < k) g = dout[((size_t)b * k + id) * ld_out + d] / (float)counts[(size_t)b * k + id];
    if (force && j < k && counts[(size_t)b * k + j] == 0) g += dout[((size_t)b * k + j) * ld_out + d];
  }
  dfeat[((size_t)b * n + j) * ld_in + d] = g;
}

size_t sg_kmeans_smem(int n, int dim, int k) {
  return ((size_t)n * dim + (size_t)k * dim + n) * sizeof(float) + ((size_t)n + k) * sizeof(int) + 40 * sizeof(float);
}

int sg_feco_kmeans_launch(const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed, int max_iter, float tol,
                          int* ids, cudaStream_t st) {
  const size_t smem = sg_kmeans_smem(n, dim, k);
  if (smem > 200 * 1024) { sg_set_error("FeCo k-means: utterance too long for the shared-memory kernel (n=%d dim=%d k=%d)", n, dim, k); return SG_EUNSUPPORTED; }
  SG_CUDA_CHECK(cudaFuncSetAttribute(feco_kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  feco_kmeans_kernel<<<B, KM_THREADS, smem, st>>>(feat, ld, n, dim, k, (uint32_t)seed, (uint32_t)(seed >> 32), max_iter, tol, ids);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_feco_means_fwd_launch(const float* feat, int ld_in, const int* ids, int B, int n, int dim, int k, int force,
                             float* out, int ld_out, int* counts, cudaStream_t st) {
  feco_means_fwd_kernel<<<dim3(k, B), 32, 0, st>>>(feat, ld_in, ids, n, dim, k, force, out, ld_out, counts);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_feco_means_bwd_launch(const float* dout, int ld_out, const int* ids, const int* counts, int B, int n, int dim, int k,
                             int force, float* dfeat, int ld_in, cudaStream_t st) {
  feco_means_bwd_kernel<<<dim3(n, B), 32, 0, st>>>(dout, ld_out, ids, counts, n, dim, k, force, dfeat, ld_in);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
