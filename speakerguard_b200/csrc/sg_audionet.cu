// A1: AudioNet log-mel front end and its adjoint; max-pool / global-max helpers of the CNN; CW2's
// tanh-space Adam update and best-example tracking.
//
// Front end (reference model/_audionet/Preprocessor.py:85-112): pre-emphasis w[n] = x[n+1] - 0.97 x[n],
// torch.stft(n_fft 1024, hop 160, win 800 periodic Hann, centred, reflect pad), |X|^2, 32 Slaney mel
// filters over 513 bins, 10 log10(clamp(., 1e-16)).  One warp per frame; the 1024-point real FFT is a
// 512-point complex FFT (8 x 8 x 8 radix passes, 16 complex values per lane, two shared-memory
// exchanges; index maps and paddings prototyped in tools/fft512_proto.py).
// The adjoint recomputes the forward per frame, writes window-weighted frame gradients, and a gather
// kernel overlap-adds them (deterministic) through the reflect padding and the pre-emphasis.
#include <math.h>
#include <string.h>

#include "sg_audionet.cuh"

#define AN_THREADS 256
#define AN_WARPS 8
#define AN_SCRATCH 1152            // floats per warp: re[576] + im[576]

// =============================================================================================
// host: tables
// =============================================================================================
int sg_an_tables_build(SgAnTables* t) {
  memset(t, 0, sizeof(*t));
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < AN_WIN; ++i) t->window[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / AN_WIN));   // periodic Hann
  for (int lane = 0; lane < 32; ++lane)
    for (int h = 0; h < 2; ++h) {
      for (int k0 = 0; k0 < 8; ++k0) {                       // pass A: W_64^(n1*k0), n1 = (lane+32h)>>3
        int n1 = (lane + 32 * h) >> 3;
        double a = -2.0 * PI * (n1 * k0) / 64.0;
        t->twA[k0][h][lane] = make_float2((float)cos(a), (float)sin(a));
      }
      for (int k1 = 0; k1 < 8; ++k1) {                       // pass B: W_512^(n2*(k0+8k1)), k0=(lane>>3)+4h, n2=lane&7
        int k0 = (lane >> 3) + 4 * h, n2 = lane & 7;
        double a = -2.0 * PI * (n2 * (k0 + 8 * k1)) / 512.0;
        t->twB[k1][h][lane] = make_float2((float)cos(a), (float)sin(a));
      }
    }
  for (int r = 0; r < 16; ++r)
    for (int lane = 0; lane < 32; ++lane) {
      double a = 2.0 * PI * (lane + 32 * r) / 1024.0;
      t->twU[r][lane] = make_float2((float)cos(a), (float)sin(a));
    }
  // Slaney mel filterbank = librosa.filters.mel(sr=16000, n_fft=1024, n_mels=32, fmin=0, fmax=8000)
  auto hz2mel = [](double f) { return f >= 1000.0 ? 15.0 + log(f / 1000.0) / (log(6.4) / 27.0) : f / (200.0 / 3.0); };
  auto mel2hz = [](double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : m * 200.0 / 3.0; };
  double melf[AN_MELS + 2];
  const double m_lo = hz2mel(0.0), m_hi = hz2mel(8000.0);
  for (int i = 0; i < AN_MELS + 2; ++i) melf[i] = mel2hz(m_lo + (m_hi - m_lo) * i / (AN_MELS + 1));
  static double w[AN_MELS][AN_BINS];
  for (int c = 0; c < AN_MELS; ++c) {
    const double enorm = 2.0 / (melf[c + 2] - melf[c]);
    for (int b = 0; b < AN_BINS; ++b) {
      const double f = 8000.0 * b / (AN_BINS - 1);
      const double lower = (f - melf[c]) / (melf[c + 1] - melf[c]), upper = (melf[c + 2] - f) / (melf[c + 2] - melf[c + 1]);
      double v = lower < upper ? lower : upper;
      w[c][b] = v > 0.0 ? v * enorm : 0.0;
    }
  }
  for (int b = 0; b < 512; ++b) { t->bin_c0[b] = t->bin_c1[b] = 0; t->bin_w0[b] = t->bin_w1[b] = 0.f; }
  int off = 0, maxlen = 0;
  for (int c = 0; c < AN_MELS; ++c) {
    if (w[c][0] != 0.0 || w[c][512] != 0.0) return SG_EINVAL;      // bins 0 and 512 are assumed weightless
    int lo = AN_BINS, hi = -1;
    for (int b = 0; b < 512; ++b)
      if (w[c][b] > 0.0) { if (b < lo) lo = b; hi = b; }
    // float4 groups: the window starts at a multiple of 4 and is zero-padded to whole groups (an_mel reads float4 pairs)
    const int lo4 = hi >= lo ? (lo & ~3) : 0;
    const int groups = hi >= lo ? (hi - lo4) / 4 + 1 : 0;
    t->mel_lo[c] = lo4; t->mel_len[c] = groups; t->mel_off[c] = off;
    if (off + 4 * groups > AN_MELW) return SG_EINVAL;
    for (int i = 0; i < 4 * groups; ++i) {
      const int b = lo4 + i;
      const double wv = (b >= lo && b <= hi && b < 512) ? w[c][b] : 0.0;
      t->mel_w[off + i] = (float)wv;
      if (wv > 0.0) {
        if (t->bin_w0[b] == 0.f) { t->bin_c0[b] = c; t->bin_w0[b] = (float)wv; }
        else if (t->bin_w1[b] == 0.f) { t->bin_c1[b] = c; t->bin_w1[b] = (float)wv; }
        else return SG_EINVAL;
      }
    }
    off += 4 * groups;
    if (groups > maxlen) maxlen = groups;
  }
  t->mel_maxlen = maxlen;
  return SG_OK;
}

// =============================================================================================
// device helpers
// =============================================================================================
__device__ __forceinline__ float2 an_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 an_cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 an_csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ void an_fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = an_cadd(a0, a2), t1 = an_csub(a0, a2), t2 = an_cadd(a1, a3), t3 = an_csub(a1, a3);
  a0 = an_cadd(t0, t2); a2 = an_csub(t0, t2);
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);
}
__device__ __forceinline__ void an_fft8(float2 (&v)[8]) {
  an_fft4(v[0], v[2], v[4], v[6]);
  an_fft4(v[1], v[3], v[5], v[7]);
  const float h = 0.70710678118654752440f;
  float2 o1 = make_float2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
  float2 o2 = make_float2(v[5].y, -v[5].x);
  float2 o3 = make_float2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = an_cadd(e0, o0); v[4] = an_csub(e0, o0);
  v[1] = an_cadd(e1, o1); v[5] = an_csub(e1, o1);
  v[2] = an_cadd(e2, o2); v[6] = an_csub(e2, o2);
  v[3] = an_cadd(e3, o3); v[7] = an_csub(e3, o3);
}

// 512-point complex FFT across a warp.  In: z[h][n0] = element 64*n0 + lane + 32*h.
// Out: z[h][k2] = Z[lane + 32*h + 64*k2].  sre/sim: 576 floats each.
__device__ __forceinline__ void warp_fft512(float2 (&z)[2][8], const SgAnTables* T, float* sre, float* sim, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    an_fft8(z[h]);
#pragma unroll
    for (int k0 = 1; k0 < 8; ++k0) z[h][k0] = an_cmul(z[h][k0], T->twA[k0][h][lane]);
#pragma unroll
    for (int k0 = 0; k0 < 8; ++k0) { sre[k0 * 72 + lane + 32 * h] = z[h][k0].x; sim[k0 * 72 + lane + 32 * h] = z[h][k0].y; }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int base = ((lane >> 3) + 4 * h) * 72 + (lane & 7);
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) z[h][n1] = make_float2(sre[base + n1 * 8], sim[base + n1 * 8]);
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    an_fft8(z[h]);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) z[h][k1] = an_cmul(z[h][k1], T->twB[k1][h][lane]);
    const int wb = ((lane >> 3) + 4 * h) * 8 + (lane & 7);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) { sre[k1 * 65 + wb] = z[h][k1].x; sim[k1 * 65 + wb] = z[h][k1].y; }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int rb = ((lane >> 3) + 4 * h) * 65 + (lane & 7) * 8;
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) z[h][n2] = make_float2(sre[rb + n2], sim[rb + n2]);
  }
  __syncwarp();
  an_fft8(z[0]);
  an_fft8(z[1]);
}

__device__ __forceinline__ void fft512_out_to_smem(const float2 (&z)[2][8], float* sre, float* sim, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) { sre[lane + 32 * h + 64 * k2] = z[h][k2].x; sim[lane + 32 * h + 64 * k2] = z[h][k2].y; }
}

// pre-emphasised, reflect-padded sample at padded position q (frame-local i in [112,912) -> q = 160 t + i - 512)
__device__ __forceinline__ float an_sample(const float* __restrict__ xb, int M, int j) {
  j = j < 0 ? -j : (j >= M ? 2 * (M - 1) - j : j);                 // torch 'reflect' (no edge repeat)
  return __ldg(xb + j + 1) - 0.97f * __ldg(xb + j);                 // Preprocessor.py:85-86
}

// loads + windows frame t into z, runs the FFT and the real-FFT untangle; X[r] = bin lane + 32 r
__device__ __forceinline__ void an_frame_spectrum(float2 (&X)[16], const float* __restrict__ xb, int M, int t,
                                                  const SgAnTables* T, float* sre, float* sim, int lane) {
  float2 z[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int n0 = 0; n0 < 8; ++n0) {
      const int i = 2 * (64 * n0 + lane + 32 * h);                 // even sample index in the 1024 frame
      float2 v = make_float2(0.f, 0.f);
      if (i >= AN_WOFF && i < AN_WOFF + AN_WIN) {                   // AN_WOFF even: the pair is inside or outside together
        const int j = AN_HOP * t + i - AN_NFFT / 2;
        v.x = an_sample(xb, M, j) * T->window[i - AN_WOFF];
        v.y = an_sample(xb, M, j + 1) * T->window[i + 1 - AN_WOFF];
      }
      z[h][n0] = v;
    }
  warp_fft512(z, T, sre, sim, lane);
  fft512_out_to_smem(z, sre, sim, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int k = lane + 32 * r, km = (512 - k) & 511;
    float2 zk = make_float2(sre[k], sim[k]), zm = make_float2(sre[km], -sim[km]);
    float2 cs = T->twU[r][lane];
    float2 a = make_float2(0.5f * (1.f - cs.y), -0.5f * cs.x), b = make_float2(0.5f * (1.f + cs.y), 0.5f * cs.x);
    X[r] = an_cadd(an_cmul(a, zk), an_cmul(b, zm));
  }
  __syncwarp();
}

__device__ __forceinline__ void an_copy_tables(SgAnTables* dst, const SgAnTables* __restrict__ src) {
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(SgAnTables) / 16); i += blockDim.x) d[i] = s[i];
}

// lane c: mel energy of filter c; windows are float4-aligned and zero-padded (P must be finite up to index 515)
__device__ __forceinline__ float an_mel(const SgAnTables* T, const float* P, int lane) {
  const int lo = T->mel_lo[lane], len = T->mel_len[lane], off = T->mel_off[lane];
  const float4* w4 = reinterpret_cast<const float4*>(&T->mel_w[off]);
  const float4* p4 = reinterpret_cast<const float4*>(&P[lo]);
  float acc = 0.f;
  for (int i = 0; i < T->mel_maxlen; ++i)
    if (i < len) {
      const float4 w = w4[i], p = p4[i];
      acc = fmaf(w.x, p.x, acc); acc = fmaf(w.y, p.y, acc); acc = fmaf(w.z, p.z, acc); acc = fmaf(w.w, p.w, acc);
    }
  return acc;
}

#define AN_LOGSCALE 4.342944819032518f      // 10 / ln(10)

// =============================================================================================
// forward: x [B,N] -> log-mel [B,T,32]
// =============================================================================================
__global__ void __launch_bounds__(AN_THREADS)
an_logmel_fwd_kernel(const float* __restrict__ x, int N, int T_frames, int frames_per_cta, float* __restrict__ feat,
                     const SgAnTables* __restrict__ gT, float* __restrict__ stash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SgAnTables* T = reinterpret_cast<SgAnTables*>(smem_raw);
  an_copy_tables(T, gT);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sre = reinterpret_cast<float*>(smem_raw + sizeof(SgAnTables)) + warp * AN_SCRATCH;
  float* sim = sre + 576;
  const int b = blockIdx.y, M = N - 1;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * frames_per_cta, f1 = min(f0 + frames_per_cta, T_frames);
  for (int t = f0 + warp; t < f1; t += AN_WARPS) {
    float2 X[16];
    an_frame_spectrum(X, xb, M, t, T, sre, sim, lane);
#pragma unroll
    for (int r = 0; r < 16; ++r) sre[lane + 32 * r] = X[r].x * X[r].x + X[r].y * X[r].y;   // power, Preprocessor.py:28-37
    if (lane < 4) sre[512 + lane] = 0.f;                           // the zero-weighted tail of the last float4 group
    __syncwarp();
    const float me = an_mel(T, sre, lane);
    feat[((size_t)b * T_frames + t) * AN_MELS + lane] = AN_LOGSCALE * logf(fmaxf(me, 1e-16f));   // Preprocessor.py:111
    if (stash != nullptr) {
      // forward -> adjoint hand-over inside the CW2 loop: the spectrum and the mel energies of this frame (4.1 KB), so that
      // the adjoint neither reloads the frame nor repeats the forward FFT
      float* sp = stash + ((size_t)b * T_frames + t) * AN_STASH_FLOATS;
      float4* s4 = reinterpret_cast<float4*>(sp);
#pragma unroll
      for (int r = 0; r < 8; ++r) __stcs(s4 + r * 32 + lane, make_float4(X[2 * r].x, X[2 * r].y, X[2 * r + 1].x, X[2 * r + 1].y));
      __stcs(sp + 1024 + lane, me);
    }
    __syncwarp();
  }
}

// =============================================================================================
// backward, stage 1: d(log-mel) -> window-weighted frame gradients dgw [B,T,800]
// =============================================================================================
__global__ void __launch_bounds__(AN_THREADS)
an_logmel_bwd_frames_kernel(const float* __restrict__ x, int N, int T_frames, int frames_per_cta,
                            const float* __restrict__ dfeat, float* __restrict__ dgw, const SgAnTables* __restrict__ gT,
                            const float* __restrict__ stash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SgAnTables* T = reinterpret_cast<SgAnTables*>(smem_raw);
  an_copy_tables(T, gT);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sre = reinterpret_cast<float*>(smem_raw + sizeof(SgAnTables)) + warp * AN_SCRATCH;
  float* sim = sre + 576;
  const int b = blockIdx.y, M = N - 1;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * frames_per_cta, f1 = min(f0 + frames_per_cta, T_frames);
  for (int t = f0 + warp; t < f1; t += AN_WARPS) {
    float2 X[16];
    float me;
    if (stash != nullptr) {
      const float* sp = stash + ((size_t)b * T_frames + t) * AN_STASH_FLOATS;
      const float4* s4 = reinterpret_cast<const float4*>(sp);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = __ldcs(s4 + r * 32 + lane);
        X[2 * r] = make_float2(v.x, v.y); X[2 * r + 1] = make_float2(v.z, v.w);
      }
      me = __ldcs(sp + 1024 + lane);
    } else {
      an_frame_spectrum(X, xb, M, t, T, sre, sim, lane);
#pragma unroll
      for (int r = 0; r < 16; ++r) sre[lane + 32 * r] = X[r].x * X[r].x + X[r].y * X[r].y;
      if (lane < 4) sre[512 + lane] = 0.f;
      __syncwarp();
      me = an_mel(T, sre, lane);
    }
    const float dF = __ldg(dfeat + ((size_t)b * T_frames + t) * AN_MELS + lane);
    const float dmel = (me > 1e-16f) ? AN_LOGSCALE * dF / me : 0.f;   // clamp passes gradient only above the floor
    __syncwarp();
    sim[lane] = dmel;                                               // dmel[32] (sim is free after the untangle)
    __syncwarp();
    float2 dX[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int k = lane + 32 * r;
      const float dP = T->bin_w0[k] * sim[T->bin_c0[k]] + T->bin_w1[k] * sim[T->bin_c1[k]];
      dX[r] = make_float2(2.f * dP * X[r].x, 2.f * dP * X[r].y);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 16; ++r) { sre[lane + 32 * r] = dX[r].x; sim[lane + 32 * r] = dX[r].y; }
    __syncwarp();
    // untangle adjoint: dZ[k] = ((1-s) + i c)/2 dX[k] + ((1+s) - i c)/2 conj(dX[512-k]), dZ[0] = 0; the
    // FFT input layout z[h][n0] = element 64 n0 + lane + 32 h is bin index lane + 32 (2 n0 + h)
    float2 z[2][8];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int k = lane + 32 * r, km = (512 - k) & 511;
      float2 xm = make_float2(sre[km], -sim[km]);
      float2 cs = T->twU[r][lane];
      float2 ca = make_float2(0.5f * (1.f - cs.y), 0.5f * cs.x), cb = make_float2(0.5f * (1.f + cs.y), -0.5f * cs.x);
      float2 dz = an_cadd(an_cmul(ca, dX[r]), an_cmul(cb, xm));
      if (k == 0) dz = make_float2(0.f, 0.f);
      z[r & 1][r >> 1] = make_float2(dz.x, -dz.y);                  // conj -> FFT -> conj = unnormalised inverse
    }
    __syncwarp();
    warp_fft512(z, T, sre, sim, lane);
    fft512_out_to_smem(z, sre, sim, lane);
    __syncwarp();
    // dg[2n] = Re, dg[2n+1] = -Im(conj);  keep only the 800 windowed samples, times the window
    float* o = dgw + ((size_t)b * T_frames + t) * AN_WIN;
    for (int n = (AN_WOFF >> 1) + lane; n < ((AN_WOFF + AN_WIN) >> 1); n += 32) {
      const int i = 2 * n - AN_WOFF;
      *reinterpret_cast<float2*>(o + i) = make_float2(sre[n] * T->window[i], -sim[n] * T->window[i + 1]);
    }
    __syncwarp();
  }
}

// backward, stage 2: gather overlap-add through the reflect padding and the pre-emphasis -> dx [B,N]
__device__ __forceinline__ float an_dpad(const float* __restrict__ g, int T_frames, int q) {
  // sum over frames t of dgw[t][q - 160 t], q = shifted padded position (frame t covers [160 t, 160 t + 800)): exactly the five
  // frames t_hi - 4 .. t_hi with t_hi = q / 160 (offsets (q % 160) + 160 k < 800), those inside [0, T) - five independent,
  // predicated loads, added in increasing frame order
  const int t_hi = q / AN_HOP, r = q - AN_HOP * t_hi;
  float v[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int t = t_hi - 4 + k;
    v[k] = (t >= 0 && t < T_frames) ? __ldg(g + (size_t)t * AN_WIN + r + AN_HOP * (4 - k)) : 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) s += v[k];
  return s;
}
__device__ __forceinline__ float an_dw(const float* __restrict__ g, int T_frames, int M, int j) {
  // total gradient of the pre-emphasised sample w[j]: direct + left / right reflections
  const int qoff = AN_NFFT / 2 - AN_WOFF;                           // 400
  const int qmax = AN_HOP * (T_frames - 1) + AN_WIN - 1;
  float s = an_dpad(g, T_frames, j + qoff);
  if (j >= 1 && j <= qoff) s += an_dpad(g, T_frames, qoff - j);
  const int jm = 2 * (M - 1) - j;
  if (jm >= M && jm + qoff <= qmax) s += an_dpad(g, T_frames, jm + qoff);
  return s;
}
__global__ void an_overlap_add_kernel(const float* __restrict__ dgw, int N, int T_frames, float* __restrict__ dx,
                                      float scale, int accumulate) {
  const int b = blockIdx.y, M = N - 1, lane = threadIdx.x & 31;
  const float* g = dgw + (size_t)b * T_frames * AN_WIN;
  // x[n] feeds w[n-1] = x[n] - 0.97 x[n-1] and w[n]: dx[n] = dw[n-1] - 0.97 dw[n].  Every thread forms dw[n] once; dw[n-1] is
  // the left neighbour's value (lane 0 computes it itself).  The loop bound is warp-uniform so the shuffle is always full.
  const int nblk = (N + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (int i = 0; i < nblk; ++i) {
    const int n = (i * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const float cur = (n <= M - 1) ? an_dw(g, T_frames, M, n) : 0.f;
    float prev = __shfl_up_sync(0xffffffffu, cur, 1);
    if (lane == 0) prev = (n >= 1 && n - 1 <= M - 1) ? an_dw(g, T_frames, M, n - 1) : 0.f;
    if (n < N) {
      float v = 0.f;
      if (n >= 1) v += prev;                                          // w[n-1] = x[n] - 0.97 x[n-1]
      if (n <= M - 1) v -= 0.97f * cur;
      const size_t gi = (size_t)b * N + n;
      dx[gi] = accumulate ? dx[gi] + scale * v : scale * v;
    }
  }
}

// =============================================================================================
// CNN helpers: MaxPool1d(2,2) over channels-last [B,T,C] and the global max over time
// =============================================================================================
// One thread per pooled position and 4 channels (C is a multiple of 4: 32 / 64 / 128), 32-bit index arithmetic: the scalar
// forms with 64-bit div / mod per element ran at 1.4 TB/s.
__global__ void maxpool2_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int C) {
  const int To = T / 2, C4 = C >> 2;
  const unsigned n = (unsigned)B * To * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4, bt = i / C4, t = bt % To, b = bt / To;
    const float4* p = reinterpret_cast<const float4*>(in + ((size_t)b * T + 2 * t) * C) + c4;
    const float4 u = p[0], v = p[C4];
    reinterpret_cast<float4*>(out)[i] = make_float4(fmaxf(u.x, v.x), fmaxf(u.y, v.y), fmaxf(u.z, v.z), fmaxf(u.w, v.w));
  }
}
// routes the gradient to the arg-max (first element on ties, like torch) and applies the ReLU mask of `in`; a trailing odd
// frame receives no gradient
__global__ void maxpool2_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, float* __restrict__ din,
                                    int B, int T, int C) {
  const int To = T / 2, C4 = C >> 2, Th = (T + 1) / 2;
  const unsigned n = (unsigned)B * Th * C4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned c4 = i % C4, bt = i / C4, t = bt % Th, b = bt / Th;
    float4* q = reinterpret_cast<float4*>(din + ((size_t)b * T + 2 * t) * C) + c4;
    if ((int)t >= To) { q[0] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }     // the odd last frame
    const float4* p = reinterpret_cast<const float4*>(in + ((size_t)b * T + 2 * t) * C) + c4;
    const float4 u = p[0], v = p[C4];
    const float4 g = reinterpret_cast<const float4*>(dout + ((size_t)b * To + t) * C)[c4];
    float4 a, d;
    a.x = (u.x >= v.x && u.x > 0.f) ? g.x : 0.f;  d.x = (!(u.x >= v.x) && v.x > 0.f) ? g.x : 0.f;
    a.y = (u.y >= v.y && u.y > 0.f) ? g.y : 0.f;  d.y = (!(u.y >= v.y) && v.y > 0.f) ? g.y : 0.f;
    a.z = (u.z >= v.z && u.z > 0.f) ? g.z : 0.f;  d.z = (!(u.z >= v.z) && v.z > 0.f) ? g.z : 0.f;
    a.w = (u.w >= v.w && u.w > 0.f) ? g.w : 0.f;  d.w = (!(u.w >= v.w) && v.w > 0.f) ? g.w : 0.f;
    q[0] = a; q[C4] = d;
  }
}
// global max over time (audionet_csine.py:203): out [B,C], arg [B,C]
__global__ void globalmax_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int* __restrict__ arg, int T, int Tv, int C) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  float best = -INFINITY; int bi = 0;
  for (int t = 0; t < Tv; ++t) {
    const float v = in[((size_t)b * T + t) * C + c];
    if (v > best) { best = v; bi = t; }
  }
  out[(size_t)b * C + c] = best;
  arg[(size_t)b * C + c] = bi;
}
__global__ void globalmax_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout, const int* __restrict__ arg,
                                     float* __restrict__ din, int B, int T, int C) {
  const size_t n = (size_t)B * T * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t bt = i / C;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const bool hit = arg[(size_t)b * C + c] == t;
    din[i] = (hit && in[i] > 0.f) ? dout[(size_t)b * C + c] : 0.f;   // ReLU mask of conv8's output
  }
}
// argmax decision over logits (audionet_csine.py:255-256)
__global__ void argmax_rows_kernel(const float* __restrict__ s, long long* __restrict__ dec, int B, int S, float threshold) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int n = lane; n < S; n += 32) { float v = s[(size_t)b * S + n]; if (v > best) { best = v; bi = n; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) dec[b] = (best > threshold) ? (long long)bi : -1LL;
}

// =============================================================================================
// CW2 (attack/CW2.py:41-132): tanh-space iterate, L2 term, Adam on the modifier, best tracking
// =============================================================================================
// input = tanh(w + atanh(0.999999 x));  l2part[b][chunk] = sum (input - x)^2 over the chunk (deterministic 2-stage)
__global__ void cw2_prepare_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ inp,
                                   float* __restrict__ l2part, int N) {
  __shared__ float red[8];
  const int b = blockIdx.y;
  const size_t off = (size_t)b * N;
  float s = 0.f;
  const int per = (N + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per, hi = min(lo + per, N);
  for (int n = lo + threadIdx.x; n < hi; n += blockDim.x) {
    const float xv = x[off + n];
    const float v = tanhf(w[off + n] + atanhf(xv * 0.999999f));
    inp[off + n] = v;
    const float d = v - xv;
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += red[i];
    l2part[(size_t)b * gridDim.x + blockIdx.x] = a;
  }
}
__global__ void cw2_l2_reduce_kernel(const float* __restrict__ l2part, float* __restrict__ loss2, int B, int chunks) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a = 0.f;
  for (int c = 0; c < chunks; ++c) a += l2part[(size_t)b * chunks + c];
  loss2[b] = a;
}
// grad wrt modifier = (c_b * g_model + 2 (input - x)) * (1 - input^2); torch.optim.Adam (no amsgrad, no decay)
__global__ void cw2_adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ x,
                                const float* __restrict__ inp, const float* __restrict__ gmodel, const float* __restrict__ cst,
                                int N, size_t total, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const float in = inp[i];
    const float g = (cst[b] * gmodel[i] + 2.f * (in - x[i])) * (1.f - in * in);
    const float mn = beta1 * m[i] + (1.f - beta1) * g;
    const float vn = beta2 * v[i] + (1.f - beta2) * g * g;
    m[i] = mn; v[i] = vn;
    const float denom = sqrtf(vn) / bc2_sqrt + eps;
    w[i] = w[i] - (lr / bc1) * (mn / denom);
  }
}
// IF-BRANCH-1/2 of attack/CW2.py:102-111, on the device.  Row copy first (uses the old global best), then scalars.
__global__ void cw2_track_copy_kernel(const float* __restrict__ inp, float* __restrict__ best_x, const float* __restrict__ loss1,
                                      const float* __restrict__ loss2, const float* __restrict__ gbest_l2, int N) {
  const int b = blockIdx.y;
  if (!(loss1[b] <= 0.f && loss2[b] < gbest_l2[b])) return;
  const size_t off = (size_t)b * N;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) best_x[off + n] = inp[off + n];
}
__global__ void cw2_track_update_kernel(const float* __restrict__ loss1, const float* __restrict__ loss2,
                                        const long long* __restrict__ dec, float* __restrict__ best_l2,
                                        long long* __restrict__ best_score, float* __restrict__ gbest_l2,
                                        long long* __restrict__ gbest_score, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float l1 = loss1[b], l2 = loss2[b];
  if (l1 <= 0.f && l2 < best_l2[b]) { best_l2[b] = l2; best_score[b] = dec[b]; }
  if (l1 <= 0.f && l2 < gbest_l2[b]) { gbest_l2[b] = l2; gbest_score[b] = dec[b]; }
}
// per-utterance binary search on c (attack/CW2.py:113-123); best_score == -2 means "never succeeded"
__global__ void cw2_search_update_kernel(float* __restrict__ cst, float* __restrict__ lower, float* __restrict__ upper,
                                         const long long* __restrict__ best_score, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (best_score[b] != -2) {
    upper[b] = fminf(upper[b], cst[b]);
    if (upper[b] < 1e9f) cst[b] = (lower[b] + upper[b]) / 2.f;
  } else {
    lower[b] = fmaxf(lower[b], cst[b]);
    if (upper[b] < 1e9f) cst[b] = (lower[b] + upper[b]) / 2.f;
    else cst[b] *= 10.f;
  }
}

// =============================================================================================
// host launchers
// =============================================================================================
static size_t an_smem() { return sizeof(SgAnTables) + AN_WARPS * AN_SCRATCH * sizeof(float); }

int sg_an_init() {
  SG_CUDA_CHECK(cudaFuncSetAttribute(an_logmel_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)an_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(an_logmel_bwd_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)an_smem()));
  return SG_OK;
}
static int an_fpc(int B, int T) {
  int fpc = 64;
  while (fpc > 8 && (long long)B * ((T + fpc - 1) / fpc) < 444) fpc >>= 1;
  return fpc;
}
int sg_an_logmel_fwd_launch(const SgAnTables* dT, const float* x, int B, int N, int T, float* feat, cudaStream_t st, float* stash) {
  const int fpc = an_fpc(B, T);
  an_logmel_fwd_kernel<<<dim3((T + fpc - 1) / fpc, B), AN_THREADS, an_smem(), st>>>(x, N, T, fpc, feat, dT, stash);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_an_logmel_bwd_launch(const SgAnTables* dT, const float* x, int B, int N, int T, const float* dfeat, float* dgw,
                            float* dx, float scale, int accumulate, cudaStream_t st, const float* stash) {
  const int fpc = an_fpc(B, T);
  an_logmel_bwd_frames_kernel<<<dim3((T + fpc - 1) / fpc, B), AN_THREADS, an_smem(), st>>>(x, N, T, fpc, dfeat, dgw, dT, stash);
  SG_LAUNCH_CHECK();
  an_overlap_add_kernel<<<dim3((N + 1023) / 1024, B), 256, 0, st>>>(dgw, N, T, dx, scale, accumulate);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
static int ew_blocks(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }
int sg_maxpool2_fwd_launch(const float* in, float* out, int B, int T, int C, cudaStream_t st) {
  maxpool2_fwd_kernel<<<ew_blocks((size_t)B * (T / 2) * C), 256, 0, st>>>(in, out, B, T, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_maxpool2_bwd_launch(const float* in, const float* dout, float* din, int B, int T, int C, cudaStream_t st) {
  maxpool2_bwd_kernel<<<ew_blocks((size_t)B * T * C), 256, 0, st>>>(in, dout, din, B, T, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_globalmax_fwd_launch(const float* in, float* out, int* arg, int B, int T, int Tv, int C, cudaStream_t st) {
  globalmax_fwd_kernel<<<B, 32, 0, st>>>(in, out, arg, T, Tv, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_globalmax_bwd_launch(const float* in, const float* dout, const int* arg, float* din, int B, int T, int C, cudaStream_t st) {
  globalmax_bwd_kernel<<<ew_blocks((size_t)B * T * C), 256, 0, st>>>(in, dout, arg, din, B, T, C);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_argmax_rows_launch(const float* s, long long* dec, int B, int S, float threshold, cudaStream_t st) {
  argmax_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(s, dec, B, S, threshold);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_cw2_prepare_launch(const float* x, const float* w, float* inp, float* l2part, float* loss2, int B, int N, cudaStream_t st) {
  cw2_prepare_kernel<<<dim3(SG_CW2_CHUNKS, B), 256, 0, st>>>(x, w, inp, l2part, N);
  SG_LAUNCH_CHECK();
  cw2_l2_reduce_kernel<<<(B + 127) / 128, 128, 0, st>>>(l2part, loss2, B, SG_CW2_CHUNKS);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_cw2_adam_launch(float* w, float* m, float* v, const float* x, const float* inp, const float* gmodel, const float* cst,
                       int B, int N, float lr, int step, cudaStream_t st) {
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
  const float bc1 = 1.f - powf(b1, (float)step), bc2s = sqrtf(1.f - powf(b2, (float)step));
  cw2_adam_kernel<<<ew_blocks((size_t)B * N), 256, 0, st>>>(w, m, v, x, inp, gmodel, cst, N, (size_t)B * N, lr, b1, b2, eps, bc1, bc2s);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_cw2_track_launch(const float* inp, float* best_x, const float* loss1, const float* loss2, const long long* dec,
                        float* best_l2, long long* best_score, float* gbest_l2, long long* gbest_score, int B, int N, cudaStream_t st) {
  cw2_track_copy_kernel<<<dim3(8, B), 256, 0, st>>>(inp, best_x, loss1, loss2, gbest_l2, N);
  SG_LAUNCH_CHECK();
  cw2_track_update_kernel<<<(B + 127) / 128, 128, 0, st>>>(loss1, loss2, dec, best_l2, best_score, gbest_l2, gbest_score, B);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_cw2_search_update_launch(float* cst, float* lower, float* upper, const long long* best_score, int B, cudaStream_t st) {
  cw2_search_update_kernel<<<(B + 127) / 128, 128, 0, st>>>(cst, lower, upper, best_score, B);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
