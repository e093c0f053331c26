// AudioNet in training mode: what `outputs = model(x_batch); loss.backward()` does inside
// adver_train.py:183-221 / natural_train.py:127-160 when the module is in train() mode - BatchNorm with batch statistics
// (and the momentum update of the running statistics), the input gradient through those statistics (the PGD attack inside
// the training loop runs on the train-mode model), and the gradients of every parameter.
//
// Layout as in the inference path (sg_api_audionet.cu): channels-last [B*T_l, C_l] activations, conv = 3 (5 for the banded
// pre-filter) row-shifted GEMMs with 'same' padding inside each utterance.  BatchNorm cannot be folded any more, so every
// conv stage is: conv + bias -> z, two-pass batch statistics over the valid rows, normalise (+ ReLU) -> r, optional pool.
// The backward pass per stage: ReLU/pool adjoint, BN adjoint (two reductions + one apply), bias / weight gradients (a
// split-K GEMM over the B*T rows), dgrad with the flipped taps.
#include <math.h>
#include <string.h>

#include "sg_handle.cuh"
#include "sg_iv.cuh"

#define TR_ROWS_PER_CHUNK 256     // rows per partial sum of the BN / bias reductions
#define TR_WG_ROWS 2048           // rows per CTA of the weight-gradient GEMM
#define TR_MAXC 128

// ---------------------------------------------------------------------------------------------
// parameter packing (PyTorch layouts -> the GEMM layouts of the conv kernel), run every forward
// ---------------------------------------------------------------------------------------------
__global__ void tr_pack_conv_kernel(const float* __restrict__ w, float* __restrict__ W, float* __restrict__ Wb, int ci, int co) {
  const int n = co * ci * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = i % 3, c = (i / 3) % ci, o = i / (3 * ci);
    const float v = w[i];
    W[((size_t)k * ci + c) * co + o] = v;
    Wb[((size_t)k * co + o) * ci + c] = v;
  }
}
// Conv2d(1,1,5x5,pad 2) over [F=32, T] as 5 time taps of a banded 32x32 matrix; bias scalar -> vector
__global__ void tr_pack_band_kernel(const float* __restrict__ w25, const float* __restrict__ b, float* __restrict__ W1,
                                    float* __restrict__ W1b, float* __restrict__ b1v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 5 * 32 * 32; i += gridDim.x * blockDim.x) {
    const int fo = i % 32, fi = (i / 32) % 32, j = i / 1024;
    const int ii = fi - fo + 2;
    const float v = (ii >= 0 && ii < 5) ? w25[ii * 5 + j] : 0.f;
    W1[((size_t)j * 32 + fi) * 32 + fo] = v;
    W1b[((size_t)j * 32 + fo) * 32 + fi] = v;
    if (i < 32) b1v[i] = b[0];
  }
}
__global__ void tr_pack_fc_kernel(const float* __restrict__ fw, const float* __restrict__ fb, float* __restrict__ Wfc,
                                  float* __restrict__ Wfcb, float* __restrict__ bfc, int C, int Cp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Cp * 32; i += gridDim.x * blockDim.x) {
    const int c = i % 32, o = i / 32;
    const float v = o < C ? fw[(size_t)o * 32 + c] : 0.f;
    Wfc[(size_t)c * Cp + o] = v;
    Wfcb[(size_t)o * 32 + c] = v;
    if (c == 0) bfc[o] = o < C ? fb[o] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// column reductions over the valid rows (t < t_valid of every utterance), two-stage and deterministic
//   MODE 0: sum z           MODE 1: sum (z - mean)^2        MODE 2: sum dy and sum dy * zhat (two outputs)
// grid (nchunk, C/32), block (32, 8)
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void tr_col_partial_kernel(const float* __restrict__ a, const float* __restrict__ z, int C, int rows, int T, int t_valid,
                                      const float* __restrict__ mean_col, const float* __restrict__ rstd_col,
                                      float* __restrict__ part1, float* __restrict__ part2) {
  __shared__ float s1[8][33], s2[8][33];
  const int col = blockIdx.y * 32 + threadIdx.x;
  const int r0 = blockIdx.x * TR_ROWS_PER_CHUNK;
  float acc1 = 0.f, acc2 = 0.f;
  if (col < C) {
    const float mu = (MODE >= 1) ? mean_col[col] : 0.f;
    const float rs = (MODE == 2) ? rstd_col[col] : 0.f;
    for (int r = r0 + threadIdx.y; r < min(r0 + TR_ROWS_PER_CHUNK, rows); r += 8) {
      if (r % T >= t_valid) continue;
      const float v = a[(size_t)r * C + col];
      if (MODE == 0) acc1 += v;
      else if (MODE == 1) { const float d = v - mu; acc1 = fmaf(d, d, acc1); }
      else { acc1 += v; acc2 = fmaf(v, (z[(size_t)r * C + col] - mu) * rs, acc2); }
    }
  }
  s1[threadIdx.y][threadIdx.x] = acc1; s2[threadIdx.y][threadIdx.x] = acc2;
  __syncthreads();
  if (threadIdx.y == 0 && col < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t1 += s1[i][threadIdx.x]; t2 += s2[i][threadIdx.x]; }
    part1[(size_t)blockIdx.x * C + col] = t1;
    if (MODE == 2) part2[(size_t)blockIdx.x * C + col] = t2;
  }
}

// one thread per statistics group (gsz adjacent columns share one BatchNorm channel: 32 for the BatchNorm2d(1) pre-filter)
// stage 0: mean      stage 1: variance -> rstd, running statistics      stage 2: BN-adjoint sums -> dgamma, dbeta, m1, m2
// stage 3: plain column sums (bias gradients)
__global__ void tr_col_finalize_kernel(int stage, const float* __restrict__ part1, const float* __restrict__ part2, int nchunk, int C,
                                       int gsz, double count, float eps, float momentum, float* mean_col, float* rstd_col,
                                       float* run_mean, float* run_var, const float* __restrict__ gamma, float* dgamma, float* dbeta,
                                       float* m1_col, float* m2_col, float* out_sum) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= C / gsz) return;
  double t1 = 0.0, t2 = 0.0;
  for (int c = g * gsz; c < (g + 1) * gsz; ++c)
    for (int k = 0; k < nchunk; ++k) {
      t1 += (double)part1[(size_t)k * C + c];
      if (stage == 2) t2 += (double)part2[(size_t)k * C + c];
    }
  if (stage == 0) {
    const float mu = (float)(t1 / count);
    for (int c = g * gsz; c < (g + 1) * gsz; ++c) mean_col[c] = mu;
  } else if (stage == 1) {
    const double var_b = t1 / count;
    const float rs = (float)(1.0 / sqrt(var_b + (double)eps));
    for (int c = g * gsz; c < (g + 1) * gsz; ++c) rstd_col[c] = rs;
    if (momentum > 0.f && run_mean && run_var) {        // torch.nn.BatchNorm: unbiased variance in the running estimate
      run_mean[g] = (1.f - momentum) * run_mean[g] + momentum * mean_col[g * gsz];
      run_var[g] = (1.f - momentum) * run_var[g] + momentum * (float)(var_b * count / (count - 1.0));
    }
  } else if (stage == 2) {
    if (dbeta) dbeta[g] = (float)t1;
    if (dgamma) dgamma[g] = (float)t2;
    for (int c = g * gsz; c < (g + 1) * gsz; ++c) { m1_col[c] = (float)(t1 / count); m2_col[c] = (float)(t2 / count); }
  } else {
    out_sum[g] = (float)t1;
  }
}

// r = gamma * (z - mean) * rstd + beta, optionally ReLU
__global__ void tr_bn_apply_kernel(const float* __restrict__ z, float* __restrict__ r, size_t n, int C, int gsz,
                                   const float* __restrict__ mean_col, const float* __restrict__ rstd_col,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), g = c / gsz;
    float y = fmaf((z[i] - mean_col[c]) * rstd_col[c], gamma[g], beta[g]);
    r[i] = relu ? fmaxf(y, 0.f) : y;
  }
}
// dz = gamma * rstd * (dy - mean(dy) - zhat * mean(dy * zhat)) on the valid rows, 0 elsewhere (in place on dy)
__global__ void tr_bn_bwd_apply_kernel(float* __restrict__ dy, const float* __restrict__ z, size_t n, int C, int gsz, int T, int t_valid,
                                       const float* __restrict__ mean_col, const float* __restrict__ rstd_col,
                                       const float* __restrict__ gamma, const float* __restrict__ m1_col, const float* __restrict__ m2_col) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t row = i / C;
    if ((int)(row % T) >= t_valid) { dy[i] = 0.f; continue; }
    const float zh = (z[i] - mean_col[c]) * rstd_col[c];
    dy[i] = gamma[c / gsz] * rstd_col[c] * (dy[i] - m1_col[c] - zh * m2_col[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dWp[(k*ci + c), o] = sum_p in[p + tap_base + k, c] * dz[p, o]   ('same' padding inside each utterance)
// 64 x 64 output tile per CTA over TR_WG_ROWS rows; partial sums per row chunk, reduced by splitk_reduce
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tr_wgrad_kernel(const float* __restrict__ in, int ci, const float* __restrict__ dz, int co, int rows, int T, int taps, int tap_base,
                float* __restrict__ part) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int M = taps * ci;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64, p_begin = blockIdx.z * TR_WG_ROWS;
  const int p_end = min(p_begin + TR_WG_ROWS, rows);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int p0 = p_begin; p0 < p_end; p0 += 16) {
#pragma unroll
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int kk = e >> 6, mm = e & 63;
      const int p = p0 + kk, m = m0 + mm;
      float av = 0.f, bv = 0.f;
      if (p < p_end) {
        if (m < M) {
          const int k = m / ci, c = m - k * ci;
          const int off = tap_base + k, t = p % T;
          if (t + off >= 0 && t + off < T) av = in[(size_t)(p + off) * ci + c];
        }
        if (n0 + mm < co) bv = dz[(size_t)p * co + n0 + mm];
      }
      As[kk][mm] = av;
      Bs[kk][mm] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* o = part + (size_t)blockIdx.z * M * co;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < co) o[(size_t)m * co + n] = acc[i][j];
    }
}
// packed [3*ci, co] -> PyTorch [co, ci, 3]
__global__ void tr_unpack_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dw, int ci, int co) {
  const int n = co * ci * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int k = i % 3, c = (i / 3) % ci, o = i / (3 * ci);
    dw[i] = dWp[((size_t)k * ci + c) * co + o];
  }
}
// banded [5*32, 32] -> the 25 weights of the 5x5 kernel
__global__ void tr_unband_wgrad_kernel(const float* __restrict__ dWp, float* __restrict__ dw25) {
  const int i = threadIdx.x;        // ii*5 + j
  if (i >= 25) return;
  const int ii = i / 5, j = i % 5;
  double acc = 0.0;
  for (int fo = 0; fo < 32; ++fo) {
    const int fi = fo + ii - 2;
    if (fi >= 0 && fi < 32) acc += (double)dWp[((size_t)j * 32 + fi) * 32 + fo];
  }
  dw25[i] = (float)acc;
}
// fc: dW[o, c] = sum_b dlogits[b, o] emb[b, c], db[o] = sum_b dlogits[b, o]; one warp per class
__global__ void tr_fc_wgrad_kernel(const float* __restrict__ dlogits, int ld, const float* __restrict__ emb, int B, int C,
                                   float* __restrict__ dW, float* __restrict__ db) {
  const int o = blockIdx.x, c = threadIdx.x;
  if (o >= C) return;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dlogits[(size_t)b * ld + o];
    acc = fmaf(d, emb[(size_t)b * 32 + c], acc);
    accb += d;
  }
  dW[(size_t)o * 32 + c] = acc;
  if (c == 0) db[o] = accb;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct TrWs {
  int T[8];                       // time length at the input of conv stage l (l = 0..6); [7]: valid outputs of conv8
  float *z[8], *r[8], *p[7];      // index 0: pre-filter; 1..7: conv2..conv8 (z pre-BN, r post-BN/ReLU, p pooled)
  float *mean[8], *rstd[8];
  float *W1, *W1b, *b1v, *W[7], *Wb[7], *Wfc, *Wfcb, *bfc;
  float *part1, *part2, *m1, *m2, *wpart, *wred, *g0, *g1, *emb, *demb;
  int* arg;
  int nchunk_max, wchunk_max;
  size_t bytes;
};
static TrWs tr_ws_layout(void* base, int B, int N, int Cp) {
  TrWs w;
  memset(&w, 0, sizeof(w));
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloat) { float* q = (float*)(p + off); off += (nfloat * sizeof(float) + 255) / 256 * 256; return q; };
  const int T0 = 1 + (N - 1) / AN_HOP;
  int t = T0;
  for (int l = 0; l < 7; ++l) { w.T[l] = t; if (kAnPool[l]) t = t / 2; }
  w.T[7] = w.T[6] - 2;
  const size_t R0 = (size_t)B * T0;
  w.z[0] = take(R0 * 32); w.r[0] = take(R0 * 32);
  size_t gmax = R0 * 32;
  for (int l = 0; l < 7; ++l) {
    const size_t n = (size_t)B * w.T[l] * kAnCout[l];
    w.z[l + 1] = take(n); w.r[l + 1] = take(n);
    w.p[l] = kAnPool[l] ? take((size_t)B * (w.T[l] / 2) * kAnCout[l]) : w.r[l + 1];
    if (n > gmax) gmax = n;
    const size_t nin = (size_t)B * w.T[l] * kAnCin[l];
    if (nin > gmax) gmax = nin;
  }
  for (int l = 0; l < 8; ++l) { w.mean[l] = take(TR_MAXC); w.rstd[l] = take(TR_MAXC); }
  w.W1 = take(5 * 32 * 32); w.W1b = take(5 * 32 * 32); w.b1v = take(32);
  for (int l = 0; l < 7; ++l) { w.W[l] = take((size_t)3 * kAnCin[l] * kAnCout[l]); w.Wb[l] = take((size_t)3 * kAnCin[l] * kAnCout[l]); }
  w.Wfc = take((size_t)32 * Cp); w.Wfcb = take((size_t)32 * Cp); w.bfc = take(Cp);
  w.nchunk_max = (int)((R0 + TR_ROWS_PER_CHUNK - 1) / TR_ROWS_PER_CHUNK);
  w.wchunk_max = (int)((R0 + TR_WG_ROWS - 1) / TR_WG_ROWS);
  w.part1 = take((size_t)w.nchunk_max * TR_MAXC); w.part2 = take((size_t)w.nchunk_max * TR_MAXC);
  w.m1 = take(TR_MAXC); w.m2 = take(TR_MAXC);
  w.wpart = take((size_t)w.wchunk_max * 384 * 128); w.wred = take((size_t)384 * 128);
  w.g0 = take(gmax); w.g1 = take(gmax);
  w.emb = take((size_t)B * 32); w.demb = take((size_t)B * 32); w.arg = (int*)take((size_t)B * 32);
  w.bytes = off;
  return w;
}

extern "C" size_t sg_audionet_train_ws_bytes(const sg_handle* h, int B, int N) {
  if (!h || !h->an || B < 1 || N < AN_NFFT) return 0;
  return tr_ws_layout(nullptr, B, N, h->an->Cp).bytes;
}

static int tr_check(sg_handle* h, const sg_audionet_train_tensors* p, int B, int N) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->an) { sg_set_error("AudioNet not loaded (call sg_load_audionet first: the front-end tables and class count come from it)"); return SG_ESTATE; }
  if (!p || B < 2 || N < 2 * AN_NFFT) { sg_set_error("AudioNet training needs parameters, B >= 2 and N >= %d samples (B=%d N=%d)", 2 * AN_NFFT, B, N); return SG_EINVAL; }
  if (!p->conv1_w || !p->conv1_b || !p->fc_w || !p->fc_b) { sg_set_error("sg_audionet_train: null parameter pointer"); return SG_EINVAL; }
  for (int l = 0; l < 7; ++l) if (!p->conv_w[l] || !p->conv_b[l]) { sg_set_error("sg_audionet_train: null conv parameter (stage %d)", l + 2); return SG_EINVAL; }
  for (int l = 0; l < 8; ++l) if (!p->bn_gamma[l] || !p->bn_beta[l]) { sg_set_error("sg_audionet_train: null BatchNorm parameter (%d)", l + 1); return SG_EINVAL; }
  return SG_OK;
}

#define TR_K(call) do { h->launches += 1; PROF(h, SG_PROF_AUDIONET, st, (call)); } while (0)
static int launch_ok() { SG_LAUNCH_CHECK(); return SG_OK; }
static int ew_blocks(size_t n) { size_t b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

static void tr_conv_args(SgConvArgs& a, const float* A, int cin, const float* W, const float* bias, float* out, int cout, int rows,
                         int taps, int tap_base, int tap_step, int T, int epi, const float* mask, int ldmask) {
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = cin; a.W = W; a.bias = bias; a.out = out; a.ldo = cout; a.rows = rows; a.N = cout; a.cin = cin;
  a.taps = taps; a.tap_base = tap_base; a.tap_step = tap_step; a.same_utt = 1; a.T = T; a.t_valid = T; a.epilogue = epi;
  a.mask = mask; a.ldmask = ldmask;
}

// batch statistics of z over the valid rows + running-statistics update + normalise (+ ReLU)
static int tr_bn_forward(sg_handle* h, const TrWs& w, int li, const float* z, float* r, int rows, int C, int gsz, int T, int t_valid,
                         int B, const float* gamma, const float* beta, float* run_mean, float* run_var, float momentum, float eps,
                         int relu, cudaStream_t st) {
  const int nchunk = (rows + TR_ROWS_PER_CHUNK - 1) / TR_ROWS_PER_CHUNK;
  const dim3 grid(nchunk, (C + 31) / 32), block(32, 8);
  const double count = (double)B * t_valid * gsz;
  const int ngrp = C / gsz;
  tr_col_partial_kernel<0><<<grid, block, 0, st>>>(z, nullptr, C, rows, T, t_valid, nullptr, nullptr, w.part1, nullptr);
  TR_K(launch_ok());
  tr_col_finalize_kernel<<<(ngrp + 127) / 128, 128, 0, st>>>(0, w.part1, nullptr, nchunk, C, gsz, count, eps, momentum, w.mean[li], w.rstd[li],
                                                           nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  TR_K(launch_ok());
  tr_col_partial_kernel<1><<<grid, block, 0, st>>>(z, nullptr, C, rows, T, t_valid, w.mean[li], nullptr, w.part1, nullptr);
  TR_K(launch_ok());
  tr_col_finalize_kernel<<<(ngrp + 127) / 128, 128, 0, st>>>(1, w.part1, nullptr, nchunk, C, gsz, count, eps, momentum, w.mean[li], w.rstd[li],
                                                           run_mean, run_var, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  TR_K(launch_ok());
  tr_bn_apply_kernel<<<ew_blocks((size_t)rows * C), 256, 0, st>>>(z, r, (size_t)rows * C, C, gsz, w.mean[li], w.rstd[li], gamma, beta, relu);
  TR_K(launch_ok());
  return SG_OK;
}

extern "C" int sg_audionet_train_fwd(sg_handle* h, const sg_audionet_train_tensors* p, const float* feat, int B, int N, float momentum,
                                     float eps, void* ws, float* logits, sg_stream stream) {
  SG_TRY(tr_check(h, p, B, N));
  if (!feat || !ws || !logits) { sg_set_error("sg_audionet_train_fwd: null pointer"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  SgAudioNet* an = h->an;
  const TrWs w = tr_ws_layout(ws, B, N, an->Cp);
  if (w.T[7] < 1) { sg_set_error("utterance too short for AudioNet's conv8 (N=%d)", N); return SG_EINVAL; }
  if (eps <= 0.f) eps = 1e-5f;
  // current parameters -> GEMM layouts
  tr_pack_band_kernel<<<20, 256, 0, st>>>(p->conv1_w, p->conv1_b, w.W1, w.W1b, w.b1v);
  TR_K(launch_ok());
  for (int l = 0; l < 7; ++l) {
    tr_pack_conv_kernel<<<ew_blocks((size_t)3 * kAnCin[l] * kAnCout[l]), 256, 0, st>>>(p->conv_w[l], w.W[l], w.Wb[l], kAnCin[l], kAnCout[l]);
    TR_K(launch_ok());
  }
  tr_pack_fc_kernel<<<ew_blocks((size_t)an->Cp * 32), 256, 0, st>>>(p->fc_w, p->fc_b, w.Wfc, w.Wfcb, w.bfc, an->C, an->Cp);
  TR_K(launch_ok());

  SgConvArgs a;
  tr_conv_args(a, feat, 32, w.W1, w.b1v, w.z[0], 32, B * w.T[0], 5, -2, 1, w.T[0], SG_EPI_BIAS, nullptr, 0);
  SG_TRY(sg_run_conv(h, a, false, SG_PROF_AUDIONET, st));
  SG_TRY(tr_bn_forward(h, w, 0, w.z[0], w.r[0], B * w.T[0], 32, 32, w.T[0], w.T[0], B, p->bn_gamma[0], p->bn_beta[0], p->bn_mean[0],
                       p->bn_var[0], momentum, eps, 0, st));
  const float* in = w.r[0];
  for (int l = 0; l < 7; ++l) {
    const int rows = B * w.T[l], co = kAnCout[l];
    tr_conv_args(a, in, kAnCin[l], w.W[l], p->conv_b[l], w.z[l + 1], co, rows, 3, -kAnPad[l], 1, w.T[l], SG_EPI_BIAS, nullptr, 0);
    SG_TRY(sg_run_conv(h, a, false, SG_PROF_AUDIONET, st));
    const int tv = (l == 6) ? w.T[7] : w.T[l];
    SG_TRY(tr_bn_forward(h, w, l + 1, w.z[l + 1], w.r[l + 1], rows, co, 1, w.T[l], tv, B, p->bn_gamma[l + 1], p->bn_beta[l + 1],
                         p->bn_mean[l + 1], p->bn_var[l + 1], momentum, eps, 1, st));
    if (kAnPool[l]) TR_K(sg_maxpool2_fwd_launch(w.r[l + 1], w.p[l], B, w.T[l], co, st));
    in = w.p[l];
  }
  TR_K(sg_globalmax_fwd_launch(w.r[7], w.emb, w.arg, B, w.T[6], w.T[7], 32, st));
  memset(&a, 0, sizeof(a));
  a.A = w.emb; a.lda = 32; a.W = w.Wfc; a.bias = w.bfc; a.out = logits; a.ldo = an->Cp; a.rows = B; a.N = an->Cp; a.cin = 32;
  a.taps = 1; a.epilogue = SG_EPI_BIAS; a.T = 1;
  return sg_run_conv(h, a, false, SG_PROF_AUDIONET, st);
}

// BN adjoint of one stage, in place on dy; optionally the BatchNorm parameter gradients and the conv-bias gradient
static int tr_bn_backward(sg_handle* h, const TrWs& w, int li, float* dy, const float* z, int rows, int C, int gsz, int T, int t_valid,
                          int B, const float* gamma, float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  const int nchunk = (rows + TR_ROWS_PER_CHUNK - 1) / TR_ROWS_PER_CHUNK;
  const dim3 grid(nchunk, (C + 31) / 32), block(32, 8);
  const double count = (double)B * t_valid * gsz;
  const int ngrp = C / gsz;
  tr_col_partial_kernel<2><<<grid, block, 0, st>>>(dy, z, C, rows, T, t_valid, w.mean[li], w.rstd[li], w.part1, w.part2);
  TR_K(launch_ok());
  tr_col_finalize_kernel<<<(ngrp + 127) / 128, 128, 0, st>>>(2, w.part1, w.part2, nchunk, C, gsz, count, 0.f, 0.f, w.mean[li], w.rstd[li],
                                                           nullptr, nullptr, gamma, dgamma, dbeta, w.m1, w.m2, nullptr);
  TR_K(launch_ok());
  tr_bn_bwd_apply_kernel<<<ew_blocks((size_t)rows * C), 256, 0, st>>>(dy, z, (size_t)rows * C, C, gsz, T, t_valid, w.mean[li], w.rstd[li],
                                                                      gamma, w.m1, w.m2);
  TR_K(launch_ok());
  if (dbias) {     // analytically zero in front of a train-mode BatchNorm; computed like autograd does
    tr_col_partial_kernel<0><<<grid, block, 0, st>>>(dy, nullptr, C, rows, T, T, nullptr, nullptr, w.part1, nullptr);
    TR_K(launch_ok());
    tr_col_finalize_kernel<<<(ngrp + 127) / 128, 128, 0, st>>>(3, w.part1, nullptr, nchunk, C, gsz, count, 0.f, 0.f, nullptr, nullptr, nullptr,
                                                             nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dbias);
    TR_K(launch_ok());
  }
  return SG_OK;
}

static int tr_wgrad(sg_handle* h, const TrWs& w, const float* in, int ci, const float* dz, int co, int rows, int T, int taps, int tap_base,
                    cudaStream_t st) {
  const int M = taps * ci, nchunk = (rows + TR_WG_ROWS - 1) / TR_WG_ROWS;
  tr_wgrad_kernel<<<dim3((M + 63) / 64, (co + 63) / 64, nchunk), 256, 0, st>>>(in, ci, dz, co, rows, T, taps, tap_base, w.wpart);
  TR_K(launch_ok());
  TR_K(sg_splitk_reduce_launch(w.wpart, nchunk, M, co, w.wred, co, st));
  return SG_OK;
}

extern "C" int sg_audionet_train_bwd(sg_handle* h, const sg_audionet_train_tensors* p, const float* feat, const float* dlogits, int B,
                                     int N, void* ws, float* dfeat, const sg_audionet_train_tensors* g, sg_stream stream) {
  SG_TRY(tr_check(h, p, B, N));
  if (!feat || !dlogits || !ws || (!dfeat && !g)) { sg_set_error("sg_audionet_train_bwd: null pointer"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  SgAudioNet* an = h->an;
  const TrWs w = tr_ws_layout(ws, B, N, an->Cp);
  SgConvArgs a;
  if (g) {
    tr_fc_wgrad_kernel<<<an->C, 32, 0, st>>>(dlogits, an->Cp, w.emb, B, an->C, g->fc_w, g->fc_b);
    TR_K(launch_ok());
  }
  memset(&a, 0, sizeof(a));
  a.A = dlogits; a.lda = an->Cp; a.W = w.Wfcb; a.out = w.demb; a.ldo = 32; a.rows = B; a.N = 32; a.cin = an->Cp;
  a.taps = 1; a.epilogue = SG_EPI_NONE; a.T = 1;
  SG_TRY(sg_run_conv(h, a, false, SG_PROF_AUDIONET, st));
  TR_K(sg_globalmax_bwd_launch(w.r[7], w.demb, w.arg, w.g0, B, w.T[6], 32, st));        // d(BN output of conv8), ReLU-masked
  float* gin = w.g0;
  float* gout = w.g1;
  for (int l = 6; l >= 0; --l) {
    const int rows = B * w.T[l], ci = kAnCin[l], co = kAnCout[l];
    const int tv = (l == 6) ? w.T[7] : w.T[l];
    SG_TRY(tr_bn_backward(h, w, l + 1, gin, w.z[l + 1], rows, co, 1, w.T[l], tv, B, p->bn_gamma[l + 1], g ? g->bn_gamma[l + 1] : nullptr,
                          g ? g->bn_beta[l + 1] : nullptr, g ? g->conv_b[l] : nullptr, st));
    const float* lin = l > 0 ? w.p[l - 1] : w.r[0];
    if (g) {
      SG_TRY(tr_wgrad(h, w, lin, ci, gin, co, rows, w.T[l], 3, -kAnPad[l], st));
      tr_unpack_wgrad_kernel<<<ew_blocks((size_t)3 * ci * co), 256, 0, st>>>(w.wred, g->conv_w[l], ci, co);
      TR_K(launch_ok());
    }
    const bool prev_pooled = l > 0 && kAnPool[l - 1];
    const bool mask_here = l > 0 && !prev_pooled;
    tr_conv_args(a, gin, co, w.Wb[l], nullptr, gout, ci, rows, 3, kAnPad[l], -1, w.T[l], mask_here ? SG_EPI_MASK : SG_EPI_NONE,
                 mask_here ? w.r[l] : nullptr, mask_here ? kAnCout[l - 1] : 0);
    SG_TRY(sg_run_conv(h, a, false, SG_PROF_AUDIONET, st));
    float* t = gin; gin = gout; gout = t;
    if (prev_pooled) {
      TR_K(sg_maxpool2_bwd_launch(w.r[l], gin, gout, B, w.T[l - 1], kAnCout[l - 1], st));
      t = gin; gin = gout; gout = t;
    }
  }
  // pre-filter: BatchNorm2d(1) over all 32 mel columns, no ReLU
  const int rows0 = B * w.T[0];
  SG_TRY(tr_bn_backward(h, w, 0, gin, w.z[0], rows0, 32, 32, w.T[0], w.T[0], B, p->bn_gamma[0], g ? g->bn_gamma[0] : nullptr,
                        g ? g->bn_beta[0] : nullptr, g ? g->conv1_b : nullptr, st));
  if (g) {
    SG_TRY(tr_wgrad(h, w, feat, 32, gin, 32, rows0, w.T[0], 5, -2, st));
    tr_unband_wgrad_kernel<<<1, 32, 0, st>>>(w.wred, g->conv1_w);
    TR_K(launch_ok());
  }
  if (dfeat) {
    tr_conv_args(a, gin, 32, w.W1b, nullptr, dfeat, 32, rows0, 5, 2, -1, w.T[0], SG_EPI_NONE, nullptr, 0);
    SG_TRY(sg_run_conv(h, a, false, SG_PROF_AUDIONET, st));
  }
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults' update rule, adver_train.py:119: no amsgrad, L2 weight decay folded into the gradient) on one
// flat tensor: p -= lr * m_hat / (sqrt(v_hat) + eps)
// ---------------------------------------------------------------------------------------------
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                 size_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = grad[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, p[i], gi);
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);           // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}
extern "C" int sg_adam_step(sg_handle* h, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int step, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) { sg_set_error("sg_adam_step: bad argument"); return SG_EINVAL; }
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  h->launches += 1;
  adam_step_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                  bc1, sqrtf(bc2));
  SG_LAUNCH_CHECK();
  return SG_OK;
}
