// C-ABI for the AudioNet path (reference model/audionet_csine.py:133-257) and CW2
// (attack/CW2.py:41-132): weight folding/packing, workspace layout, forward / backward
// orchestration over the conv-as-GEMM kernel ('same' padding mode) and the whole CW2 loop.
#include <math.h>
#include <string.h>

#include "sg_handle.cuh"
#include "sg_head.cuh"

void sg_audionet_free(sg_handle* h) {
  if (h && h->an) { delete h->an; h->an = nullptr; }            // device buffers are in h->allocs
}

static int an_up(sg_handle* h, float** dst, const std::vector<float>& v) { return sg_dev_upload(h, dst, v); }
// K-major copy of a packed [K, N] matrix: [N, K]
static int an_up_t(sg_handle* h, float** dst, const std::vector<float>& v, size_t K, size_t N) {
  std::vector<float> t(v.size());
  for (size_t k = 0; k < K; ++k)
    for (size_t n = 0; n < N; ++n) t[n * K + k] = v[k * N + n];
  return sg_dev_upload(h, dst, t);
}

extern "C" int sg_load_audionet(sg_handle* h, const sg_audionet_weights* w) {
  SG_TRY(sg_check_handle(h, false));
  if (!w || w->num_class < 1) { sg_set_error("sg_load_audionet: bad argument"); return SG_EINVAL; }
  if (h->an) { sg_set_error("sg_load_audionet: already loaded"); return SG_ESTATE; }
  SG_CUDA_CHECK(cudaSetDevice(h->device));
  SgAudioNet* an = new SgAudioNet();
  h->an = an;
  {
    SgAnTables* host = new SgAnTables();
    int r = sg_an_tables_build(host);
    if (r != SG_OK) { delete host; sg_set_error("AudioNet table construction failed"); return r; }
    cudaError_t e = cudaMalloc((void**)&an->d_tables, sizeof(SgAnTables));
    if (e == cudaSuccess) e = cudaMemcpy(an->d_tables, host, sizeof(SgAnTables), cudaMemcpyHostToDevice);
    delete host;
    if (e != cudaSuccess) { sg_set_error("AudioNet table upload failed: %s", cudaGetErrorString(e)); return SG_ECUDA; }
    h->allocs.push_back(an->d_tables);
    SG_TRY(sg_an_init());
  }
  const double eps = w->bn_eps > 0.f ? w->bn_eps : 1e-5;
  {  // conv1: Conv2d(1,1,5x5,pad 2) + BatchNorm2d(1) over [F=32, T]  ->  5 time taps of a banded 32x32 matrix
    const double sc = w->bn_gamma[0][0] / sqrt((double)w->bn_var[0][0] + eps);
    std::vector<float> W1((size_t)5 * 32 * 32, 0.f), W1b((size_t)5 * 32 * 32, 0.f), b1(32);
    for (int j = 0; j < 5; ++j)
      for (int fi = 0; fi < 32; ++fi)
        for (int fo = 0; fo < 32; ++fo) {
          const int i = fi - fo + 2;
          if (i < 0 || i >= 5) continue;
          const float v = (float)(w->conv1_w[i * 5 + j] * sc);
          W1[((size_t)j * 32 + fi) * 32 + fo] = v;
          W1b[((size_t)j * 32 + fo) * 32 + fi] = v;
        }
    for (int fo = 0; fo < 32; ++fo) b1[fo] = (float)(((double)w->conv1_b[0] - w->bn_mean[0][0]) * sc + w->bn_beta[0][0]);
    SG_TRY(an_up(h, &an->W1, W1)); SG_TRY(an_up(h, &an->W1b, W1b)); SG_TRY(an_up(h, &an->b1, b1));
    SG_TRY(an_up_t(h, &an->W1k, W1, 5 * 32, 32)); SG_TRY(an_up_t(h, &an->W1bk, W1b, 5 * 32, 32));
  }
  for (int l = 0; l < 7; ++l) {   // Conv1d k=3 + BatchNorm1d (affine) folded: BN directly follows the conv
    const int ci = kAnCin[l], co = kAnCout[l];
    std::vector<float> W((size_t)3 * ci * co), Wb((size_t)3 * co * ci), b(co);
    for (int o = 0; o < co; ++o) {
      const double sc = w->bn_gamma[l + 1][o] / sqrt((double)w->bn_var[l + 1][o] + eps);
      for (int c = 0; c < ci; ++c)
        for (int k = 0; k < 3; ++k) {
          const float v = (float)(w->conv_w[l][((size_t)o * ci + c) * 3 + k] * sc);
          W[((size_t)k * ci + c) * co + o] = v;
          Wb[((size_t)k * co + o) * ci + c] = v;
        }
      b[o] = (float)(((double)w->conv_b[l][o] - w->bn_mean[l + 1][o]) * sc + w->bn_beta[l + 1][o]);
    }
    SG_TRY(an_up(h, &an->W[l], W)); SG_TRY(an_up(h, &an->Wb[l], Wb)); SG_TRY(an_up(h, &an->bias[l], b));
    SG_TRY(an_up_t(h, &an->Wk[l], W, (size_t)3 * ci, co)); SG_TRY(an_up_t(h, &an->Wbk[l], Wb, (size_t)3 * co, ci));
  }
  {
    const int C = w->num_class, Cp = (C + 15) / 16 * 16;
    an->C = C; an->Cp = Cp;
    std::vector<float> Wfc((size_t)32 * Cp, 0.f), Wfcb((size_t)Cp * 32, 0.f), b(Cp, 0.f);
    for (int o = 0; o < C; ++o) {
      for (int c = 0; c < 32; ++c) { Wfc[(size_t)c * Cp + o] = w->fc_w[(size_t)o * 32 + c]; Wfcb[(size_t)o * 32 + c] = w->fc_w[(size_t)o * 32 + c]; }
      b[o] = w->fc_b[o];
    }
    SG_TRY(an_up(h, &an->Wfc, Wfc)); SG_TRY(an_up(h, &an->Wfcb, Wfcb)); SG_TRY(an_up(h, &an->bfc, b));
  }
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct AnWs {
  int T[8];                 // time length at the input of conv stage l (l = 0..6) and of the global max (7)
  float *c1, *a[7], *p[7], *emb, *logits, *dlogits, *demb, *g0, *g1, *dfeat, *feat, *dgw;
  int* arg;
  // CW2 state
  float *inp, *w, *m, *v, *gmodel, *l2part, *loss1, *loss2, *dscores, *cst, *lower, *upper, *best_l2, *gbest_l2, *lossmean;
  float* anstash;           // log-mel forward -> adjoint hand-over (spectrum + mel energies per frame)
  long long *dec, *best_score, *gbest_score;
  size_t bytes;
};

static AnWs an_ws_layout(void* base, int B, int N, int Cp, bool cw2) {
  AnWs w;
  memset(&w, 0, sizeof(w));
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloat) { float* q = (float*)(p + off); off += (nfloat * sizeof(float) + 255) / 256 * 256; return q; };
  const int T0 = 1 + (N - 1) / AN_HOP;
  int t = T0;
  for (int l = 0; l < 7; ++l) { w.T[l] = t; if (kAnPool[l]) t = t / 2; }
  w.T[7] = w.T[6] - 2;                                          // conv8: k=3, no padding
  w.feat = take((size_t)B * T0 * 32); w.dfeat = take((size_t)B * T0 * 32);
  w.c1 = take((size_t)B * T0 * 32);
  size_t gmax = (size_t)B * T0 * 32;
  for (int l = 0; l < 7; ++l) {
    const size_t n = (size_t)B * w.T[l] * kAnCout[l];
    w.a[l] = take(n);
    w.p[l] = kAnPool[l] ? take((size_t)B * (w.T[l] / 2) * kAnCout[l]) : w.a[l];
    if (n > gmax) gmax = n;
    const size_t nin = (size_t)B * w.T[l] * kAnCin[l];
    if (nin > gmax) gmax = nin;
  }
  w.g0 = take(gmax); w.g1 = take(gmax);
  w.emb = take((size_t)B * 32); w.demb = take((size_t)B * 32); w.arg = (int*)take((size_t)B * 32);
  w.logits = take((size_t)B * Cp); w.dlogits = take((size_t)B * Cp);
  w.dgw = take((size_t)B * T0 * AN_WIN);
  if (cw2) {
    const size_t BN = (size_t)B * N;
    w.inp = take(BN); w.w = take(BN); w.m = take(BN); w.v = take(BN); w.gmodel = take(BN);
    w.l2part = take((size_t)B * SG_CW2_CHUNKS); w.loss1 = take(B); w.loss2 = take(B); w.dscores = take((size_t)B * Cp);
    w.cst = take(B); w.lower = take(B); w.upper = take(B); w.best_l2 = take(B); w.gbest_l2 = take(B); w.lossmean = take(4);
    w.dec = (long long*)take((size_t)B * 2); w.best_score = (long long*)take((size_t)B * 2); w.gbest_score = (long long*)take((size_t)B * 2);
    w.anstash = take((size_t)B * T0 * AN_STASH_FLOATS);
  }
  w.bytes = off;
  return w;
}

extern "C" size_t sg_audionet_ws_bytes(const sg_handle* h, int B, int N, int for_cw2) {
  if (!h || !h->an || B < 1 || N < AN_NFFT) return 0;
  return an_ws_layout(nullptr, B, N, h->an->Cp, for_cw2 != 0).bytes;
}
extern "C" int sg_audionet_num_frames(int N) { return 1 + (N - 1) / AN_HOP; }

static int an_check(sg_handle* h, int B, int N) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->an) { sg_set_error("AudioNet weights not loaded (call sg_load_audionet first)"); return SG_ESTATE; }
  if (B < 1 || N < 2 * AN_NFFT) { sg_set_error("AudioNet needs B >= 1 and N >= %d samples (got B=%d N=%d)", 2 * AN_NFFT, B, N); return SG_EINVAL; }
  return SG_OK;
}

// The convolutions run on the tensor cores (conv_tc_kernel's utterance-tiled mode, TF32 operands, fp32 storage) when the
// handle's precision is not fp32; fp32 is the FFMA parity mode.
static void an_conv_args(SgConvArgs& a, const float* A, int cin, const float* W, const float* Wk, const float* bias, float* out, int cout, int rows,
                         int taps, int tap_base, int tap_step, int T, int epi, const float* mask, int ldmask) {
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = cin; a.W = W; a.Wk = Wk; a.bias = bias; a.out = out; a.ldo = cout; a.rows = rows; a.N = cout; a.cin = cin;
  a.taps = taps; a.tap_base = tap_base; a.tap_step = tap_step; a.same_utt = 1; a.T = T; a.t_valid = T; a.epilogue = epi;
  a.mask = mask; a.ldmask = ldmask;
}

// extract_emb (audionet_csine.py:176-207): conv1 .. conv8 + max over time -> emb [B,32] (also left in w.emb / w.arg)
static int an_emb_fwd(sg_handle* h, const float* feat, int B, const AnWs& w, float* emb, cudaStream_t st) {
  SgAudioNet* an = h->an;
  SgConvArgs a;
  an_conv_args(a, feat, 32, an->W1, an->W1k, an->b1, w.c1, 32, B * w.T[0], 5, -2, 1, w.T[0], SG_EPI_BIAS, nullptr, 0);
  SG_TRY(sg_run_conv(h, a, true, SG_PROF_AUDIONET, st));
  const float* in = w.c1;
  for (int l = 0; l < 7; ++l) {
    an_conv_args(a, in, kAnCin[l], an->W[l], an->Wk[l], an->bias[l], w.a[l], kAnCout[l], B * w.T[l], 3, -kAnPad[l], 1, w.T[l],
                 SG_EPI_BIAS_RELU, nullptr, 0);
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_AUDIONET, st));
    if (kAnPool[l]) {
      h->launches += 1;
      PROF(h, SG_PROF_AUDIONET, st, sg_maxpool2_fwd_launch(w.a[l], w.p[l], B, w.T[l], kAnCout[l], st));
    }
    in = w.p[l];
  }
  h->launches += 1;
  PROF(h, SG_PROF_AUDIONET, st, sg_globalmax_fwd_launch(w.a[6], emb, w.arg, B, w.T[6], w.T[7], 32, st));   // valid conv8 outputs only
  return SG_OK;
}
// predict_from_embeddings (audionet_csine.py:210-211): logits = fc(emb)
static int an_fc_fwd(sg_handle* h, const float* emb, int B, float* logits, cudaStream_t st) {
  SgAudioNet* an = h->an;
  SgConvArgs a;
  memset(&a, 0, sizeof(a));
  a.A = emb; a.lda = 32; a.W = an->Wfc; a.bias = an->bfc; a.out = logits; a.ldo = an->Cp; a.rows = B; a.N = an->Cp; a.cin = 32;
  a.taps = 1; a.epilogue = SG_EPI_BIAS; a.T = 1;
  return sg_run_conv(h, a, false, SG_PROF_AUDIONET, st);
}
static int an_fc_bwd(sg_handle* h, const float* dlogits, int B, float* demb, cudaStream_t st) {
  SgAudioNet* an = h->an;
  SgConvArgs a;
  memset(&a, 0, sizeof(a));
  a.A = dlogits; a.lda = an->Cp; a.W = an->Wfcb; a.out = demb; a.ldo = 32; a.rows = B; a.N = 32; a.cin = an->Cp;
  a.taps = 1; a.epilogue = SG_EPI_NONE; a.T = 1;
  return sg_run_conv(h, a, false, SG_PROF_AUDIONET, st);
}
static int an_cnn_fwd(sg_handle* h, const float* feat, int B, const AnWs& w, float* logits, cudaStream_t st) {
  SG_TRY(an_emb_fwd(h, feat, B, w, w.emb, st));
  return an_fc_fwd(h, w.emb, B, logits, st);
}

static int an_emb_bwd(sg_handle* h, const float* demb, int B, const AnWs& w, float* dfeat, cudaStream_t st);
static int an_cnn_bwd(sg_handle* h, const float* dlogits, int B, const AnWs& w, float* dfeat, cudaStream_t st) {
  SG_TRY(an_fc_bwd(h, dlogits, B, w.demb, st));
  return an_emb_bwd(h, w.demb, B, w, dfeat, st);
}
static int an_emb_bwd(sg_handle* h, const float* demb, int B, const AnWs& w, float* dfeat, cudaStream_t st) {
  SgAudioNet* an = h->an;
  SgConvArgs a;
  h->launches += 1;
  PROF(h, SG_PROF_AUDIONET, st, sg_globalmax_bwd_launch(w.a[6], demb, w.arg, w.g0, B, w.T[6], 32, st));   // dA8 (ReLU-masked)
  float* gin = w.g0;
  float* gout = w.g1;
  for (int l = 6; l >= 0; --l) {
    // dgrad of conv stage l: dIn[s] = sum_k W_k^T dOut[s - k + pad]; output = gradient wrt p[l-1] (or c1)
    const bool prev_pooled = l > 0 && kAnPool[l - 1];
    const bool mask_here = l > 0 && !prev_pooled;       // ReLU of stage l-1 sits directly under this conv
    an_conv_args(a, gin, kAnCout[l], an->Wb[l], an->Wbk[l], nullptr, gout, kAnCin[l], B * w.T[l], 3, kAnPad[l], -1, w.T[l],
                 mask_here ? SG_EPI_MASK : SG_EPI_NONE, mask_here ? w.a[l - 1] : nullptr, mask_here ? kAnCout[l - 1] : 0);
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_AUDIONET, st));
    float* t = gin; gin = gout; gout = t;
    if (prev_pooled) {                                  // un-pool into stage l-1's resolution + its ReLU mask
      h->launches += 1;
      PROF(h, SG_PROF_AUDIONET, st, sg_maxpool2_bwd_launch(w.a[l - 1], gin, gout, B, w.T[l - 1], kAnCout[l - 1], st));
      t = gin; gin = gout; gout = t;
    }
  }
  an_conv_args(a, gin, 32, an->W1b, an->W1bk, nullptr, dfeat, 32, B * w.T[0], 5, 2, -1, w.T[0], SG_EPI_NONE, nullptr, 0);
  return sg_run_conv(h, a, true, SG_PROF_AUDIONET, st);
}

// ---------------------------------------------------------------------------------------------
// stage entry points
// ---------------------------------------------------------------------------------------------
extern "C" int sg_audionet_logmel_fwd(sg_handle* h, const float* x, int B, int N, float* feat, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!x || !feat) { sg_set_error("sg_audionet_logmel_fwd: null pointer"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_AUDIONET, (cudaStream_t)stream, sg_an_logmel_fwd_launch(h->an->d_tables, x, B, N, sg_audionet_num_frames(N), feat, (cudaStream_t)stream));
  return SG_OK;
}
extern "C" int sg_audionet_logmel_bwd(sg_handle* h, const float* x, int B, int N, const float* dfeat, void* ws, float* dx,
                                      float scale, int accumulate, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!x || !dfeat || !ws || !dx) { sg_set_error("sg_audionet_logmel_bwd: null pointer"); return SG_EINVAL; }
  AnWs w = an_ws_layout(ws, B, N, h->an->Cp, false);
  h->launches += 2;
  PROF(h, SG_PROF_AUDIONET, (cudaStream_t)stream, sg_an_logmel_bwd_launch(h->an->d_tables, x, B, N, sg_audionet_num_frames(N), dfeat, w.dgw, dx, scale, accumulate, (cudaStream_t)stream));
  return SG_OK;
}
extern "C" int sg_audionet_cnn_fwd(sg_handle* h, const float* feat, int B, int N, void* ws, float* logits, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!feat || !ws || !logits) { sg_set_error("sg_audionet_cnn_fwd: null pointer"); return SG_EINVAL; }
  AnWs w = an_ws_layout(ws, B, N, h->an->Cp, false);
  if (w.T[7] < 1) { sg_set_error("utterance too short for AudioNet's conv8 (N=%d)", N); return SG_EINVAL; }
  return an_cnn_fwd(h, feat, B, w, logits, (cudaStream_t)stream);
}
extern "C" int sg_audionet_cnn_bwd(sg_handle* h, const float* dlogits, int B, int N, void* ws, float* dfeat, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!dlogits || !ws || !dfeat) { sg_set_error("sg_audionet_cnn_bwd: null pointer"); return SG_EINVAL; }
  AnWs w = an_ws_layout(ws, B, N, h->an->Cp, false);
  return an_cnn_bwd(h, dlogits, B, w, dfeat, (cudaStream_t)stream);
}
extern "C" int sg_audionet_emb_fwd(sg_handle* h, const float* feat, int B, int N, void* ws, float* emb, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!feat || !ws || !emb) { sg_set_error("sg_audionet_emb_fwd: null pointer"); return SG_EINVAL; }
  AnWs w = an_ws_layout(ws, B, N, h->an->Cp, false);
  if (w.T[7] < 1) { sg_set_error("utterance too short for AudioNet's conv8 (N=%d)", N); return SG_EINVAL; }
  return an_emb_fwd(h, feat, B, w, emb, (cudaStream_t)stream);
}
extern "C" int sg_audionet_emb_bwd(sg_handle* h, const float* demb, int B, int N, void* ws, float* dfeat, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!demb || !ws || !dfeat) { sg_set_error("sg_audionet_emb_bwd: null pointer"); return SG_EINVAL; }
  AnWs w = an_ws_layout(ws, B, N, h->an->Cp, false);
  return an_emb_bwd(h, demb, B, w, dfeat, (cudaStream_t)stream);
}
extern "C" int sg_audionet_fc_fwd(sg_handle* h, const float* emb, int B, float* logits, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->an || !emb || !logits || B < 1) { sg_set_error("sg_audionet_fc_fwd: AudioNet not loaded or bad argument"); return SG_EINVAL; }
  return an_fc_fwd(h, emb, B, logits, (cudaStream_t)stream);
}
extern "C" int sg_audionet_fc_bwd(sg_handle* h, const float* dlogits, int B, float* demb, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->an || !dlogits || !demb || B < 1) { sg_set_error("sg_audionet_fc_bwd: AudioNet not loaded or bad argument"); return SG_EINVAL; }
  return an_fc_bwd(h, dlogits, B, demb, (cudaStream_t)stream);
}
extern "C" int sg_audionet_num_class_padded(const sg_handle* h) { return (h && h->an) ? h->an->Cp : 0; }
extern "C" int sg_argmax_decide(sg_handle* h, const float* scores, int B, int S, int ld, float threshold, int64_t* decisions,
                                sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!scores || !decisions || B < 1 || S < 1 || ld != S) { sg_set_error("sg_argmax_decide: bad argument (ld must equal S)"); return SG_EINVAL; }
  h->launches += 1;
  return sg_argmax_rows_launch(scores, (long long*)decisions, B, S, threshold, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// CW2 against AudioNet: the whole attack on the device (attack/CW2.py:41-132)
// ---------------------------------------------------------------------------------------------
__global__ void cw2_init_kernel(float* cst, float* lower, float* upper, float* gbest_l2, long long* gbest_score, float c0, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  cst[b] = c0; lower[b] = 0.f; upper[b] = 1e10f; gbest_l2[b] = INFINITY; gbest_score[b] = -2;
}
__global__ void cw2_step_init_kernel(float* best_l2, long long* best_score, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  best_l2[b] = INFINITY; best_score[b] = -2;
}
// logits [B,Cp] -> compact scores [B,C] view is not needed: the loss kernel takes S = C with row stride Cp via a copy
__global__ void compact_rows_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, int B, int S) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * S; i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[(i / S) * ld + (i % S)];
}
__global__ void expand_rows_kernel(const float* __restrict__ in, int S, float* __restrict__ out, int ld, int B) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)B * ld; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / ld; const int c = (int)(i % ld);
    out[i] = c < S ? in[b * S + c] : 0.f;
  }
}
__global__ void cw2_mean_loss_kernel(const float* cst, const float* l1, const float* l2, float* out, int B) {
  // mean over the batch of c*loss1 + loss2 in fp64 (np.mean of python floats), single thread: B is small
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += (double)(cst[b] * l1[b] + l2[b]);
    out[0] = (float)(s / B);
  }
}
__global__ void cw2_success_kernel(const long long* gbest_score, long long* success, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) success[b] = gbest_score[b] != -2 ? 1 : 0;
}

extern "C" int sg_cw2_audionet_run(sg_handle* h, const float* x, const int64_t* y, int B, int N, const sg_cw2_params* p, void* ws,
                                   float* best_x, int64_t* success, float* final_const, sg_stream stream) {
  SG_TRY(an_check(h, B, N));
  if (!x || !y || !p || !ws || !best_x || !success) { sg_set_error("sg_cw2_audionet_run: null argument"); return SG_EINVAL; }
  if (p->loss.loss != SG_LOSS_MARGIN || !p->loss.clip_max) { sg_set_error("CW2 uses the clipped margin loss (attack/CW2.py:39)"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  SgAudioNet* an = h->an;
  const int C = an->C, Cp = an->Cp, T0 = sg_audionet_num_frames(N), nb = (B + 127) / 128;
  AnWs w = an_ws_layout(ws, B, N, Cp, true);
  if (w.T[7] < 1) { sg_set_error("utterance too short for AudioNet (N=%d)", N); return SG_EINVAL; }
  const size_t BN = (size_t)B * N;
  float* scores = w.g1;                                         // [B,C] compact copy of the logits (g1 is free in the forward)
  SG_CUDA_CHECK(cudaMemcpyAsync(best_x, x, BN * sizeof(float), cudaMemcpyDeviceToDevice, st));   // global_best_adver_x = x.clone()
  cw2_init_kernel<<<nb, 128, 0, st>>>(w.cst, w.lower, w.upper, w.gbest_l2, w.gbest_score, p->initial_const, B);
  SG_LAUNCH_CHECK();
  h->cw2_iters = 0;
  for (int bs = 0; bs < p->binary_search_steps; ++bs) {
    SG_CUDA_CHECK(cudaMemsetAsync(w.w, 0, BN * sizeof(float), st));
    SG_CUDA_CHECK(cudaMemsetAsync(w.m, 0, BN * sizeof(float), st));
    SG_CUDA_CHECK(cudaMemsetAsync(w.v, 0, BN * sizeof(float), st));
    cw2_step_init_kernel<<<nb, 128, 0, st>>>(w.best_l2, w.best_score, B);
    SG_LAUNCH_CHECK();
    bool cont = true;
    double prev = INFINITY;
    for (int it = 0; it <= p->max_iter && cont; ++it) {
      h->launches += 2;
      PROF(h, SG_PROF_CW2, st, sg_cw2_prepare_launch(x, w.w, w.inp, w.l2part, w.loss2, B, N, st));
      h->launches += 1;
      const bool grad = it < p->max_iter;
      PROF(h, SG_PROF_AUDIONET, st, sg_an_logmel_fwd_launch(an->d_tables, w.inp, B, N, T0, w.feat, st, grad ? w.anstash : nullptr));
      SG_TRY(an_cnn_fwd(h, w.feat, B, w, w.logits, st));
      compact_rows_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(w.logits, Cp, scores, B, C);
      SG_LAUNCH_CHECK();
      h->launches += 3;
      SG_TRY(sg_argmax_rows_launch(scores, w.dec, B, C, p->decision_threshold, st));
      PROF(h, SG_PROF_LOSS, st, sg_loss_launch(scores, (const long long*)y, B, C, p->loss, w.loss1, grad ? w.dscores : nullptr, st));
      if (grad) {
        h->cw2_iters += 1;
        expand_rows_kernel<<<(B * Cp + 255) / 256, 256, 0, st>>>(w.dscores, C, w.dlogits, Cp, B);
        SG_LAUNCH_CHECK();
        SG_TRY(an_cnn_bwd(h, w.dlogits, B, w, w.dfeat, st));
        h->launches += 3;
        PROF(h, SG_PROF_AUDIONET, st, sg_an_logmel_bwd_launch(an->d_tables, w.inp, B, N, T0, w.dfeat, w.dgw, w.gmodel, 1.0f, 0, st, w.anstash));
        PROF(h, SG_PROF_CW2, st, sg_cw2_adam_launch(w.w, w.m, w.v, x, w.inp, w.gmodel, w.cst, B, N, p->lr, it + 1, st));
      }
      if (p->stop_early && p->stop_early_iter > 0 && it % p->stop_early_iter == 0) {   // attack/CW2.py:96-100
        cw2_mean_loss_kernel<<<1, 32, 0, st>>>(w.cst, w.loss1, w.loss2, w.lossmean, B);
        SG_LAUNCH_CHECK();
        float mean_h = 0.f;
        SG_CUDA_CHECK(cudaMemcpyAsync(&mean_h, w.lossmean, sizeof(float), cudaMemcpyDeviceToHost, st));
        SG_CUDA_CHECK(cudaStreamSynchronize(st));
        if ((double)mean_h > 0.9999 * prev) cont = false;
        prev = mean_h;
      }
      h->launches += 2;
      PROF(h, SG_PROF_CW2, st, sg_cw2_track_launch(w.inp, best_x, w.loss1, w.loss2, w.dec, w.best_l2, w.best_score, w.gbest_l2,
                                                   w.gbest_score, B, N, st));
    }
    h->launches += 1;
    SG_TRY(sg_cw2_search_update_launch(w.cst, w.lower, w.upper, w.best_score, B, st));
  }
  cw2_success_kernel<<<nb, 128, 0, st>>>(w.gbest_score, (long long*)success, B);
  SG_LAUNCH_CHECK();
  if (final_const) SG_CUDA_CHECK(cudaMemcpyAsync(final_const, w.cst, (size_t)B * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SG_OK;
}

extern "C" long long sg_cw2_last_iterations(const sg_handle* h) { return h ? h->cw2_iters : 0; }

// ---------------------------------------------------------------------------------------------
// FeCo (defense/feature_level.py:18-50, :168-217)
// ---------------------------------------------------------------------------------------------
// (launchers of sg_kmeans.cu: declared in sg_head.cuh)
static int feco_check(sg_handle* h, int B, int n, int dim, int k) {
  SG_TRY(sg_check_handle(h, false));
  if (B < 1 || n < 1 || dim < 1 || dim > 32 || k < 1 || k > n) { sg_set_error("FeCo: need B >= 1, 1 <= k <= n, 1 <= dim <= 32 (B=%d n=%d dim=%d k=%d)", B, n, dim, k); return SG_EINVAL; }
  return SG_OK;
}
extern "C" int sg_feco_kmeans(sg_handle* h, const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed,
                              int max_iter, float tol, int32_t* ids, sg_stream stream) {
  SG_TRY(feco_check(h, B, n, dim, k));
  if (!feat || !ids || ld < dim || max_iter < 1) { sg_set_error("sg_feco_kmeans: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  return sg_feco_kmeans_launch(feat, ld, B, n, dim, k, seed, max_iter, tol, (int*)ids, (cudaStream_t)stream, nullptr);
}
extern "C" int sg_feco_kmeans_keyed(sg_handle* h, const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed,
                                    int max_iter, float tol, int32_t* ids, uint32_t pass, uint32_t utt_offset, uint32_t copy_rows,
                                    sg_stream stream) {
  SG_TRY(feco_check(h, B, n, dim, k));
  if (!feat || !ids || ld < dim || max_iter < 1) { sg_set_error("sg_feco_kmeans_keyed: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  return sg_feco_kmeans_launch(feat, ld, B, n, dim, k, seed, max_iter, tol, (int*)ids, (cudaStream_t)stream, nullptr, pass, utt_offset, copy_rows);
}
extern "C" int sg_feco_means_fwd(sg_handle* h, const float* feat, int ld, const int32_t* ids, int B, int n, int dim, int k,
                                 int force, float* out, int32_t* counts, sg_stream stream) {
  SG_TRY(feco_check(h, B, n, dim, k));
  if (!feat || !ids || !out || !counts || ld < dim) { sg_set_error("sg_feco_means_fwd: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  return sg_feco_means_fwd_launch(feat, ld, (const int*)ids, B, n, dim, k, force, out, dim, (int*)counts, (cudaStream_t)stream);
}
extern "C" int sg_feco_means_bwd(sg_handle* h, const float* dout, const int32_t* ids, const int32_t* counts, int B, int n, int dim,
                                 int k, int force, float* dfeat, sg_stream stream) {
  SG_TRY(feco_check(h, B, n, dim, k));
  if (!dout || !ids || !counts || !dfeat) { sg_set_error("sg_feco_means_bwd: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  return sg_feco_means_bwd_launch(dout, dim, (const int*)ids, (const int*)counts, B, n, dim, k, force, dfeat, dim, (cudaStream_t)stream);
}
