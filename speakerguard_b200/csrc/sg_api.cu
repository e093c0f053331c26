// C-ABI of libsgb200 (include/sgb200.h): handle, weight packing, workspace layout, and the
// orchestration of the per-pass kernel sequence and of the whole-attack loops.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include <cuda_bf16.h>

#include "sg_handle.cuh"

// ---------------------------------------------------------------------------------------------
// error string
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void sg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* sg_last_error(void) { return g_err; }
extern "C" int sg_version(void) { return SGB200_VERSION; }

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
static const int kTaps[5] = {5, 5, 7, 1, 1};
static const int kDil[5] = {1, 2, 3, 1, 1};
static const int kCin[5] = {30, 512, 512, 512, 512};
static const int kCinP[5] = {32, 512, 512, 512, 512};
static const int kCout[5] = {512, 512, 512, 512, 1500};
static const int kCoutP[5] = {512, 512, 512, 512, SG_C5P};

static const char* kProfNames[SG_PROF_COUNT] = {
    "mfcc_fwd", "mfcc_bwd", "cmvn", "tdnn_fwd", "tdnn_dgrad", "pool", "head_gemm", "head", "loss", "step", "audionet", "cw2", "iv_gemm", "iv", "tdnn_dgrad5_pool", "feco"};

int sg_dev_upload(sg_handle* h, float** dst, const std::vector<float>& src) {
  SG_CUDA_CHECK(cudaMalloc((void**)dst, src.size() * sizeof(float)));
  h->allocs.push_back(*dst);
  SG_CUDA_CHECK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
  return SG_OK;
}

static int upload_bf16(sg_handle* h, void** dst, const std::vector<float>& src) {
  std::vector<__nv_bfloat16> tmp(src.size());
  for (size_t i = 0; i < src.size(); ++i) tmp[i] = __float2bfloat16_rn(src[i]);
  SG_CUDA_CHECK(cudaMalloc(dst, tmp.size() * sizeof(__nv_bfloat16)));
  h->allocs.push_back(*dst);
  SG_CUDA_CHECK(cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  return SG_OK;
}

extern "C" int sg_create(sg_handle** out, int device) {
  if (!out) { sg_set_error("sg_create: out is NULL"); return SG_EINVAL; }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    sg_set_error("sg_create: no CUDA device (%s); libsgb200 has no CPU fallback", cudaGetErrorString(e));
    return SG_ECUDA;
  }
  if (device < 0 || device >= n) { sg_set_error("sg_create: device %d out of range [0,%d)", device, n); return SG_EINVAL; }
  SG_CUDA_CHECK(cudaSetDevice(device));
  sg_handle* h = new sg_handle();
  h->device = device;
  if (const char* e = getenv("SGB200_POOL_FUSION")) h->pool_fusion = atoi(e) != 0;   // A/B switches for bench.py runs
  if (const char* e = getenv("SGB200_FEAT_STASH")) h->feat_stash = atoi(e) != 0;
  if (const char* e = getenv("SGB200_L1_TAP_FORM")) h->l1_tap_form = atoi(e) != 0;
  if (const char* e = getenv("SGB200_CUDA_GRAPH")) h->use_graph = atoi(e) != 0;
  if (const char* e = getenv("SGB200_CMVN_FUSION")) h->cmvn_fusion = atoi(e) != 0;
  if (const char* e = getenv("SGB200_ROW_COMPACTION")) h->row_compaction = atoi(e) != 0;
  SgFeatTables* host = new SgFeatTables();
  int r = sg_feat_tables_build(host);
  if (r != SG_OK) { delete host; delete h; sg_set_error("sg_create: feature table construction failed"); return r; }
  cudaError_t ce = cudaMalloc((void**)&h->d_tables, sizeof(SgFeatTables));
  if (ce == cudaSuccess) ce = cudaMemcpy(h->d_tables, host, sizeof(SgFeatTables), cudaMemcpyHostToDevice);
  delete host;
  if (ce != cudaSuccess) { sg_set_error("sg_create: table upload failed: %s", cudaGetErrorString(ce)); delete h; return SG_ECUDA; }
  h->allocs.push_back(h->d_tables);
  r = sg_feat_init();
  if (r == SG_OK) r = sg_kmeans_init();
  if (r == SG_OK) r = sg_conv_tc_warm();
  if (r != SG_OK) { sg_destroy(h); return r; }
  *out = h;
  return SG_OK;
}

extern "C" void sg_destroy(sg_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (cudaEvent_t e : h->prof.ev) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) if (h->pgd_graph.exec[i]) cudaGraphExecDestroy(h->pgd_graph.exec[i]);
  if (h->pgd_graph.cap_stream) cudaStreamDestroy(h->pgd_graph.cap_stream);
  sg_comm_destroy(h);
  sg_audionet_free(h);
  sg_iv_free(h);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

extern "C" int sg_set_precision(sg_handle* h, int precision) {
  if (!h) { sg_set_error("null handle"); return SG_EINVAL; }
  if (precision < SG_PREC_FP32 || precision > SG_PREC_BF16) { sg_set_error("unknown precision %d", precision); return SG_EINVAL; }
  h->precision = precision;
  return SG_OK;
}
extern "C" int sg_get_precision(const sg_handle* h) { return h ? h->precision : SG_EINVAL; }
extern "C" int sg_set_option(sg_handle* h, int option, int value) {
  if (!h) { sg_set_error("null handle"); return SG_EINVAL; }
  if (option == SG_OPT_POOL_FUSION) { h->pool_fusion = value != 0; return SG_OK; }
  if (option == SG_OPT_FEAT_STASH) { h->feat_stash = value != 0; return SG_OK; }
  if (option == SG_OPT_L1_TAP_FORM) { h->l1_tap_form = value != 0; return SG_OK; }
  if (option == SG_OPT_CUDA_GRAPH) { h->use_graph = value != 0; return SG_OK; }
  if (option == SG_OPT_CMVN_FUSION) { h->cmvn_fusion = value != 0; return SG_OK; }
  if (option == SG_OPT_ROW_COMPACTION) { h->row_compaction = value != 0; return SG_OK; }
  if (option == SG_OPT_UTT_OFFSET) {
    if (value < 0) { sg_set_error("SG_OPT_UTT_OFFSET must be >= 0"); return SG_EINVAL; }
    h->utt_offset = value; return SG_OK;
  }
  sg_set_error("unknown option %d", option);
  return SG_EINVAL;
}
extern "C" int sg_profile_enable(sg_handle* h, int enable) {
  if (!h) { sg_set_error("null handle"); return SG_EINVAL; }
  h->prof.on = enable != 0;
  h->prof.used = 0;
  return SG_OK;
}
extern "C" int sg_profile_read(sg_handle* h, int category, double* total_ms, long long* launches) {
  if (!h || category < 0 || category >= SG_PROF_COUNT || !total_ms || !launches) { sg_set_error("sg_profile_read: bad argument"); return SG_EINVAL; }
  double tot = 0.0; long long n = 0;
  for (size_t i = 0; i < h->prof.used; ++i) {
    if (h->prof.cat[i] != category) continue;
    SG_CUDA_CHECK(cudaEventSynchronize(h->prof.ev[2 * i + 1]));
    float ms = 0.f;
    SG_CUDA_CHECK(cudaEventElapsedTime(&ms, h->prof.ev[2 * i], h->prof.ev[2 * i + 1]));
    tot += ms; ++n;
  }
  *total_ms = tot; *launches = n;
  return SG_OK;
}
extern "C" int sg_profile_dump(sg_handle* h, int* categories, int* tags, float* ms, int capacity, int* count) {
  if (!h || !count || capacity < 0) { sg_set_error("sg_profile_dump: bad argument"); return SG_EINVAL; }
  const size_t n = h->prof.used < (size_t)capacity ? h->prof.used : (size_t)capacity;
  for (size_t i = 0; i < n; ++i) {
    SG_CUDA_CHECK(cudaEventSynchronize(h->prof.ev[2 * i + 1]));
    float t = 0.f;
    SG_CUDA_CHECK(cudaEventElapsedTime(&t, h->prof.ev[2 * i], h->prof.ev[2 * i + 1]));
    if (categories) categories[i] = h->prof.cat[i];
    if (tags) tags[i] = h->prof.tag[i];
    if (ms) ms[i] = t;
  }
  *count = (int)h->prof.used;
  return SG_OK;
}
extern "C" const char* sg_profile_name(int category) {
  return (category >= 0 && category < SG_PROF_COUNT) ? kProfNames[category] : "";
}

extern "C" long long sg_launch_count(const sg_handle* h) { return h ? h->launches : 0; }
extern "C" void sg_reset_launch_count(sg_handle* h) { if (h) h->launches = 0; }

// ---------------------------------------------------------------------------------------------
// weights: fold eval-mode BatchNorm (affine=False) of layer l into layer l+1 (exact: valid
// convolutions), pack [taps*cin, cout] for the forward and [taps*cout, cin] for dgrad
// ---------------------------------------------------------------------------------------------
// PLDA back-end shared by the x-vector and i-vector systems (plda.py:27-51, :140-190)
int sg_load_backend(sg_handle* h, const float* plda_mean, const float* plda_transform, const float* plda_psi,
                    const float* enroll, int L, int S) {
  const int Lp = (L + 31) / 32 * 32;
  h->L = L; h->Lp = Lp; h->S = S;
  std::vector<float> mean(plda_mean, plda_mean + L), T(plda_transform, plda_transform + (size_t)L * L);
  std::vector<float> Tt((size_t)L * L), ip(L), pr(L), ivg(L);
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) Tt[(size_t)j * L + i] = T[(size_t)i * L + j];
  float ld_given = 0.f, ld_without = 0.f;
  for (int i = 0; i < L; ++i) {
    const float psi = plda_psi[i];
    ip[i] = 1.0f / (psi + 1.0f);
    pr[i] = psi / (psi + 1.0f);
    const float vg = 1.0f + psi / (psi + 1.0f);
    ivg[i] = 1.0f / vg;
    ld_given += logf(vg);
    ld_without += logf(psi + 1.0f);
  }
  SG_TRY(sg_dev_upload(h, &h->plda_mean, mean));
  SG_TRY(sg_dev_upload(h, &h->plda_T, T));
  SG_TRY(sg_dev_upload(h, &h->plda_Tt, Tt));
  SG_TRY(sg_dev_upload(h, &h->inv_psi1, ip));
  SG_TRY(sg_dev_upload(h, &h->psi_ratio, pr));
  SG_TRY(sg_dev_upload(h, &h->inv_var_given, ivg));
  std::vector<float> en(enroll, enroll + (size_t)S * L);
  SG_TRY(sg_dev_upload(h, &h->enroll, en));
  h->H.L = L; h->H.Lp = Lp;
  h->H.plda_mean = h->plda_mean; h->H.plda_T = h->plda_T; h->H.plda_Tt = h->plda_Tt;
  h->H.inv_psi1 = h->inv_psi1; h->H.psi_ratio = h->psi_ratio; h->H.inv_var_given = h->inv_var_given;
  h->H.logdet_given = ld_given; h->H.logdet_without = ld_without;
  h->H.log2pi_L = logf(2.0f * 3.1415926f) * (float)L;            // plda.py:179
  h->backend_loaded = true;
  return SG_OK;
}

extern "C" int sg_load_xv(sg_handle* h, const sg_xv_weights* w) {
  if (!h || !w) { sg_set_error("sg_load_xv: null argument"); return SG_EINVAL; }
  if (h->xv_loaded || h->backend_loaded) { sg_set_error("sg_load_xv: a model is already loaded on this handle"); return SG_ESTATE; }
  if (w->L < 1 || w->L > 512 || w->S < 1) { sg_set_error("sg_load_xv: need 1 <= L <= 512, S >= 1 (L=%d S=%d)", w->L, w->S); return SG_EINVAL; }
  for (int l = 0; l < 5; ++l)
    if (!w->tdnn_w[l] || !w->tdnn_b[l] || !w->bn_mean[l] || !w->bn_var[l]) { sg_set_error("sg_load_xv: null TDNN pointer (layer %d)", l + 1); return SG_EINVAL; }
  if (!w->fc1_w || !w->fc1_b || !w->emb_mean || !w->lda || !w->plda_mean || !w->plda_transform || !w->plda_psi || !w->enroll) {
    sg_set_error("sg_load_xv: null head pointer"); return SG_EINVAL;
  }
  SG_CUDA_CHECK(cudaSetDevice(h->device));
  const float eps = w->bn_eps > 0.f ? w->bn_eps : 1e-5f;
  const int L = w->L, Lp = (L + 31) / 32 * 32, S = w->S;
  h->L = L; h->Lp = Lp; h->S = S;
  for (int l = 0; l < 5; ++l) {
    const int K = kTaps[l], ci = kCin[l], cip = kCinP[l], co = kCout[l], cop = kCoutP[l];
    std::vector<float> Wf((size_t)K * cip * cop, 0.f), Wb((size_t)K * cop * cip, 0.f), bias(cop, 0.f);
    std::vector<float> Wfk((size_t)cop * K * cip, 0.f), Wbk((size_t)cip * K * cop, 0.f);
    std::vector<double> istd(ci, 1.0), mu(ci, 0.0);
    if (l > 0)
      for (int c = 0; c < ci; ++c) {
        istd[c] = 1.0 / sqrt((double)w->bn_var[l - 1][c] + (double)eps);
        mu[c] = w->bn_mean[l - 1][c];
      }
    for (int o = 0; o < co; ++o) {
      double bacc = w->tdnn_b[l][o];
      for (int c = 0; c < ci; ++c)
        for (int k = 0; k < K; ++k) {
          const double wv = (double)w->tdnn_w[l][((size_t)o * ci + c) * K + k] * istd[c];
          bacc -= wv * mu[c];
          Wf[((size_t)k * cip + c) * cop + o] = (float)wv;
          Wb[((size_t)k * cop + o) * cip + c] = (float)wv;
          Wfk[(size_t)o * (K * cip) + (size_t)k * cip + c] = (float)wv;
          Wbk[(size_t)c * (K * cop) + (size_t)k * cop + o] = (float)wv;
        }
      bias[o] = (float)bacc;
    }
    SG_TRY(sg_dev_upload(h, &h->Wf[l], Wf));
    SG_TRY(sg_dev_upload(h, &h->Wb[l], Wb));
    SG_TRY(sg_dev_upload(h, &h->Wfk[l], Wfk));
    SG_TRY(sg_dev_upload(h, &h->Wbk[l], Wbk));
    SG_TRY(upload_bf16(h, &h->Wfk_h[l], Wfk));
    SG_TRY(upload_bf16(h, &h->Wbk_h[l], Wbk));
    if (l == 0) {
      // per-tap form of the layer-1 dgrad (embed_bwd): G[t, k*32 + c] = sum_o dA1[t, o] w[o][c][k], dx[s, c] = sum_k G[s - k*d, k*32 + c]
      std::vector<float> Wg((size_t)K * cip * cop, 0.f);
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int k = 0; k < K; ++k) Wg[((size_t)k * cip + c) * cop + o] = w->tdnn_w[l][((size_t)o * ci + c) * K + k];
      SG_TRY(upload_bf16(h, &h->Wg1_h, Wg));
    }
    SG_TRY(sg_dev_upload(h, &h->bias[l], bias));
  }
  {
    std::vector<float> m5(SG_C5P, 0.f), i5(SG_C5P, 0.f);
    for (int c = 0; c < SG_C5; ++c) { m5[c] = w->bn_mean[4][c]; i5[c] = (float)(1.0 / sqrt((double)w->bn_var[4][c] + (double)eps)); }
    SG_TRY(sg_dev_upload(h, &h->bn5_mean, m5));
    SG_TRY(sg_dev_upload(h, &h->bn5_istd, i5));
  }
  {  // fc1 (xvecTDNN.py:36, :63) with emb_mean folded into the bias (xvector_extract.py:41-43)
    std::vector<float> Wfc((size_t)SG_STATS * SG_EMB, 0.f), Wfcb((size_t)SG_EMB * SG_STATS, 0.f), b(SG_EMB);
    for (int o = 0; o < SG_EMB; ++o) {
      for (int j = 0; j < 2 * SG_C5; ++j) {
        const int jp = j < SG_C5 ? j : SG_C5P + (j - SG_C5);
        const float v = w->fc1_w[(size_t)o * (2 * SG_C5) + j];
        Wfc[(size_t)jp * SG_EMB + o] = v;
        Wfcb[(size_t)o * SG_STATS + jp] = v;
      }
      b[o] = w->fc1_b[o] - w->emb_mean[o];
    }
    SG_TRY(sg_dev_upload(h, &h->Wfc, Wfc));
    SG_TRY(sg_dev_upload(h, &h->Wfc_b, Wfcb));
    h->Wfc_k = h->Wfc_b;    // [512, 3072] is the K-major form of the forward matrix
    h->Wfc_bk = h->Wfc;     // and vice versa
    SG_TRY(sg_dev_upload(h, &h->bfc, b));
  }
  {  // LDA (model/iv_plda.py:423-435): [L, 513], offset in the last column
    std::vector<float> Wl((size_t)SG_EMB * Lp, 0.f), Wlb((size_t)Lp * SG_EMB, 0.f), b(Lp, 0.f);
    for (int i = 0; i < L; ++i) {
      for (int c = 0; c < SG_EMB; ++c) {
        const float v = w->lda[(size_t)i * (SG_EMB + 1) + c];
        Wl[(size_t)c * Lp + i] = v;
        Wlb[(size_t)i * SG_EMB + c] = v;
      }
      b[i] = w->lda[(size_t)i * (SG_EMB + 1) + SG_EMB];
    }
    SG_TRY(sg_dev_upload(h, &h->Wlda, Wl));
    SG_TRY(sg_dev_upload(h, &h->Wlda_b, Wlb));
    h->Wlda_k = h->Wlda_b;  // [Lp, 512]
    h->Wlda_bk = h->Wlda;   // [512, Lp]
    SG_TRY(sg_dev_upload(h, &h->blda, b));
  }
  SG_TRY(sg_load_backend(h, w->plda_mean, w->plda_transform, w->plda_psi, w->enroll, L, S));
  h->xv_loaded = true;
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------------
struct XvWs {
  uint32_t* bits[4];
  float *r[5], *G0, *G1, *G2, *stats, *dstats, *save_mean, *save_std, *ab, *e1, *de1, *e2, *de2, *tsave, *scal;
  float* splitk; size_t splitk_floats;   // fp32 partials of the head's split-K contractions (8 slices x B rows padded to 128 x 512)
  // attack-loop extras
  float *raw, *draw, *feat, *dfeat, *emb, *demb, *scores, *dscores, *loss, *xbuf, *grad, *stash;
  float *xbuf2, *x0c;               // graph replay: the iterate ping-pongs between xbuf / xbuf2, x0 and y are copied in so that
  long long *dec, *yc;              // the captured kernels only reference workspace addresses
  uint32_t* ctl;                    // {pass, seed_lo, seed_hi} read by the MFCC kernels
  float* xtile;                     // EOT samples as batch rows: the iterate repeated eot_batch times
  int *km_ids, *km_cnt;             // FeCo inside the loop: cluster ids [B][m], member counts [B][k]
  size_t bytes;
};

static XvWs xv_ws_layout(void* base, int B, int T, int Lp, int L, int S, bool attack, int N) {
  XvWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloat) { float* q = (float*)(p + off); off += (nfloat * sizeof(float) + 255) / 256 * 256; return q; };
  const size_t R = (size_t)B * T;
  // TDNN activations / gradient ping-pong: fp32, or bf16 in SG_PREC_BF16 mode (then half of each region is used;
  // the layout does not depend on the precision so a workspace stays valid across sg_set_precision)
  auto take_act = [&](size_t nelem) { return take(nelem); };
  for (int l = 0; l < 4; ++l) w.r[l] = take_act(R * SG_C1);
  w.r[4] = take_act(R * SG_C5P);
  for (int l = 0; l < 4; ++l) w.bits[l] = (uint32_t*)take(R * (SG_C1 / 32));
  w.G0 = take_act(R * SG_C5P); w.G1 = take_act(R * SG_C1); w.G2 = take_act(R * SG_C1);
  w.stats = take((size_t)B * SG_STATS); w.dstats = take((size_t)B * SG_STATS);
  w.save_mean = take((size_t)B * SG_C5P); w.save_std = take((size_t)B * SG_C5P);
  w.ab = take((size_t)B * SG_C5P * 2);
  w.e1 = take((size_t)B * SG_EMB); w.de1 = take((size_t)B * SG_EMB);
  w.e2 = take((size_t)B * Lp); w.de2 = take((size_t)B * Lp); w.tsave = take((size_t)B * Lp);
  w.scal = take((size_t)B * 4);
  w.splitk_floats = (size_t)8 * ((B + 127) / 128) * 128 * SG_EMB;
  w.splitk = take(w.splitk_floats);
  w.raw = w.draw = w.feat = w.dfeat = w.emb = w.demb = w.scores = w.dscores = w.loss = w.xbuf = w.grad = w.stash = nullptr;
  w.dec = w.yc = nullptr; w.xbuf2 = w.x0c = nullptr; w.ctl = nullptr; w.xtile = nullptr; w.km_ids = w.km_cnt = nullptr;
  if (attack) {
    w.raw = take(R * SG_FLD); w.draw = take(R * SG_FLD); w.feat = take(R * SG_FLD); w.dfeat = take(R * SG_FLD);
    w.emb = take((size_t)B * L); w.demb = take((size_t)B * L);
    w.scores = take((size_t)B * S); w.dscores = take((size_t)B * S); w.loss = take(B);
    w.dec = (long long*)take((size_t)B * 2);
    w.xbuf = take((size_t)B * N); w.grad = take((size_t)B * N);
    w.stash = take(sg_feat_stash_floats(B, T));      // forward -> adjoint hand-over of the MFCC kernels
    w.xbuf2 = take((size_t)B * N); w.x0c = take((size_t)B * N);
    w.yc = (long long*)take((size_t)B * 2); w.ctl = (uint32_t*)take(64);
    w.xtile = take((size_t)B * N);
    w.km_ids = (int*)take(R); w.km_cnt = (int*)take(R);
  }
  w.bytes = off;
  return w;
}

extern "C" size_t sg_xv_ws_bytes(const sg_handle* h, int B, int T) {
  if (!h || !h->xv_loaded || B < 1 || T < 1) return 0;
  return xv_ws_layout(nullptr, B, T, h->Lp, h->L, h->S, false, 0).bytes;
}
extern "C" size_t sg_pgd_ws_bytes(const sg_handle* h, int B, int N) {
  if (!h || !h->xv_loaded || B < 1 || N < SG_WIN) return 0;
  return xv_ws_layout(nullptr, B, sg_num_frames(N), h->Lp, h->L, h->S, true, N).bytes;
}

// ---------------------------------------------------------------------------------------------
// argument checks
// ---------------------------------------------------------------------------------------------
int sg_check_handle(sg_handle* h, bool need_xv) {
  if (!h) { sg_set_error("null handle"); return SG_EINVAL; }
  // a handle is bound to one device: make it current so that launches, events and per-device kernel attributes of this call
  // land there even when the caller's current device is another one (free when it already is)
  SG_CUDA_CHECK(cudaSetDevice(h->device));
  if (need_xv && !h->xv_loaded) { sg_set_error("x-vector weights not loaded (call sg_load_xv first)"); return SG_ESTATE; }
  return SG_OK;
}
static int check_wave(int B, int N) {
  if (B < 1) { sg_set_error("batch must be >= 1 (B=%d)", B); return SG_EINVAL; }
  if (N < SG_WIN) { sg_set_error("waveform shorter than one 25 ms window (N=%d < %d), kaldi.py:141", N, SG_WIN); return SG_EINVAL; }
  return SG_OK;
}
static int check_dither(int mode, const float* dither) {
  if (mode < SG_DITHER_OFF || mode > SG_DITHER_PHILOX) { sg_set_error("unknown dither mode %d", mode); return SG_EINVAL; }
  if (mode == SG_DITHER_TENSOR && !dither) { sg_set_error("SG_DITHER_TENSOR needs a dither tensor"); return SG_EINVAL; }
  return SG_OK;
}

// pass counter (low 32 bits) + the handle's global utterance offset (high 32 bits), as sg_feat.cu's make_dither() expects
static inline uint64_t dither_pass(const sg_handle* h, uint64_t pass) {
  return (pass & 0xffffffffull) | ((uint64_t)(uint32_t)h->utt_offset << 32);
}

extern "C" int sg_num_frames(int N) { return (N + SG_SHIFT / 2) / SG_SHIFT; }

// ---------------------------------------------------------------------------------------------
// stage entry points
// ---------------------------------------------------------------------------------------------
extern "C" int sg_mfcc_fwd(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                           uint64_t seed, uint64_t pass, float* raw, int ld, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false)); SG_TRY(check_wave(B, N)); SG_TRY(check_dither(dither_mode, dither));
  if (!x || !raw || ld < SG_NCEP || ld > 32) { sg_set_error("sg_mfcc_fwd: bad pointer or ld (%d not in [30,32])", ld); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_MFCC_FWD, (cudaStream_t)stream, sg_feat_fwd_launch(h->d_tables, x, B, N, sg_num_frames(N), dither_mode, dither, seed, dither_pass(h, pass), raw, ld, (cudaStream_t)stream));
  return SG_OK;
}

extern "C" int sg_mfcc_bwd(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                           uint64_t seed, uint64_t pass, const float* draw, int ld, float* grad, float scale,
                           int accumulate, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false)); SG_TRY(check_wave(B, N)); SG_TRY(check_dither(dither_mode, dither));
  if (!x || !draw || !grad || ld < SG_NCEP || ld > 32) { sg_set_error("sg_mfcc_bwd: bad pointer or ld (%d)", ld); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_MFCC_BWD, (cudaStream_t)stream, sg_feat_bwd_launch(h->d_tables, x, B, N, sg_num_frames(N), dither_mode, dither, seed, dither_pass(h, pass), draw, ld, grad,
                            scale, accumulate, (cudaStream_t)stream));
  return SG_OK;
}

extern "C" int sg_dither_fill(sg_handle* h, int B, int N, uint64_t seed, uint64_t pass, float* out, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false)); SG_TRY(check_wave(B, N));
  if (!out) { sg_set_error("sg_dither_fill: null output"); return SG_EINVAL; }
  h->launches += 1;
  return sg_dither_fill_launch(B, sg_num_frames(N), seed, dither_pass(h, pass), out, (cudaStream_t)stream);
}

extern "C" int sg_cmvn_fwd(sg_handle* h, const float* raw, int ld_in, float* out, int ld_out, int B, int T, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!raw || !out || B < 1 || T < 1 || ld_in < SG_NCEP || ld_out < SG_NCEP || ld_in > 32 || ld_out > 32) { sg_set_error("sg_cmvn_fwd: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_CMVN, (cudaStream_t)stream, sg_cmvn_launch(raw, ld_in, out, ld_out, B, T, 0, (cudaStream_t)stream));
  return SG_OK;
}
extern "C" int sg_cmvn_bwd(sg_handle* h, const float* dout, int ld_in, float* draw, int ld_out, int B, int T, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!dout || !draw || B < 1 || T < 1 || ld_in < SG_NCEP || ld_out < SG_NCEP || ld_in > 32 || ld_out > 32) { sg_set_error("sg_cmvn_bwd: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_CMVN, (cudaStream_t)stream, sg_cmvn_launch(dout, ld_in, draw, ld_out, B, T, 1, (cudaStream_t)stream));
  return SG_OK;
}

// ---- TDNN -------------------------------------------------------------------------------------
int sg_run_conv(sg_handle* h, const SgConvArgs& a, bool tensor_ok, int cat, cudaStream_t st) {
  h->launches += 1;
  if (h->precision != SG_PREC_FP32 && tensor_ok) {
    PROF(h, cat, st, sg_conv_tc(a, h->precision, st));
    h->launches += sg_conv_tc_extra_launches();                  // the split-K reduction, when one was launched
    return SG_OK;
  }
  PROF(h, cat, st, sg_conv_simt(a, st));
  return SG_OK;
}

static void tdnn_valid(int T, int tv[5]) {
  int t = T;
  for (int l = 0; l < 5; ++l) { t -= (kTaps[l] - 1) * kDil[l]; tv[l] = t; }
}

// frames per utterance kept from layer 3's output on (0: no compaction).  Needs the tensor-core path (3-D TMA stores) and
// >= 128 rows per utterance on both sides, so that a 128-row box touches at most two utterances.
static int xv_compact_T(const sg_handle* h, int T) {
  if (h->precision == SG_PREC_FP32 || !h->row_compaction) return 0;
  const int tc = T - (kTaps[0] - 1) * kDil[0] - (kTaps[1] - 1) * kDil[1] - (kTaps[2] - 1) * kDil[2];
  return (tc >= 128 && tc < T) ? tc : 0;
}

static int embed_fwd(sg_handle* h, const float* feat, int B, int T, const XvWs& w, float* emb, cudaStream_t st) {
  int tv[5];
  tdnn_valid(T, tv);
  if (tv[4] < 2) { sg_set_error("need at least 32 frames for the TDNN + unbiased std (T=%d)", T); return SG_EINVAL; }
  const int R = B * T;
  // Row compaction (tensor-core modes): layers 4 and 5 are 1 x 1, so from layer 3's output on only the tv[2] valid frames of
  // every utterance need to exist.  Layer 3's epilogue stores them as [B][Tc] rows, and layers 4, 5, the pooling and their
  // adjoints run on B * Tc instead of B * T rows (270 / 300 at 3 s).
  const int Tc = xv_compact_T(h, T);
  const float* in = feat;
  int lda = SG_FLD;
  for (int l = 0; l < 5; ++l) {
    SgConvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = in; a.lda = lda; a.W = h->Wf[l]; a.Wk = h->Wfk[l]; a.bias = h->bias[l]; a.out = w.r[l]; a.ldo = kCoutP[l];
    a.rows = R; a.N = kCoutP[l]; a.cin = kCinP[l]; a.taps = kTaps[l]; a.tap_step = kDil[l];
    a.epilogue = SG_EPI_BIAS_RELU; a.T = T; a.t_valid = tv[l];
    if (Tc && l == 2) { a.out_T = Tc; a.out_Tstride = Tc; a.bits_T = Tc; }
    if (Tc && l >= 3) { a.rows = B * Tc; a.T = Tc; }
    if (l < 4 && h->precision != SG_PREC_FP32) { a.bits_out = w.bits[l]; a.ldbits = SG_C1 / 32; }
    if (h->precision == SG_PREC_BF16) {                  // bf16 activations; layer 1 still reads the fp32 features
      a.out_bf16 = 1;
      if (l > 0) { a.op_bf16 = 1; a.Wk = (const float*)h->Wfk_h[l]; }
    }
    h->prof.next_tag = l + 1;
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_TDNN_FWD, st));
    in = w.r[l]; lda = kCoutP[l];
  }
  h->launches += 1;
  PROF(h, SG_PROF_POOL, st, sg_pool_fwd_launch(w.r[4], h->precision == SG_PREC_BF16, B, Tc ? Tc : T, tv[4], h->bn5_mean, h->bn5_istd, w.stats, w.save_mean, w.save_std, st));
  {
    SgConvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = w.stats; a.lda = SG_STATS; a.W = h->Wfc; a.Wk = h->Wfc_k; a.bias = h->bfc; a.out = w.e1; a.ldo = SG_EMB;
    a.rows = B; a.N = SG_EMB; a.cin = SG_STATS; a.taps = 1; a.tap_step = 0; a.epilogue = SG_EPI_BIAS; a.T = 1;
    a.splitk_ws = w.splitk; a.splitk_floats = w.splitk_floats;
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_HEAD_GEMM, st));
    a.A = w.e1; a.lda = SG_EMB; a.W = h->Wlda; a.Wk = h->Wlda_k; a.bias = h->blda; a.out = w.e2; a.ldo = h->Lp;
    a.N = h->Lp; a.cin = SG_EMB;
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_HEAD_GEMM, st));
  }
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, st, sg_head_fwd_launch(h->H, w.e2, B, w.tsave, w.scal, emb, st));
  return SG_OK;
}

static int embed_bwd(sg_handle* h, const float* demb, int B, int T, const XvWs& w, float* dfeat, cudaStream_t st) {
  int tv[5];
  tdnn_valid(T, tv);
  const int R = B * T;
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, st, sg_head_bwd_launch(h->H, demb, B, w.tsave, w.scal, w.de2, st));
  {
    SgConvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = w.de2; a.lda = h->Lp; a.W = h->Wlda_b; a.Wk = h->Wlda_bk; a.out = w.de1; a.ldo = SG_EMB;
    a.rows = B; a.N = SG_EMB; a.cin = h->Lp; a.taps = 1; a.epilogue = SG_EPI_NONE; a.T = 1;
    a.splitk_ws = w.splitk; a.splitk_floats = w.splitk_floats;
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_HEAD_GEMM, st));
    a.A = w.de1; a.lda = SG_EMB; a.W = h->Wfc_b; a.Wk = h->Wfc_bk; a.out = w.dstats; a.ldo = SG_STATS; a.N = SG_STATS; a.cin = SG_EMB;
    SG_TRY(sg_run_conv(h, a, true, SG_PROF_HEAD_GEMM, st));
  }
  // bf16 mode: the pooling adjoint is applied to the staged r5 tiles inside the layer-5 dgrad (no dA5 round trip through HBM)
  const int Tc = xv_compact_T(h, T);
  const bool fuse_pool = h->precision == SG_PREC_BF16 && h->pool_fusion && T >= 128;
  h->launches += 1;
  if (fuse_pool) PROF(h, SG_PROF_POOL, st, sg_pool_bwd_params_launch(B, tv[4], h->bn5_istd, w.dstats, w.save_mean, w.save_std, w.ab, st));
  else PROF(h, SG_PROF_POOL, st, sg_pool_bwd_launch(w.r[4], h->precision == SG_PREC_BF16, B, Tc ? Tc : T, tv[4], h->bn5_istd, w.dstats, w.save_mean, w.save_std, w.G0, st));
  // dgrad chain: dA_l (pre-ReLU grad of layer l) -> dA_{l-1}
  const float* gin = w.G0;
  float* bufs[2] = {w.G1, w.G2};
  for (int l = 4; l >= 0; --l) {
    SgConvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = gin; a.lda = kCoutP[l]; a.W = h->Wb[l]; a.Wk = h->Wbk[l]; a.rows = R; a.cin = kCoutP[l]; a.taps = kTaps[l];
    a.tap_step = -kDil[l]; a.T = T;
    if (Tc && l >= 3) { a.rows = B * Tc; a.T = Tc; }                // compact rows: the adjoints of the 1 x 1 layers
    if (h->precision == SG_PREC_BF16) { a.op_bf16 = 1; a.out_bf16 = l > 0; a.Wk = (const float*)h->Wbk_h[l]; }
    if (l == 4 && fuse_pool) { a.A = w.r[4]; a.xf_ab = w.ab; a.xf_ld = SG_C5P; a.xf_tv = tv[4]; }
    h->prof.next_tag = l + 1;
    if (l > 0) {
      float* out = bufs[(4 - l) & 1];
      a.out = out; a.ldo = kCinP[l]; a.N = kCinP[l];
      a.epilogue = SG_EPI_MASK; a.mask = w.r[l - 1]; a.ldmask = kCoutP[l - 1]; a.t_valid = tv[l - 1];
      if (h->precision != SG_PREC_FP32) { a.bits_in = w.bits[l - 1]; a.ldbits = SG_C1 / 32; }
      if (Tc && l == 3) {
        // layer 4's adjoint hands dA3 back in layer 3's row space ([B][T] rows): its epilogue spreads the compact rows out
        // again, and the T - Tc tail rows of every utterance, which layer 3's adjoint reads as taps, are cleared first
        const size_t es = a.out_bf16 ? 2 : 4, rowb = (size_t)kCinP[l] * es;
        SG_CUDA_CHECK(cudaMemset2DAsync((char*)out + (size_t)Tc * rowb, (size_t)T * rowb, 0, (size_t)(T - Tc) * rowb, (size_t)B, st));
        a.out_T = Tc; a.out_Tstride = T;
      }
      SG_TRY(sg_run_conv(h, a, true, (l == 4 && fuse_pool) ? SG_PROF_TDNN_BWD_POOL : SG_PROF_TDNN_BWD, st));
      gin = out;
    } else if (h->precision == SG_PREC_BF16 && h->l1_tap_form) {
      // layer-1 dgrad in per-tap form: one K = 512 contraction into G [R, taps * 32] (8 k-blocks per tile instead of 40:
      // the N = 32 contraction pays the same ~650 cycles per k-block as a 256-column one), then the shifted sum over taps
      float* G = w.G0;                                   // free here: dA5 (or nothing, with the pooling fusion) was consumed
      a.Wk = (const float*)h->Wg1_h; a.W = nullptr; a.taps = 1; a.tap_step = 0;
      a.out = G; a.ldo = kTaps[0] * SG_FLD; a.N = kTaps[0] * SG_FLD; a.epilogue = SG_EPI_NONE;
      SG_TRY(sg_run_conv(h, a, true, SG_PROF_TDNN_BWD, st));
      h->launches += 1;
      h->prof.next_tag = 1;
      PROF(h, SG_PROF_TDNN_BWD, st, sg_tap_gather_launch(G, kTaps[0] * SG_FLD, dfeat, SG_FLD, R, kTaps[0], kDil[0], st));
    } else {
      a.out = dfeat; a.ldo = SG_FLD; a.N = SG_FLD; a.epilogue = SG_EPI_NONE;
      SG_TRY(sg_run_conv(h, a, true, SG_PROF_TDNN_BWD, st));
    }
  }
  return SG_OK;
}

extern "C" int sg_xv_embed_fwd(sg_handle* h, const float* feat, int B, int T, void* ws, float* emb, sg_stream stream) {
  SG_TRY(sg_check_handle(h, true));
  if (!feat || !ws || !emb || B < 1) { sg_set_error("sg_xv_embed_fwd: bad argument"); return SG_EINVAL; }
  XvWs w = xv_ws_layout(ws, B, T, h->Lp, h->L, h->S, false, 0);
  return embed_fwd(h, feat, B, T, w, emb, (cudaStream_t)stream);
}
extern "C" int sg_xv_embed_bwd(sg_handle* h, const float* demb, int B, int T, void* ws, float* dfeat, sg_stream stream) {
  SG_TRY(sg_check_handle(h, true));
  if (!demb || !ws || !dfeat || B < 1) { sg_set_error("sg_xv_embed_bwd: bad argument"); return SG_EINVAL; }
  XvWs w = xv_ws_layout(ws, B, T, h->Lp, h->L, h->S, false, 0);
  return embed_bwd(h, demb, B, T, w, dfeat, (cudaStream_t)stream);
}

// generic contraction, either path (tests): W is [taps*cin, N], Wk its K-major copy [N, taps*cin]
extern "C" int sg_debug_conv(sg_handle* h, int precision, const float* A, int lda, const float* W, const float* Wk,
                             const float* bias, float* out, int ldo, int rows, int N, int cin, int taps, int tap_step,
                             int epilogue, const float* mask, int ldmask, int T, int t_valid, int op_bf16, int out_bf16,
                             sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!A || !out || rows < 1 || N < 1) { sg_set_error("sg_debug_conv: bad argument"); return SG_EINVAL; }
  SgConvArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.W = W; a.Wk = Wk; a.bias = bias; a.out = out; a.ldo = ldo; a.rows = rows; a.N = N; a.cin = cin;
  a.taps = taps; a.tap_step = tap_step; a.epilogue = epilogue; a.mask = mask; a.ldmask = ldmask; a.T = T; a.t_valid = t_valid;
  a.op_bf16 = op_bf16; a.out_bf16 = out_bf16;
  h->launches += 1;
  if (precision == SG_PREC_FP32) return sg_conv_simt(a, (cudaStream_t)stream);
  return sg_conv_tc(a, precision, (cudaStream_t)stream);
}

// ---- scoring / loss ----------------------------------------------------------------------------
extern "C" int sg_plda_score_fwd(sg_handle* h, const float* emb, int B, const float* enroll, int S, float threshold,
                                 float* scores, int64_t* decisions, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->backend_loaded) { sg_set_error("sg_plda_score_fwd: no PLDA back-end loaded (sg_load_xv / sg_load_iv)"); return SG_ESTATE; }
  if (!emb || !scores || B < 1) { sg_set_error("sg_plda_score_fwd: bad argument"); return SG_EINVAL; }
  if (!enroll) { enroll = h->enroll; S = h->S; }
  if (S < 1) { sg_set_error("sg_plda_score_fwd: S must be >= 1"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, (cudaStream_t)stream, sg_score_fwd_launch(h->H, emb, B, enroll, S, threshold, scores, (long long*)decisions, (cudaStream_t)stream));
  return SG_OK;
}
extern "C" int sg_plda_score_bwd(sg_handle* h, const float* emb, const float* dscores, int B, const float* enroll, int S,
                                 float* demb, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->backend_loaded) { sg_set_error("sg_plda_score_bwd: no PLDA back-end loaded (sg_load_xv / sg_load_iv)"); return SG_ESTATE; }
  if (!emb || !dscores || !demb || B < 1) { sg_set_error("sg_plda_score_bwd: bad argument"); return SG_EINVAL; }
  if (!enroll) { enroll = h->enroll; S = h->S; }
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, (cudaStream_t)stream, sg_score_bwd_launch(h->H, emb, dscores, B, enroll, S, demb, (cudaStream_t)stream));
  return SG_OK;
}
static int check_loss(const sg_loss_params* lp, int S) {
  if (!lp) { sg_set_error("null loss params"); return SG_EINVAL; }
  if (lp->loss == SG_LOSS_CE && lp->task != SG_TASK_CSI) { sg_set_error("CrossEntropy only supports the CSI task (attack/utils.py:12)"); return SG_EINVAL; }
  if (lp->task == SG_TASK_SV && S != 1) { sg_set_error("SV task needs exactly one enrolled speaker (S=%d)", S); return SG_EINVAL; }
  return SG_OK;
}
extern "C" int sg_loss_fwd_bwd(sg_handle* h, const float* scores, const int64_t* y, int B, int S, const sg_loss_params* lp,
                               float* loss, float* dscores, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false)); SG_TRY(check_loss(lp, S));
  if (!scores || !y || !loss || B < 1 || S < 1) { sg_set_error("sg_loss_fwd_bwd: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_LOSS, (cudaStream_t)stream, sg_loss_launch(scores, (const long long*)y, B, S, *lp, loss, dscores, (cudaStream_t)stream));
  return SG_OK;
}

extern "C" int sg_step_linf(sg_handle* h, float* x, const float* x0, const float* grad, size_t n, float step,
                            float grad_sign, float eps, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!x || !x0 || !grad) { sg_set_error("sg_step_linf: null pointer"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_STEP, (cudaStream_t)stream, sg_step_linf_launch(x, x0, grad, n, step * grad_sign, eps, (cudaStream_t)stream));
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
// fused passes
// ---------------------------------------------------------------------------------------------
// fused-CMVN scratch of the handle (grown outside any stream capture: called at the top of the entry points)
static int cmvn_scratch_reserve(sg_handle* h, int B) {
  if (!h->cmvn_fusion || B <= h->cm_cap) return SG_OK;
  float* part = nullptr; unsigned int* cnt = nullptr;
  SG_CUDA_CHECK(cudaMalloc(&part, sg_feat_cmvn_part_floats(B) * sizeof(float)));
  SG_CUDA_CHECK(cudaMalloc(&cnt, (size_t)B * sizeof(unsigned int)));
  SG_CUDA_CHECK(cudaMemset(cnt, 0, (size_t)B * sizeof(unsigned int)));
  h->allocs.push_back(part); h->allocs.push_back(cnt);            // an outgrown pair stays allocated until sg_destroy
  h->cm_part = part; h->cm_count = cnt; h->cm_cap = B;
  return SG_OK;
}

// FeCo inside the loop (sg_pgd_params::feco_ratio): k cluster means per utterance replace the m raw frames
struct FecoSpec {
  int k;                  // 0: no defense
  int max_iter; float tol;
  uint64_t seed;
  uint32_t pass;          // pass index of the run (graph replay: the round inside the iteration, the control block adds the rest):
  const uint32_t* ctl;    // the kernel draws a fresh clustering per pass
  uint32_t copy_rows;     // EOT copies as batch rows: utterances per copy (0: rows are utterances)
};
static const FecoSpec kNoFeco = {0, 0, 0.f, 0, 0, nullptr, 0};

static int forward_pass(sg_handle* h, const float* x, int B, int N, int m, int mode, const float* dither, uint64_t seed,
                        uint64_t pass, float thr, const XvWs& w, float* emb, float* scores, long long* dec, cudaStream_t st,
                        float* stash = nullptr, const FecoSpec& fc = kNoFeco) {
  int T = m;                                                       // frames the TDNN sees
  if (fc.k > 0) {
    // raw MFCC -> k-means ids -> cluster means (w.draw as the staging tensor: free in the forward) -> CMVN over the k means
    h->launches += 5;
    PROF(h, SG_PROF_MFCC_FWD, st, sg_feat_fwd_launch(h->d_tables, x, B, N, m, mode, dither, seed, dither_pass(h, pass), w.raw, SG_FLD, st, stash));
    PROF(h, SG_PROF_FECO, st, sg_feco_kmeans_launch(w.raw, SG_FLD, B, m, SG_NCEP, fc.k, fc.seed, fc.max_iter, fc.tol, w.km_ids, st, fc.ctl, fc.pass,
                                                    (uint32_t)h->utt_offset, fc.copy_rows));
    PROF(h, SG_PROF_FECO, st, sg_feco_means_fwd_launch(w.raw, SG_FLD, w.km_ids, B, m, SG_NCEP, fc.k, 1, w.draw, SG_FLD, w.km_cnt, st));
    PROF(h, SG_PROF_CMVN, st, sg_cmvn_launch(w.draw, SG_FLD, w.feat, SG_FLD, B, fc.k, 0, st));
    T = fc.k;
  } else if (h->cmvn_fusion && sg_feat_cmvn_fusable(m)) {
    // CMVN inside the MFCC kernel (one CTA cluster per utterance): no raw-feature round trip, one launch less
    h->launches += 2;
    PROF(h, SG_PROF_MFCC_FWD, st, sg_feat_fwd_launch(h->d_tables, x, B, N, m, mode, dither, seed, dither_pass(h, pass), w.feat, SG_FLD, st, stash, 1,
                                                     h->cm_part, h->cm_count));
  } else {
    h->launches += 3;
    PROF(h, SG_PROF_MFCC_FWD, st, sg_feat_fwd_launch(h->d_tables, x, B, N, m, mode, dither, seed, dither_pass(h, pass), w.raw, SG_FLD, st, stash));
    PROF(h, SG_PROF_CMVN, st, sg_cmvn_launch(w.raw, SG_FLD, w.feat, SG_FLD, B, m, 0, st));
  }
  SG_TRY(embed_fwd(h, w.feat, B, T, w, emb, st));
  PROF(h, SG_PROF_HEAD, st, sg_score_fwd_launch(h->H, emb, B, h->enroll, h->S, thr, scores, dec, st));
  return SG_OK;
}

extern "C" int sg_xv_forward(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                             uint64_t seed, uint64_t pass, float decision_threshold, void* ws, float* scores,
                             int64_t* decisions, float* emb, sg_stream stream) {
  SG_TRY(sg_check_handle(h, true)); SG_TRY(check_wave(B, N)); SG_TRY(check_dither(dither_mode, dither));
  if (!x || !ws || !scores) { sg_set_error("sg_xv_forward: bad argument"); return SG_EINVAL; }
  const int m = sg_num_frames(N);
  XvWs w = xv_ws_layout(ws, B, m, h->Lp, h->L, h->S, true, N);
  SG_TRY(cmvn_scratch_reserve(h, B));
  return forward_pass(h, x, B, N, m, dither_mode, dither, seed, pass, decision_threshold, w, emb ? emb : w.emb, scores,
                      (long long*)decisions, (cudaStream_t)stream);
}

// FeCo geometry of a run: k cluster means for m frames (0: no defense)
static int feco_k(const sg_pgd_params* p, int m) { return p->feco_ratio > 0.f ? (int)((float)m * p->feco_ratio) : 0; }
static FecoSpec feco_spec(const sg_pgd_params* p, int m, uint64_t pass, const uint32_t* ctl) {
  FecoSpec fc;
  fc.k = feco_k(p, m); fc.max_iter = p->feco_max_iter > 0 ? p->feco_max_iter : 100; fc.tol = p->feco_tol > 0.f ? p->feco_tol : 0.01f;
  fc.seed = p->seed; fc.pass = (uint32_t)pass; fc.copy_rows = 0;
  fc.ctl = ctl;
  return fc;
}

// One PGD iteration (attack/FGSM.py:44-68 for iter < max_iter): E gradient samples (EOT) and the sign step, reading the iterate
// from `cur` and writing the next one to `other` (E == 1: fused into the MFCC adjoint) or back into `cur` (E > 1).
// The E samples run as E / Eb passes of B * Eb rows (Eb = eot_batch copies of the batch per pass, adaptive_attack/EOT.py:30-42).
// `it` only selects the dither slice / loss-history row of the launch-by-launch path; with a device control block (ctl) the
// pass counter comes from there and the function is iteration-independent, which is what makes it capturable.
static int pgd_iteration(sg_handle* h, float* cur, float* other, const float* x0, const long long* y, const float* dither, int B,
                         int N, int m, const sg_pgd_params* p, float grad_sign, const XvWs& w, float* sc, long long* dec,
                         float* loss_hist, int it, uint32_t* ctl, cudaStream_t st) {
  const int E = p->eot_size, Eb = p->eot_batch > 1 ? p->eot_batch : 1, rounds = E / Eb, Bv = B * Eb;
  struct CopyRowsGuard { ~CopyRowsGuard() { sg_feat_set_copy_rows(0); } } copy_rows_guard;   // whatever path this function leaves by
  const size_t dstride = (size_t)B * m * SG_WIN;
  float* const stash = h->feat_stash ? w.stash : nullptr;
  const int k = feco_k(p, m), T = k > 0 ? k : m;
  for (int e = 0; e < rounds; ++e) {
    const uint64_t pass = ctl ? (uint64_t)e : (uint64_t)it * rounds + e;
    const float* dth = dither ? dither + ((uint64_t)it * E + e) * dstride : nullptr;
    const float* xin = cur;
    const long long* yin = y;
    float* scb = sc;
    long long* decb = dec;
    if (Eb > 1) {
      h->launches += 1;
      PROF(h, SG_PROF_STEP, st, sg_tile_rows_launch(cur, w.xtile, (size_t)N, B, Eb, y, w.yc, st));
      xin = w.xtile; yin = w.yc; scb = w.scores; decb = w.dec;
    }
    FecoSpec fc = feco_spec(p, m, pass, ctl);
    if (Eb > 1) { fc.copy_rows = (uint32_t)B; sg_feat_set_copy_rows(B); }   // noise / clustering keyed by (copy, global utterance)
    SG_TRY(forward_pass(h, xin, Bv, N, m, p->dither_mode, dth, p->seed, pass, p->decision_threshold, w, w.emb, scb, decb, st, stash, fc));
    float* lossp = (loss_hist && e == 0) ? loss_hist + (size_t)it * B : w.loss;
    h->launches += 3;
    PROF(h, SG_PROF_LOSS, st, sg_loss_launch(scb, yin, Bv, h->S, p->loss, lossp, w.dscores, st));
    PROF(h, SG_PROF_HEAD, st, sg_score_bwd_launch(h->H, w.emb, w.dscores, Bv, h->enroll, h->S, w.demb, st));
    SG_TRY(embed_bwd(h, w.demb, Bv, T, w, w.dfeat, st));
    const int fuse_cmvn = k == 0 && h->cmvn_fusion && sg_feat_cmvn_fusable(m);
    const float* dr = w.dfeat;                                     // fused: the MFCC adjoint applies dx = dy - mean(dy) itself
    if (k > 0) {
      // CMVN adjoint over the k means (w.feat is free in the backward), then the cluster-mean adjoint back to the m frames
      h->launches += 1;
      PROF(h, SG_PROF_CMVN, st, sg_cmvn_launch(w.dfeat, SG_FLD, w.feat, SG_FLD, Bv, k, 1, st));
      PROF(h, SG_PROF_FECO, st, sg_feco_means_bwd_launch(w.feat, SG_FLD, w.km_ids, w.km_cnt, Bv, m, SG_NCEP, k, 1, w.draw, SG_FLD, st));
      dr = w.draw;
    } else if (!fuse_cmvn) {
      PROF(h, SG_PROF_CMVN, st, sg_cmvn_launch(w.dfeat, SG_FLD, w.draw, SG_FLD, Bv, m, 1, st));
      dr = w.draw;
    } else {
      h->launches -= 1;
    }
    h->launches += 1;
    if (E == 1) {
      PROF(h, SG_PROF_MFCC_BWD, st, sg_feat_bwd_step_launch(h->d_tables, cur, B, N, m, p->dither_mode, dth, p->seed, dither_pass(h, pass), dr, SG_FLD, x0,
                                     other, p->step_size * grad_sign, p->epsilon, st, stash, fuse_cmvn));
    } else if (Eb == 1) {
      PROF(h, SG_PROF_MFCC_BWD, st, sg_feat_bwd_launch(h->d_tables, cur, B, N, m, p->dither_mode, dth, p->seed, dither_pass(h, pass), dr, SG_FLD, w.grad,
                                1.0f / (float)E, e > 0, st, stash, fuse_cmvn));
    } else {
      // one gradient row per copy, then the copies of an utterance are summed in copy order into the [B, N] accumulator
      h->launches += 1;
      PROF(h, SG_PROF_MFCC_BWD, st, sg_feat_bwd_launch(h->d_tables, xin, Bv, N, m, p->dither_mode, dth, p->seed, dither_pass(h, pass), dr, SG_FLD, w.grad,
                                1.0f / (float)E, 0, st, stash, 0));
      PROF(h, SG_PROF_STEP, st, sg_reduce_rows_launch(w.grad, w.xbuf2, (size_t)N, B, Eb, e > 0, st));
    }
    sg_feat_set_copy_rows(0);
  }
  if (E > 1) {
    h->launches += 1;
    PROF(h, SG_PROF_STEP, st, sg_step_linf_launch(cur, x0, Eb > 1 ? w.xbuf2 : w.grad, (size_t)B * N, p->step_size * grad_sign, p->epsilon, st));
  }
  if (ctl) { h->launches += 1; SG_TRY(sg_feat_ctl_tick_launch(ctl, (uint32_t)rounds, st)); }
  return SG_OK;
}

// capture pgd_iteration for both ping-pong parities; on any failure the capture is abandoned and the caller falls back
static int pgd_capture(sg_handle* h, int B, int N, int m, const sg_pgd_params* p, float grad_sign, const XvWs& w,
                       const unsigned long long* key, cudaStream_t st) {
  SgPgdGraph& g = h->pgd_graph;
  for (int i = 0; i < 2; ++i) if (g.exec[i]) { cudaGraphExecDestroy(g.exec[i]); g.exec[i] = nullptr; }
  g.valid = false;
  const long long launches0 = h->launches;
  // capture on a private stream: the caller's stream may be the legacy default stream, which cannot be captured; captured
  // work does not execute, so no ordering with the caller's stream is needed here
  if (!g.cap_stream && cudaStreamCreateWithFlags(&g.cap_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return SG_ECUDA; }
  st = g.cap_stream;
  for (int par = 0; par < 2; ++par) {
    float* cur = par == 0 ? w.xbuf : w.xbuf2;
    float* other = par == 0 ? w.xbuf2 : w.xbuf;
    if (p->eot_size > 1) { cur = w.xbuf; other = w.xbuf2; }              // E > 1 steps in place: one parity only
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return SG_ECUDA; }
    sg_feat_set_ctl(w.ctl);
    int r = pgd_iteration(h, cur, other, w.x0c, w.yc, nullptr, B, N, m, p, grad_sign, w, w.scores, w.dec, nullptr, 0, w.ctl, st);
    sg_feat_set_ctl(nullptr);
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (r != SG_OK || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return SG_ECUDA; }
    e = cudaGraphInstantiate(&g.exec[par], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { g.exec[par] = nullptr; cudaGetLastError(); return SG_ECUDA; }
  }
  g.kernels = (int)((h->launches - launches0) / 2);
  h->launches = launches0;                                               // captured, not launched
  memcpy(g.key, key, sizeof(g.key));
  g.valid = true;
  return SG_OK;
}

extern "C" int sg_pgd_run(sg_handle* h, float* x_adv, const float* x0, const int64_t* y, const float* dither, int B, int N,
                          const sg_pgd_params* p, void* ws, int64_t* decisions, float* scores, float* loss_hist,
                          sg_stream stream) {
  SG_TRY(sg_check_handle(h, true)); SG_TRY(check_wave(B, N));
  if (!x_adv || !x0 || !y || !p || !ws) { sg_set_error("sg_pgd_run: null argument"); return SG_EINVAL; }
  SG_TRY(check_dither(p->dither_mode, dither)); SG_TRY(check_loss(&p->loss, h->S));
  if (p->max_iter < 0 || p->eot_size < 1) { sg_set_error("sg_pgd_run: max_iter >= 0 and eot_size >= 1 required"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const int m = sg_num_frames(N), E = p->eot_size;
  const int Eb = p->eot_batch > 1 ? p->eot_batch : 1, rounds = E / Eb;
  if (E % Eb != 0) { sg_set_error("sg_pgd_run: eot_batch (%d) must divide eot_size (%d)", Eb, E); return SG_EINVAL; }
  if (Eb > 1 && (p->dither_mode == SG_DITHER_TENSOR || loss_hist)) {
    sg_set_error("sg_pgd_run: eot_batch > 1 needs dither OFF / PHILOX and no loss history"); return SG_EUNSUPPORTED;
  }
  const int fk = feco_k(p, m);
  if (p->feco_ratio < 0.f || p->feco_ratio > 1.f || (p->feco_ratio > 0.f && (fk < 1 || B < 2))) {
    sg_set_error("sg_pgd_run: FeCo needs 0 < feco_ratio <= 1, >= 1 cluster and a batch of >= 2 (ratio %g, frames %d, B %d)", p->feco_ratio, m, B);
    return SG_EINVAL;
  }
  XvWs w = xv_ws_layout(ws, B * Eb, m, h->Lp, h->L, h->S, true, N);
  SG_TRY(cmvn_scratch_reserve(h, B));
  const size_t dstride = (size_t)B * m * SG_WIN;
  // grad_sign of attack/utils.py:114 follows the requested loss NAME (SV / OSI with loss='Entropy' run the margin loss with
  // the cross-entropy sign), so the caller passes it; 0 derives it from the effective loss
  if (p->grad_sign != 0.f && p->grad_sign != 1.f && p->grad_sign != -1.f) { sg_set_error("sg_pgd_run: grad_sign must be +1, -1 or 0 (derive), got %g", p->grad_sign); return SG_EINVAL; }
  const float grad_sign = p->grad_sign != 0.f ? p->grad_sign : ((p->loss.loss == SG_LOSS_CE) ? (p->loss.targeted ? -1.f : 1.f) : -1.f);
  float* sc = scores ? scores : w.scores;
  long long* dec = decisions ? (long long*)decisions : w.dec;
  const size_t xbytes = (size_t)B * N * sizeof(float);

  // ---- graph replay: every iteration is the same kernel sequence on workspace addresses ------------------------------
  bool graphed = false;
  if (h->use_graph && !h->prof.on && !loss_hist && p->dither_mode != SG_DITHER_TENSOR && p->max_iter >= 2) {
    unsigned long long key[16] = {(unsigned long long)(uintptr_t)ws, (unsigned long long)B, (unsigned long long)N,
                                  (unsigned long long)E, (unsigned long long)p->dither_mode, 0, 0, 0, 0, 0, 0, 0,
                                  (unsigned long long)(uintptr_t)h->cm_part, (unsigned long long)Eb, 0, 0};
    memcpy(&key[14], &p->feco_ratio, sizeof(float)); memcpy(&key[15], &p->feco_tol, sizeof(float));
    key[15] |= (unsigned long long)(uint32_t)p->feco_max_iter << 32;
    static_assert(sizeof(sg_loss_params) == 24, "sg_loss_params fills key[7..9]");
    memcpy(&key[5], &p->epsilon, sizeof(float)); memcpy(&key[6], &p->step_size, sizeof(float));
    memcpy(&key[7], &p->loss, sizeof(sg_loss_params));
    memcpy(&key[10], &p->decision_threshold, sizeof(float));
    key[11] = (unsigned long long)(grad_sign > 0.f) | ((unsigned long long)h->precision << 1) | ((unsigned long long)h->pool_fusion << 3) |
              ((unsigned long long)h->feat_stash << 4) | ((unsigned long long)h->l1_tap_form << 5) | ((unsigned long long)h->cmvn_fusion << 6) | ((unsigned long long)h->row_compaction << 7) |
              ((unsigned long long)(uint32_t)h->utt_offset << 8);
    if (!h->pgd_graph.valid || memcmp(h->pgd_graph.key, key, sizeof(key)) != 0) {
      if (pgd_capture(h, B, N, m, p, grad_sign, w, key, st) != SG_OK) h->use_graph = 0;    // capture unavailable: launch-by-launch from now on
    }
    if (h->pgd_graph.valid && h->use_graph) {
      SG_CUDA_CHECK(cudaMemcpyAsync(w.xbuf, x_adv, xbytes, cudaMemcpyDeviceToDevice, st));
      SG_CUDA_CHECK(cudaMemcpyAsync(w.x0c, x0, xbytes, cudaMemcpyDeviceToDevice, st));
      SG_CUDA_CHECK(cudaMemcpyAsync(w.yc, y, (size_t)B * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      h->launches += 1;
      SG_TRY(sg_feat_ctl_init_launch(w.ctl, p->seed, 0u, st));
      for (int it = 0; it < p->max_iter; ++it) {
        SG_CUDA_CHECK(cudaGraphLaunch(h->pgd_graph.exec[E > 1 ? 0 : (it & 1)], st));
        h->launches += h->pgd_graph.kernels;                               // kernels inside one replayed iteration
      }
      float* cur = (E > 1 || (p->max_iter & 1) == 0) ? w.xbuf : w.xbuf2;
      // final evaluation pass (attack/FGSM.py:44-57 with iter == max_iter): pass counter max_iter * E from the control block
      sg_feat_set_ctl(w.ctl);
      int r = forward_pass(h, cur, B, N, m, p->dither_mode, nullptr, p->seed, 0, p->decision_threshold, w, w.emb, sc, dec, st, nullptr,
                           feco_spec(p, m, 0, w.ctl));
      sg_feat_set_ctl(nullptr);
      SG_TRY(r);
      h->launches += 1;
      PROF(h, SG_PROF_LOSS, st, sg_loss_launch(sc, (const long long*)y, B, h->S, p->loss, w.loss, nullptr, st));
      SG_CUDA_CHECK(cudaMemcpyAsync(x_adv, cur, xbytes, cudaMemcpyDeviceToDevice, st));
      graphed = true;
    }
  }
  if (graphed) return SG_OK;

  // ---- launch by launch -------------------------------------------------------------------------------------------------
  float* cur = x_adv;
  float* other = w.xbuf;
  for (int it = 0; it < p->max_iter; ++it) {
    SG_TRY(pgd_iteration(h, cur, other, x0, (const long long*)y, dither, B, N, m, p, grad_sign, w, sc, dec, loss_hist, it, nullptr, st));
    if (E == 1) { float* t = cur; cur = other; other = t; }
  }
  // final evaluation pass (attack/FGSM.py:44-57 with iter == max_iter)
  {
    const uint64_t pass = (uint64_t)p->max_iter * rounds;
    const float* dth = dither ? dither + (uint64_t)p->max_iter * E * dstride : nullptr;
    SG_TRY(forward_pass(h, cur, B, N, m, p->dither_mode, dth, p->seed, pass, p->decision_threshold, w, w.emb, sc, dec, st, nullptr,
                        feco_spec(p, m, pass, nullptr)));
    float* lossp = loss_hist ? loss_hist + (size_t)p->max_iter * B : w.loss;
    h->launches += 1;
    PROF(h, SG_PROF_LOSS, st, sg_loss_launch(sc, (const long long*)y, B, h->S, p->loss, lossp, nullptr, st));
  }
  if (cur != x_adv) SG_CUDA_CHECK(cudaMemcpyAsync(x_adv, cur, xbytes, cudaMemcpyDeviceToDevice, st));
  return SG_OK;
}
