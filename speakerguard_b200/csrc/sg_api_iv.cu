// C-ABI of the i-vector system (reference model/iv_plda.py: add_delta :248-293, extract_emb :380-396,
// process_emb :411-443; model/_iv_plda/gmm.py:120-171; model/_iv_plda/ivector_extract.py:94-114).
//
// Per pass (B utterances, T frames, padded to Tp = 16-multiple rows each, R = B*Tp):
//   Xa  [R, Fa]      features + a ones column (the zeroth-order statistic rides along with the first-order ones)
//   Q   [R, Kq]      packed quadratic expansion  -> post = softmax(Q Wq + gconst)           [R, C]
//   FsT [B, Fa, C]   = Xa_b^T post_b   (rows 0..F-1: F_c^T, row F: N_c)                      batched GEMM
//   Lpk [B, Pp]      = N U             (packed upper triangle of sum_c N_c T_c' S_c^-1 T_c)
//   lin [B, Dp]      = vec(F) Wlin     (sum_c T_c' S_c^-1 F_c)
//   iv  = (I + L)^-1 (lin + offset e0) - offset e0 - emb_mean ; e2 = LDA iv ; emb = PLDA transform (shared head)
// The backward pass replays the chain with the transposed operands; the SPD solve keeps its fp64 factor.
#include <math.h>
#include <string.h>

#include <thread>
#include <vector>

#include "sg_handle.cuh"
#include "sg_head.cuh"
#include "sg_iv.cuh"

#define IV_SPLITS 32

struct SgIv {
  int C, F, D, L, Lp, Fa, Dp, P, Pp, Kq;
  float offset;
  float *Wq, *WqT, *gconst;        // [Kq, C], [C, Kq], [C]
  uint32_t* qidx;                  // [Kq]: factor pair of every column of the quadratic expansion (i | j << 16; see quad_expand_fwd_kernel)
  float *Wq3K, *WqT3K;             // 3xTF32 operands (K-major): [C, 3Kq] = [hi|lo|hi](WqT), [Kq, 3C] = [hi|lo|hi](Wq)
  float *U;                        // [C, Pp]: packed upper triangles of U_c = T_c' S_c^-1 T_c
  float *UT3;                      // 3xTF32 operand of the L assembly, K-major [Pp, 3C] = [hi|lo|hi](U^T); built on first use
  float *Tt;                       // [Dp, F*C]: Tt[d, f*C + c] = T_c[f, d]   (adjoint of the L assembly)
  float* WlinT3;                   // 3xTF32 operand of the linear term computed transposed, K-major [Dp, 3 F C] = [hi | lo | hi](Wlin^T); built on first use
  float *Wlin3K, *Tt3K;            // 3xTF32 operands of the two [B, Dp] x [Dp, F*C] adjoint contractions, K-major [F*C, K3p]; built on first use
  int K3p;                         // 3 Dp rounded up to the tensor core's k-block
  float *Wlin, *WlinT;             // [F*C, Dp], [Dp, F*C]
  float *Wlda, *Wlda_b, *blda;     // [Dp, Lp], [Lp, Dp], [Lp]
  float* emb_mean;                 // [Dp]
};

void sg_iv_free(sg_handle* h) {
  delete h->iv;                    // device buffers are owned by h->allocs
  h->iv = nullptr;
}

static inline int up16(int v) { return (v + 15) / 16 * 16; }

template <typename Fn>
static void parallel_for(int n, Fn fn) {
  int nt = (int)std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 32 ? 32 : nt);
  if (nt > n) nt = n;
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([=]() { for (int i = t; i < n; i += nt) fn(i); });
  for (auto& x : th) x.join();
}

extern "C" int sg_load_iv(sg_handle* h, const sg_iv_weights* w) {
  if (!h || !w) { sg_set_error("sg_load_iv: null argument"); return SG_EINVAL; }
  if (h->iv || h->backend_loaded) { sg_set_error("sg_load_iv: a model is already loaded on this handle"); return SG_ESTATE; }
  if (w->C < 16 || w->C % 16 != 0 || w->F < 1 || w->F > 128 || w->D < 1 || w->D > 800 || w->L < 1 || w->L > 512 || w->S < 1) {
    sg_set_error("sg_load_iv: need C %% 16 == 0, 1 <= F <= 128, 1 <= D <= 800, 1 <= L <= 512, S >= 1 (C=%d F=%d D=%d L=%d S=%d)",
                 w->C, w->F, w->D, w->L, w->S);
    return SG_EINVAL;
  }
  if (!w->gmm_gconsts || !w->gmm_means_invcovars || !w->gmm_invcovars || !w->ive_T || !w->ive_sigma_inv || !w->emb_mean ||
      !w->lda || !w->plda_mean || !w->plda_transform || !w->plda_psi || !w->enroll) {
    sg_set_error("sg_load_iv: null weight pointer"); return SG_EINVAL;
  }
  SG_CUDA_CHECK(cudaSetDevice(h->device));
  SgIv* m = new SgIv();
  memset(m, 0, sizeof(*m));
  h->iv = m;
  const int C = w->C, F = w->F, D = w->D, L = w->L;
  m->C = C; m->F = F; m->D = D; m->L = L; m->Lp = (L + 31) / 32 * 32;
  m->Fa = up16(F + 1); m->Dp = up16(D); m->P = D * (D + 1) / 2; m->Pp = (m->P + 511) / 512 * 512;   // 512: split-K friendly
  m->Kq = (F + F * (F + 1) / 2 + 255) / 256 * 256;   // 256: whole N-tiles for the tensor-core dgrad
  m->offset = w->ive_offset;
  const int Kq = m->Kq, Pp = m->Pp, Dp = m->Dp, Lp = m->Lp;

  {  // UBM: ll = gconst + (S^-1 mu).x - 1/2 x' S^-1 x on the packed expansion (gmm.py:120-131)
    std::vector<float> Wq((size_t)Kq * C, 0.f), WqT((size_t)C * Kq, 0.f);
    for (int c = 0; c < C; ++c) {
      const float* P = w->gmm_invcovars + (size_t)c * F * F;
      for (int f = 0; f < F; ++f) Wq[(size_t)f * C + c] = w->gmm_means_invcovars[(size_t)c * F + f];
      int p = F;
      for (int i = 0; i < F; ++i)
        for (int j = i; j < F; ++j, ++p)
          Wq[(size_t)p * C + c] = (i == j) ? -0.5f * P[i * F + i] : -0.5f * (P[i * F + j] + P[j * F + i]);
      for (int k = 0; k < Kq; ++k) WqT[(size_t)c * Kq + k] = Wq[(size_t)k * C + c];
    }
    SG_TRY(sg_dev_upload(h, &m->Wq, Wq));
    SG_TRY(sg_dev_upload(h, &m->WqT, WqT));
    {  // split operands: x*w ~ x_lo*w_hi + x_hi*w_lo + x_hi*w_hi (small terms first), the hi factors exactly representable in tf32;
      // activations are laid out [lo|hi|hi] along K, weights [hi|lo|hi]
      auto hi = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
      std::vector<float> A3((size_t)C * 3 * Kq), B3((size_t)Kq * 3 * C);
      for (int c = 0; c < C; ++c)
        for (int k = 0; k < Kq; ++k) {
          const float v = WqT[(size_t)c * Kq + k], vh = hi(v);
          float* r = &A3[(size_t)c * 3 * Kq];
          r[k] = vh; r[Kq + k] = v - vh; r[2 * Kq + k] = vh;
          float* q = &B3[(size_t)k * 3 * C];
          q[c] = vh; q[C + c] = v - vh; q[2 * C + c] = vh;
        }
      SG_TRY(sg_dev_upload(h, &m->Wq3K, A3));
      SG_TRY(sg_dev_upload(h, &m->WqT3K, B3));
    }
    {
      std::vector<uint32_t> qi((size_t)Kq, 0xffffu);                 // padding columns
      for (int f = 0; f < F; ++f) qi[f] = (uint32_t)f | (0xffffu << 16);
      int k = F;
      for (int i = 0; i < F; ++i)
        for (int j = i; j < F; ++j) qi[k++] = (uint32_t)i | ((uint32_t)j << 16);
      std::vector<float> qf((size_t)Kq);
      memcpy(qf.data(), qi.data(), (size_t)Kq * sizeof(float));
      float* dq = nullptr;
      SG_TRY(sg_dev_upload(h, &dq, qf));
      m->qidx = reinterpret_cast<uint32_t*>(dq);
    }
    SG_TRY(sg_dev_upload(h, &m->gconst, std::vector<float>(w->gmm_gconsts, w->gmm_gconsts + C)));
  }
  {  // extractor (ivector_extract.py:94-107): G_c = T_c' S_c^-1 [D,F], U_c = G_c T_c [D,D]
    std::vector<float> U((size_t)C * Pp, 0.f), Wlin((size_t)F * C * Dp, 0.f);
    parallel_for(C, [&](int c) {
      const float* T = w->ive_T + (size_t)c * F * D;          // [F, D]
      const float* S = w->ive_sigma_inv + (size_t)c * F * F;  // [F, F]
      std::vector<float> G((size_t)D * F, 0.f);
      for (int g = 0; g < F; ++g)
        for (int d = 0; d < D; ++d) {
          const float t = T[(size_t)g * D + d];
          float* gr = &G[(size_t)d * F];
          const float* sr = S + (size_t)g * F;
          for (int f = 0; f < F; ++f) gr[f] += t * sr[f];
        }
      for (int f = 0; f < F; ++f)
        for (int d = 0; d < D; ++d) Wlin[((size_t)f * C + c) * Dp + d] = G[(size_t)d * F + f];
      float* u = &U[(size_t)c * Pp];
      for (int i = 0; i < D; ++i) {
        float* ur = u + ((size_t)i * D - (size_t)i * (i - 1) / 2) - i;   // ur[j] = packed (i, j)
        for (int f = 0; f < F; ++f) {
          const float g = G[(size_t)i * F + f];
          const float* tr = T + (size_t)f * D;
          for (int j = i; j < D; ++j) ur[j] += g * tr[j];
        }
      }
    });
    SG_TRY(sg_dev_upload(h, &m->U, U));
    SG_TRY(sg_dev_upload(h, &m->Wlin, Wlin));
    {
      std::vector<float> Tt((size_t)Dp * F * C, 0.f);
      parallel_for(C, [&](int c) {
        const float* T = w->ive_T + (size_t)c * F * D;
        for (int f = 0; f < F; ++f)
          for (int d = 0; d < D; ++d) Tt[(size_t)d * F * C + (size_t)f * C + c] = T[(size_t)f * D + d];
      });
      SG_TRY(sg_dev_upload(h, &m->Tt, Tt));
    }
    {
      std::vector<float> WlinT((size_t)Dp * F * C, 0.f);
      const size_t K = (size_t)F * C;
      parallel_for(Dp, [&](int d) { for (size_t k = 0; k < K; ++k) WlinT[(size_t)d * K + k] = Wlin[k * Dp + d]; });
      SG_TRY(sg_dev_upload(h, &m->WlinT, WlinT));
    }
  }
  {  // LDA [L, D+1], offset in the last column (model/iv_plda.py:423-435)
    std::vector<float> Wl((size_t)Dp * Lp, 0.f), Wlb((size_t)Lp * Dp, 0.f), b(Lp, 0.f), mean(Dp, 0.f);
    for (int i = 0; i < L; ++i) {
      for (int d = 0; d < D; ++d) {
        const float v = w->lda[(size_t)i * (D + 1) + d];
        Wl[(size_t)d * Lp + i] = v;
        Wlb[(size_t)i * Dp + d] = v;
      }
      b[i] = w->lda[(size_t)i * (D + 1) + D];
    }
    for (int d = 0; d < D; ++d) mean[d] = w->emb_mean[d];
    SG_TRY(sg_dev_upload(h, &m->Wlda, Wl));
    SG_TRY(sg_dev_upload(h, &m->Wlda_b, Wlb));
    SG_TRY(sg_dev_upload(h, &m->blda, b));
    SG_TRY(sg_dev_upload(h, &m->emb_mean, mean));
  }
  SG_TRY(sg_load_backend(h, w->plda_mean, w->plda_transform, w->plda_psi, w->enroll, L, w->S));
  return SG_OK;
}

// ---------------------------------------------------------------------------------------------
struct IvWs {
  float *Xa, *XaT, *Q, *post, *dpost, *FsT, *dFsT, *dFs, *Lpk, *lin, *dlin, *wfull, *iv, *div, *e2, *de2, *tsave, *scal, *dXa, *part, *dll3, *N3, *d3;
  double* fac;
  size_t bytes;
};
// split-K scratch: the SIMT skinny contractions (IV_SPLITS slices of [B, max(C, Dp)]) or the tensor-core linear term
// (32 slices of [Dp rounded to 128, B rounded to 32] plus its transposed result)
static size_t iv_part_floats(const SgIv* m, int B) {
  const size_t simt = (size_t)IV_SPLITS * B * (m->C > m->Dp ? m->C : m->Dp);
  const size_t Bp = ((size_t)B + 31) / 32 * 32, mrows = ((size_t)m->Dp + 127) / 128 * 128;
  const size_t tc = 32 * mrows * Bp + (size_t)m->Dp * Bp;
  return simt > tc ? simt : tc;
}
static IvWs iv_ws_layout(void* base, const SgIv* m, int B, int T) {
  IvWs w;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t nfloat) { float* q = (float*)(p + off); off += (nfloat * sizeof(float) + 255) / 256 * 256; return q; };
  const size_t Tp = up16(T), R = (size_t)B * Tp;
  w.Xa = take(R * m->Fa); w.XaT = take(R * m->Fa); w.dXa = take(R * m->Fa);
  w.Q = take(R * m->Kq * 3);                            // [R, Kq], or [R, 3Kq] split operand in tensor-core mode
  w.dll3 = take(R * m->C * 3);
  w.post = take(R * m->C); w.dpost = take(R * m->C);
  w.FsT = take((size_t)B * m->Fa * m->C); w.dFsT = take((size_t)B * m->Fa * m->C); w.dFs = take((size_t)B * m->Fa * m->C);
  w.Lpk = take((size_t)B * m->Pp); w.N3 = take((size_t)B * 3 * m->C);
  w.d3 = take((size_t)B * ((3 * m->Dp + 31) / 32 * 32));
  w.lin = take((size_t)B * m->Dp); w.dlin = take((size_t)B * m->Dp);
  w.wfull = take((size_t)B * m->Dp); w.iv = take((size_t)B * m->Dp); w.div = take((size_t)B * m->Dp);
  w.e2 = take((size_t)B * m->Lp); w.de2 = take((size_t)B * m->Lp); w.tsave = take((size_t)B * m->Lp);
  w.scal = take((size_t)B * 4);
  w.fac = (double*)take((size_t)B * m->D * m->D * 2);
  w.part = take(iv_part_floats(m, B));
  w.bytes = off;
  return w;
}

static int check_iv(sg_handle* h) {
  SG_TRY(sg_check_handle(h, false));
  if (!h->iv) { sg_set_error("i-vector weights not loaded (call sg_load_iv first)"); return SG_ESTATE; }
  return SG_OK;
}

extern "C" size_t sg_iv_ws_bytes(const sg_handle* h, int B, int T) {
  if (!h || !h->iv || B < 1 || T < 1) return 0;
  return iv_ws_layout(nullptr, h->iv, B, T).bytes;
}

static SgConvArgs gemm_args(const float* A, int lda, const float* W, const float* Wk, const float* bias, float* out, int ldo,
                            int rows, int N, int K) {
  SgConvArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.lda = lda; a.W = W; a.Wk = Wk; a.bias = bias; a.out = out; a.ldo = ldo; a.rows = rows; a.N = N; a.cin = K;
  a.taps = 1; a.T = 1; a.epilogue = bias ? SG_EPI_BIAS : SG_EPI_NONE;
  return a;
}
// fp32 FFMA for the contractions of this path.  The UBM log-likelihoods cancel O(100) terms against each other, so plain
// tf32 / bf16 operands are not offered; with a tensor-core precision set on the handle the two large contractions (frames x
// components) run as 3xTF32 instead: both operands split into tf32-exact hi + lo parts, three tcgen05 products, fp32 accumulate
static int iv_gemm_tc3(sg_handle* h, const float* A3, int K3, const float* W3k, const float* bias, float* out, int ldo,
                       int rows, int N, cudaStream_t st) {
  SgConvArgs a;
  memset(&a, 0, sizeof(a));
  a.A = A3; a.lda = K3; a.Wk = W3k; a.bias = bias; a.out = out; a.ldo = ldo; a.rows = rows; a.N = N; a.cin = K3;
  a.taps = 1; a.T = 1; a.epilogue = bias ? SG_EPI_BIAS : SG_EPI_NONE;
  h->launches += 1;
  PROF(h, SG_PROF_IV_GEMM, st, sg_conv_tc(a, SG_PREC_TF32, st));
  return SG_OK;
}
// the 2 GB split operand of the L assembly is only needed in tensor-core mode: built from U on the device at first use
static int iv_build_ut3(sg_handle* h, cudaStream_t st) {
  SgIv* m = h->iv;
  if (m->UT3) return SG_OK;
  SG_CUDA_CHECK(cudaMalloc((void**)&m->UT3, (size_t)m->Pp * 3 * m->C * sizeof(float)));
  h->allocs.push_back(m->UT3);
  return sg_build_ut3_launch(m->U, m->UT3, m->C, m->Pp, st);
}
// ~0.7 GB each at C = 2048, D = 400: only built when a tensor-core precision is set and the adjoint runs
static int iv_build_w3(sg_handle* h, cudaStream_t st) {
  SgIv* m = h->iv;
  if (m->Wlin3K) return SG_OK;
  const size_t N = (size_t)m->F * m->C;
  m->K3p = (3 * m->Dp + 31) / 32 * 32;
  SG_CUDA_CHECK(cudaMalloc((void**)&m->Wlin3K, N * m->K3p * sizeof(float)));
  h->allocs.push_back(m->Wlin3K);
  SG_CUDA_CHECK(cudaMalloc((void**)&m->Tt3K, N * m->K3p * sizeof(float)));
  h->allocs.push_back(m->Tt3K);
  SG_TRY(sg_build_w3_launch(m->Wlin, (size_t)m->Dp, 1, m->Wlin3K, (int)N, m->Dp, m->K3p, st));     // Wlin [F*C, Dp]
  SG_TRY(sg_build_w3_launch(m->Tt, 1, N, m->Tt3K, (int)N, m->Dp, m->K3p, st));                     // Tt [Dp, F*C]
  return SG_OK;
}
static int iv_build_wlint3(sg_handle* h, cudaStream_t st) {
  SgIv* m = h->iv;
  if (m->WlinT3) return SG_OK;
  const size_t K = (size_t)m->F * m->C;
  SG_CUDA_CHECK(cudaMalloc((void**)&m->WlinT3, (size_t)m->Dp * 3 * K * sizeof(float)));
  h->allocs.push_back(m->WlinT3);
  return sg_build_w3_launch(m->WlinT, K, 1, m->WlinT3, m->Dp, (int)K, (int)(3 * K), st);              // WlinT [Dp, F*C]
}
static int iv_gemm(sg_handle* h, const SgConvArgs& a, cudaStream_t st) {
  h->launches += 1;
  PROF(h, SG_PROF_IV_GEMM, st, sg_conv_simt(a, st));
  return SG_OK;
}
// skinny contraction (rows = B utterances, long K): split K across IV_SPLITS CTAs per tile, then add the partial sums
static int iv_gemm_splitk(sg_handle* h, SgConvArgs a, float* part, cudaStream_t st) {
  int splits = IV_SPLITS;
  while (splits > 1 && (a.cin % (16 * splits) != 0)) splits >>= 1;
  if (splits == 1) return iv_gemm(h, a, st);
  float* out = a.out;
  const int ldo = a.ldo;
  const int kc = a.cin / splits;
  a.cin = kc; a.nbatch = splits; a.strideA = kc; a.strideW = (long long)kc * a.N; a.strideO = (long long)a.rows * a.N;
  a.out = part; a.ldo = a.N; a.bias = nullptr; a.epilogue = SG_EPI_NONE;
  SG_TRY(iv_gemm(h, a, st));
  h->launches += 1;
  PROF(h, SG_PROF_IV_GEMM, st, sg_splitk_reduce_launch(part, splits, a.rows, a.N, out, ldo, st));
  return SG_OK;
}
#define IV_K(call) do { h->launches += 1; PROF(h, SG_PROF_IV, st, (call)); } while (0)

static int iv_embed_fwd(sg_handle* h, const float* feat, int ld, int B, int T, const IvWs& w, float* emb, cudaStream_t st) {
  const SgIv* m = h->iv;
  const int Tp = up16(T), R = B * Tp, C = m->C, F = m->F, Fa = m->Fa;
  IV_K(sg_pad_aug_launch(feat, ld, w.Xa, Fa, B, T, Tp, F, st));
  IV_K(sg_transpose_batched_launch(w.Xa, w.XaT, Tp, Fa, Fa, Tp, (size_t)Tp * Fa, (size_t)Tp * Fa, B, st));
  const bool tc = h->precision != SG_PREC_FP32;
  if (tc) {
    IV_K(sg_quad_expand_launch(w.Xa, Fa, w.Q, 3 * m->Kq, R, F, 1, m->Kq, m->qidx, st));
    SG_TRY(iv_gemm_tc3(h, w.Q, 3 * m->Kq, m->Wq3K, m->gconst, w.post, C, R, C, st));
  } else {
    IV_K(sg_quad_expand_launch(w.Xa, Fa, w.Q, m->Kq, R, F, 0, 0, nullptr, st));
    SG_TRY(iv_gemm(h, gemm_args(w.Q, m->Kq, m->Wq, m->WqT, m->gconst, w.post, C, R, C, m->Kq), st));
  }
  IV_K(sg_softmax_rows_launch(w.post, nullptr, w.post, R, C, T, Tp, 0, 0, st));
  {  // Baum-Welch statistics (gmm.py:166-171), one GEMM per utterance: [Fa, Tp] x [Tp, C]
    SgConvArgs a = gemm_args(w.XaT, Tp, w.post, nullptr, nullptr, w.FsT, C, Fa, C, Tp);
    a.nbatch = B; a.strideA = (long long)Fa * Tp; a.strideW = (long long)Tp * C; a.strideO = (long long)Fa * C;
    SG_TRY(iv_gemm(h, a, st));
  }
  if (tc) {      // L assembly N x U as 3xTF32: [B, 3C] x [3C, Pp]
    SG_TRY(iv_build_ut3(h, st));
    IV_K(sg_split3_rows_launch(w.FsT + (size_t)F * C, Fa * C, w.N3, B, C, st));
    SG_TRY(iv_gemm_tc3(h, w.N3, 3 * C, m->UT3, nullptr, w.Lpk, m->Pp, B, m->Pp, st));
  } else {
    SG_TRY(iv_gemm(h, gemm_args(w.FsT + (size_t)F * C, Fa * C, m->U, nullptr, nullptr, w.Lpk, m->Pp, B, m->Pp, C), st));
  }
  {
    // linear term vec(F) Wlin.  Tensor-core form: computed transposed, lin' [Dp, B] = Wlin' [Dp, 3 F C] x vec(F)' - the weights
    // are the row operand (4 row tiles at D = 400), the B utterances the 32-padded column operand, K split 32 ways (128 CTAs
    // stream the 0.7 GB split weights once); both split operands of the pass live in the expansion buffer, dead by now.
    const size_t K = (size_t)F * C;
    const int Bp = (B + 31) / 32 * 32, mrows = (m->Dp + 127) / 128 * 128;
    const size_t need = (size_t)32 * mrows * Bp, part_floats = iv_part_floats(m, B);
    if (tc && K % 32 == 0 && Bp <= 512 && need + (size_t)m->Dp * Bp <= part_floats && (size_t)Bp * 3 * K <= (size_t)R * m->Kq * 3) {
      SG_TRY(iv_build_wlint3(h, st));
      float* F3 = w.Q;
      float* linT = w.part + need;
      if (Bp > B) SG_CUDA_CHECK(cudaMemsetAsync(F3 + (size_t)B * 3 * K, 0, (size_t)(Bp - B) * 3 * K * sizeof(float), st));
      IV_K(sg_split3_rows_ld_launch(w.FsT, Fa * C, F3, (int)(3 * K), B, (int)K, st));
      SgConvArgs a;
      memset(&a, 0, sizeof(a));
      a.A = m->WlinT3; a.lda = (int)(3 * K); a.Wk = F3; a.out = linT; a.ldo = Bp; a.rows = m->Dp; a.N = Bp; a.cin = (int)(3 * K);
      a.taps = 1; a.T = 1; a.epilogue = SG_EPI_NONE; a.splitk_ws = w.part; a.splitk_floats = need;
      h->launches += 2;
      PROF(h, SG_PROF_IV_GEMM, st, sg_conv_tc(a, SG_PREC_TF32, st));
      IV_K(sg_transpose_batched_launch(linT, w.lin, m->Dp, B, Bp, m->Dp, 0, 0, 1, st));
    } else {
      SG_TRY(iv_gemm_splitk(h, gemm_args(w.FsT, Fa * C, m->Wlin, m->WlinT, nullptr, w.lin, m->Dp, B, m->Dp, F * C), w.part, st));
    }
  }
  IV_K(sg_chol_solve_launch(w.Lpk, m->Pp, w.lin, m->Dp, m->offset, m->emb_mean, w.fac, w.wfull, w.iv, B, m->D, st));
  SG_TRY(iv_gemm(h, gemm_args(w.iv, m->Dp, m->Wlda, m->Wlda_b, m->blda, w.e2, m->Lp, B, m->Lp, m->Dp), st));
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, st, sg_head_fwd_launch(h->H, w.e2, B, w.tsave, w.scal, emb, st));
  return SG_OK;
}

static int iv_embed_bwd(sg_handle* h, const float* demb, int B, int T, const IvWs& w, float* dfeat, int ld, cudaStream_t st) {
  const SgIv* m = h->iv;
  const int Tp = up16(T), R = B * Tp, C = m->C, F = m->F, Fa = m->Fa;
  h->launches += 1;
  PROF(h, SG_PROF_HEAD, st, sg_head_bwd_launch(h->H, demb, B, w.tsave, w.scal, w.de2, st));
  SG_TRY(iv_gemm(h, gemm_args(w.de2, m->Lp, m->Wlda_b, m->Wlda, nullptr, w.div, m->Dp, B, m->Dp, m->Lp), st));
  IV_K(sg_chol_solve_bwd_launch(w.fac, w.wfull, w.div, m->Dp, w.dlin, nullptr, m->Pp, B, m->D, st));
  SG_CUDA_CHECK(cudaMemsetAsync(w.dFsT, 0, (size_t)B * Fa * C * sizeof(float), st));
  // dL = -lambda w' contracts with U_c = G_c T_c without forming it: dN_c = -(G_c' lambda) . (T_c w) = -dF_c . (T_c w)
  if (h->precision != SG_PREC_FP32 && (F * C) % 32 == 0) {      // both as 3xTF32: [B, K3p] x [K3p, F*C]
    SG_TRY(iv_build_w3(h, st));
    IV_K(sg_split3_rows_ld_launch(w.dlin, m->Dp, w.d3, m->K3p, B, m->Dp, st));
    SG_TRY(iv_gemm_tc3(h, w.d3, m->K3p, m->Wlin3K, nullptr, w.dFsT, Fa * C, B, F * C, st));
    IV_K(sg_split3_rows_ld_launch(w.wfull, m->Dp, w.d3, m->K3p, B, m->Dp, st));
    SG_TRY(iv_gemm_tc3(h, w.d3, m->K3p, m->Tt3K, nullptr, w.dFs, Fa * C, B, F * C, st));
  } else {
    SG_TRY(iv_gemm(h, gemm_args(w.dlin, m->Dp, m->WlinT, m->Wlin, nullptr, w.dFsT, Fa * C, B, F * C, m->Dp), st));
    SG_TRY(iv_gemm(h, gemm_args(w.wfull, m->Dp, m->Tt, nullptr, nullptr, w.dFs, Fa * C, B, F * C, m->Dp), st));
  }
  IV_K(sg_dn_from_df_launch(w.dFsT, w.dFs, B, F, Fa, C, st));
  // dFs_b = dFsT_b^T  [C, Fa]: the K-major operand of d post, and the first-order statistics' direct path to x further down
  IV_K(sg_transpose_batched_launch(w.dFsT, w.dFs, Fa, C, C, Fa, (size_t)Fa * C, (size_t)Fa * C, B, st));
  const int K3a = (3 * Fa + 31) / 32 * 32;
  if (h->precision != SG_PREC_FP32 && C % 32 == 0 && (size_t)(R + B * C) * K3a <= (size_t)R * m->Kq * 3) {
    // d post_b = Xa_b dFs_b' as a batched 3xTF32 contraction on tcgen05: one [C, 3 Fa] operand per utterance, tiles of 128
    // frames of one utterance.  Both split operands live in the (dead at this point) buffer of the quadratic expansion.
    float* X3 = w.Q;
    float* W3 = w.Q + (size_t)R * K3a;
    IV_K(sg_split3_rows_ld_launch(w.Xa, Fa, X3, K3a, R, Fa, st));
    IV_K(sg_build_w3_launch(w.dFs, (size_t)Fa, 1, W3, B * C, Fa, K3a, st));
    SgConvArgs a;
    memset(&a, 0, sizeof(a));
    a.A = X3; a.lda = K3a; a.Wk = W3; a.out = w.dpost; a.ldo = C; a.rows = R; a.N = C; a.cin = K3a;
    a.taps = 1; a.T = Tp; a.same_utt = 1; a.w_per_utt = 1; a.epilogue = SG_EPI_NONE;
    h->launches += 1;
    PROF(h, SG_PROF_IV_GEMM, st, sg_conv_tc(a, SG_PREC_TF32, st));
  } else
  {  // d post_b = Xa_b dFsT_b : [Tp, Fa] x [Fa, C]
    SgConvArgs a = gemm_args(w.Xa, Fa, w.dFsT, nullptr, nullptr, w.dpost, C, Tp, C, Fa);
    a.nbatch = B; a.strideA = (long long)Tp * Fa; a.strideW = (long long)Fa * C; a.strideO = (long long)Tp * C;
    SG_TRY(iv_gemm(h, a, st));
  }
  if (h->precision != SG_PREC_FP32) {
    IV_K(sg_softmax_rows_launch(w.post, w.dpost, w.dll3, R, C, T, Tp, 1, 1, st));
    SG_TRY(iv_gemm_tc3(h, w.dll3, 3 * C, m->WqT3K, nullptr, w.Q, m->Kq, R, m->Kq, st));
  } else {
    IV_K(sg_softmax_rows_launch(w.post, w.dpost, w.dpost, R, C, T, Tp, 1, 0, st));
    SG_TRY(iv_gemm(h, gemm_args(w.dpost, C, m->WqT, m->Wq, nullptr, w.Q, m->Kq, R, m->Kq, C), st));
  }
  // the first-order statistics also depend on x directly: dXa_b = post_b dFs_b
  {
    SgConvArgs a = gemm_args(w.post, C, w.dFs, nullptr, nullptr, w.dXa, Fa, Tp, Fa, C);
    a.nbatch = B; a.strideA = (long long)Tp * C; a.strideW = (long long)C * Fa; a.strideO = (long long)Tp * Fa;
    SG_TRY(iv_gemm(h, a, st));
  }
  IV_K(sg_quad_expand_bwd_launch(w.Q, m->Kq, w.Xa, Fa, w.dXa, dfeat, ld, B, T, Tp, F, m->qidx, st));
  return SG_OK;
}

extern "C" int sg_iv_embed_fwd(sg_handle* h, const float* feat, int ld, int B, int T, void* ws, float* emb, sg_stream stream) {
  SG_TRY(check_iv(h));
  if (!feat || !ws || !emb || B < 1 || T < 1 || ld < h->iv->F) { sg_set_error("sg_iv_embed_fwd: bad argument"); return SG_EINVAL; }
  return iv_embed_fwd(h, feat, ld, B, T, iv_ws_layout(ws, h->iv, B, T), emb, (cudaStream_t)stream);
}
extern "C" int sg_iv_embed_bwd(sg_handle* h, const float* demb, int B, int T, void* ws, float* dfeat, int ld, sg_stream stream) {
  SG_TRY(check_iv(h));
  if (!demb || !ws || !dfeat || B < 1 || T < 1 || ld < h->iv->F) { sg_set_error("sg_iv_embed_bwd: bad argument"); return SG_EINVAL; }
  return iv_embed_bwd(h, demb, B, T, iv_ws_layout(ws, h->iv, B, T), dfeat, ld, (cudaStream_t)stream);
}

// intermediate results of the last sg_iv_embed_fwd on this workspace (tests / diagnostics)
extern "C" int sg_iv_stage_read(sg_handle* h, const void* ws, int B, int T, int stage, float* out, sg_stream stream) {
  SG_TRY(check_iv(h));
  if (!ws || !out || B < 1 || T < 1) { sg_set_error("sg_iv_stage_read: bad argument"); return SG_EINVAL; }
  const SgIv* m = h->iv;
  const IvWs w = iv_ws_layout((void*)ws, m, B, T);
  const int Tp = up16(T);
  cudaStream_t st = (cudaStream_t)stream;
  switch (stage) {
    case SG_IV_STAGE_POST:     // [B, T, C]
      SG_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)T * m->C * 4, w.post, (size_t)Tp * m->C * 4, (size_t)T * m->C * 4, B, cudaMemcpyDeviceToDevice, st));
      break;
    case SG_IV_STAGE_STATS:    // [B, F+1, C]: first-order statistics transposed, zeroth-order in the last row
      SG_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)(m->F + 1) * m->C * 4, w.FsT, (size_t)m->Fa * m->C * 4, (size_t)(m->F + 1) * m->C * 4, B, cudaMemcpyDeviceToDevice, st));
      break;
    case SG_IV_STAGE_IVECTOR:  // [B, D]: solution of the linear system (prior offset still in coordinate 0, mean not removed)
      SG_CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)m->D * 4, w.wfull, (size_t)m->Dp * 4, (size_t)m->D * 4, B, cudaMemcpyDeviceToDevice, st));
      break;
    default:
      sg_set_error("sg_iv_stage_read: unknown stage %d", stage);
      return SG_EINVAL;
  }
  return SG_OK;
}

// ---- feature-level pieces ----------------------------------------------------------------------
extern "C" int sg_add_delta_fwd(sg_handle* h, const float* in, int ld_in, float* out, int ld_out, int B, int T, int F, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!in || !out || B < 1 || T < 1 || F < 1 || ld_in < F || ld_out < 3 * F) { sg_set_error("sg_add_delta_fwd: bad argument"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  IV_K(sg_delta_launch(in, ld_in, out, ld_out, B, T, F, 0, st));
  return SG_OK;
}
extern "C" int sg_add_delta_bwd(sg_handle* h, const float* dout, int ld_in, float* din, int ld_out, int B, int T, int F, sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!dout || !din || B < 1 || T < 1 || F < 1 || ld_in < 3 * F || ld_out < F) { sg_set_error("sg_add_delta_bwd: bad argument"); return SG_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  IV_K(sg_delta_launch(dout, ld_in, din, ld_out, B, T, F, 1, st));
  return SG_OK;
}
extern "C" int sg_cmvn_cols(sg_handle* h, const float* in, int ld_in, float* out, int ld_out, int ncol, int B, int T, int backward,
                            sg_stream stream) {
  SG_TRY(sg_check_handle(h, false));
  if (!in || !out || B < 1 || T < 1 || ncol < 1 || ld_in < ncol || ld_out < ncol) { sg_set_error("sg_cmvn_cols: bad argument"); return SG_EINVAL; }
  h->launches += 1;
  PROF(h, SG_PROF_CMVN, (cudaStream_t)stream, sg_cmvn_cols_launch(in, ld_in, out, ld_out, B, T, ncol, backward ? 1 : 0, (cudaStream_t)stream));
  return SG_OK;
}
