// Internal declarations for the AudioNet path (sg_audionet.cu / sg_api_audionet.cu).
#pragma once
#include "sg_common.cuh"

#define AN_NFFT 1024
#define AN_HOP 160
#define AN_WIN 800
#define AN_WOFF 112                 // (1024 - 800) / 2: the window is centred in the FFT frame
#define AN_MELS 32
#define AN_BINS 513
#define AN_MELW 1280              // packed mel weights, float4 groups (zero-padded)
#define SG_CW2_CHUNKS 16

struct alignas(16) SgAnTables {
  float window[AN_WIN];
  float2 twA[8][2][32];
  float2 twB[8][2][32];
  float2 twU[16][32];
  int mel_lo[32], mel_len[32], mel_off[32];
  float mel_w[AN_MELW];
  int bin_c0[512], bin_c1[512];
  float bin_w0[512], bin_w1[512];
  int mel_maxlen;
};

int sg_an_tables_build(SgAnTables* host_out);
int sg_an_init();
#define AN_STASH_FLOATS 1056             // per frame: spectrum X[16][32] float2 (as float4 pairs) + mel energies [32]
int sg_an_logmel_fwd_launch(const SgAnTables* dT, const float* x, int B, int N, int T, float* feat, cudaStream_t st,
                            float* stash = nullptr);
int sg_an_logmel_bwd_launch(const SgAnTables* dT, const float* x, int B, int N, int T, const float* dfeat, float* dgw,
                            float* dx, float scale, int accumulate, cudaStream_t st, const float* stash = nullptr);
int sg_maxpool2_fwd_launch(const float* in, float* out, int B, int T, int C, cudaStream_t st);
int sg_maxpool2_bwd_launch(const float* in, const float* dout, float* din, int B, int T, int C, cudaStream_t st);
int sg_globalmax_fwd_launch(const float* in, float* out, int* arg, int B, int T, int Tv, int C, cudaStream_t st);
int sg_globalmax_bwd_launch(const float* in, const float* dout, const int* arg, float* din, int B, int T, int C, cudaStream_t st);
int sg_argmax_rows_launch(const float* s, long long* dec, int B, int S, float threshold, cudaStream_t st);
int sg_cw2_prepare_launch(const float* x, const float* w, float* inp, float* l2part, float* loss2, int B, int N, cudaStream_t st);
int sg_cw2_adam_launch(float* w, float* m, float* v, const float* x, const float* inp, const float* gmodel, const float* cst,
                       int B, int N, float lr, int step, cudaStream_t st);
int sg_cw2_track_launch(const float* inp, float* best_x, const float* loss1, const float* loss2, const long long* dec,
                        float* best_l2, long long* best_score, float* gbest_l2, long long* gbest_score, int B, int N, cudaStream_t st);
int sg_cw2_search_update_launch(float* cst, float* lower, float* upper, const long long* best_score, int B, cudaStream_t st);

// conv stack after the 5x5 pre-filter: (C_in, C_out, pad, pool)   audionet_csine.py:66-118
static const int kAnCin[7] = {32, 64, 128, 128, 128, 128, 64};
static const int kAnCout[7] = {64, 128, 128, 128, 128, 64, 32};
static const int kAnPad[7] = {1, 1, 1, 1, 1, 1, 0};
static const int kAnPool[7] = {1, 0, 0, 1, 0, 1, 0};

struct SgAudioNet {
  SgAnTables* d_tables = nullptr;
  int C = 0, Cp = 0;                    // classes, padded to a multiple of 16
  float *W1 = nullptr, *W1b = nullptr, *b1 = nullptr;         // banded 5x5 pre-filter as a 5-tap 32->32 conv
  float *W[7] = {}, *Wb[7] = {}, *bias[7] = {};               // [3*cin, cout], [3*cout, cin], [cout]
  float *Wfc = nullptr, *Wfcb = nullptr, *bfc = nullptr;      // [32, Cp], [Cp, 32], [Cp]
  // K-major copies for the tensor-core path (tf32 / bf16 precision modes): [cout, taps*cin] and [cin, taps*cout]
  float *W1k = nullptr, *W1bk = nullptr, *Wk[7] = {}, *Wbk[7] = {};
};

