// Internal declarations shared by sg_head.cu / sg_feat.cu / sg_api.cu.
#pragma once
#include "sg_common.cuh"

struct SgHeadConst {           // device pointers + scalars describing the PLDA back-end
  int L, Lp;                   // embedding dim and its padding to a multiple of 16
  const float* plda_mean;      // [L]
  const float* plda_T;         // [L][L]   transform (row i = output i)
  const float* plda_Tt;        // [L][L]   transposed copy (coalesced mat-vec in the forward)
  const float* inv_psi1;       // 1 / (psi + 1)
  const float* psi_ratio;      // psi / (psi + 1)
  const float* inv_var_given;  // 1 / (1 + psi / (psi + 1))
  float logdet_given, logdet_without, log2pi_L;   // plda.py:176-186 constants (fp32 like the reference)
};

int sg_feat_init();
int sg_feat_fwd_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode, const float* dither,
                       uint64_t seed, uint64_t pass, float* raw, int ld, cudaStream_t st, float* stash = nullptr, int cmvn = 0,
                       float* cmvn_part = nullptr, unsigned int* cmvn_count = nullptr);
size_t sg_feat_cmvn_part_floats(int B);                        // fused-CMVN scratch: this many floats + B zeroed counters
int sg_feat_cmvn_fusable(int m);                               // m <= 300 frames: CMVN can run inside the MFCC kernels (cmvn = 1 below)
int sg_feat_bwd_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode, const float* dither,
                       uint64_t seed, uint64_t pass, const float* draw, int ld, float* grad, float scale,
                       int accumulate, cudaStream_t st, const float* stash = nullptr, int cmvn = 0);
int sg_feat_bwd_step_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode,
                            const float* dither, uint64_t seed, uint64_t pass, const float* draw, int ld,
                            const float* x0, float* x_out, float step, float eps, cudaStream_t st, const float* stash = nullptr,
                            int cmvn = 0);
void sg_feat_set_copy_rows(int rows);                          // EOT copies as batch rows: utterances per copy for the following launches (0: off)
void sg_feat_set_ctl(const uint32_t* ctl);                     // device {pass, seed_lo, seed_hi} for the following launches (null: immediates)
int sg_feat_ctl_init_launch(uint32_t* ctl, uint64_t seed, uint32_t pass, cudaStream_t st);
int sg_feat_ctl_tick_launch(uint32_t* ctl, uint32_t n, cudaStream_t st);
size_t sg_feat_stash_floats(int B, int m);   // per-frame forward stash consumed by the adjoint (attack loop only)
int sg_dither_fill_launch(int B, int m, uint64_t seed, uint64_t pass, float* out, cudaStream_t st);
int sg_tap_gather_launch(const float* G, int ldg, float* out, int ldo, size_t rows, int taps, int dil, cudaStream_t st);
int sg_step_linf_launch(float* x, const float* x0, const float* grad, size_t n, float step, float eps, cudaStream_t st);
int sg_tile_rows_launch(const float* in, float* out, size_t row_floats, int B, int copies, const long long* yin, long long* yout,
                        cudaStream_t st);
int sg_reduce_rows_launch(const float* rows, float* acc, size_t row_floats, int B, int copies, int accumulate, cudaStream_t st);
// FeCo (sg_kmeans.cu); ctl: device {pass, seed_lo, seed_hi} mixed into the k-means seed (graph replay), or null
int sg_kmeans_init();
int sg_feco_kmeans_launch(const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed, int max_iter, float tol,
                          int* ids, cudaStream_t st, const uint32_t* ctl, uint32_t pass = 0, uint32_t b_off = 0, uint32_t copy_rows = 0);
int sg_feco_means_fwd_launch(const float* feat, int ld_in, const int* ids, int B, int n, int dim, int k, int force,
                             float* out, int ld_out, int* counts, cudaStream_t st);
int sg_feco_means_bwd_launch(const float* dout, int ld_out, const int* ids, const int* counts, int B, int n, int dim, int k,
                             int force, float* dfeat, int ld_in, cudaStream_t st);
int sg_cmvn_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int backward, cudaStream_t st);

int sg_conv_tc_warm();                                         // per-device one-time setup of the tcgen05 kernels (outside any capture)
int sg_pool_fwd_launch(const void* r5, int bf16, int B, int T, int Tv, const float* bn_mean, const float* bn_istd,
                       float* stats, float* save_mean, float* save_std, cudaStream_t st);
int sg_pool_bwd_launch(const void* r5, int bf16, int B, int T, int Tv, const float* bn_istd, const float* dstats,
                       const float* save_mean, const float* save_std, void* dA5, cudaStream_t st);
int sg_pool_bwd_params_launch(int B, int Tv, const float* bn_istd, const float* dstats, const float* save_mean,
                              const float* save_std, float* ab, cudaStream_t st);
int sg_head_fwd_launch(const SgHeadConst& H, const float* e2, int B, float* tsave, float* scal, float* emb, cudaStream_t st);
int sg_head_bwd_launch(const SgHeadConst& H, const float* dq, int B, const float* tsave, const float* scal, float* de2, cudaStream_t st);
int sg_score_fwd_launch(const SgHeadConst& H, const float* emb, int B, const float* enroll, int S, float threshold,
                        float* scores, long long* decisions, cudaStream_t st);
int sg_score_bwd_launch(const SgHeadConst& H, const float* emb, const float* dscores, int B, const float* enroll, int S,
                        float* demb, cudaStream_t st);
int sg_loss_launch(const float* scores, const long long* y, int B, int S, const sg_loss_params& lp, float* loss,
                   float* dscores, cudaStream_t st);
