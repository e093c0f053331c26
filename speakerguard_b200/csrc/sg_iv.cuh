// Internal declarations for the i-vector path (sg_iv.cu / sg_api_iv.cu).
#pragma once
#include "sg_common.cuh"

int sg_delta_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int F, int backward, cudaStream_t st);
int sg_cmvn_cols_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int ncol, int backward, cudaStream_t st);
int sg_pad_aug_launch(const float* feat, int ld, float* xa, int Fa, int B, int T, int Tp, int F, cudaStream_t st);
int sg_quad_expand_launch(const float* xa, int ldx, float* q, int ldq, int rows, int F, int split3, int kseg, const uint32_t* idx,
                          cudaStream_t st);
int sg_quad_expand_bwd_launch(const float* dq, int ldq, const float* xa, int ldx, const float* add, float* dx, int lddx,
                              int B, int T, int Tp, int F, const uint32_t* idx, cudaStream_t st);
int sg_softmax_rows_launch(const float* a, const float* b, float* out, int rows, int C, int T, int Tp, int backward,
                           int split3, cudaStream_t st);
int sg_chol_solve_launch(const float* Lp, int ldp, const float* rhs, int ldr, float offset, const float* emb_mean, double* fac,
                         float* wfull, float* iv, int B, int D, cudaStream_t st);
int sg_chol_solve_bwd_launch(const double* fac, const float* w, const float* dw, int ldr, float* drhs, float* dLp, int ldp,
                             int B, int D, cudaStream_t st);
int sg_transpose_batched_launch(const float* in, float* out, int R, int C, int ld_in, int ld_out, size_t stride_in,
                                size_t stride_out, int nbatch, cudaStream_t st);
int sg_splitk_reduce_launch(const float* part, int splits, int rows, int N, float* out, int ldo, cudaStream_t st);
int sg_split3_rows_launch(const float* in, int ld, float* out, int rows, int C, cudaStream_t st);
int sg_build_ut3_launch(const float* U, float* UT3, int C, int Pp, cudaStream_t st);
int sg_split3_rows_ld_launch(const float* in, int ld, float* out, int ldo, int rows, int C, cudaStream_t st);
int sg_build_w3_launch(const float* src, size_t sn, size_t sd, float* W3, int N, int K, int K3p, cudaStream_t st);
int sg_dn_from_df_launch(float* dFsT, const float* a, int B, int F, int Fa, int C, cudaStream_t st);
