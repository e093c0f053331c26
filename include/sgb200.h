/* libsgb200 -- C-ABI of the B200-native engine for SpeakerGuard's attack-iteration hot path.
 *
 * Every entry point replaces a piece of the reference's Python/PyTorch path (citations are
 * relative to the SpeakerGuard repository; "kaldi.py" = torchaudio/compliance/kaldi.py).
 * Conventions:
 *   - plain C types only; all tensor arguments are raw DEVICE pointers to fp32 / int64 data
 *     owned by the caller (PyTorch's caching allocator), except sg_load_* which take HOST pointers;
 *   - every function returns 0 (SG_OK) or a negative SG_E* code; the message is read with
 *     sg_last_error() (thread-local); no C++ exception crosses the boundary;
 *   - functions enqueue work on `stream` (a cudaStream_t passed as void*) and do not synchronise;
 *   - a handle is bound to one device, holds the packed weights/tables, is not re-entrant; every entry point
 *     that takes a handle makes that device current (cudaSetDevice) and leaves it current;
 *   - the library never allocates per-call memory: workspaces are sized by sg_*_ws_bytes() and
 *     passed in by the caller.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with SG_ECUDA.
 */
#ifndef SGB200_H_
#define SGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGB200_VERSION 200

typedef struct sg_handle sg_handle;
typedef void* sg_stream; /* cudaStream_t */

enum { SG_OK = 0, SG_EINVAL = -1, SG_ECUDA = -2, SG_ESTATE = -3, SG_EUNSUPPORTED = -4 };

/* arithmetic of the TDNN contractions (everything else is always fp32) */
enum { SG_PREC_FP32 = 0 /* FFMA, parity mode */, SG_PREC_TF32 = 1 /* tcgen05 kind::tf32 */,
       SG_PREC_BF16 = 2 /* tcgen05 kind::f16: TDNN activations, gradients and weights in bf16, fp32 accumulate */ };
/* dither of kaldi.py:179-181 (dither = 1.0 hard-coded at model/xv_plda.py:119) */
enum { SG_DITHER_OFF = 0, SG_DITHER_TENSOR = 1 /* caller supplies N(0,1) [B,m,400] */,
       SG_DITHER_PHILOX = 2 /* counter-based, regenerated identically in the adjoint */ };
enum { SG_LOSS_CE = 0, SG_LOSS_MARGIN = 1 };       /* attack/utils.py:7-29, :31-102 */
enum { SG_TASK_CSI = 0, SG_TASK_SV = 1, SG_TASK_OSI = 2 };

/* ---- lifecycle / diagnostics ---------------------------------------------------------------- */
int sg_version(void);
const char* sg_last_error(void);
int sg_create(sg_handle** out, int device);
void sg_destroy(sg_handle* h);
int sg_set_precision(sg_handle* h, int precision);
int sg_get_precision(const sg_handle* h);
/* Engine options (no reference counterpart).  SG_OPT_POOL_FUSION (default 1): in SG_PREC_BF16 the adjoint of the
 * statistics pooling (xvecTDNN.py:62) is applied to the staged activation tiles inside the layer-5 dgrad contraction
 * (packed bf16x2 arithmetic on the shared-memory tile) instead of a separate pass that writes and re-reads dA5.
 * Measured on B200 (B = 1024, 3 s): contraction +0.32 ms, pooling pass -0.41 ms per pass; 0 selects the two-kernel
 * form (the tests compare the two: feature gradients within 1e-2 of the per-utterance max, cosine > 0.9999). */
#define SG_OPT_POOL_FUSION 1
/* SG_OPT_FEAT_STASH (default 1): inside sg_pgd_run the MFCC forward (F1) leaves each frame's spectrum, windowed frame and
 * mel energies in a 4 KB workspace slot that the MFCC adjoint (F2) of the same pass reads back instead of recomputing the
 * forward FFT and regenerating the dither; 0 selects the recomputing adjoint (bit-identical results). */
#define SG_OPT_FEAT_STASH 2
/* SG_OPT_L1_TAP_FORM (default 1; measured: layer-1 dgrad 236 -> ~125 us at B = 1024 x 3 s): in SG_PREC_BF16 the input gradient of the first TDNN layer (xvecTDNN.py:16, 30 <- 512
 * channels, 5 taps) is computed as one K = 512 contraction into per-tap partial sums followed by the shifted sum over the
 * taps, instead of a K = 2560 contraction with 32 output columns; same arithmetic, fp32 summation order differs. */
#define SG_OPT_L1_TAP_FORM 3
/* SG_OPT_UTT_OFFSET (default 0): global index of the first utterance of the batches given to this handle.  The
 * SG_DITHER_PHILOX noise of sample j of frame f of utterance b in pass p is philox(seed; p, value + b, f, j), so when the
 * utterance axis is split contiguously over G handles (x[r*B/G:(r+1)*B/G] on GPU r) and handle r sets
 * value = r*B/G, the sharded attack reproduces the unsharded one bit for bit. */
#define SG_OPT_UTT_OFFSET 4
/* SG_OPT_CUDA_GRAPH (default 1): sg_pgd_run captures one PGD iteration (forward, backward, fused sign step: ~26 kernels) per
 * ping-pong parity into a CUDA graph and replays it max_iter times; seed and pass counter are read from a device control
 * block so the same graphs serve every later attack with the same shapes / workspace / parameters.  Iterates are
 * bit-identical to the launch-by-launch path (0).  Not used with a dither tensor, a loss history or profiling on. */
#define SG_OPT_CUDA_GRAPH 5
/* SG_OPT_CMVN_FUSION (default 0): in sg_pgd_run / sg_xv_forward, utterances of <= 300 frames (every CMVN window is the whole
 * utterance, model/iv_plda.py:321-337) get their CMVN inside the MFCC kernel (every CTA publishes the column sums of its 64
 * frames, the CTA that finishes the utterance last subtracts the mean from rows that are still in L2) and its adjoint inside
 * the MFCC adjoint (a prologue that forms the column means of d feat): two launches and the raw-feature round trip less per
 * pass, bit-identical to the separate sg_cmvn_fwd / sg_cmvn_bwd stages (same summation order).  Measured on B200 (PGD-100,
 * B = 1024, 3 s): MFCC forward +4.8 ms, adjoint +1.8 ms, CMVN launches -4.6 ms per step, i.e. 2 ms slower than the separate
 * kernels once those issue all their loads up front (10.8 -> 4.6 ms per step); a thread-block-cluster variant with a DSMEM
 * exchange was slower still (+19 ms: every CTA waits for its cluster at the barrier).  Kept as an option, off by default. */
#define SG_OPT_CMVN_FUSION 6
/* SG_OPT_ROW_COMPACTION (default 1, tensor-core precisions): layer 3 of the TDNN stores only the frames that are still valid
 * after its dilated taps (T - 30: 270 of 300 at 3 s), so the two 1 x 1 layers, the statistics pooling and their adjoints
 * contract 10 % fewer rows; layer 4's adjoint spreads its result back into layer 3's row space.  Values are unchanged (the
 * dropped frames were masked before). */
#define SG_OPT_ROW_COMPACTION 7
int sg_set_option(sg_handle* h, int option, int value);

/* ---- x-vector / PLDA system: weights ---------------------------------------------------------
 * Replaces the state built by model/xv_plda.py:17-48 (xvectorExtractor, PLDA, parse_mean_file,
 * parse_transform_mat_file, parse_enroll_model_file).  HOST pointers, fp32, PyTorch layouts:
 * conv weight [C_out][C_in][k] (model/_xv_plda/xvecTDNN.py:16-34), fc1 [512][3000],
 * lda [L][513] (offset in the last column, model/iv_plda.py:423-435), plda_transform [L][L]
 * (model/_xv_plda/plda.py:27-51), enroll [S][L].  BatchNorm running stats are folded into the
 * next layer (eval mode, affine=False). */
typedef struct {
  const float* tdnn_w[5];
  const float* tdnn_b[5];
  const float* bn_mean[5];
  const float* bn_var[5];
  const float* fc1_w;
  const float* fc1_b;
  const float* emb_mean;
  const float* lda;
  const float* plda_mean;
  const float* plda_transform;
  const float* plda_psi;
  const float* enroll;
  int L;
  int S;
  float bn_eps;
} sg_xv_weights;
int sg_load_xv(sg_handle* h, const sg_xv_weights* w);

/* ---- feature extraction ----------------------------------------------------------------------
 * sg_num_frames: kaldi.py:70 (snip_edges=False).
 * sg_mfcc_fwd: xv_plda.raw (model/xv_plda.py:107-156) = check_input_range x 2^15
 *   (model/utils.py:7-19) + torchaudio kaldi.mfcc (kaldi.py:669-813).  x [B,N] in [-1,1];
 *   raw [B,m,ld] (ld >= 30; columns 30..ld-1 are written as zero).
 * sg_mfcc_bwd: its adjoint (what autograd computes in adaptive_attack/EOT.py:35);
 *   grad [B,N] = (accumulate ? grad : 0) + scale * d(raw)/dx^T draw.
 * sg_dither_fill: materialises the SG_DITHER_PHILOX noise as [B,m,400] (tests / oracle). */
int sg_num_frames(int N);
int sg_mfcc_fwd(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                uint64_t seed, uint64_t pass, float* raw, int ld, sg_stream stream);
int sg_mfcc_bwd(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                uint64_t seed, uint64_t pass, const float* draw, int ld, float* grad, float scale,
                int accumulate, sg_stream stream);
int sg_dither_fill(sg_handle* h, int B, int N, uint64_t seed, uint64_t pass, float* out,
                   sg_stream stream);
/* iv_plda.cmvn (model/iv_plda.py:296-377): sliding mean-only CMVN, window 300, centred. */
int sg_cmvn_fwd(sg_handle* h, const float* raw, int ld_in, float* out, int ld_out, int B, int T,
                sg_stream stream);
int sg_cmvn_bwd(sg_handle* h, const float* dout, int ld_in, float* draw, int ld_out, int B, int T,
                sg_stream stream);

/* ---- TDNN + embedding head -------------------------------------------------------------------
 * sg_xv_embed_fwd: xv_plda.extract_emb (model/xv_plda.py:159-174) = xvecTDNN.embedding
 *   (xvecTDNN.py:46-64) + iv_plda.process_emb (model/iv_plda.py:411-443; length-norm with the
 *   norm detached, xvector_extract.py:33; PLDA transform plda.py:73-97).
 *   feat [B,T,32] (CMVN features, row stride 32) -> emb [B,L].  Activations stay in `ws`.
 * sg_xv_embed_bwd: demb [B,L] -> dfeat [B,T,32], using the activations left in `ws`. */
size_t sg_xv_ws_bytes(const sg_handle* h, int B, int T);
int sg_xv_embed_fwd(sg_handle* h, const float* feat, int B, int T, void* ws, float* emb,
                    sg_stream stream);
int sg_xv_embed_bwd(sg_handle* h, const float* demb, int B, int T, void* ws, float* dfeat,
                    sg_stream stream);

/* ---- scoring, decision, loss -----------------------------------------------------------------
 * sg_plda_score_fwd: iv_plda.scoring_trials -> PLDA.ComputeScores (model/iv_plda.py:399-408,
 *   plda.py:140-190) and the decision rule of model/defended_model.py:167-170.
 *   enroll == NULL uses the handle's enrolled set.  decisions may be NULL.
 * sg_loss_fwd_bwd: SEC4SR_CrossEntropy / SEC4SR_MarginLoss (attack/utils.py:7-102) and the
 *   gradient of sum(loss) wrt scores (loss.backward(ones), adaptive_attack/EOT.py:35). */
int sg_plda_score_fwd(sg_handle* h, const float* emb, int B, const float* enroll, int S,
                      float threshold, float* scores, int64_t* decisions, sg_stream stream);
int sg_plda_score_bwd(sg_handle* h, const float* emb, const float* dscores, int B,
                      const float* enroll, int S, float* demb, sg_stream stream);
typedef struct {
  int loss;         /* SG_LOSS_* */
  int task;         /* SG_TASK_* */
  int targeted;
  int clip_max;     /* SEC4SR_MarginLoss(clip_max) */
  float confidence;
  float threshold;  /* SV / OSI */
} sg_loss_params;
int sg_loss_fwd_bwd(sg_handle* h, const float* scores, const int64_t* y, int B, int S,
                    const sg_loss_params* lp, float* loss, float* dscores, sg_stream stream);

/* ---- the attack step and whole-attack loops --------------------------------------------------
 * sg_step_linf: attack/FGSM.py:62-68 with the bounds of attack/PGD.py:48-49 recomputed from
 *   x0: x <- min(max(x + step*sign(grad)*grad_sign, max(x0-eps,-1)), min(x0+eps,1)).
 *   eps = +inf gives FGSM's [-1,1] box (attack/FGSM.py:74-81).
 * sg_pgd_run: FGSM.attack_batch (attack/FGSM.py:38-70) with EOT.forward
 *   (adaptive_attack/EOT.py:16-54) inlined: max_iter gradient passes + the final evaluation
 *   pass, entirely on the device.  x_adv is in/out (caller initialises it with x0 or a random
 *   start, attack/PGD.py:58-61).  dither (SG_DITHER_TENSOR) is [(passes), B, m, 400] with
 *   passes = max_iter*eot_size + 1.  loss_hist (nullable) is [max_iter+1, B]. */
int sg_step_linf(sg_handle* h, float* x, const float* x0, const float* grad, size_t n, float step,
                 float grad_sign, float eps, sg_stream stream);
typedef struct {
  int max_iter;
  float epsilon;      /* bounds; +inf for FGSM */
  float step_size;
  int eot_size;       /* >= 1 gradient samples averaged per iteration */
  int dither_mode;
  uint64_t seed;
  sg_loss_params loss;
  float decision_threshold; /* model.threshold; -inf for CSI */
  float grad_sign;    /* +1 / -1: the sign resolve_loss returns (attack/utils.py:114).  It follows the loss NAME the caller
                         asked for, not the loss that is finally used: task SV / OSI with loss='Entropy' runs the margin
                         loss but keeps the cross-entropy sign (+1 untargeted, -1 targeted).  0 = derive it from `loss`
                         (CE: +1 / -1 by `targeted`; margin: -1), which is only right when name and loss agree. */
  /* EOT samples as batch rows (adaptive_attack/EOT.py:30-42: x_batch.repeat(EOT_batch_size, 1, 1)): eot_batch copies of the
   * batch run through one pass as B * eot_batch rows, row e * B + b = copy e of utterance b, each with its own dither and its
   * own FeCo clustering; eot_size / eot_batch passes per iteration.  0 or 1 = one copy per pass.  The workspace must then be
   * sized by sg_pgd_ws_bytes(h, B * eot_batch, N).  Needs dither_mode OFF or PHILOX. */
  int eot_batch;
  /* FeCo feature compression between the raw MFCC and CMVN (model/defended_model.py:46-65 with defense = [[1, FeCo]],
   * defense/feature_level.py:18-50): k = (int)(frames * feco_ratio) cluster means per utterance, clustering re-drawn for
   * every pass (the randomness EOT averages over).  feco_ratio 0 = no defense.  Needs >= 2 rows per pass (the reference
   * drops empty clusters for a batch of one, which changes the frame count). */
  float feco_ratio;
  int feco_max_iter;  /* Lloyd iteration cap (kmeans_ids default: 100) */
  float feco_tol;     /* stop when <= feco_tol * frames change cluster (libKMCUDA default: 0.01) */
} sg_pgd_params;
size_t sg_pgd_ws_bytes(const sg_handle* h, int B, int N);
int sg_pgd_run(sg_handle* h, float* x_adv, const float* x0, const int64_t* y, const float* dither,
               int B, int N, const sg_pgd_params* p, void* ws, int64_t* decisions, float* scores,
               float* loss_hist, sg_stream stream);
/* one forward pass wav -> scores/decisions through the same fused path (model.make_decision) */
int sg_xv_forward(sg_handle* h, const float* x, int B, int N, int dither_mode, const float* dither,
                  uint64_t seed, uint64_t pass, float decision_threshold, void* ws, float* scores,
                  int64_t* decisions, float* emb, sg_stream stream);

/* ---- AudioNet (log-mel CNN, CSI-NE) and CW2 ------------------------------------------------------
 * sg_load_audionet: weights of audionet_csine (model/audionet_csine.py:58-121), HOST pointers in
 *   PyTorch layouts: conv1_w [5][5] (Conv2d 1->1), conv_w[l] [C_out][C_in][3] for conv2..conv8,
 *   BatchNorm (eval, affine) running stats / gamma / beta with index 0 = conv1's BatchNorm2d(1),
 *   fc_w [num_class][32].  BatchNorm is folded into the convolutions.
 * sg_audionet_logmel_fwd/bwd: Preprocessor.forward (model/_audionet/Preprocessor.py:85-112) and its
 *   adjoint; x [B,N] in [-1,1] -> feat [B,T,32], T = sg_audionet_num_frames(N).
 * sg_audionet_cnn_fwd/bwd: extract_emb + fc (audionet_csine.py:176-224); logits have row stride
 *   sg_audionet_num_class_padded() (classes padded to a multiple of 16 with zero weights).
 * sg_argmax_decide: make_decision (audionet_csine.py:243-257).
 * sg_cw2_audionet_run: CW2.attack_batch (attack/CW2.py:41-132) entirely on the device: tanh-space
 *   iterate, clipped margin loss, L2 term, Adam (lr, betas 0.9/0.999, eps 1e-8), best-example
 *   tracking, per-utterance binary search on c; the only host syncs are the early-stop checks every
 *   stop_early_iter iterations (one 4-byte read per check; at BASELINE's stop_early_iter = max_iter = 1000 that is two per
 *   search step, which is why the check stays on the host).  best_x [B,N] out, success [B] (0/1) out.
 * Precision: sg_set_precision(h, SG_PREC_FP32) runs the CNN's convolutions and their adjoints as fp32 FFMA (parity mode);
 *   TF32 / BF16 run them on the tensor cores with TF32 operands and fp32 storage (conv_tc_kernel's utterance-tiled mode:
 *   3-D TMA maps (channel, frame, utterance), frames outside the utterance zero-fill = the 'same' padding of
 *   audionet_csine.py:66-118); the log-mel front end, pooling, fc layer and CW2 arithmetic are fp32 in every mode.
 *   Inside sg_cw2_audionet_run the log-mel forward hands its spectrum and mel energies to the adjoint (4.2 KB per frame in
 *   the workspace) instead of the adjoint repeating the FFT. */
typedef struct {
  const float* conv1_w;
  const float* conv1_b;
  const float* conv_w[7];
  const float* conv_b[7];
  const float* bn_mean[8];
  const float* bn_var[8];
  const float* bn_gamma[8];
  const float* bn_beta[8];
  const float* fc_w;
  const float* fc_b;
  int num_class;
  float bn_eps;
} sg_audionet_weights;
typedef struct {
  int binary_search_steps;
  int max_iter;
  int stop_early;
  int stop_early_iter;
  float lr;
  float initial_const;
  sg_loss_params loss;        /* SG_LOSS_MARGIN with clip_max = 1 */
  float decision_threshold;   /* -inf for CSI */
} sg_cw2_params;
int sg_load_audionet(sg_handle* h, const sg_audionet_weights* w);
int sg_audionet_num_frames(int N);
int sg_audionet_num_class_padded(const sg_handle* h);
size_t sg_audionet_ws_bytes(const sg_handle* h, int B, int N, int for_cw2);
int sg_audionet_logmel_fwd(sg_handle* h, const float* x, int B, int N, float* feat, sg_stream stream);
int sg_audionet_logmel_bwd(sg_handle* h, const float* x, int B, int N, const float* dfeat, void* ws,
                           float* dx, float scale, int accumulate, sg_stream stream);
int sg_audionet_cnn_fwd(sg_handle* h, const float* feat, int B, int N, void* ws, float* logits,
                        sg_stream stream);
int sg_audionet_cnn_bwd(sg_handle* h, const float* dlogits, int B, int N, void* ws, float* dfeat,
                        sg_stream stream);
/* The same CNN split at the embedding (forward(..., return_emb=True), embedding(), predict_from_embeddings:
 * audionet_csine.py:176-229): sg_audionet_emb_fwd/bwd = extract_emb, feat [B,T,32] <-> emb [B,32] (max over time of conv8);
 * sg_audionet_fc_fwd/bwd = the final fc, emb [B,32] <-> logits [B,Cp]. */
int sg_audionet_emb_fwd(sg_handle* h, const float* feat, int B, int N, void* ws, float* emb, sg_stream stream);
int sg_audionet_emb_bwd(sg_handle* h, const float* demb, int B, int N, void* ws, float* dfeat, sg_stream stream);
int sg_audionet_fc_fwd(sg_handle* h, const float* emb, int B, float* logits, sg_stream stream);
int sg_audionet_fc_bwd(sg_handle* h, const float* dlogits, int B, float* demb, sg_stream stream);
int sg_argmax_decide(sg_handle* h, const float* scores, int B, int S, int ld, float threshold,
                     int64_t* decisions, sg_stream stream);
int sg_cw2_audionet_run(sg_handle* h, const float* x, const int64_t* y, int B, int N,
                        const sg_cw2_params* p, void* ws, float* best_x, int64_t* success,
                        float* final_const, sg_stream stream);

/* gradient iterations executed by the last sg_cw2_audionet_run on this handle, summed over the binary-search steps (early
 * stop, attack/CW2.py:96-100, can end a search step before max_iter) */
long long sg_cw2_last_iterations(const sg_handle* h);

/* ---- FeCo feature compression (replaces libKMCUDA) ----------------------------------------------
 * sg_feco_kmeans: per-utterance Lloyd k-means over the frames (defense/feature_level.py:192-193:
 *   kmeans_cuda(x, k, yinyang_t=0, metric='L2'); k-means++ seeding, stops when <= tol*n frames
 *   change cluster, libKMCUDA's default tol 0.01).  feat [B,n,ld] -> ids [B,n] int32 in [0,k).
 * sg_feco_means_fwd/bwd: the differentiable cluster means built from the ids (:202-217);
 *   empty cluster i -> feat[i] when force.  out [B,k,dim], counts [B,k], dfeat [B,n,dim]. */
int sg_feco_kmeans(sg_handle* h, const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed,
                   int max_iter, float tol, int32_t* ids, sg_stream stream);
/* the same with the random stream of row r keyed as inside sg_pgd_run: rows are `copy_rows` utterances repeated (row r = copy
 * r / copy_rows of utterance r % copy_rows; 0: rows are utterances), utterance indices are offset by `utt_offset` (a shard of
 * a larger batch), and `pass` selects the clustering of that pass of an attack */
int sg_feco_kmeans_keyed(sg_handle* h, const float* feat, int ld, int B, int n, int dim, int k, uint64_t seed,
                         int max_iter, float tol, int32_t* ids, uint32_t pass, uint32_t utt_offset, uint32_t copy_rows,
                         sg_stream stream);
int sg_feco_means_fwd(sg_handle* h, const float* feat, int ld, const int32_t* ids, int B, int n, int dim,
                      int k, int force, float* out, int32_t* counts, sg_stream stream);
int sg_feco_means_bwd(sg_handle* h, const float* dout, const int32_t* ids, const int32_t* counts, int B,
                      int n, int dim, int k, int force, float* dfeat, sg_stream stream);

/* ---- i-vector system (BASELINE config 5) --------------------------------------------------------
 * iv_plda (model/iv_plda.py:17-153): 24 MFCC (the first 24 columns of sg_mfcc_fwd) -> add_delta
 * (:248-293) -> sliding CMVN over the 72 columns (:296-377) -> full-covariance UBM posteriors and
 * Baum-Welch statistics (model/_iv_plda/gmm.py:120-171) -> i-vector (model/_iv_plda/
 * ivector_extract.py:94-114) -> process_emb (:411-443) -> the PLDA back-end shared with xv_plda
 * (sg_plda_score_fwd / sg_plda_score_bwd, sg_loss_fwd_bwd).
 * sg_iv_weights: every matrix dense row-major fp32 (the reference's packed Kaldi text matrices
 * are unpacked by the host class):
 *   gmm_gconsts [C], gmm_means_invcovars [C,F], gmm_invcovars [C,F,F]      (gmm.py:73-118)
 *   ive_T [C,F,D], ive_sigma_inv [C,F,F], ive_offset                       (ivector_extract.py:25-92)
 *   emb_mean [D], lda [L,D+1] (offset in the last column), plda_* and enroll [S,L] as sg_xv_weights.
 * C must be a multiple of 16.  Handle precision SG_PREC_FP32: every contraction of this path is fp32 FFMA.
 * SG_PREC_TF32 (the host class's 'tf32x3'): the UBM log-likelihood contraction, its adjoint and the L = N x U assembly run
 * on tcgen05 as split-TF32 (operands split hi + lo, three products per algorithmic product, fp32 accumulate: posteriors
 * within 5e-5 of fp64); the remaining small GEMMs stay FFMA.  The per-utterance SPD solve is always an fp64 Cholesky.
 * sg_iv_embed_fwd: feat [B,T,ld] CMVN'd features -> emb [B,L] (extract_emb, model/iv_plda.py:380-396).
 * sg_iv_embed_bwd: adjoint, using what the forward left in `ws` (sg_iv_ws_bytes(h,B,T) bytes);
 *   dfeat [B,T,ld], columns >= F written as zero.
 * sg_iv_stage_read: intermediate results of the last forward on `ws` (tests, diagnostics).
 * sg_add_delta_fwd/bwd: order-2, window-3 deltas with replicated edges, [B,T,F] <-> [B,T,3F].
 * sg_cmvn_cols: sg_cmvn_fwd/bwd (backward != 0: adjoint) for any number of columns. */
typedef struct {
  int C, F, D, L, S;
  const float* gmm_gconsts;
  const float* gmm_means_invcovars;
  const float* gmm_invcovars;
  const float* ive_T;
  const float* ive_sigma_inv;
  float ive_offset;
  const float* emb_mean;
  const float* lda;
  const float* plda_mean;
  const float* plda_transform;
  const float* plda_psi;
  const float* enroll;
} sg_iv_weights;
enum { SG_IV_STAGE_POST = 0, SG_IV_STAGE_STATS = 1, SG_IV_STAGE_IVECTOR = 2 };
int sg_load_iv(sg_handle* h, const sg_iv_weights* w);
size_t sg_iv_ws_bytes(const sg_handle* h, int B, int T);
int sg_iv_embed_fwd(sg_handle* h, const float* feat, int ld, int B, int T, void* ws, float* emb,
                    sg_stream stream);
int sg_iv_embed_bwd(sg_handle* h, const float* demb, int B, int T, void* ws, float* dfeat, int ld,
                    sg_stream stream);
int sg_iv_stage_read(sg_handle* h, const void* ws, int B, int T, int stage, float* out,
                     sg_stream stream);
int sg_add_delta_fwd(sg_handle* h, const float* in, int ld_in, float* out, int ld_out, int B, int T,
                     int F, sg_stream stream);
int sg_add_delta_bwd(sg_handle* h, const float* dout, int ld_in, float* din, int ld_out, int B, int T,
                     int F, sg_stream stream);
int sg_cmvn_cols(sg_handle* h, const float* in, int ld_in, float* out, int ld_out, int ncol, int B,
                 int T, int backward, sg_stream stream);

/* ---- AudioNet in training mode -----------------------------------------------
 * What `outputs = model(x_batch); loss.backward(); optimizer.step()` runs for an audionet_csine in
 * train() mode (adver_train.py:183-221, natural_train.py:127-160; model/audionet_csine.py:58-121):
 * BatchNorm with batch statistics and the momentum update of the running ones, the input gradient through
 * those statistics (the attack inside the training loop runs on the train-mode model) and the gradient of
 * every parameter.  sg_load_audionet must have been called (front-end tables, class count).
 * sg_audionet_train_tensors: DEVICE pointers in PyTorch's own layouts - conv1_w [1,1,5,5], conv1_b [1],
 *   conv_w[l] [C_out,C_in,3], conv_b[l] [C_out] (conv2..conv8), bn_gamma/bn_beta/bn_mean/bn_var[0..7]
 *   ([1] for the BatchNorm2d(1) of the pre-filter, then [C_out]), fc_w [C,32], fc_b [C].  As parameters:
 *   the live values (bn_mean / bn_var are the running statistics, updated in place when momentum > 0;
 *   they may be null when momentum == 0).  As gradients: where to write each gradient (bn_mean / bn_var unused).
 * sg_audionet_train_fwd: feat [B,T,32] log-mel (sg_audionet_logmel_fwd) -> logits [B,Cp]; eps <= 0 -> 1e-5.
 * sg_audionet_train_bwd: dlogits [B,Cp] -> dfeat [B,T,32] (may be null) and parameter gradients g (may be
 *   null: input gradient only, the attack's case), using what the forward left in ws
 *   (sg_audionet_train_ws_bytes(h,B,N) bytes).
 * sg_adam_step: torch.optim.Adam's update (no amsgrad; weight_decay added to the gradient) on one flat
 *   tensor; step counts from 1. */
typedef struct {
  float* conv1_w; float* conv1_b;
  float* conv_w[7]; float* conv_b[7];
  float* bn_gamma[8]; float* bn_beta[8];
  float* bn_mean[8]; float* bn_var[8];
  float* fc_w; float* fc_b;
} sg_audionet_train_tensors;
size_t sg_audionet_train_ws_bytes(const sg_handle* h, int B, int N);
int sg_audionet_train_fwd(sg_handle* h, const sg_audionet_train_tensors* p, const float* feat, int B, int N,
                          float momentum, float eps, void* ws, float* logits, sg_stream stream);
int sg_audionet_train_bwd(sg_handle* h, const sg_audionet_train_tensors* p, const float* feat,
                          const float* dlogits, int B, int N, void* ws, float* dfeat,
                          const sg_audionet_train_tensors* g, sg_stream stream);
int sg_adam_step(sg_handle* h, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int step, sg_stream stream);

/* ---- caller I/O -------------------------------------------------------------
 * sg_pcm16_quantize: save_audio's conversion (attackMain.py:154-160) for a whole batch on the device:
 *   per utterance, if 0.9*max <= 1 and 0.9*min >= -1 multiply by 2^15; then numpy's astype(int16)
 *   (truncate toward zero through int32, keep the low 16 bits).  adv [B,N] float -> pcm [B,N] int16 (device);
 *   scaled [B] (optional, device) receives 1 where the 2^15 scale was applied.
 * sg_wav_write_batch: scipy.io.wavfile.write (attackMain.py:166) for B mono PCM16 files from one host
 *   buffer pcm [B,N], by `nthreads` host threads (0 = all cores); parent directories are created
 *   (attackMain.py:161-164).  Files are byte-identical to scipy's.
 * sg_wav_read_batch: Dataset.__getitem__ (dataset/Dataset.py:72-84) for B files into one host buffer
 *   out [B,wav_length]: first channel, crop at starts[i] (or centred when starts is null / negative) or
 *   zero-pad at the end; normalize != 0 -> [-1,1) scale, else int16 range; lens [B] (optional) = frames per file.
 * The two host functions do no device work and need no handle. */
int sg_pcm16_quantize(sg_handle* h, const float* adv, int B, int N, int16_t* pcm, int32_t* scaled,
                      sg_stream stream);
int sg_wav_write_batch(const char* const* paths, const int16_t* pcm, int B, int N, int fs, int nthreads);
int sg_wav_read_batch(const char* const* paths, int B, int wav_length, const int64_t* starts,
                      int normalize, float* out, int32_t* lens, int nthreads);

/* ---- the path's one collective ----------------------------------------------------
 * Utterances shard independently over the GPUs (no data-path collective); what is reduced over NVLink / NVSwitch is a
 * handful of metric scalars at the end of an attack (success count, sums of SNR / L2 / Linf, utterance count: reference
 * metric/metric.py:10-42 semantics).  NCCL is bound with dlopen at run time (inside PyTorch: the copy torch loaded).
 * sg_comm_unique_id: rank 0 fills 128 bytes (ncclUniqueId) and ships them to the other ranks by any host channel.
 * sg_comm_init: one communicator per handle (= per GPU / process).
 * sg_allreduce_metrics: in-place sum over the ranks of a DEVICE fp64 vector, asynchronous on `stream`. */
int sg_comm_unique_id(void* id128);
int sg_comm_init(sg_handle* h, const void* id128, int rank, int world);
int sg_allreduce_metrics(sg_handle* h, double* v, int n, sg_stream stream);
int sg_comm_destroy(sg_handle* h);
int sg_comm_nccl_version(void);

/* ---- test hook: one conv-as-GEMM launch on either arithmetic path -----------------------------
 * out[p,n] = epi(sum_{tap,c} A[p + tap*tap_step, c] * W[tap*cin + c, n]); W is [taps*cin, N]
 * (FFMA path), Wk its K-major copy [N, taps*cin] (tcgen05 path); epilogue 0 bias, 1 bias+ReLU,
 * 2 ReLU-mask (mask > 0 and row %% T < t_valid), 3 none.  op_bf16: A and Wk hold bf16 (tcgen05
 * kind::f16); out_bf16: out holds bf16.  This is the kernel behind the TDNN layers
 * (xvecTDNN.py:16-53) and their dgrad. */
int sg_debug_conv(sg_handle* h, int precision, const float* A, int lda, const float* W, const float* Wk,
                  const float* bias, float* out, int ldo, int rows, int N, int cin, int taps,
                  int tap_step, int epilogue, const float* mask, int ldmask, int T, int t_valid,
                  int op_bf16, int out_bf16, sg_stream stream);

/* ---- device-side timing ------------------------------------------------------------------------
 * With profiling enabled every kernel launch is bracketed by CUDA events on its stream and
 * attributed to a category; sg_profile_read() synchronises on the recorded events and returns the
 * summed duration and the launch count since sg_profile_enable() was last called.  bench.py
 * uses this for the per-kernel share of a step and for the roofline of the dominant kernel. */
enum { SG_PROF_MFCC_FWD = 0, SG_PROF_MFCC_BWD, SG_PROF_CMVN, SG_PROF_TDNN_FWD, SG_PROF_TDNN_BWD,
       SG_PROF_POOL, SG_PROF_HEAD_GEMM, SG_PROF_HEAD, SG_PROF_LOSS, SG_PROF_STEP, SG_PROF_AUDIONET,
       SG_PROF_CW2, SG_PROF_IV_GEMM, SG_PROF_IV,
       SG_PROF_TDNN_BWD_POOL /* layer-5 dgrad with the pooling adjoint fused in (SG_OPT_POOL_FUSION) */,
       SG_PROF_FECO /* FeCo k-means, cluster means and their adjoint inside sg_pgd_run */, SG_PROF_COUNT };
int sg_profile_enable(sg_handle* h, int enable);
int sg_profile_read(sg_handle* h, int category, double* total_ms, long long* launches);
const char* sg_profile_name(int category);
/* every launch recorded since sg_profile_enable(), in launch order: category, tag (TDNN layer 1..5 for the contraction
 * launches, else 0) and duration in ms; at most `capacity` entries are written, *count receives the number recorded. */
int sg_profile_dump(sg_handle* h, int* categories, int* tags, float* ms, int capacity, int* count);

/* kernel-launch counter (bench.py's gpu_launches): kernels launched by this library since the
 * last sg_reset_launch_count() on the calling thread's handle. */
long long sg_launch_count(const sg_handle* h);
void sg_reset_launch_count(sg_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SGB200_H_ */
