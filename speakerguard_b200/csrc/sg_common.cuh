// Shared declarations for libsgb200 (internal; the public C-ABI is include/sgb200.h).
#pragma once
#include <stddef.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/sgb200.h"

#define SG_WIN 400      // 25 ms @ 16 kHz          (kaldi.py:125-151)
#define SG_SHIFT 160    // 10 ms
#define SG_NFFT 512     // round_to_power_of_two
#define SG_HALO 120     // WIN/2 - SHIFT/2: left reflect pad (kaldi.py:71)
#define SG_NMEL 30
#define SG_NCEP 30
#define SG_FLD 32       // internal feature row stride (30 cepstra + 2 zero pads)
#define SG_EPS 1.1920928955078125e-07f

#define SG_C1 512
#define SG_C5 1500
#define SG_C5P 1536     // layer-5 channels padded to a multiple of 128
#define SG_STATS (2 * SG_C5P)
#define SG_EMB 512

void sg_set_error(const char* fmt, ...);

// One-time per-DEVICE setup (cudaFuncSetAttribute opt-ins and __constant__ uploads belong to a device's context, so a
// process-global "done" flag would leave every device after the first one unconfigured).  Returns true the first time
// it is called with `mask` while `device` is current; thread-safe.
#include <atomic>
static inline bool sg_first_on_device(std::atomic<unsigned long long>* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  return (mask->fetch_or(bit) & bit) == 0;
}

#define SG_CUDA_CHECK(expr)                                                          \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      sg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                   __LINE__);                                                        \
      return SG_ECUDA;                                                               \
    }                                                                                \
  } while (0)

#define SG_LAUNCH_CHECK()                                                            \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      sg_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),       \
                   __FILE__, __LINE__);                                              \
      return SG_ECUDA;                                                               \
    }                                                                                \
  } while (0)

// ---- feature tables (device resident, built once per handle by sg_feat_tables_build) --------
// tables of the half-warp-per-frame MFCC kernels (sg_feat.cu, "V2"): 16 lanes own a frame, lane l holds the FFT elements
// 16 a + l and, after the transform, the bins l + 16 i
#define SG_M2_ITERS 16
struct alignas(16) SgFeatTables2 {
  // common
  float window[SG_WIN];          // Povey window
  float2 tw16[16][16];           // [k1][b] = exp(-2 pi i b k1 / 256): twiddles between the two radix-16 passes
  float2 untw[16][16];           // [i][l] = (cos, sin)(2 pi (l + 16 i) / 512): real-FFT untangle
  // forward only (the forward kernel copies the struct up to `binw`)
  float4 m2_w[SG_M2_ITERS][16];  // mel weights: lane l runs through the float4 groups of filter l, then of filter 29 - l
  int m2_len0[16], m2_lo0[16], m2_lo1s[16];   // groups of the first filter, its first bin, first bin of the second minus 4 * len0
  int m2_iters;                  // max over lanes of the two filters' groups
  int pad_[3];
  float dct_kn[32][36];          // as below
  // adjoint only
  float2 binw[256];              // per FFT bin: weights towards the (<= 2) filters it feeds
  int binc[256];                 // their indices c0 | c1 << 8 (31 = none)
  float dct_nk[32][36];
};
#define SG_T2_FWD_BYTES (offsetof(SgFeatTables2, binw))
#define SG_T2_COMMON_BYTES (offsetof(SgFeatTables2, m2_w))

struct alignas(16) SgFeatTables {
  float window[SG_WIN];          // Povey window
  float2 tw[24][32];             // per-lane FFT twiddles: [0..7] pass A, [8..15] pass B, [16..23] untangle
  int mel_lo[32];                // first FFT bin of filter c, rounded down to a multiple of 4 (float4 loads)
  int mel_len[32];               // number of float4 groups
  int mel_off[32];               // offset into mel_w (floats, multiple of 4)
  float mel_w[768];              // packed weights, zero-padded to float4 groups
  int bin_c0[256], bin_c1[256];  // per FFT bin: the (<=2) filters it feeds (31 = none)
  float bin_w0[256], bin_w1[256];
  float dct_kn[32][36];          // [k][n] = D[n][k] * lifter[k]: forward, lane k reads float4 over n (stride 36: conflict-free)
  float dct_nk[32][36];          // [n][k] (k >= 1; column 0 zero: C0 is the log-energy): adjoint, lane n reads float4 over k
  int mel_maxlen;                // max number of float4 groups
  int pad_[3];
  SgFeatTables2 v2;              // LAST member: the V1 kernels copy only the bytes before it into shared memory
};
#define SG_FEAT_V1_BYTES (offsetof(SgFeatTables, v2))

int sg_feat_tables_build(SgFeatTables* host_out);

// ---- GEMM-as-convolution (SIMT fp32 path) ----------------------------------------------------
enum SgEpilogue {
  SG_EPI_BIAS = 0,        // out = acc + bias
  SG_EPI_BIAS_RELU = 1,   // out = relu(acc + bias)
  SG_EPI_MASK = 2,        // out = acc * (mask > 0) * (t < t_valid)      (dgrad through ReLU)
  SG_EPI_NONE = 3         // out = acc
};

struct SgConvArgs {
  const float* A; int lda;          // activations [rows, lda], K (channels) contiguous
  const float* W;                   // packed weights [taps*cin, N] row-major (SIMT path)
  const float* Wk;                  // K-major copy [N, taps*cin] (tensor-core path); may be null for SIMT-only calls
  const float* bias;                // [N] or null
  float* out; int ldo;              // [rows, ldo]
  int rows;                         // M
  int N;                            // output columns (<= ldo)
  int cin;                          // K per tap (multiple of 16)
  int taps; int tap_step;           // source row of tap k = p + tap_base + k*tap_step (may be negative)
  int tap_base; int same_utt;       // same_utt: taps that leave the utterance (row / T changes) read zero ('same' padding)
  int w_per_utt;                    // tensor-core path with same_utt: Wk holds one [N, taps*cin] K-major operand PER UTTERANCE (batched contraction)
  int epilogue;
  const float* mask; int ldmask;    // SG_EPI_MASK: post-ReLU activation of the producing layer
  int T; int t_valid;               // rows per utterance / valid rows (SG_EPI_MASK)
  // tensor-core path only: 1 bit per output element instead of re-reading the fp32 activation as a mask
  uint32_t* bits_out;               // SG_EPI_BIAS_RELU: bit j of word [row, col/32] = (out[row, col] > 0)
  const uint32_t* bits_in;          // SG_EPI_MASK: replaces `mask` when non-null
  int ldbits;                       // words per row
  // bf16 mode (tensor-core path): A / Wk hold __nv_bfloat16 when op_bf16, out is __nv_bfloat16 when out_bf16
  int op_bf16; int out_bf16;
  // tensor-core path, bf16 only: fuse the statistics-pooling adjoint into this (layer-5 dgrad) contraction.  A is then the
  // stored activation r5 and xf_ab [rows / T][xf_ld] holds (alpha', beta) pairs: dA5 = (t < xf_tv && r > 0) ? alpha' + beta r : 0
  const float* xf_ab; int xf_ld; int xf_tv;
  // tensor-core path: re-strided output.  The rows are frames of utterances with T rows each; with out_T > 0 the output is
  // [rows / T][out_T frames][N] with out_Tstride rows per utterance in memory: out_T < T drops the frames >= out_T of every
  // utterance (compaction to the valid frames), out_T == T with a larger out_Tstride spreads a compact tensor out again.
  // bits_T > 0: the ReLU-bit rows written by the forward epilogue use the same compact indexing (frame < bits_T kept).
  int out_T, out_Tstride, bits_T;
  // tensor-core path, small problems: scratch for split-K partials (fp32, splitk_floats elements); null = no split
  float* splitk_ws; size_t splitk_floats;
  // batched GEMM (SIMT path only): blockIdx.z selects an item; element strides, 0 = shared operand
  int nbatch; long long strideA, strideW, strideO;
};

int sg_conv_simt(const SgConvArgs& a, cudaStream_t st);
int sg_conv_tc(const SgConvArgs& a, int precision, cudaStream_t st);   // tcgen05 path (sg_tdnn_tc.cu)
int sg_conv_tc_extra_launches();                                       // kernels the last sg_conv_tc on this thread launched beyond the contraction

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
