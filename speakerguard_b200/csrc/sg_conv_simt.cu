// fp32 (FFMA) GEMM-as-dilated-convolution: the parity-mode path for the TDNN layers
// (reference model/_xv_plda/xvecTDNN.py:16-53, forward; autograd dgrad in the backward) and for
// the dense projections of the head (fc1 :63, LDA model/iv_plda.py:423-435).
//
//   out[p, n] = epi( sum_{tap, c} A[p + tap*tap_step, c] * W[tap*cin + c, n] )
//
// Activations are channels-last [rows, lda] with every utterance occupying T consecutive rows,
// so a dilated tap is just a row offset (no im2col); rows outside [0, rows) read as zero.
// 128 x 128 x 16 tiles, 256 threads, 8 x 8 register micro-tiles, register-staged double buffering.
// A warp covers 4 x 8 micro-tiles (not 2 x 16): its A fragment is 64 contiguous bytes and its B fragment 128, so every LDS.128
// of the inner loop is one shared-memory wavefront (the 2 x 16 arrangement ran the data pipe at 68 % and the FMA pipe at 60 %).
// The loop carries its tap / channel position and the weight pointer forward instead of dividing the chunk index each round.
#include "sg_common.cuh"
#include <cstdlib>

#define BM 128
#define BK 16

// BN = 128 / 64 / 32 output columns per CTA (the narrow tiles serve the 32- and 64-channel AudioNet stages and the skinny
// i-vector contractions); every thread owns 8 rows x NT = BN/16 columns
template <int BN>
__global__ void __launch_bounds__(256, 2)
conv_simt_kernel(SgConvArgs a) {
  constexpr int NT = BN / 16;            // columns per thread: 8 (two float4), 4 (one float4) or 2 (one float2)
  constexpr int H = NT == 8 ? 2 : 1;     // column groups per thread
  constexpr int W = NT == 8 ? 4 : NT;    // columns per group
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  a.A += (long long)blockIdx.z * a.strideA;
  a.W += (long long)blockIdx.z * a.strideW;
  a.out += (long long)blockIdx.z * a.strideO;
  const int n0 = blockIdx.x * BN;
  const int p0 = blockIdx.y * BM;
  // warps 4 (rows) x 2 (columns), lanes 4 x 8 inside a warp
  const int ty = ((tid >> 6) << 2) | ((tid >> 3) & 3), tx = (((tid >> 5) & 1) << 3) | (tid & 7);
  // loaders
  const int a_row = tid & 127, a_k = (tid >> 7) * 8;
  const int b_k = tid >> 4, b_col = (tid & 15) * NT;

  float acc[8][NT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;

  const int kchunks = a.cin / BK;
  const int a_t = a.same_utt ? (p0 + a_row) % a.T : 0;          // frame index of this loader's row
  const int nk = a.taps * kchunks;

  float4 ra[2], rb[2];
  // position of the next chunk to load: tap, first channel, row offset of the tap, weight row (tap * cin + c0 = chunk * BK)
  int ld_c0 = 0, ld_off = a.tap_base;
  const float* a_src = a.A + ((long long)p0 + a_row) * a.lda + a_k;      // dereferenced only for rows inside [0, rows)
  const float* w_src = a.W + (size_t)b_k * a.N + n0 + b_col;
  const bool b_in0 = NT >= 4 ? n0 + b_col + 3 < a.N : n0 + b_col + 1 < a.N;
  const bool b_in1 = NT == 8 && n0 + b_col + 7 < a.N;
  auto gload = [&]() {
    const long long sr = (long long)p0 + a_row + ld_off;
    const bool inside = !a.same_utt || (a_t + ld_off >= 0 && a_t + ld_off < a.T);
    if (sr >= 0 && sr < a.rows && inside) {
      const float4* src = reinterpret_cast<const float4*>(a_src + (long long)ld_off * a.lda + ld_c0);
      ra[0] = __ldg(src); ra[1] = __ldg(src + 1);
    } else {
      ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    rb[0] = rb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NT >= 4) {
      if (b_in0) rb[0] = __ldg(reinterpret_cast<const float4*>(w_src));
      if (b_in1) rb[1] = __ldg(reinterpret_cast<const float4*>(w_src + 4));
    } else if (b_in0) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(w_src));
      rb[0].x = t.x; rb[0].y = t.y;
    }
    w_src += (size_t)BK * a.N;
    ld_c0 += BK;
    if (ld_c0 == a.cin) { ld_c0 = 0; ld_off += a.tap_step; }
  };
  auto sstore = [&](int buf) {
    As[buf][a_k + 0][a_row] = ra[0].x; As[buf][a_k + 1][a_row] = ra[0].y;
    As[buf][a_k + 2][a_row] = ra[0].z; As[buf][a_k + 3][a_row] = ra[0].w;
    As[buf][a_k + 4][a_row] = ra[1].x; As[buf][a_k + 5][a_row] = ra[1].y;
    As[buf][a_k + 6][a_row] = ra[1].z; As[buf][a_k + 7][a_row] = ra[1].w;
    if (NT >= 4) {
      *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = rb[0];
      if (NT == 8) *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col + 4]) = rb[1];
    } else {
      *reinterpret_cast<float2*>(&Bs[buf][b_k][b_col]) = make_float2(rb[0].x, rb[0].y);
    }
  };

  gload();
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[NT];
      if (NT == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        bv[NT - 4] = b1.x; bv[NT - 3] = b1.y; bv[NT - 2] = b1.z; bv[NT - 1] = b1.w;
      } else if (NT == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        bv[0] = b0.x; bv[1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = p0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= a.rows) continue;
    bool row_ok = true;
    if (a.epilogue == SG_EPI_MASK) row_ok = (row % a.T) < a.t_valid;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const int col = n0 + (NT == 8 ? (h == 0 ? tx * 4 : 64 + tx * 4) : tx * W);
      if (col >= a.N) continue;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = acc[i][h * W + j];
      if (a.epilogue == SG_EPI_BIAS || a.epilogue == SG_EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (col + j < a.N) v[j] += a.bias ? __ldg(a.bias + col + j) : 0.f;
      }
      if (a.epilogue == SG_EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (a.epilogue == SG_EPI_MASK) {
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const bool on = row_ok && (col + j < a.N) && (__ldg(a.mask + (size_t)row * a.ldmask + col + j) > 0.f);
          v[j] = on ? v[j] : 0.f;
        }
      }
      float* o = a.out + (size_t)row * a.ldo + col;
      if (W == 4 && col + 3 < a.N) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else if (W == 2 && col + 1 < a.N) {
        *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (col + j < a.N) o[j] = v[j];
      }
    }
  }
}

int sg_conv_simt(const SgConvArgs& a, cudaStream_t st) {
  if (a.cin % BK != 0 || a.N % 4 != 0 || a.lda % 4 != 0 || a.ldo % 4 != 0) {
    sg_set_error("sg_conv_simt: cin %% 16, N %% 4, lda %% 4, ldo %% 4 required (cin=%d N=%d lda=%d ldo=%d)",
                 a.cin, a.N, a.lda, a.ldo);
    return SG_EINVAL;
  }
  if (a.nbatch > 1 && (a.strideA % 4 != 0 || a.strideW % 4 != 0 || a.strideO % 4 != 0 || a.nbatch > 65535)) {
    sg_set_error("sg_conv_simt: batch strides must be multiples of 4 floats and nbatch <= 65535");
    return SG_EINVAL;
  }
  // Tile width: the one with the least padded columns per unit of measured throughput.  The 8 x 8 micro-tile of the 128-wide
  // kernel does 64 FMAs per 4 fragment loads, the 8 x 4 and 8 x 2 ones 32 per 3 and 16 per 3; measured on the N = 80 statistics
  // adjoint of the i-vector path (2048-deep contraction, 256 batches): 1.77 ms at 128 (48 padded columns), 2.00 ms at 64,
  // 2.27 ms at 32 - i.e. 0.89 and 0.59 of the wide kernel's rate per computed column.
  auto padded = [&](int t) { return (a.N + t - 1) / t * t; };
  const float c128 = (float)padded(128), c64 = padded(64) / 0.89f, c32 = padded(32) / 0.59f;
  int bn = 128;                                  // a narrow tile has to win by 10 % (N = 400 at 64: 1.22 ms against 1.01 ms at 128)
  if (c64 < 0.9f * c128 && c64 <= c32) bn = 64;
  else if (c32 < 0.9f * c128 && c32 < c64) bn = 32;
  static const int force_bn = [] { const char* e = getenv("SGB200_SIMT_BN"); return e ? atoi(e) : 0; }();   // A/B switch
  if (force_bn == 32 || force_bn == 64 || force_bn == 128) bn = force_bn;
  dim3 grid((a.N + bn - 1) / bn, (a.rows + BM - 1) / BM, a.nbatch > 1 ? a.nbatch : 1);
  if (bn == 32) conv_simt_kernel<32><<<grid, 256, 0, st>>>(a);
  else if (bn == 64) conv_simt_kernel<64><<<grid, 256, 0, st>>>(a);
  else conv_simt_kernel<128><<<grid, 256, 0, st>>>(a);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
