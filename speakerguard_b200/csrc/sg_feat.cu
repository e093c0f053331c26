// F1 / F2: Kaldi-compatible MFCC forward and its adjoint (+ fused L-inf sign step), sliding CMVN.
//
// Restates xv_plda.raw (reference model/xv_plda.py:107-156) = torchaudio kaldi.mfcc
// (kaldi.py:669-813 -> fbank :514-645 -> _get_window :154-217 -> _get_strided :44-83) and
// iv_plda.cmvn (model/iv_plda.py:296-377) as hand-written sm_100a kernels.  One warp owns one
// 25 ms frame: reflect-padded gather, dither, DC removal, raw log-energy, pre-emphasis, Povey
// window, a 512-point real FFT done as a 256-point complex FFT in registers + shared memory
// (8 x 8 x 4 radix passes), sparse triangular mel filterbank, log, DCT + lifter.
// The adjoint recomputes the forward per frame (bytes are scarce, flops are free), walks the
// chain backwards, and overlap-adds the frame gradients in shared memory in a fixed order
// (deterministic), so the waveform gradient / the updated iterate is written exactly once.
#include <math.h>
#include <string.h>

#include "sg_common.cuh"

// =============================================================================================
// host: constant tables
// =============================================================================================
int sg_feat_tables_build(SgFeatTables* t) {
  memset(t, 0, sizeof(*t));
  const double PI = 3.14159265358979323846;
  for (int j = 0; j < SG_WIN; ++j) {
    double h = 0.5 - 0.5 * cos(2.0 * PI * j / (SG_WIN - 1));     // kaldi.py:98-100 (povey)
    t->window[j] = (float)pow(h, 0.85);
  }
  for (int lane = 0; lane < 32; ++lane) {
    for (int k0 = 0; k0 < 8; ++k0) {                              // pass A: W_64^(n1*k0)
      int n1 = lane >> 2;
      double a = -2.0 * PI * (n1 * k0) / 64.0;
      t->tw[k0][lane] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k1 = 0; k1 < 8; ++k1) {                              // pass B: W_256^(n2*(k0+8*k1))
      int k0 = lane >> 2, n2 = lane & 3;
      double a = -2.0 * PI * (n2 * (k0 + 8 * k1)) / 256.0;
      t->tw[8 + k1][lane] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int i = 0; i < 8; ++i) {                                 // untangle: (cos, sin)(2 pi k / 512)
      int k = lane + 32 * i;
      double a = 2.0 * PI * k / 512.0;
      t->tw[16 + i][lane] = make_float2((float)cos(a), (float)sin(a));
    }
  }
  // mel filterbank, kaldi.py:436-511 with low 20 Hz, high 7600 Hz, 30 bins, vtln_warp 1
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double ml = mel(20.0), mh = mel(7600.0);
  const double delta = (mh - ml) / (SG_NMEL + 1);
  static double w[SG_NMEL][256];
  for (int c = 0; c < SG_NMEL; ++c) {
    double left = ml + c * delta, center = ml + (c + 1.0) * delta, right = ml + (c + 2.0) * delta;
    for (int b = 0; b < 256; ++b) {
      double mf = mel((16000.0 / SG_NFFT) * b);
      double up = (mf - left) / (center - left), down = (right - mf) / (right - center);
      double v = up < down ? up : down;
      w[c][b] = v > 0.0 ? v : 0.0;
    }
  }
  int off = 0, maxlen = 0;
  for (int b = 0; b < 256; ++b) { t->bin_c0[b] = t->bin_c1[b] = 31; }
  for (int c = 0; c < 32; ++c) {
    int lo = 0, hi = -1;
    if (c < SG_NMEL) {
      lo = 256;
      for (int b = 0; b < 256; ++b)
        if (w[c][b] > 0.0) { if (b < lo) lo = b; hi = b; }
      if (hi < 0) lo = 0;
    }
    if (hi >= 252) return SG_EINVAL;                      // float4 groups must stay inside P[0..255]
    const int lo4 = lo & ~3;
    const int groups = hi >= lo ? (hi - lo4) / 4 + 1 : 0;
    t->mel_lo[c] = lo4; t->mel_len[c] = groups; t->mel_off[c] = off;
    if (off + 4 * groups > 768) return SG_EINVAL;
    for (int i = 0; i < 4 * groups; ++i) {
      const int b = lo4 + i;
      const double wv = (b >= lo && b <= hi) ? w[c][b] : 0.0;
      t->mel_w[off + i] = (float)wv;
      if (wv > 0.0) {
        if (t->bin_c0[b] == 31) { t->bin_c0[b] = c; t->bin_w0[b] = (float)wv; }
        else if (t->bin_c1[b] == 31) { t->bin_c1[b] = c; t->bin_w1[b] = (float)wv; }
        else return SG_EINVAL;   // triangles overlap at most pairwise
      }
    }
    off += 4 * groups;
    if (groups > maxlen) maxlen = groups;
  }
  t->mel_maxlen = maxlen;
  // DCT-II (ortho) with kaldi's first column and the lifter folded in: kaldi.py:648-666, :788-796
  for (int n = 0; n < SG_NMEL; ++n)
    for (int k = 0; k < SG_NCEP; ++k) {
      double d = (k == 0) ? sqrt(1.0 / SG_NMEL) : sqrt(2.0 / SG_NMEL) * cos(PI / SG_NMEL * (n + 0.5) * k);
      double lift = 1.0 + 0.5 * 22.0 * sin(PI * k / 22.0);
      t->dct_kn[k][n] = (float)(d * lift);
      t->dct_nk[n][k] = (k == 0) ? 0.f : (float)(d * lift);
    }
  // ---- V2 (half-warp per frame): the same values, re-indexed -------------------------------------
  SgFeatTables2& u = t->v2;
  memcpy(u.window, t->window, sizeof(u.window));
  memcpy(u.dct_kn, t->dct_kn, sizeof(u.dct_kn));
  memcpy(u.dct_nk, t->dct_nk, sizeof(u.dct_nk));
  for (int k1 = 0; k1 < 16; ++k1)
    for (int b = 0; b < 16; ++b) {
      const double a = -2.0 * PI * (b * k1) / 256.0;
      u.tw16[k1][b] = make_float2((float)cos(a), (float)sin(a));
    }
  for (int i = 0; i < 16; ++i)
    for (int l = 0; l < 16; ++l) {
      const double a = 2.0 * PI * (l + 16 * i) / 512.0;
      u.untw[i][l] = make_float2((float)cos(a), (float)sin(a));
    }
  int iters = 0;
  for (int l = 0; l < 16; ++l) {
    const int c0 = l, c1 = SG_NMEL - 1 - l;
    const int g0 = l < 15 ? t->mel_len[c0] : 0, g1 = l < 15 ? t->mel_len[c1] : 0;
    if (g0 + g1 > SG_M2_ITERS) return SG_EINVAL;
    u.m2_len0[l] = g0; u.m2_lo0[l] = l < 15 ? t->mel_lo[c0] : 0; u.m2_lo1s[l] = l < 15 ? t->mel_lo[c1] - 4 * g0 : 0;
    for (int i = 0; i < g0 + g1; ++i) {
      const float* w4 = i < g0 ? &t->mel_w[t->mel_off[c0] + 4 * i] : &t->mel_w[t->mel_off[c1] + 4 * (i - g0)];
      u.m2_w[i][l] = make_float4(w4[0], w4[1], w4[2], w4[3]);
    }
    if (g0 + g1 > iters) iters = g0 + g1;
  }
  u.m2_iters = iters;
  for (int k = 0; k < 256; ++k) {
    u.binw[k] = make_float2(t->bin_w0[k], t->bin_w1[k]);
    u.binc[k] = t->bin_c0[k] | (t->bin_c1[k] << 8);
  }
  return SG_OK;
}

// =============================================================================================
// device helpers
// =============================================================================================
#define FEAT_THREADS 256
#define FEAT_WARPS 8
#define WARP_SCRATCH 832                  // floats per warp: re[288] + im[288] + P[256]
#define ACC_LEN 1520                      // 7*160 + 400: padded span of 8 consecutive frames
#define ACC_RING 2048                     // circular accumulator of the adjoint (>= ACC_LEN, power of two)
// cmvn_colsum_order: column sums over the frames of an utterance of T <= CMN_WIN frames are formed as
//   sum over chunks c of CMVN_CHUNK frames ( sum over r < 8 ( sum over i < 8 of x[64 c + r + 8 i] ) ), every sum left to right from 0.f,
// by all three implementations (cmvn_kernel, the cluster epilogue of mfcc_fwd_kernel<true>, the prologue of mfcc_bwd_kernel),
// so the fused attack loop and the stage-by-stage API produce identical bits.
#define CMVN_CHUNK 64
#define CMVN_MAX_CHUNKS 8                 // (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK = 5, padded
#define CMN_WIN 300

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i*t3
  a3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i*t3
}

// natural-order in, natural-order out: v[k] = sum_n v[n] W_8^(nk)
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  fft4(v[0], v[2], v[4], v[6]);
  fft4(v[1], v[3], v[5], v[7]);
  const float h = 0.70710678118654752440f;
  float2 o1 = make_float2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));     // * W_8^1
  float2 o2 = make_float2(v[5].y, -v[5].x);                                   // * W_8^2 = -i
  float2 o3 = make_float2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));    // * W_8^3
  float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// 256-point complex FFT across one warp.  In: z[r] = element 32*r + lane.  Out: z[4*j + k2] =
// Z[k0 + 8*(c + 4*j) + 64*k2] with lane = k0 + 8*c.  sre/sim: >= 288 floats each, warp private.
// (index maps prototyped and checked in tools/fft_proto.py)
__device__ __forceinline__ void warp_fft256(float2 (&z)[8], const float2 (*tw)[32], float* sre,
                                            float* sim, int lane) {
  fft8(z);
#pragma unroll
  for (int k0 = 1; k0 < 8; ++k0) z[k0] = cmul(z[k0], tw[k0][lane]);
#pragma unroll
  for (int k0 = 0; k0 < 8; ++k0) { sre[k0 * 36 + lane] = z[k0].x; sim[k0 * 36 + lane] = z[k0].y; }
  __syncwarp();
  {
    const int base = (lane >> 2) * 36 + (lane & 3);
#pragma unroll
    for (int n1 = 0; n1 < 8; ++n1) z[n1] = make_float2(sre[base + n1 * 4], sim[base + n1 * 4]);
  }
  __syncwarp();
  fft8(z);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) z[k1] = cmul(z[k1], tw[8 + k1][lane]);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) { sre[k1 * 33 + lane] = z[k1].x; sim[k1 * 33 + lane] = z[k1].y; }
  __syncwarp();
  {
    const int k0 = lane & 7, c = lane >> 3;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        int a = (c + 4 * j) * 33 + k0 * 4 + n2;
        z[4 * j + n2] = make_float2(sre[a], sim[a]);
      }
  }
  __syncwarp();
  fft4(z[0], z[1], z[2], z[3]);
  fft4(z[4], z[5], z[6], z[7]);
}

// scatter the FFT output to natural order in shared memory: s[k] for k in [0,256)
__device__ __forceinline__ void fft_out_to_smem(const float2 (&z)[8], float* sre, float* sim, int lane) {
  const int k0 = lane & 7, c = lane >> 3;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      int k = k0 + 8 * (c + 4 * j) + 64 * k2;
      sre[k] = z[4 * j + k2].x;
      sim[k] = z[4 * j + k2].y;
    }
}

// ---- cache prefetch hints: the per-frame working set of the NEXT frame a warp will process is requested while the current
// one is computed (one 128-byte line per lane), so the dependent loads at the top of the next iteration hit L1 / L2
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- split CTA barrier (mbarrier): arrive now, wait later -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
}

// ---- Philox4x32-7 (counter-based dither; 7 rounds pass BigCrush, Salmon et al. 2011) -----------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// The philox dither is a statistically-equivalent stand-in for torch.randn (SG_DITHER_TENSOR is the
// bit-parity mode), so the fast intrinsics are fine here; forward, adjoint and sg_dither_fill share
// this function and therefore see identical noise.
// N(0,1) pairs from 16-bit uniforms: one philox call (4 words) feeds four sample pairs.  The dither is one quantisation
// step of the int16-range waveform; its distribution, not its resolution, is what the features see (radius: 65536 levels up
// to 4.8 sigma, angle: 65536 directions).
__device__ __forceinline__ float2 box_muller16(uint32_t w) {
  const float u1 = __uint_as_float(0x3f800000u | ((w & 0xffffu) << 7)) - 0.99999237060546875f;    // (lo16 + 0.5) / 65536 in (0, 1)
  const float u2 = __uint_as_float(0x3f800000u | ((w >> 16) << 7)) - 1.0f;                       // hi16 / 65536 in [0, 1)
  // the bare MUFU forms (u1 is a normal number well inside the range of lg2.approx, the angle lies in [0, 2 pi))
  float lg, rs, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
  const float t2 = -1.3862943611198906f * lg;                                                     // -2 ln u1 > 0
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t2));
  const float r = t2 * rs;
  const float ang = 6.283185307179586f * u2;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
  return make_float2(r * cs, r * sn);
}

struct DitherSpec {
  int mode;              // SG_DITHER_*
  const float* tensor;   // [B, m, 400] for this pass
  uint32_t seed_lo, seed_hi;
  uint32_t pass;
  uint32_t b_off;        // global index of utterance 0 (SG_OPT_UTT_OFFSET): the philox counter uses b + b_off
  uint32_t copy_rows;    // EOT copies as batch rows (sg_pgd_params::eot_batch): row r = copy r / copy_rows of utterance r % copy_rows,
                         // keyed as (utterance + b_off) + (copy << 24) so that a sharded run draws the unsharded one's noise; 0 = off
  // CUDA-graph replay (sg_pgd_run): seed and pass counter live in device memory so that one captured iteration can be
  // replayed for every iteration of every attack; ctl = {pass, seed_lo, seed_hi}, `pass` above is then added to ctl[0]
  const uint32_t* ctl;
};

__device__ __forceinline__ uint32_t dither_key(const DitherSpec& D, int b) {
  if (D.copy_rows == 0u) return (uint32_t)b + D.b_off;
  const uint32_t e = (uint32_t)b / D.copy_rows;
  return ((uint32_t)b - e * D.copy_rows) + D.b_off + (e << 24);
}
// resolve the device-resident part of a DitherSpec (warp-uniform loads, once per kernel)
__device__ __forceinline__ void dither_resolve(DitherSpec& D) {
  if (D.ctl != nullptr) { D.pass += __ldg(D.ctl); D.seed_lo = __ldg(D.ctl + 1); D.seed_hi = __ldg(D.ctl + 2); }
}

// Per-lane ownership of a frame: sample j = 64*n0 + 2*lane + e, n0 in [0,7), e in {0,1};
// valid iff j < 400 (n0 == 6 only for lane < 8).  The FFT packs z[n] = g[2n] + i g[2n+1], so
// the lane's pair (n0) is exactly FFT input element n = 32*n0 + lane.
struct Frame {
  float fe[7], fo[7];    // DC-removed samples (even / odd of each pair)
  float sumsq;           // raw energy sum (kaldi.py:116-122)
  float2 X[8];           // spectrum bins k = lane + 32*i
};

__device__ __forceinline__ void load_frame(Frame& F, const float* __restrict__ xb, int N, int b, int m,
                                           int fr, const DitherSpec& D, int lane) {
  const int p0 = fr * SG_SHIFT - SG_HALO + 2 * lane;
  float nz[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) nz[i] = 0.f;
  if (D.mode == SG_DITHER_TENSOR) {
    const float* dp = D.tensor + ((size_t)b * m + fr) * SG_WIN + 2 * lane;
#pragma unroll
    for (int n0 = 0; n0 < 7; ++n0)
      if (n0 < 6 || lane < 8) {
        float2 d = *reinterpret_cast<const float2*>(dp + 64 * n0);
        nz[2 * n0] = d.x; nz[2 * n0 + 1] = d.y;
      }
  } else if (D.mode == SG_DITHER_PHILOX) {
    // one dither stream for both kernel generations (defined by the V2 ownership: call q, lane l = lane & 15, word u -> sample
    // pair a = 4 q + u, samples 32 a + 2 l, +1): this lane's pair n0 is a = 2 n0 + (lane >> 4)
    const int hi = lane >> 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 r = philox4x32_10(make_uint4(q * 16 + (lane & 15), (uint32_t)fr, dither_key(D, b), D.pass), make_uint2(D.seed_lo, D.seed_hi));
      const float2 a = box_muller16(hi ? r.y : r.x);
      nz[4 * q] = a.x; nz[4 * q + 1] = a.y;
      if (q < 3) { const float2 c = box_muller16(hi ? r.w : r.z); nz[4 * q + 2] = c.x; nz[4 * q + 3] = c.y; }
    }
  }
  float s = 0.f;
  const bool interior = (fr * SG_SHIFT - SG_HALO >= 0) && (fr * SG_SHIFT - SG_HALO + SG_WIN <= N);   // warp-uniform
  if (interior) {
#pragma unroll
    for (int n0 = 0; n0 < 7; ++n0) {
      const bool valid = (n0 < 6) || (lane < 8);
      float ve = 0.f, vo = 0.f;
      if (valid) {
        ve = __ldg(xb + p0 + 64 * n0) * 32768.0f + nz[2 * n0];
        vo = __ldg(xb + p0 + 64 * n0 + 1) * 32768.0f + nz[2 * n0 + 1];
      }
      F.fe[n0] = ve; F.fo[n0] = vo;
      s += ve + vo;
    }
  } else
#pragma unroll
  for (int n0 = 0; n0 < 7; ++n0) {
    const bool valid = (n0 < 6) || (lane < 8);
    float ve = 0.f, vo = 0.f;
    if (valid) {
      int pe = p0 + 64 * n0, po = pe + 1;
      pe = pe < 0 ? -pe - 1 : (pe >= N ? 2 * N - 1 - pe : pe);     // kaldi.py:69-77 reflect pad
      po = po < 0 ? -po - 1 : (po >= N ? 2 * N - 1 - po : po);
      ve = __ldg(xb + pe) * 32768.0f + nz[2 * n0];                 // model/utils.py:14, kaldi.py:181
      vo = __ldg(xb + po) * 32768.0f + nz[2 * n0 + 1];
    }
    F.fe[n0] = ve; F.fo[n0] = vo;
    s += ve + vo;
  }
  const float mean = warp_sum(s) * (1.0f / SG_WIN);                // kaldi.py:183-186
  float q = 0.f;
#pragma unroll
  for (int n0 = 0; n0 < 7; ++n0) {
    const bool valid = (n0 < 6) || (lane < 8);
    F.fe[n0] = valid ? F.fe[n0] - mean : 0.f;
    F.fo[n0] = valid ? F.fo[n0] - mean : 0.f;
    q += F.fe[n0] * F.fe[n0] + F.fo[n0] * F.fo[n0];
  }
  F.sumsq = warp_sum(q);
}

// pre-emphasis + window + real FFT; leaves the spectrum in F.X and |X|^2 in P[0..255] (smem)
__device__ __forceinline__ void frame_spectrum(Frame& F, const SgFeatTables* T, float* sre, float* sim,
                                               float* P, int lane) {
  float2 z[8];
#pragma unroll
  for (int n0 = 0; n0 < 7; ++n0) {
    // previous sample of the even element: odd element of lane-1 (same pair row), or of lane 31
    // in the previous row; j == 0 replicates itself (kaldi.py:193-198)
    float up = __shfl_up_sync(0xffffffffu, F.fo[n0], 1);
    float wrap = __shfl_sync(0xffffffffu, n0 > 0 ? F.fo[n0 > 0 ? n0 - 1 : 0] : F.fe[0], 31);
    float prev = lane > 0 ? up : (n0 > 0 ? wrap : F.fe[0]);
    const bool valid = (n0 < 6) || (lane < 8);
    float2 w2 = valid ? *reinterpret_cast<const float2*>(&T->window[64 * n0 + 2 * lane]) : make_float2(0.f, 0.f);
    z[n0] = make_float2((F.fe[n0] - 0.97f * prev) * w2.x, (F.fo[n0] - 0.97f * F.fe[n0]) * w2.y);
  }
  z[7] = make_float2(0.f, 0.f);
  warp_fft256(z, T->tw, sre, sim, lane);
  fft_out_to_smem(z, sre, sim, lane);
  __syncwarp();
  // untangle: X[k] = a_k Z[k] + b_k conj(Z[256-k]),  a_k = ((1-s) - i c)/2, b_k = ((1+s) + i c)/2
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = lane + 32 * i;
    const int km = (256 - k) & 255;
    float2 zk = make_float2(sre[k], sim[k]);
    float2 zm = make_float2(sre[km], -sim[km]);
    float2 cs = T->tw[16 + i][lane];
    float2 a = make_float2(0.5f * (1.f - cs.y), -0.5f * cs.x);
    float2 bq = make_float2(0.5f * (1.f + cs.y), 0.5f * cs.x);
    float2 X = cadd(cmul(a, zk), cmul(bq, zm));
    F.X[i] = X;
    P[k] = X.x * X.x + X.y * X.y;                                   // kaldi.py:616-618
  }
  __syncwarp();
}

// lane c (< 30): mel energy (before the log); kaldi.py:621-630.  Windows are float4-aligned and zero-padded.
__device__ __forceinline__ float mel_energy(const SgFeatTables* T, const float* P, int lane) {
  const int lo = T->mel_lo[lane], len = T->mel_len[lane], off = T->mel_off[lane];
  const float4* w4 = reinterpret_cast<const float4*>(&T->mel_w[off]);
  const float4* p4 = reinterpret_cast<const float4*>(&P[lo]);
  float acc = 0.f;
  for (int i = 0; i < T->mel_maxlen; ++i)
    if (i < len) {
      const float4 w = w4[i], p = p4[i];
      acc = fmaf(w.x, p.x, acc); acc = fmaf(w.y, p.y, acc); acc = fmaf(w.z, p.z, acc); acc = fmaf(w.w, p.w, acc);
    }
  return acc;
}

__device__ __forceinline__ void copy_tables(SgFeatTables* dst, const SgFeatTables* __restrict__ src) {
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < (int)(SG_FEAT_V1_BYTES / 16); i += blockDim.x) d[i] = s[i];
}
static_assert(SG_FEAT_V1_BYTES % 16 == 0 && sizeof(SgFeatTables2) % 16 == 0, "feature tables must be int4-copyable");

// ---- per-frame stash: what the adjoint needs from the forward (spectrum, DC-removed dithered frame, mel energies,
// raw energy).  Written by mfcc_fwd_kernel when the attack loop asks for it, so that mfcc_bwd_kernel neither regenerates
// the dither nor repeats the forward FFT: 4 KB per frame of extra HBM traffic each way instead of ~1 200 instructions
// per frame-warp (the adjoint is issue-bound, not bandwidth-bound).  Layout (floats): X[8][32] float2 | f[7][32] float2
// | mel[32] | sumsq | pad.
#define SG_STASH_FLOATS 1024
__device__ __forceinline__ void stash_store(float* __restrict__ sp, const Frame& F, float me, int lane) {
  float2* s2 = reinterpret_cast<float2*>(sp);
#pragma unroll
  for (int i = 0; i < 8; ++i) __stcs(s2 + i * 32 + lane, F.X[i]);
#pragma unroll
  for (int n0 = 0; n0 < 7; ++n0) __stcs(s2 + 256 + n0 * 32 + lane, make_float2(F.fe[n0], F.fo[n0]));
  __stcs(sp + 960 + lane, me);
  if (lane == 0) __stcs(sp + 992, F.sumsq);
}
__device__ __forceinline__ float stash_load(const float* __restrict__ sp, Frame& F, int lane) {
  const float2* s2 = reinterpret_cast<const float2*>(sp);
#pragma unroll
  for (int i = 0; i < 8; ++i) F.X[i] = __ldcs(s2 + i * 32 + lane);
#pragma unroll
  for (int n0 = 0; n0 < 7; ++n0) { const float2 v = __ldcs(s2 + 256 + n0 * 32 + lane); F.fe[n0] = v.x; F.fo[n0] = v.y; }
  F.sumsq = __ldcs(sp + 992);
  return __ldcs(sp + 960 + lane);
}

// =============================================================================================
// F1: waveform -> raw MFCC
// =============================================================================================
// CMVN = true (utterances of <= CMN_WIN frames, where every CMVN window is the whole utterance: model/iv_plda.py:321-337):
// every warp keeps the running column sum of the cepstra it produced; each CTA (CMVN_CHUNK frames) publishes its column sums
// in `cm.part`, and the CTA that finishes an utterance LAST (a counter per utterance) forms the mean in chunk order and
// subtracts it from all rows of the utterance, which are still in L2.  `raw` therefore receives the CMVN output: no separate
// CMVN launch, no raw-feature round trip through HBM and - unlike a cluster barrier - no CTA ever waits for another.
// The summation order (cmvn_colsum_order) is the one the stand-alone cmvn_kernel uses, so both routes give identical bits,
// and it does not depend on which CTA happens to be last.
struct CmvnScratch {
  float* part;            // [B][CMVN_MAX_CHUNKS][32] column sums per chunk
  unsigned int* count;    // [B] CTAs finished per utterance; zero between launches (the last CTA resets it)
};
template <bool CMVN, int MINB>
__global__ void __launch_bounds__(FEAT_THREADS, MINB)
mfcc_fwd_kernel(const float* __restrict__ x, int N, int m, int frames_per_cta, DitherSpec D,
                float* __restrict__ raw, int ld, const SgFeatTables* __restrict__ gT, float* __restrict__ stash, CmvnScratch cm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int cm_last;                                          // CMVN: this CTA finished its utterance last
  SgFeatTables* T = reinterpret_cast<SgFeatTables*>(smem_raw);
  copy_tables(T, gT);
  dither_resolve(D);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scratch = reinterpret_cast<float*>(smem_raw + SG_FEAT_V1_BYTES) + warp * WARP_SCRATCH;
  float *sre = scratch, *sim = scratch + 288, *P = scratch + 576;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * frames_per_cta;
  const int f1 = min(f0 + frames_per_cta, m);
  float colsum = 0.f;                                              // CMVN: sum of column `lane` over this warp's frames, in frame order
  for (int fr = f0 + warp; fr < f1; fr += FEAT_WARPS) {
    Frame F;
    if (fr + FEAT_WARPS < f1 && lane < 14) {                       // next frame's 400 samples: <= 14 lines of 128 bytes
      int p = (fr + FEAT_WARPS) * SG_SHIFT - SG_HALO + 32 * lane;
      p = p < 0 ? 0 : (p >= N ? N - 1 : p);                        // (edge frames read reflected samples: nearby lines anyway)
      prefetch_l1(xb + p);
    }
    load_frame(F, xb, N, b, m, fr, D, lane);
    const float logE = logf(fmaxf(F.sumsq, SG_EPS));               // kaldi.py:119
    frame_spectrum(F, T, sre, sim, P, lane);
    float me = mel_energy(T, P, lane);
    if (stash != nullptr) stash_store(stash + ((size_t)b * m + fr) * SG_STASH_FLOATS, F, me, lane);
    float lm = logf(fmaxf(me, SG_EPS));                            // kaldi.py:631-633
    __syncwarp();
    P[lane] = (lane < SG_NMEL) ? lm : 0.f;                         // P is free after mel_energy: lm[32] for the DCT
    __syncwarp();
    float c = 0.f;
    {
      const float4* d4 = reinterpret_cast<const float4*>(&T->dct_kn[lane][0]);
      const float4* l4 = reinterpret_cast<const float4*>(P);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 d = d4[g], l = l4[g];
        c = fmaf(l.x, d.x, c); c = fmaf(l.y, d.y, c); c = fmaf(l.z, d.z, c); c = fmaf(l.w, d.w, c);
      }
    }
    if (lane == 0) c = logE;                                       // kaldi.py:799-800
    if (lane >= SG_NCEP) c = 0.f;
    if (lane < ld) raw[((size_t)b * m + fr) * ld + lane] = c;
    if (CMVN) colsum += c;
    __syncwarp();
  }
  if (CMVN) {
    // model/iv_plda.py:296-377 with T <= 300: y[t] = x[t] - mean_t(x)
    const float* fb = reinterpret_cast<const float*>(smem_raw + SG_FEAT_V1_BYTES);
    const int nch = gridDim.x;
    __syncthreads();                                               // every warp is done with its scratch
    sre[lane] = colsum;                                            // warp partial (scratch of warp w starts at fb + w * WARP_SCRATCH)
    __syncthreads();
    if (warp == 0) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < FEAT_WARPS; ++w) a += fb[w * WARP_SCRATCH + lane];
      __stcg(cm.part + ((size_t)b * CMVN_MAX_CHUNKS + blockIdx.x) * 32 + lane, a);
    }
    __threadfence();                                               // this CTA's rows and chunk sums are visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) cm_last = (atomicAdd(cm.count + b, 1u) == (unsigned)nch - 1u);   // ... before it is counted
    __syncthreads();
    if (cm_last) {
      __threadfence();
      float tot = 0.f;
      for (int c = 0; c < nch; ++c) tot += __ldcg(cm.part + ((size_t)b * CMVN_MAX_CHUNKS + c) * 32 + lane);   // chunk order
      const float mu = tot / (float)m;
      // rows written by the other CTAs are read past L1 (they are in L2); 8 rows per warp in flight
      if (lane < ld) {
        float* q0 = raw + (size_t)b * m * ld + lane;
        for (int t0 = warp; t0 < m; t0 += 8 * FEAT_WARPS) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int t = t0 + u * FEAT_WARPS; v[u] = t < m ? __ldcg(q0 + (size_t)t * ld) : 0.f; }
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int t = t0 + u * FEAT_WARPS; if (t < m) q0[(size_t)t * ld] = (lane < SG_NCEP) ? v[u] - mu : 0.f; }
        }
      }
      if (threadIdx.x == 0) cm.count[b] = 0u;                      // ready for the next launch on this stream
    }
  }
}

// =============================================================================================
// F2: d(raw MFCC) -> d(waveform), optionally fused with the L-inf sign step
// =============================================================================================
struct BwdOut {
  int mode;               // 0: write gradient, 1: fused sign step
  float* grad;            // mode 0: [B,N]
  float scale;            // mode 0
  int accumulate;         // mode 0
  const float* x0;        // mode 1: clean waveform (bounds)
  float* x_out;           // mode 1: updated iterate (must not alias the input waveform)
  float step;             // mode 1: step_size * grad_sign
  float eps;              // mode 1
  int cmvn;               // `draw` is d(CMVN output) of an utterance of <= CMN_WIN frames: the kernel applies the CMVN adjoint
                          // dx = dy - mean_t(dy) itself (model/iv_plda.py:296-377 is self-adjoint when every window is [0, T))
};

template <bool STASH>
__global__ void __launch_bounds__(FEAT_THREADS, 3)
mfcc_bwd_kernel(const float* __restrict__ x, int N, int m, int own_frames, DitherSpec D,
                const float* __restrict__ draw, int ld, BwdOut O, const SgFeatTables* __restrict__ gT,
                const float* __restrict__ stash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SgFeatTables* T = reinterpret_cast<SgFeatTables*>(smem_raw);
  float* fbase = reinterpret_cast<float*>(smem_raw + SG_FEAT_V1_BYTES);
  float* framebuf = fbase + FEAT_WARPS * WARP_SCRATCH;             // [8][400]: frame gradients of the current group
  float* acc = framebuf + FEAT_WARPS * SG_WIN;                     // [ACC_RING]: circular overlap-add accumulator, index = padded position & (ACC_RING - 1)
  __shared__ __align__(8) uint64_t ola_done;                       // split barrier: "group g's frame gradients have been consumed"
  copy_tables(T, gT);
  dither_resolve(D);
  for (int i = threadIdx.x; i < ACC_RING; i += FEAT_THREADS) acc[i] = 0.f;
  if (threadIdx.x == 0) mbar_init(&ola_done, FEAT_THREADS);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* scratch = fbase + warp * WARP_SCRATCH;
  float *sre = scratch, *sim = scratch + 288, *P = scratch + 576;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * own_frames;
  const int f1 = min(f0 + own_frames, m);
  const int fs = max(f0 - 2, 0);                                   // 2 halo frames on the left
  const int pmax = SG_SHIFT * (m - 1) + (SG_WIN - SG_HALO) - 1;    // last padded position touched
  const int own_lo = SG_SHIFT * f0 - SG_HALO;
  const int own_hi = (f1 == m) ? pmax + 1 : SG_SHIFT * f1 - SG_HALO;

  // fused CMVN adjoint: column means of the utterance's d(feat), summed in cmvn_colsum_order (identical bits to cmvn_kernel);
  // every CTA of the utterance recomputes them (m x 32 floats from L2) instead of a separate launch and a dRaw round trip
  float cm_mu = 0.f;
  if (O.cmvn) {
    const float* db = draw + (size_t)b * m * ld;
    const int nch = (m + CMVN_CHUNK - 1) / CMVN_CHUNK;
    const bool cin = lane < SG_NCEP && lane < ld;
    // all loads first (d(feat) was just written: L2 hits), then the sums in the fixed order
    float v[(CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK][CMVN_CHUNK / FEAT_WARPS];
#pragma unroll
    for (int c = 0; c < (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK; ++c)
#pragma unroll
      for (int i = 0; i < CMVN_CHUNK / FEAT_WARPS; ++i) {
        const int t = c * CMVN_CHUNK + warp + FEAT_WARPS * i;
        v[c][i] = (cin && t < m) ? __ldg(db + (size_t)t * ld + lane) : 0.f;
      }
#pragma unroll
    for (int c = 0; c < (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK; ++c) {
      float s0 = 0.f;
#pragma unroll
      for (int i = 0; i < CMVN_CHUNK / FEAT_WARPS; ++i) s0 += v[c][i];   // frames beyond m contribute +0.f (exact)
      if (c < nch) scratch[c * 32 + lane] = s0;
    }
    __syncthreads();
    float tot = 0.f;
    for (int c = 0; c < nch; ++c) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < FEAT_WARPS; ++w) a += fbase[w * WARP_SCRATCH + c * 32 + lane];
      tot += a;
    }
    cm_mu = tot / (float)m;
    __syncthreads();                                               // the scratch is reused by the frame computation
  }

  // Frame groups of up to FEAT_WARPS frames.  The remainder goes FIRST so that the last group is full (or the only one):
  // every sample whose right-reflected partner exists is then finalised in the last group, where the partner's sum is
  // complete.  The first group keeps >= 2 frames so that it finalises all samples n < 120, whose left-reflected partners
  // only need frame 0 (a remainder of 1 is split 5 + 4 with the following group).
  const int nfrm = f1 - fs, rem = nfrm % FEAT_WARPS;
  const int gs0 = nfrm <= FEAT_WARPS ? nfrm : (rem == 0 ? FEAT_WARPS : (rem == 1 ? 5 : rem));
  const int gs1 = (nfrm > FEAT_WARPS && rem == 1) ? 4 : FEAT_WARPS;
  int gi = 0;
  int gsz = gs0;
  for (int gf = fs; gf < f1; gf += gsz, gsz = (gi == 0 ? gs1 : FEAT_WARPS), ++gi) {
    const int fr = (warp < gsz) ? gf + warp : f1;                  // warps beyond the group idle (fr >= f1)
    float* const gbuf = framebuf;
    float* const mybuf = gbuf + warp * SG_WIN;
    // prefetch the iterate / clean waveform of the samples this thread will finalise after the group's frames
    // (the loads then complete under the FFT work instead of stalling the whole CTA in the finalise phase)
    float pre_x[6], pre_x0[6];
    if (O.mode == 1) {
      const int base0 = SG_SHIFT * gf - SG_HALO;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int n = base0 + threadIdx.x + i * FEAT_THREADS;
        const bool ok = (threadIdx.x + i * FEAT_THREADS < ACC_LEN) && n >= own_lo && n < own_hi && n >= 0 && n < N;
        pre_x[i] = ok ? __ldg(xb + n) : 0.f;
        pre_x0[i] = ok ? __ldg(O.x0 + (size_t)b * N + n) : 0.f;
      }
    }
    if (STASH) {                                                   // the frame this warp takes in the NEXT group: 4 KB of stash (HBM) -> L2
      const int nf = gf + gsz + warp;
      if (nf < f1) {
        prefetch_l2(stash + ((size_t)b * m + nf) * SG_STASH_FLOATS + 32 * lane);
        if (lane == 0) prefetch_l2(draw + ((size_t)b * m + nf) * ld);
      }
    }
    if (fr < f1) {
      // ---- the forward of this frame: read back from the stash, or recomputed ----------------
      Frame F;
      float me;
      if (STASH) {
        me = stash_load(stash + ((size_t)b * m + fr) * SG_STASH_FLOATS, F, lane);
      } else {
        load_frame(F, xb, N, b, m, fr, D, lane);
        frame_spectrum(F, T, sre, sim, P, lane);
        me = mel_energy(T, P, lane);
      }
      // ---- backward: cepstra -> log-mel -> mel -> power -> spectrum -----------------------
      const float dC = (lane < SG_NCEP) ? __ldg(draw + ((size_t)b * m + fr) * ld + lane) - cm_mu : 0.f;
      const float dE = __shfl_sync(0xffffffffu, dC, 0);            // C0 <- log-energy
      __syncwarp();
      P[lane] = dC;                                                // P is free after mel_energy: dC[32] (k >= 30 zero)
      __syncwarp();
      float dM = 0.f;
      {
        const float4* d4 = reinterpret_cast<const float4*>(&T->dct_nk[lane][0]);   // column 0 is zero (C0 <- log-energy)
        const float4* c4 = reinterpret_cast<const float4*>(P);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 d = d4[g], cc = c4[g];
          dM = fmaf(cc.x, d.x, dM); dM = fmaf(cc.y, d.y, dM); dM = fmaf(cc.z, d.z, dM); dM = fmaf(cc.w, d.w, dM);
        }
      }
      const float dmel = (lane < SG_NMEL && me > SG_EPS) ? dM / me : 0.f;
      __syncwarp();
      P[lane] = dmel;                                              // reuse as dmel[32]
      __syncwarp();
      float2 dX[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        float dP = T->bin_w0[k] * P[T->bin_c0[k]] + T->bin_w1[k] * P[T->bin_c1[k]];
        dX[i] = make_float2(2.f * dP * F.X[i].x, 2.f * dP * F.X[i].y);
        sre[k] = dX[i].x; sim[k] = dX[i].y;
      }
      __syncwarp();
      // adjoint of the untangle step (derivation + check: tools/fft_proto.py):
      //   dZ[k] = ((1-s) + i c)/2 * dX[k] + ((1+s) - i c)/2 * conj(dX[256-k]),  dZ[0] = 0
      float2 z[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = lane + 32 * i;
        const int km = (256 - k) & 255;
        float2 xm = make_float2(sre[km], -sim[km]);
        float2 cs = T->tw[16 + i][lane];
        float2 ca = make_float2(0.5f * (1.f - cs.y), 0.5f * cs.x);
        float2 cb = make_float2(0.5f * (1.f + cs.y), -0.5f * cs.x);
        float2 dz = cadd(cmul(ca, dX[i]), cmul(cb, xm));
        if (k == 0) dz = make_float2(0.f, 0.f);
        z[i] = make_float2(dz.x, -dz.y);                           // conj -> forward FFT -> conj = inverse
      }
      __syncwarp();
      warp_fft256(z, T->tw, sre, sim, lane);
      fft_out_to_smem(z, sre, sim, lane);
      __syncwarp();
      // back in sample ownership: dg[2n] = Re dz[n], dg[2n+1] = Im dz[n] (conj: -sim)
      float dge[7], dgo[7];
#pragma unroll
      for (int n0 = 0; n0 < 7; ++n0) {
        const bool valid = (n0 < 6) || (lane < 8);
        const int n = 32 * n0 + lane;
        float2 w2 = valid ? *reinterpret_cast<const float2*>(&T->window[64 * n0 + 2 * lane]) : make_float2(0.f, 0.f);
        dge[n0] = sre[n] * w2.x;
        dgo[n0] = -sim[n] * w2.y;
      }
      // pre-emphasis adjoint: g[j] = f[j] - 0.97 f[max(j-1,0)]
      float s = 0.f;
      const float esc = (F.sumsq > SG_EPS) ? 2.f * dE / F.sumsq : 0.f;   // d log(sum f^2)
      float dfe[7], dfo[7];
#pragma unroll
      for (int n0 = 0; n0 < 7; ++n0) {
        const bool valid = (n0 < 6) || (lane < 8);
        float dn = __shfl_down_sync(0xffffffffu, dge[n0], 1);      // even sample of lane+1
        float wrap = __shfl_sync(0xffffffffu, dge[n0 < 6 ? n0 + 1 : 6], 0);
        float next = lane < 31 ? dn : (n0 < 6 ? wrap : 0.f);       // dg of sample j+1 for the odd element
        if (n0 == 6 && lane == 7) next = 0.f;                      // j = 399 is the last sample
        float de = dge[n0] - 0.97f * dgo[n0];
        float dd = dgo[n0] - 0.97f * next;
        if (n0 == 0 && lane == 0) de -= 0.97f * dge[0];            // replicate pad at j = 0
        de = fmaf(esc, F.fe[n0], de);
        dd = fmaf(esc, F.fo[n0], dd);
        dfe[n0] = valid ? de : 0.f;
        dfo[n0] = valid ? dd : 0.f;
        s += dfe[n0] + dfo[n0];
      }
      const float mean = warp_sum(s) * (1.0f / SG_WIN);            // DC-removal adjoint
      // the previous group's frame gradients must have been summed by every thread before they are overwritten; the
      // arrivals happened right after that group's overlap-add, a whole frame computation ago: this wait is normally free
      if (gi > 0) mbar_wait(&ola_done, (uint32_t)((gi - 1) & 1));
#pragma unroll
      for (int n0 = 0; n0 < 7; ++n0)
        if ((n0 < 6) || (lane < 8))
          *reinterpret_cast<float2*>(&mybuf[64 * n0 + 2 * lane]) =
              make_float2((dfe[n0] - mean) * 32768.0f, (dfo[n0] - mean) * 32768.0f);
    }
    __syncthreads();
    // ---- deterministic overlap-add of this group's frames ------------------------------------
    // One full barrier per group (frame gradients complete) plus a split mbarrier (frame gradients consumed, waited for
    // just before the next group's gradients are stored).  The accumulator is a ring indexed by the padded sample
    // position, so a position is summed, finalised and cleared by one thread in one visit.  Positions >= base + fin stay in the ring
    // for the next group (its barrier orders the hand-over).  Only the utterance edges, whose reflected positions are summed
    // by other threads, need the two-phase form.
    const int base = SG_SHIFT * gf - SG_HALO;                      // padded position of ring slot q = 0 of this group
    const int nfr = gsz;
    const bool last = (gf + gsz >= f1);
    const int fin = last ? ACC_LEN : SG_SHIFT * gsz;               // positions no later group touches
    const bool edge = (gf == 0) || (last && f1 == m);              // CTA-uniform: left / right reflection partners in range
    float asum[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = threadIdx.x + i * FEAT_THREADS;
      asum[i] = 0.f;
      if (q >= ACC_LEN) continue;
      float a = acc[(base + q) & (ACC_RING - 1)];
      // frames w with 0 <= q - 160 w < 400 (at most 3: w0 - 2, w0 - 1, w0 = q / 160), added in increasing w (fixed
      // summation order); the three loads are independent, absent terms add +0.f (exact)
      const int w0 = q / SG_SHIFT, o0 = q - SG_SHIFT * w0;         // o0 in [0, 160)
      const float t2 = (w0 >= 2 && w0 - 2 < nfr && o0 + 2 * SG_SHIFT < SG_WIN) ? gbuf[(w0 - 2) * SG_WIN + o0 + 2 * SG_SHIFT] : 0.f;
      const float t1 = (w0 >= 1 && w0 - 1 < nfr) ? gbuf[(w0 - 1) * SG_WIN + o0 + SG_SHIFT] : 0.f;
      const float t0 = (w0 < nfr) ? gbuf[w0 * SG_WIN + o0] : 0.f;
      a += t2; a += t1; a += t0;
      asum[i] = a;
      acc[(base + q) & (ACC_RING - 1)] = (edge || q >= fin) ? a : 0.f;   // finalised slots are released for the ring's next lap
    }
    if (edge) __syncthreads();                                     // reflected partners are complete
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int q = threadIdx.x + i * FEAT_THREADS;
      if (q >= fin) continue;
      const int n = base + q;
      if (n < own_lo || n >= own_hi || n < 0 || n >= N) continue;
      float g = asum[i];
      if (edge) {
        if (n < SG_HALO) {                                         // left reflection: p = -n-1
          const int qm = (-n - 1) - base;
          if (qm >= 0 && qm < ACC_LEN) g += acc[(base + qm) & (ACC_RING - 1)];
        }
        const int pr = 2 * N - 1 - n;                              // right reflection
        if (pr <= pmax) {
          const int qm = pr - base;
          if (qm >= 0 && qm < ACC_LEN) g += acc[(base + qm) & (ACC_RING - 1)];
        }
      }
      const size_t gi = (size_t)b * N + n;
      if (O.mode == 0) {
        float v = O.scale * g;
        O.grad[gi] = O.accumulate ? O.grad[gi] + v : v;
      } else {
        const float xc = pre_x[i], x0 = pre_x0[i];
        const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
        float xn = xc + O.step * sg;                               // attack/FGSM.py:65
        const float lo = fmaxf(x0 - O.eps, -1.f), hi = fminf(x0 + O.eps, 1.f);   // attack/PGD.py:48-49
        O.x_out[gi] = fminf(fmaxf(xn, lo), hi);                    // attack/FGSM.py:68
      }
    }
    if (edge) {
      __syncthreads();                                             // every reflected read is done: release the finalised slots
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int q = threadIdx.x + i * FEAT_THREADS;
        if (q < fin) acc[(base + q) & (ACC_RING - 1)] = 0.f;
      }
    }
    mbar_arrive(&ola_done);                                        // this thread is done with the group's frame gradients
  }
}

// =============================================================================================
// dither materialisation, sign step
// =============================================================================================
__global__ void step_linf_kernel(float* __restrict__ x, const float* __restrict__ x0,
                                 const float* __restrict__ grad, size_t n, float step, float eps) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = grad[i], a = x0[i];
    const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
    const float xn = x[i] + step * sg;
    x[i] = fminf(fmaxf(xn, fmaxf(a - eps, -1.f)), fminf(a + eps, 1.f));
  }
}

// =============================================================================================
// CMVN (model/iv_plda.py:296-377): y[t] = x[t] - mean(x[ws(t):we(t)]), window 300 centred
// =============================================================================================
__device__ __forceinline__ void cmvn_window(int t, int T, int& ws, int& we) {
  ws = t - CMN_WIN / 2; we = ws + CMN_WIN;
  if (ws < 0) { we -= ws; ws = 0; }
  if (we > T) { ws -= (we - T); we = T; if (ws < 0) ws = 0; }
}

// one CTA per utterance, blockDim = (32 columns, 8 row lanes)
__global__ void cmvn_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int ld_out,
                            int T, int ncol, int backward) {
  __shared__ float part[(CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK][8][33];
  __shared__ float tot[32];
  const int lc = threadIdx.x, c = blockIdx.y * 32 + lc, r = threadIdx.y, b = blockIdx.x;
  const float* ib = in + (size_t)b * T * ld_in;
  float* ob = out + (size_t)b * T * ld_out;
  const bool cin = c < ncol && c < ld_in;
  if (T <= CMN_WIN) {
    // every window is [0,T): global mean.  Self-adjoint: dx = dy - mean(dy).
    const int nch = (T + CMVN_CHUNK - 1) / CMVN_CHUNK;             // cmvn_colsum_order
    float v[(CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK][CMVN_CHUNK / 8];   // this thread's rows, all loads in flight together
#pragma unroll
    for (int ch = 0; ch < (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK; ++ch)
#pragma unroll
      for (int i = 0; i < CMVN_CHUNK / 8; ++i) {
        const int t = ch * CMVN_CHUNK + r + 8 * i;
        v[ch][i] = (cin && t < T) ? ib[(size_t)t * ld_in + c] : 0.f;
      }
#pragma unroll
    for (int ch = 0; ch < (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK; ++ch) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < CMVN_CHUNK / 8; ++i) s += v[ch][i];         // rows beyond T contribute +0.f (exact)
      if (ch < nch) part[ch][r][lc] = s;
    }
    __syncthreads();
    if (r == 0) {
      float sum = 0.f;
      for (int ch = 0; ch < nch; ++ch) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += part[ch][i][lc];
        sum += a;
      }
      tot[lc] = sum / (float)T;
    }
    __syncthreads();
    const float mu = tot[lc];
    if (c < ld_out) {
#pragma unroll
      for (int ch = 0; ch < (CMN_WIN + CMVN_CHUNK - 1) / CMVN_CHUNK; ++ch)
#pragma unroll
        for (int i = 0; i < CMVN_CHUNK / 8; ++i) {
          const int t = ch * CMVN_CHUNK + r + 8 * i;
          if (t < T) ob[(size_t)t * ld_out + c] = cin ? v[ch][i] - mu : 0.f;
        }
    }
    return;
  }
  // T > 300: sliding window.  Serial per column (rare path: > 3 s utterances), r == 0 only.
  if (r != 0) return;
  if (!backward) {
    float cur = 0.f;
    int ls = -1, le = -1;
    for (int t = 0; t < T; ++t) {
      int ws, we;
      cmvn_window(t, T, ws, we);
      if (ls < 0) { for (int u = ws; u < we; ++u) cur += cin ? ib[(size_t)u * ld_in + c] : 0.f; }
      else {
        if (ws > ls) cur -= cin ? ib[(size_t)ls * ld_in + c] : 0.f;
        if (we > le) cur += cin ? ib[(size_t)le * ld_in + c] : 0.f;
      }
      ls = ws; le = we;
      if (c < ld_out) ob[(size_t)t * ld_out + c] = cin ? ib[(size_t)t * ld_in + c] - cur / (float)(we - ws) : 0.f;
    }
  } else {
    // dx[s] = dy[s] - (1/300) * sum_{t : ws(t) <= s < we(t)} dy[t]
    // t < 150 -> [0,300);  150 <= t <= T-150 -> [t-150,t+150);  t > T-150 -> [T-300,T)
    float head = 0.f, tail = 0.f;
    for (int t = 0; t < CMN_WIN / 2; ++t) head += cin ? ib[(size_t)t * ld_in + c] : 0.f;
    for (int t = T - CMN_WIN / 2 + 1; t < T; ++t) tail += cin ? ib[(size_t)t * ld_in + c] : 0.f;
    // middle set for s: t in [max(s-149,150), min(s+150, T-150)]
    float mid = 0.f;
    int lo = 150, hi = 149;                                        // current inclusive range (empty)
    for (int s = 0; s < T; ++s) {
      const int nlo = max(s - 149, 150), nhi = min(s + 150, T - 150);
      while (hi < nhi) { ++hi; mid += cin ? ib[(size_t)hi * ld_in + c] : 0.f; }
      while (lo < nlo) { mid -= cin ? ib[(size_t)lo * ld_in + c] : 0.f; ++lo; }
      float sum = mid + (s < CMN_WIN ? head : 0.f) + (s >= T - CMN_WIN ? tail : 0.f);
      if (c < ld_out) ob[(size_t)s * ld_out + c] = cin ? ib[(size_t)s * ld_in + c] - sum * (1.0f / CMN_WIN) : 0.f;
    }
  }
}

// T > 300, parallel form: per-column prefix sums (fp64, shared memory) turn every window sum into a difference, so all
// 8 row lanes work; used while (T+1) x 32 doubles fit (T <= 850), the serial path above handles longer utterances.
__global__ void cmvn_prefix_kernel(const float* __restrict__ in, int ld_in, float* __restrict__ out, int ld_out,
                                   int T, int ncol, int backward) {
  extern __shared__ double cm_prefix[];                             // [(T+1)][32]
  const int lc = threadIdx.x, c = blockIdx.y * 32 + lc, r = threadIdx.y, b = blockIdx.x;
  const float* ib = in + (size_t)b * T * ld_in;
  float* ob = out + (size_t)b * T * ld_out;
  const bool cin = c < ncol && c < ld_in;
  // prefix sums in fp64, eight segments in parallel (warp r owns frames [r * seg, (r + 1) * seg)): a local running sum per
  // segment, then the exclusive prefix of the eight segment totals is added in place.  (One warp walking all T frames was a
  // chain of T dependent fp64 adds with the CTA - alone on its SM: 128 KB of prefixes - waiting for it: 0.55 ms at B = 256,
  // T = 500, 72 columns.)
  __shared__ double seg_tot[8][32];
  const int seg = (T + 7) / 8, t0 = r * seg, t1 = min(T, t0 + seg);
  if (r == 0) cm_prefix[lc] = 0.0;
  {
    double acc = 0.0;
    int t = t0;
    for (; t + 8 <= t1; t += 8) {                                   // loads issued together, adds in order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = cin ? ib[(size_t)(t + u) * ld_in + c] : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc += (double)v[u]; cm_prefix[(size_t)(t + u + 1) * 32 + lc] = acc; }
    }
    for (; t < t1; ++t) { acc += cin ? (double)ib[(size_t)t * ld_in + c] : 0.0; cm_prefix[(size_t)(t + 1) * 32 + lc] = acc; }
    seg_tot[r][lc] = acc;
  }
  __syncthreads();
  {
    double off = 0.0;
    for (int q = 0; q < r; ++q) off += seg_tot[q][lc];
    if (r > 0) for (int t = t0; t < t1; ++t) cm_prefix[(size_t)(t + 1) * 32 + lc] += off;
  }
  __syncthreads();
  if (c >= ld_out) return;
  if (!backward) {
    for (int t = r; t < T; t += 8) {
      int ws, we;
      cmvn_window(t, T, ws, we);
      const double sum = cm_prefix[(size_t)we * 32 + lc] - cm_prefix[(size_t)ws * 32 + lc];
      ob[(size_t)t * ld_out + c] = cin ? ib[(size_t)t * ld_in + c] - (float)(sum / (double)(we - ws)) : 0.f;
    }
  } else {
    // dx[s] = dy[s] - (1/300) * sum_{t : ws(t) <= s < we(t)} dy[t]  (all windows have 300 frames when T > 300)
    // t < 150 -> [0,300);  150 <= t <= T-150 -> [t-150,t+150);  t > T-150 -> [T-300,T)
    const double head = cm_prefix[(size_t)(CMN_WIN / 2) * 32 + lc];
    const double tail = cm_prefix[(size_t)T * 32 + lc] - cm_prefix[(size_t)(T - CMN_WIN / 2 + 1) * 32 + lc];
    for (int s = r; s < T; s += 8) {
      const int lo = max(s - 149, 150), hi = min(s + 150, T - 150);
      double sum = (hi >= lo) ? cm_prefix[(size_t)(hi + 1) * 32 + lc] - cm_prefix[(size_t)lo * 32 + lc] : 0.0;
      if (s < CMN_WIN) sum += head;
      if (s >= T - CMN_WIN) sum += tail;
      ob[(size_t)s * ld_out + c] = cin ? ib[(size_t)s * ld_in + c] - (float)(sum * (1.0 / CMN_WIN)) : 0.f;
    }
  }
}

// =============================================================================================
// V2: one HALF-warp per frame (two frames per warp)
// =============================================================================================
// The V1 kernels above give a frame to a whole warp: 8 FFT elements per lane, a 256-point complex FFT as 8 x 8 x 4 with two
// shared-memory exchanges, and every table (twiddles, window, mel weights, DCT) read from shared memory once per frame.
// ncu (profiles/r2_mfcc_ncu_summary.json) puts them at 91 % of the L1/shared-memory pipe and 68 % issue utilisation at once:
// ~370 shared-memory wavefronts and ~1300 warp instructions per frame.  Here 16 lanes own a frame and hold 16 FFT elements
// each: 256 = 16 x 16 needs ONE exchange, the real-FFT untangle finds its mirror bin 256 - k in the partner lane 16 - l
// (a shuffle, no shared memory), and every table load is shared by the two frames of the warp (same address in both halves:
// a broadcast).  Ownership inside a half-warp, lane l = lane & 15:
//   samples      pair a in [0, 13): j = 32 a + 2 l (+1); a = 12 only for l < 8 (j < 400)        = FFT input element 16 a + l
//   spectrum     slot i in [0, 16): bin k = l + 16 i
//   mel filters  l and 29 - l (lane 15 idles), cepstra l and l + 16
#define F2_GROUP 16                        // frames per CTA iteration
#define F2_EX 272                          // floats per exchange plane: 16 rows x 17
#define F2_SMALL (4 * F2_EX)               // offset of the small per-frame rows: lm / dC [2][48], dmel [2][48]
#define F2_ROW 48                          // row stride of the small rows (32 used): the two half-warps land in different banks
#define F2_WARP_SCRATCH (4 * F2_EX + 4 * F2_ROW)
// plane order [re 0][re 1][im 0][im 1]: the two half-warps of an access are 272 = 16 (mod 32) floats apart, i.e. in disjoint
// banks; the power spectrum P (256 floats) reuses this half-warp's re plane once the FFT is done
#define F2_RE(scratch, hf) ((scratch) + (hf) * F2_EX)
#define F2_IM(scratch, hf) ((scratch) + (2 + (hf)) * F2_EX)
#define F2_ACC_LEN ((F2_GROUP - 1) * SG_SHIFT + SG_WIN)   // 2800: padded span of 16 consecutive frames
#define F2_ACC_RING 4096
#define F2_NQ ((F2_ACC_LEN + FEAT_THREADS - 1) / FEAT_THREADS)   // 11 positions per thread (scalar form)
#define F2_NQ4 ((F2_ACC_LEN / 4 + FEAT_THREADS - 1) / FEAT_THREADS)   // 3 float4 of positions per thread

__device__ __forceinline__ float2 cmulc(float2 a, float cr, float ci) { return make_float2(a.x * cr - a.y * ci, a.x * ci + a.y * cr); }

// natural-order in, natural-order out: v[k] = sum_n v[n] W_16^(nk); radix 4 x 4, constant twiddles
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
  for (int n0 = 0; n0 < 4; ++n0) fft4(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12]);
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
  v[5] = cmulc(v[5], c1, -s1);                                               // W^1
  v[9] = make_float2(h * (v[9].x + v[9].y), h * (v[9].y - v[9].x));          // W^2
  v[13] = cmulc(v[13], s1, -c1);                                             // W^3
  v[6] = make_float2(h * (v[6].x + v[6].y), h * (v[6].y - v[6].x));          // W^2
  v[10] = make_float2(v[10].y, -v[10].x);                                    // W^4 = -i
  v[14] = make_float2(h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y));    // W^6
  v[7] = cmulc(v[7], s1, -c1);                                               // W^3
  v[11] = make_float2(h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y));    // W^6
  v[15] = cmulc(v[15], -c1, s1);                                             // W^9
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) fft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
  float2 t[16];                                                              // v[4 k1 + k2] = X[k1 + 4 k2]: register renaming
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) t[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = t[k];
}

// 256-point complex FFT on 16 lanes.  In: z[a] = element 16 a + l.  Out: z[i] = Z[l + 16 i].  re / im: this half-warp's
// exchange planes (F2_EX floats each).  (index maps checked against numpy: tools/fft16_proto.py)
__device__ __forceinline__ void hw_fft256(float2 (&z)[16], const SgFeatTables2* T, float* re, float* im, int l) {
  // the two passes share one copy of fft16 in the instruction stream (the forward kernel's loop body is ~35 KB of code and
  // stalls on instruction fetch otherwise)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    fft16(z);
    if (pass == 0) {
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) z[k1] = cmul(z[k1], T->tw16[k1][l]);
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) { re[k1 * 17 + l] = z[k1].x; im[k1 * 17 + l] = z[k1].y; }
      __syncwarp();
#pragma unroll
      for (int b = 0; b < 16; ++b) z[b] = make_float2(re[l * 17 + b], im[l * 17 + b]);
      __syncwarp();
    }
  }
}

struct Frame2 {
  float fe[13], fo[13];  // DC-removed samples of pair a (even / odd)
  float sumsq;
};

__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load_frame2(Frame2& F, const float* __restrict__ xb, int N, int b, int m, int fr, const DitherSpec& D, int l) {
  const int p0 = fr * SG_SHIFT - SG_HALO + 2 * l;
  // the waveform loads go first: the dither arithmetic below covers their latency
  const int w0 = fr * SG_SHIFT - SG_HALO;
  const bool interior = (w0 >= 0) && (w0 + SG_WIN <= N);
  const bool vec = interior && ((reinterpret_cast<uintptr_t>(xb + w0) & 7) == 0);
#pragma unroll
  for (int a = 0; a < 13; ++a) {
    const bool valid = (a < 12) || (l < 8);
    float ve = 0.f, vo = 0.f;
    if (valid) {
      if (vec) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(xb + p0 + 32 * a));
        ve = v.x; vo = v.y;
      } else {
        int pe = p0 + 32 * a, po = pe + 1;
        pe = pe < 0 ? -pe - 1 : (pe >= N ? 2 * N - 1 - pe : pe);       // kaldi.py:69-77 reflect pad
        po = po < 0 ? -po - 1 : (po >= N ? 2 * N - 1 - po : po);
        ve = __ldg(xb + pe); vo = __ldg(xb + po);
      }
    }
    F.fe[a] = ve; F.fo[a] = vo;
  }
  float nz[26];
#pragma unroll
  for (int i = 0; i < 26; ++i) nz[i] = 0.f;
  if (D.mode == SG_DITHER_TENSOR) {
    const float* dp = D.tensor + ((size_t)b * m + fr) * SG_WIN + 2 * l;
#pragma unroll
    for (int a = 0; a < 13; ++a)
      if (a < 12 || l < 8) {
        const float2 d = *reinterpret_cast<const float2*>(dp + 32 * a);
        nz[2 * a] = d.x; nz[2 * a + 1] = d.y;
      }
  } else if (D.mode == SG_DITHER_PHILOX) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 r = philox4x32_10(make_uint4(q * 16 + l, (uint32_t)fr, dither_key(D, b), D.pass), make_uint2(D.seed_lo, D.seed_hi));
      const float2 n0 = box_muller16(r.x);
      nz[8 * q] = n0.x; nz[8 * q + 1] = n0.y;
      if (q < 3) {
        const float2 n1 = box_muller16(r.y), n2 = box_muller16(r.z), n3 = box_muller16(r.w);
        nz[8 * q + 2] = n1.x; nz[8 * q + 3] = n1.y; nz[8 * q + 4] = n2.x; nz[8 * q + 5] = n2.y; nz[8 * q + 6] = n3.x; nz[8 * q + 7] = n3.y;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 13; ++a) {
    const bool valid = (a < 12) || (l < 8);
    const float ve = valid ? fmaf(F.fe[a], 32768.0f, nz[2 * a]) : 0.f;      // model/utils.py:14, kaldi.py:181 (x * 2^15 is exact)
    const float vo = valid ? fmaf(F.fo[a], 32768.0f, nz[2 * a + 1]) : 0.f;
    F.fe[a] = ve; F.fo[a] = vo;
    s += ve + vo;
  }
  const float mean = half_sum(s) * (1.0f / SG_WIN);                    // kaldi.py:183-186
  float q = 0.f;
#pragma unroll
  for (int a = 0; a < 13; ++a) {
    const bool valid = (a < 12) || (l < 8);
    F.fe[a] = valid ? F.fe[a] - mean : 0.f;
    F.fo[a] = valid ? F.fo[a] - mean : 0.f;
    q += F.fe[a] * F.fe[a] + F.fo[a] * F.fo[a];
  }
  F.sumsq = half_sum(q);
}

// pre-emphasis + window + real FFT: spectrum bins l + 16 i in X, |X|^2 in P[0..255] (this half-warp's row, shared memory)
__device__ __forceinline__ void frame_spectrum2(const Frame2& F, float2 (&X)[16], const SgFeatTables2* T, float* re, float* im, float* P, int l) {
  float2 z[16];
#pragma unroll
  for (int a = 0; a < 13; ++a) {
    // previous sample of the even element: odd element of lane l-1 (same pair row), or of lane 15 of the previous row;
    // j == 0 replicates itself (kaldi.py:193-198)
    const float up = __shfl_up_sync(0xffffffffu, F.fo[a], 1, 16);
    const float wrap = __shfl_sync(0xffffffffu, a > 0 ? F.fo[a > 0 ? a - 1 : 0] : F.fe[0], 15, 16);
    const float prev = l > 0 ? up : (a > 0 ? wrap : F.fe[0]);
    const bool valid = (a < 12) || (l < 8);
    const float2 w2 = valid ? *reinterpret_cast<const float2*>(&T->window[32 * a + 2 * l]) : make_float2(0.f, 0.f);
    z[a] = make_float2((F.fe[a] - 0.97f * prev) * w2.x, (F.fo[a] - 0.97f * F.fe[a]) * w2.y);
  }
  z[13] = z[14] = z[15] = make_float2(0.f, 0.f);
  hw_fft256(z, T, re, im, l);
  // untangle: X[k] = a_k Z[k] + b_k conj(Z[256-k]),  a_k = ((1-s) - i c)/2, b_k = ((1+s) + i c)/2; the mirror bin
  // 256 - (l + 16 i) = (16 - l) + 16 (15 - i) lives in lane 16 - l, slot 15 - i (lane 0: own slot (16 - i) & 15)
  const int partner = (16 - l) & 15;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float mx = __shfl_sync(0xffffffffu, z[15 - i].x, partner, 16);
    float my = __shfl_sync(0xffffffffu, z[15 - i].y, partner, 16);
    if (l == 0) { mx = z[(16 - i) & 15].x; my = z[(16 - i) & 15].y; }
    const float2 zm = make_float2(mx, -my);
    const float2 cs = T->untw[i][l];
    const float2 a = make_float2(0.5f * (1.f - cs.y), -0.5f * cs.x);
    const float2 bq = make_float2(0.5f * (1.f + cs.y), 0.5f * cs.x);
    X[i] = cadd(cmul(a, z[i]), cmul(bq, zm));
    P[l + 16 * i] = X[i].x * X[i].x + X[i].y * X[i].y;                  // kaldi.py:616-618
  }
  __syncwarp();
}

// mel energies (before the log; kaldi.py:621-630) of filters l (me0) and 29 - l (me1); each filter is summed in the same
// order as in the V1 kernel (float4 groups left to right, one fma chain)
__device__ __forceinline__ void mel_energy2(const SgFeatTables2* T, const float* P, int l, float& me0, float& me1) {
  const int len0 = T->m2_len0[l], lo0 = T->m2_lo0[l], lo1s = T->m2_lo1s[l];
  float a0 = 0.f, a1 = 0.f;
  const int iters = T->m2_iters;
  for (int i = 0; i < iters; ++i) {
    const bool first = i < len0;
    const int off = min((first ? lo0 : lo1s) + 4 * i, 252);          // padding groups (zero weights) must still read finite values
    const float4 w = T->m2_w[i][l];
    const float4 p = *reinterpret_cast<const float4*>(&P[off]);
    float acc = first ? a0 : a1;
    acc = fmaf(w.x, p.x, acc); acc = fmaf(w.y, p.y, acc); acc = fmaf(w.z, p.z, acc); acc = fmaf(w.w, p.w, acc);
    if (first) a0 = acc; else a1 = acc;
  }
  me0 = a0; me1 = a1;
}

// ---- per-frame stash, V2 layout (floats): X as float4 {X[2j], X[2j+1]} [8][16] | f as float2 [13][16] | mel[32] | sumsq
#define S2_F 512
#define S2_MEL 928
#define S2_SUMSQ 960
__device__ __forceinline__ void stash2_store_f(float* __restrict__ sp, const Frame2& F, int l) {
  float2* s2 = reinterpret_cast<float2*>(sp + S2_F);
#pragma unroll
  for (int a = 0; a < 13; ++a) __stcs(s2 + a * 16 + l, make_float2(F.fe[a], F.fo[a]));
  if (l == 0) __stcs(sp + S2_SUMSQ, F.sumsq);
}
__device__ __forceinline__ void stash2_store_x(float* __restrict__ sp, const float2 (&X)[16], int l) {
  float4* s4 = reinterpret_cast<float4*>(sp);
#pragma unroll
  for (int j = 0; j < 8; ++j) __stcs(s4 + j * 16 + l, make_float4(X[2 * j].x, X[2 * j].y, X[2 * j + 1].x, X[2 * j + 1].y));
}
__device__ __forceinline__ void stash2_load_x(const float* __restrict__ sp, float2 (&X)[16], int l) {
  const float4* s4 = reinterpret_cast<const float4*>(sp);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = __ldcs(s4 + j * 16 + l);
    X[2 * j] = make_float2(v.x, v.y); X[2 * j + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void stash2_load_f(const float* __restrict__ sp, Frame2& F, int l) {
  const float2* s2 = reinterpret_cast<const float2*>(sp + S2_F);
#pragma unroll
  for (int a = 0; a < 13; ++a) { const float2 v = __ldcs(s2 + a * 16 + l); F.fe[a] = v.x; F.fo[a] = v.y; }
  F.sumsq = __ldcs(sp + S2_SUMSQ);
}

// PART 0: the forward's tables (a prefix of the struct), 1: the adjoint's (common part + tail), 2: everything.
// Thread 0 starts bulk asynchronous copies (cp.async.bulk, bytes counted by an mbarrier); the CTA goes on with its own
// set-up and calls wait_tables2 after its next __syncthreads().  (A load / store loop over 13 - 22 KB per CTA was 7 % of the
// adjoint's stall samples: four dependent L2 round trips per thread before the first frame.)
template <int PART>
__device__ __forceinline__ void stage_tables2(SgFeatTables2* dst, const SgFeatTables* __restrict__ src, uint64_t* bar) {
  if (threadIdx.x == 0) {
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const char* g = reinterpret_cast<const char*>(&src->v2);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t n0 = (uint32_t)(PART == 0 ? SG_T2_FWD_BYTES : (PART == 1 ? SG_T2_COMMON_BYTES : sizeof(SgFeatTables2)));
    const uint32_t n1 = PART == 1 ? (uint32_t)(sizeof(SgFeatTables2) - SG_T2_FWD_BYTES) : 0u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n0 + n1) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(g), "r"(n0), "r"(bar_a) : "memory");
    if (PART == 1)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(d + (uint32_t)SG_T2_FWD_BYTES), "l"(g + SG_T2_FWD_BYTES), "r"(n1), "r"(bar_a) : "memory");
  }
}
__device__ __forceinline__ void wait_tables2(uint64_t* bar) { mbar_wait(bar, 0); }
static_assert(SG_T2_FWD_BYTES % 16 == 0 && SG_T2_COMMON_BYTES % 16 == 0, "table sections must be int4-copyable");

// F1 (V2): waveform -> raw MFCC
template <int MINB>
__global__ void __launch_bounds__(FEAT_THREADS, MINB)
mfcc2_fwd_kernel(const float* __restrict__ x, int N, int m, int frames_per_cta, DitherSpec D, float* __restrict__ raw, int ld,
                 const SgFeatTables* __restrict__ gT, float* __restrict__ stash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SgFeatTables2* T = reinterpret_cast<SgFeatTables2*>(smem_raw);       // only the forward's prefix of the struct is resident
  __shared__ __align__(8) uint64_t tbar;
  stage_tables2<0>(T, gT, &tbar);
  dither_resolve(D);
  __syncthreads();
  wait_tables2(&tbar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hf = lane >> 4, l = lane & 15;
  float* scratch = reinterpret_cast<float*>(smem_raw + SG_T2_FWD_BYTES) + warp * F2_WARP_SCRATCH;
  float *re = F2_RE(scratch, hf), *im = F2_IM(scratch, hf), *P = re;
  float* lm = scratch + F2_SMALL + F2_ROW * hf;
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * frames_per_cta;
  const int f1 = min(f0 + frames_per_cta, m);
  for (int fb = f0 + 2 * warp; fb < f1; fb += F2_GROUP) {
    const bool active = fb + hf < f1;                                 // an odd tail: the upper half-warp recomputes the last frame
    const int fr = active ? fb + hf : f1 - 1;
    if (fb + F2_GROUP + hf < f1 && l < 14) {                           // next frame's 400 samples: <= 14 lines of 128 bytes
      int p = (fb + F2_GROUP + hf) * SG_SHIFT - SG_HALO + 32 * l;
      p = p < 0 ? 0 : (p >= N ? N - 1 : p);
      prefetch_l1(xb + p);
    }
    float* sp = stash != nullptr ? stash + ((size_t)b * m + fr) * SG_STASH_FLOATS : nullptr;
    float logE;
    float2 X[16];
    {
      Frame2 F;
      load_frame2(F, xb, N, b, m, fr, D, l);
      if (sp != nullptr && active) stash2_store_f(sp, F, l);
      logE = logf(fmaxf(F.sumsq, SG_EPS));                             // kaldi.py:119
      frame_spectrum2(F, X, T, re, im, P, l);
    }
    if (sp != nullptr && active) stash2_store_x(sp, X, l);
    float me0, me1;
    mel_energy2(T, P, l, me0, me1);
    if (sp != nullptr && active) {
      if (l < 15) { __stcs(sp + S2_MEL + l, me0); __stcs(sp + S2_MEL + SG_NMEL - 1 - l, me1); }
      else { __stcs(sp + S2_MEL + 30, 0.f); __stcs(sp + S2_MEL + 31, 0.f); }
    }
    if (l < 15) { lm[l] = logf(fmaxf(me0, SG_EPS)); lm[SG_NMEL - 1 - l] = logf(fmaxf(me1, SG_EPS)); }   // kaldi.py:631-633
    else { lm[30] = 0.f; lm[31] = 0.f; }
    __syncwarp();
    float c0 = 0.f, c1 = 0.f;
    {
      const float4* d0 = reinterpret_cast<const float4*>(&T->dct_kn[l][0]);
      const float4* d1 = reinterpret_cast<const float4*>(&T->dct_kn[l + 16][0]);
      const float4* l4 = reinterpret_cast<const float4*>(lm);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 v = l4[g], a = d0[g], bb = d1[g];
        c0 = fmaf(v.x, a.x, c0); c0 = fmaf(v.y, a.y, c0); c0 = fmaf(v.z, a.z, c0); c0 = fmaf(v.w, a.w, c0);
        c1 = fmaf(v.x, bb.x, c1); c1 = fmaf(v.y, bb.y, c1); c1 = fmaf(v.z, bb.z, c1); c1 = fmaf(v.w, bb.w, c1);
      }
    }
    if (l == 0) c0 = logE;                                             // kaldi.py:799-800
    if (l + 16 >= SG_NCEP) c1 = 0.f;
    if (active) {
      float* o = raw + ((size_t)b * m + fr) * ld;
      o[l] = c0;
      if (l + 16 < ld) o[l + 16] = c1;
    }
    __syncwarp();
  }
}

// F2 (V2): d(raw MFCC) -> d(waveform), optionally fused with the L-inf sign step.  Frame groups of up to 16 frames (one per
// half-warp); overlap-add, finalisation and the split barrier as in mfcc_bwd_kernel, with the ring sized for 16 frames.
template <bool STASH>
__global__ void __launch_bounds__(FEAT_THREADS, 2)
mfcc2_bwd_kernel(const float* __restrict__ x, int N, int m, int own_frames, DitherSpec D, const float* __restrict__ draw, int ld, BwdOut O,
                 const SgFeatTables* __restrict__ gT, const float* __restrict__ stash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SgFeatTables2* T = reinterpret_cast<SgFeatTables2*>(smem_raw);
  float* fbase = reinterpret_cast<float*>(smem_raw + sizeof(SgFeatTables2));
  float* framebuf = fbase + FEAT_WARPS * F2_WARP_SCRATCH;          // [16][400]: frame gradients of the current group
  float* acc = framebuf + F2_GROUP * SG_WIN;                       // [F2_ACC_RING]: circular overlap-add accumulator
  __shared__ __align__(8) uint64_t ola_done;
  __shared__ __align__(8) uint64_t tbar;
  stage_tables2<STASH ? 1 : 2>(T, gT, &tbar);
  dither_resolve(D);
  for (int i = threadIdx.x; i < F2_ACC_RING; i += FEAT_THREADS) acc[i] = 0.f;
  if (threadIdx.x == 0) mbar_init(&ola_done, FEAT_THREADS);
  __syncthreads();
  wait_tables2(&tbar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hf = lane >> 4, l = lane & 15;
  float* scratch = fbase + warp * F2_WARP_SCRATCH;
  float *re = F2_RE(scratch, hf), *im = F2_IM(scratch, hf), *P = re;
  float* rowc = scratch + F2_SMALL + F2_ROW * hf;                  // dC[32]
  float* rowm = scratch + F2_SMALL + F2_ROW * (2 + hf);            // dmel[32]
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * N;
  const int f0 = blockIdx.x * own_frames;
  const int f1 = min(f0 + own_frames, m);
  const int fs = max(f0 - 2, 0);                                   // 2 halo frames on the left
  const int pmax = SG_SHIFT * (m - 1) + (SG_WIN - SG_HALO) - 1;    // last padded position touched
  const int own_lo = SG_SHIFT * f0 - SG_HALO;
  const int own_hi = (f1 == m) ? pmax + 1 : SG_SHIFT * f1 - SG_HALO;
  // four samples per thread in the overlap-add / finalise phase when every row of the waveform tensors is 16-byte aligned
  const bool al4 = ((N & 3) == 0) && (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(O.x0) | reinterpret_cast<uintptr_t>(O.x_out) |
                                        reinterpret_cast<uintptr_t>(O.grad)) & 15) == 0);

  // (the fused CMVN adjoint of mfcc_bwd_kernel is not built here: the host keeps the V1 kernels when SG_OPT_CMVN_FUSION is on)
  const float mu0 = 0.f, mu1 = 0.f;

  // Frame groups of up to 16 frames; remainder first, first group >= 2 frames (see mfcc_bwd_kernel)
  const int nfrm = f1 - fs, rem = nfrm % F2_GROUP;
  const int gs0 = nfrm <= F2_GROUP ? nfrm : (rem == 0 ? F2_GROUP : (rem == 1 ? F2_GROUP / 2 + 1 : rem));
  const int gs1 = (nfrm > F2_GROUP && rem == 1) ? F2_GROUP / 2 : F2_GROUP;
  int gi = 0;
  int gsz = gs0;
  for (int gf = fs; gf < f1; gf += gsz, gsz = (gi == 0 ? gs1 : F2_GROUP), ++gi) {
    const int slot = 2 * warp + hf;
    const bool active = slot < gsz;
    const int base = SG_SHIFT * gf - SG_HALO;                      // padded position of ring slot q = 0 of this group
    const bool last = (gf + gsz >= f1);
    const int fin = last ? F2_ACC_LEN : SG_SHIFT * gsz;            // positions no later group touches
    const bool edge = (gf == 0) || (last && f1 == m);              // CTA-uniform: left / right reflection partners in range
    const bool vec = al4 && !edge;
    const int fr = active ? gf + slot : gf + gsz - 1;              // idle half-warps shadow the group's last frame (results dropped)
    float* const mybuf = framebuf + slot * SG_WIN;
    if (STASH) {                                                   // the frame this half-warp takes in the NEXT group: stash (HBM) -> L2
      const int nf = gf + gsz + slot;
      if (nf < f1) {
        prefetch_l2(stash + ((size_t)b * m + nf) * SG_STASH_FLOATS + 32 * l);
        prefetch_l2(stash + ((size_t)b * m + nf) * SG_STASH_FLOATS + 512 + 32 * l);
        if (l == 0) prefetch_l2(draw + ((size_t)b * m + nf) * ld);
      }
    }
    const bool work = 2 * warp < gsz;                              // warp-uniform: at least the lower half-warp has a frame
    const float* sp = STASH ? stash + ((size_t)b * m + fr) * SG_STASH_FLOATS : nullptr;
    float2 z[16];
    Frame2 F;
    float dE = 0.f;
    if (work) {
      float2 X[16];
      float me0, me1;
      if (STASH) {
        stash2_load_x(sp, X, l);
        me0 = __ldcs(sp + S2_MEL + l); me1 = __ldcs(sp + S2_MEL + l + 16);   // filters l, l + 16
      } else {
        load_frame2(F, xb, N, b, m, fr, D, l);
        frame_spectrum2(F, X, T, re, im, P, l);
        float ma, mb;
        mel_energy2(T, P, l, ma, mb);                               // filters l and 29 - l
        __syncwarp();
        if (l < 15) { rowm[l] = ma; rowm[SG_NMEL - 1 - l] = mb; } else { rowm[30] = 0.f; rowm[31] = 0.f; }
        __syncwarp();
        me0 = rowm[l]; me1 = rowm[l + 16];
        __syncwarp();
      }
      // ---- backward: cepstra -> log-mel -> mel -> power -> spectrum -----------------------
      const float* dr = draw + ((size_t)b * m + fr) * ld;
      const float dC0 = __ldg(dr + l) - mu0;
      const float dC1 = (l + 16 < SG_NCEP) ? __ldg(dr + l + 16) - mu1 : 0.f;
      dE = __shfl_sync(0xffffffffu, dC0, 0, 16);                    // C0 <- log-energy
      rowc[l] = dC0; rowc[l + 16] = dC1;
      __syncwarp();
      float dM0 = 0.f, dM1 = 0.f;
      {
        const float4* d0 = reinterpret_cast<const float4*>(&T->dct_nk[l][0]);        // column 0 is zero (C0 <- log-energy)
        const float4* d1 = reinterpret_cast<const float4*>(&T->dct_nk[l + 16][0]);
        const float4* c4 = reinterpret_cast<const float4*>(rowc);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 cc = c4[g], a = d0[g], bb = d1[g];
          dM0 = fmaf(cc.x, a.x, dM0); dM0 = fmaf(cc.y, a.y, dM0); dM0 = fmaf(cc.z, a.z, dM0); dM0 = fmaf(cc.w, a.w, dM0);
          dM1 = fmaf(cc.x, bb.x, dM1); dM1 = fmaf(cc.y, bb.y, dM1); dM1 = fmaf(cc.z, bb.z, dM1); dM1 = fmaf(cc.w, bb.w, dM1);
        }
      }
      rowm[l] = (me0 > SG_EPS) ? dM0 / me0 : 0.f;
      rowm[l + 16] = (l + 16 < SG_NMEL && me1 > SG_EPS) ? dM1 / me1 : 0.f;
      __syncwarp();
      float2 dX[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = l + 16 * i;
        const float2 w = T->binw[k];
        const int c = T->binc[k];
        const float dP = w.x * rowm[c & 255] + w.y * rowm[c >> 8];
        dX[i] = make_float2(2.f * dP * X[i].x, 2.f * dP * X[i].y);
      }
      // adjoint of the untangle step: dZ[k] = ((1-s) + i c)/2 * dX[k] + ((1+s) - i c)/2 * conj(dX[256-k]),  dZ[0] = 0
      const int partner = (16 - l) & 15;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float mx = __shfl_sync(0xffffffffu, dX[15 - i].x, partner, 16);
        float my = __shfl_sync(0xffffffffu, dX[15 - i].y, partner, 16);
        if (l == 0) { mx = dX[(16 - i) & 15].x; my = dX[(16 - i) & 15].y; }
        const float2 xm = make_float2(mx, -my);
        const float2 cs = T->untw[i][l];
        const float2 ca = make_float2(0.5f * (1.f - cs.y), 0.5f * cs.x);
        const float2 cb = make_float2(0.5f * (1.f + cs.y), -0.5f * cs.x);
        float2 dz = cadd(cmul(ca, dX[i]), cmul(cb, xm));
        if (i == 0 && l == 0) dz = make_float2(0.f, 0.f);
        z[i] = make_float2(dz.x, -dz.y);                            // conj -> forward FFT -> conj = inverse
      }
      __syncwarp();
      if (STASH) stash2_load_f(sp, F, l);                           // needed after the FFT: in flight during it
      hw_fft256(z, T, re, im, l);
    }
    // the iterate / clean waveform of the samples this thread finalises after the group's frames: requested now, so that
    // the loads complete under the window / pre-emphasis work and the barrier instead of stalling the finalise phase
    float4 px[F2_NQ4], px0[F2_NQ4];
#pragma unroll
    for (int i = 0; i < F2_NQ4; ++i) px[i] = px0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (O.mode == 1 && vec) {
#pragma unroll
      for (int i = 0; i < F2_NQ4; ++i) {
        const int q = 4 * (threadIdx.x + i * FEAT_THREADS), n = base + q;
        if (q < fin && n >= own_lo && n < own_hi && n < N) {           // (n >= 0: interior groups start at a positive position)
          px[i] = __ldg(reinterpret_cast<const float4*>(xb + n));
          px0[i] = __ldg(reinterpret_cast<const float4*>(O.x0 + (size_t)b * N + n));
        }
      }
    }
    if (work) {
      // back in sample ownership: pair a = element 16 a + l: dg[2n] = Re, dg[2n+1] = -Im of the transform
      float dge[13], dgo[13];
#pragma unroll
      for (int a = 0; a < 13; ++a) {
        const bool valid = (a < 12) || (l < 8);
        const float2 w2 = valid ? *reinterpret_cast<const float2*>(&T->window[32 * a + 2 * l]) : make_float2(0.f, 0.f);
        dge[a] = z[a].x * w2.x;
        dgo[a] = -z[a].y * w2.y;
      }
      // pre-emphasis adjoint: g[j] = f[j] - 0.97 f[max(j-1,0)]
      float s = 0.f;
      const float esc = (F.sumsq > SG_EPS) ? 2.f * dE / F.sumsq : 0.f;    // d log(sum f^2)
      float dfe[13], dfo[13];
#pragma unroll
      for (int a = 0; a < 13; ++a) {
        const bool valid = (a < 12) || (l < 8);
        const float dn = __shfl_down_sync(0xffffffffu, dge[a], 1, 16);   // even sample of lane l+1
        const float wrap = __shfl_sync(0xffffffffu, dge[a < 12 ? a + 1 : 12], 0, 16);
        float next = l < 15 ? dn : (a < 12 ? wrap : 0.f);           // dg of sample j+1 for the odd element
        if (a == 12 && l == 7) next = 0.f;                          // j = 399 is the last sample
        float de = dge[a] - 0.97f * dgo[a];
        float dd = dgo[a] - 0.97f * next;
        if (a == 0 && l == 0) de -= 0.97f * dge[0];                 // replicate pad at j = 0
        de = fmaf(esc, F.fe[a], de);
        dd = fmaf(esc, F.fo[a], dd);
        dfe[a] = valid ? de : 0.f;
        dfo[a] = valid ? dd : 0.f;
        s += dfe[a] + dfo[a];
      }
      const float mean = half_sum(s) * (1.0f / SG_WIN);            // DC-removal adjoint
      if (gi > 0) mbar_wait(&ola_done, (uint32_t)((gi - 1) & 1));  // the previous group's frame gradients have been consumed
      if (active) {
#pragma unroll
        for (int a = 0; a < 13; ++a)
          if ((a < 12) || (l < 8))
            *reinterpret_cast<float2*>(&mybuf[32 * a + 2 * l]) = make_float2((dfe[a] - mean) * 32768.0f, (dfo[a] - mean) * 32768.0f);
      }
    }
    __syncthreads();
    // ---- deterministic overlap-add of this group's frames (see mfcc_bwd_kernel) ---------------
    const int nfr = gsz;
    if (vec) {
      // interior group, aligned rows: every quantity of the scalar form below is constant over 4 consecutive positions
      // (SG_SHIFT, SG_WIN, SG_HALO, the group bounds and N are multiples of 4), so a thread takes a float4 of positions; each
      // position still receives its (up to) three terms in increasing frame order: identical sums.
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < F2_NQ4; ++i) {
        const int q = 4 * (threadIdx.x + i * FEAT_THREADS);
        if (q >= F2_ACC_LEN) continue;
        float4* ap = reinterpret_cast<float4*>(&acc[(base + q) & (F2_ACC_RING - 1)]);
        float4 a4 = *ap;
        const int w0 = q / SG_SHIFT, o0 = q - SG_SHIFT * w0;
        const float4 t2 = (w0 >= 2 && w0 - 2 < nfr && o0 + 2 * SG_SHIFT < SG_WIN) ? *reinterpret_cast<const float4*>(&framebuf[(w0 - 2) * SG_WIN + o0 + 2 * SG_SHIFT]) : zero4;
        const float4 t1 = (w0 >= 1 && w0 - 1 < nfr) ? *reinterpret_cast<const float4*>(&framebuf[(w0 - 1) * SG_WIN + o0 + SG_SHIFT]) : zero4;
        const float4 t0 = (w0 < nfr) ? *reinterpret_cast<const float4*>(&framebuf[w0 * SG_WIN + o0]) : zero4;
        a4.x += t2.x; a4.x += t1.x; a4.x += t0.x;
        a4.y += t2.y; a4.y += t1.y; a4.y += t0.y;
        a4.z += t2.z; a4.z += t1.z; a4.z += t0.z;
        a4.w += t2.w; a4.w += t1.w; a4.w += t0.w;
        *ap = (q >= fin) ? a4 : zero4;
        const int n = base + q;
        if (q >= fin || n < own_lo || n >= own_hi || n >= N) continue;
        const size_t gidx = (size_t)b * N + n;
        if (O.mode == 0) {
          float4 v = make_float4(O.scale * a4.x, O.scale * a4.y, O.scale * a4.z, O.scale * a4.w);
          float4* gp = reinterpret_cast<float4*>(O.grad + gidx);
          if (O.accumulate) { const float4 o = *gp; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *gp = v;
        } else {
          const float g[4] = {a4.x, a4.y, a4.z, a4.w};
          const float xc[4] = {px[i].x, px[i].y, px[i].z, px[i].w}, xz[4] = {px0[i].x, px0[i].y, px0[i].z, px0[i].w};
          float r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float sg = (g[e] > 0.f) ? 1.f : ((g[e] < 0.f) ? -1.f : 0.f);
            const float xn = xc[e] + O.step * sg;                     // attack/FGSM.py:65
            const float lo = fmaxf(xz[e] - O.eps, -1.f), hi = fminf(xz[e] + O.eps, 1.f);   // attack/PGD.py:48-49
            r[e] = fminf(fmaxf(xn, lo), hi);                          // attack/FGSM.py:68
          }
          *reinterpret_cast<float4*>(O.x_out + gidx) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    } else {
    float asum[F2_NQ];
#pragma unroll
    for (int i = 0; i < F2_NQ; ++i) {
      const int q = threadIdx.x + i * FEAT_THREADS;
      asum[i] = 0.f;
      if (q >= F2_ACC_LEN) continue;
      float a = acc[(base + q) & (F2_ACC_RING - 1)];
      const int w0 = q / SG_SHIFT, o0 = q - SG_SHIFT * w0;
      const float t2 = (w0 >= 2 && w0 - 2 < nfr && o0 + 2 * SG_SHIFT < SG_WIN) ? framebuf[(w0 - 2) * SG_WIN + o0 + 2 * SG_SHIFT] : 0.f;
      const float t1 = (w0 >= 1 && w0 - 1 < nfr) ? framebuf[(w0 - 1) * SG_WIN + o0 + SG_SHIFT] : 0.f;
      const float t0 = (w0 < nfr) ? framebuf[w0 * SG_WIN + o0] : 0.f;
      a += t2; a += t1; a += t0;
      asum[i] = a;
      acc[(base + q) & (F2_ACC_RING - 1)] = (edge || q >= fin) ? a : 0.f;
    }
    if (edge) __syncthreads();
#pragma unroll
    for (int i = 0; i < F2_NQ; ++i) {
      const int q = threadIdx.x + i * FEAT_THREADS;
      if (q >= fin) continue;
      const int n = base + q;
      if (n < own_lo || n >= own_hi || n < 0 || n >= N) continue;
      float g = asum[i];
      if (edge) {
        if (n < SG_HALO) {
          const int qm = (-n - 1) - base;
          if (qm >= 0 && qm < F2_ACC_LEN) g += acc[(base + qm) & (F2_ACC_RING - 1)];
        }
        const int pr = 2 * N - 1 - n;
        if (pr <= pmax) {
          const int qm = pr - base;
          if (qm >= 0 && qm < F2_ACC_LEN) g += acc[(base + qm) & (F2_ACC_RING - 1)];
        }
      }
      const size_t gidx = (size_t)b * N + n;
      if (O.mode == 0) {
        const float v = O.scale * g;
        O.grad[gidx] = O.accumulate ? O.grad[gidx] + v : v;
      } else {
        const float xc = __ldg(xb + n), x0 = __ldg(O.x0 + gidx);
        const float sg = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
        const float xn = xc + O.step * sg;                          // attack/FGSM.py:65
        const float lo = fmaxf(x0 - O.eps, -1.f), hi = fminf(x0 + O.eps, 1.f);   // attack/PGD.py:48-49
        O.x_out[gidx] = fminf(fmaxf(xn, lo), hi);                   // attack/FGSM.py:68
      }
    }
    if (edge) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < F2_NQ; ++i) {
        const int q = threadIdx.x + i * FEAT_THREADS;
        if (q < fin) acc[(base + q) & (F2_ACC_RING - 1)] = 0.f;
      }
    }
    }
    mbar_arrive(&ola_done);
  }
}

__global__ void dither_fill2_kernel(int m, DitherSpec D, float* __restrict__ out) {
  dither_resolve(D);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l = lane & 15;
  const int fr = (blockIdx.x * (blockDim.x >> 5) + warp) * 2 + (lane >> 4), b = blockIdx.y;
  if (fr >= m) return;
  float* o = out + ((size_t)b * m + fr) * SG_WIN;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 r = philox4x32_10(make_uint4(q * 16 + l, (uint32_t)fr, dither_key(D, b), D.pass), make_uint2(D.seed_lo, D.seed_hi));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = 32 * (4 * q + u) + 2 * l;
      if (j < SG_WIN) { const float2 n = box_muller16(w[u]); o[j] = n.x; o[j + 1] = n.y; }
    }
  }
}

// =============================================================================================
// host launchers (called from sg_api.cu)
// =============================================================================================
static size_t feat_fwd_smem() { return SG_FEAT_V1_BYTES + FEAT_WARPS * WARP_SCRATCH * sizeof(float); }
static size_t feat_bwd_smem() {
  return SG_FEAT_V1_BYTES + (FEAT_WARPS * WARP_SCRATCH + FEAT_WARPS * SG_WIN + ACC_RING) * sizeof(float);
}

// Device control block {pass, seed_lo, seed_hi} for the launches that follow on this thread (set around the captured /
// replayed PGD iteration by sg_api.cu; null = use the immediate seed / pass arguments).
static thread_local const uint32_t* g_feat_ctl = nullptr;
void sg_feat_set_ctl(const uint32_t* ctl) { g_feat_ctl = ctl; }
static thread_local uint32_t g_feat_copy_rows = 0;    // DitherSpec::copy_rows for the launches that follow on this thread
void sg_feat_set_copy_rows(int rows) { g_feat_copy_rows = rows > 0 ? (uint32_t)rows : 0u; }

// `pass` carries the pass counter in its low 32 bits and the handle's utterance offset in the high 32 (packed by
// sg_api.cu: dither_pass())
static DitherSpec make_dither(int mode, const float* tensor, uint64_t seed, uint64_t pass) {
  DitherSpec D;
  D.mode = mode; D.tensor = tensor;
  D.seed_lo = (uint32_t)seed; D.seed_hi = (uint32_t)(seed >> 32); D.pass = (uint32_t)pass; D.b_off = (uint32_t)(pass >> 32);
  D.ctl = g_feat_ctl; D.copy_rows = g_feat_copy_rows;
  return D;
}
__global__ void feat_ctl_init_kernel(uint32_t* ctl, uint32_t pass, uint32_t seed_lo, uint32_t seed_hi) {
  ctl[0] = pass; ctl[1] = seed_lo; ctl[2] = seed_hi;
}
__global__ void feat_ctl_tick_kernel(uint32_t* ctl, uint32_t n) { ctl[0] += n; }
int sg_feat_ctl_init_launch(uint32_t* ctl, uint64_t seed, uint32_t pass, cudaStream_t st) {
  feat_ctl_init_kernel<<<1, 1, 0, st>>>(ctl, pass, (uint32_t)seed, (uint32_t)(seed >> 32));
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_feat_ctl_tick_launch(uint32_t* ctl, uint32_t n, cudaStream_t st) {
  feat_ctl_tick_kernel<<<1, 1, 0, st>>>(ctl, n);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// resident CTAs per SM the forward kernel is compiled for: 3 (80 registers, the default: measured 60.8 vs 61.5 ms per step)
// or 4 (64 registers, a few spilled words);
// SGB200_FEAT_OCC selects (A/B switch)
static int g_fwd_occ = 3;
// SGB200_FEAT_V2 (default 1): the half-warp-per-frame kernels (mfcc2_*); 0 keeps the warp-per-frame ones.  The forward ->
// adjoint stash layout belongs to the kernel generation, so forward and adjoint always come from the same one; launches with
// the fused CMVN (SG_OPT_CMVN_FUSION) use the V1 kernels.
static int g_feat_v2 = 1;
static int g_fwd2_fpc = 64;      // SGB200_FEAT2_FPC: frames per CTA of the V2 forward (multiple of 16)
static int g_bwd2_chunks = 0;    // SGB200_FEAT2_CHUNKS: CTAs per utterance of the V2 adjoint (0 = heuristic)
static int g_fwd2_occ = 3;       // SGB200_FEAT2_OCC: resident CTAs per SM the V2 forward is compiled for (2: 128 registers, 3: 80)
static size_t feat2_fwd_smem() { return SG_T2_FWD_BYTES + FEAT_WARPS * F2_WARP_SCRATCH * sizeof(float); }
static size_t feat2_bwd_smem() { return sizeof(SgFeatTables2) + (FEAT_WARPS * F2_WARP_SCRATCH + F2_GROUP * SG_WIN + F2_ACC_RING) * sizeof(float); }

int sg_feat_init() {
  if (const char* e = getenv("SGB200_FEAT_OCC")) g_fwd_occ = atoi(e) == 4 ? 4 : 3;
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_fwd_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_fwd_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_fwd_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_fwd_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_bwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat_bwd_smem()));
  if (const char* e = getenv("SGB200_FEAT_V2")) g_feat_v2 = atoi(e) != 0;
  if (const char* e = getenv("SGB200_FEAT2_OCC")) g_fwd2_occ = atoi(e) == 2 ? 2 : 3;
  if (const char* e = getenv("SGB200_FEAT2_FPC")) { const int v = atoi(e); if (v >= 16 && v % 16 == 0) g_fwd2_fpc = v; }
  if (const char* e = getenv("SGB200_FEAT2_CHUNKS")) { const int v = atoi(e); if (v >= 0 && v <= 16) g_bwd2_chunks = v; }
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc2_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat2_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc2_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat2_fwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc2_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat2_bwd_smem()));
  SG_CUDA_CHECK(cudaFuncSetAttribute(mfcc2_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)feat2_bwd_smem()));
  return SG_OK;
}

int sg_feat_cmvn_fusable(int m) { return m >= 1 && m <= CMN_WIN; }

// fused-CMVN scratch for B utterances: [B][CMVN_MAX_CHUNKS][32] floats, then [B] counters that must be ZERO before the first launch
size_t sg_feat_cmvn_part_floats(int B) { return (size_t)B * CMVN_MAX_CHUNKS * 32; }

// cmvn != 0 (requires sg_feat_cmvn_fusable(m)): `raw` receives the CMVN output
int sg_feat_fwd_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode, const float* dither,
                       uint64_t seed, uint64_t pass, float* raw, int ld, cudaStream_t st, float* stash, int cmvn,
                       float* cmvn_part, unsigned int* cmvn_count) {
  if (cmvn) {
    if (!sg_feat_cmvn_fusable(m)) { sg_set_error("sg_feat_fwd_launch: fused CMVN needs m <= %d frames (got %d)", CMN_WIN, m); return SG_EINVAL; }
    if (!cmvn_part || !cmvn_count) { sg_set_error("sg_feat_fwd_launch: fused CMVN needs its scratch (sg_feat_cmvn_scratch_bytes)"); return SG_EINVAL; }
    const int nch = (m + CMVN_CHUNK - 1) / CMVN_CHUNK;            // <= 5
    CmvnScratch cm; cm.part = cmvn_part; cm.count = cmvn_count;
    dim3 grid(nch, B);
    if (g_fwd_occ == 4)
      mfcc_fwd_kernel<true, 4><<<grid, FEAT_THREADS, feat_fwd_smem(), st>>>(x, N, m, (int)CMVN_CHUNK, make_dither(mode, dither, seed, pass), raw, ld, dT, stash, cm);
    else
      mfcc_fwd_kernel<true, 3><<<grid, FEAT_THREADS, feat_fwd_smem(), st>>>(x, N, m, (int)CMVN_CHUNK, make_dither(mode, dither, seed, pass), raw, ld, dT, stash, cm);
    SG_LAUNCH_CHECK();
    return SG_OK;
  }
  if (g_feat_v2) {
    int fpc = g_fwd2_fpc;
    while (fpc > 16 && (long long)B * ((m + fpc - 1) / fpc) < 592) fpc >>= 1;
    dim3 grid((m + fpc - 1) / fpc, B);
    if (g_fwd2_occ == 2) mfcc2_fwd_kernel<2><<<grid, FEAT_THREADS, feat2_fwd_smem(), st>>>(x, N, m, fpc, make_dither(mode, dither, seed, pass), raw, ld, dT, stash);
    else mfcc2_fwd_kernel<3><<<grid, FEAT_THREADS, feat2_fwd_smem(), st>>>(x, N, m, fpc, make_dither(mode, dither, seed, pass), raw, ld, dT, stash);
    SG_LAUNCH_CHECK();
    return SG_OK;
  }
  CmvnScratch cm; cm.part = nullptr; cm.count = nullptr;
  int fpc = 64;
  while (fpc > 8 && (long long)B * ((m + fpc - 1) / fpc) < 592) fpc >>= 1;   // >= 4 CTAs per SM when possible
  dim3 grid((m + fpc - 1) / fpc, B);
  if (g_fwd_occ == 4)
    mfcc_fwd_kernel<false, 4><<<grid, FEAT_THREADS, feat_fwd_smem(), st>>>(x, N, m, fpc, make_dither(mode, dither, seed, pass), raw, ld, dT, stash, cm);
  else
    mfcc_fwd_kernel<false, 3><<<grid, FEAT_THREADS, feat_fwd_smem(), st>>>(x, N, m, fpc, make_dither(mode, dither, seed, pass), raw, ld, dT, stash, cm);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

static int bwd_own_frames(int B, int m) {
  // 2 CTAs fit per SM (registers): aim for >= 8 waves of 296 CTAs so the tail wave is cheap; every extra
  // chunk costs 2 halo frames, so keep chunks >= 32 frames
  int chunks = 1;
  while (chunks < 16 && (long long)B * chunks < 8 * 296 && m / (chunks + 1) >= 32) ++chunks;
  return (m + chunks - 1) / chunks;
}

static int bwd_dispatch(const SgFeatTables* dT, const float* x, int B, int N, int m, DitherSpec D, const float* draw, int ld, const BwdOut& O,
                        const float* stash, cudaStream_t st) {
  int own = bwd_own_frames(B, m);
  if (g_feat_v2 && !O.cmvn && g_bwd2_chunks > 0 && m / g_bwd2_chunks >= 32) own = (m + g_bwd2_chunks - 1) / g_bwd2_chunks;
  dim3 grid((m + own - 1) / own, B);
  if (g_feat_v2 && !O.cmvn) {
    if (stash) mfcc2_bwd_kernel<true><<<grid, FEAT_THREADS, feat2_bwd_smem(), st>>>(x, N, m, own, D, draw, ld, O, dT, stash);
    else mfcc2_bwd_kernel<false><<<grid, FEAT_THREADS, feat2_bwd_smem(), st>>>(x, N, m, own, D, draw, ld, O, dT, nullptr);
  } else {
    if (stash) mfcc_bwd_kernel<true><<<grid, FEAT_THREADS, feat_bwd_smem(), st>>>(x, N, m, own, D, draw, ld, O, dT, stash);
    else mfcc_bwd_kernel<false><<<grid, FEAT_THREADS, feat_bwd_smem(), st>>>(x, N, m, own, D, draw, ld, O, dT, nullptr);
  }
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_feat_bwd_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode, const float* dither,
                       uint64_t seed, uint64_t pass, const float* draw, int ld, float* grad, float scale,
                       int accumulate, cudaStream_t st, const float* stash, int cmvn) {
  if (cmvn && !sg_feat_cmvn_fusable(m)) { sg_set_error("sg_feat_bwd_launch: fused CMVN needs m <= %d frames (got %d)", CMN_WIN, m); return SG_EINVAL; }
  BwdOut O;
  memset(&O, 0, sizeof(O));
  O.mode = 0; O.grad = grad; O.scale = scale; O.accumulate = accumulate; O.cmvn = cmvn;
  return bwd_dispatch(dT, x, B, N, m, make_dither(mode, dither, seed, pass), draw, ld, O, stash, st);
}

int sg_feat_bwd_step_launch(const SgFeatTables* dT, const float* x, int B, int N, int m, int mode,
                            const float* dither, uint64_t seed, uint64_t pass, const float* draw, int ld,
                            const float* x0, float* x_out, float step, float eps, cudaStream_t st, const float* stash, int cmvn) {
  if (cmvn && !sg_feat_cmvn_fusable(m)) { sg_set_error("sg_feat_bwd_step_launch: fused CMVN needs m <= %d frames (got %d)", CMN_WIN, m); return SG_EINVAL; }
  BwdOut O;
  memset(&O, 0, sizeof(O));
  O.mode = 1; O.x0 = x0; O.x_out = x_out; O.step = step; O.eps = eps; O.cmvn = cmvn;
  return bwd_dispatch(dT, x, B, N, m, make_dither(mode, dither, seed, pass), draw, ld, O, stash, st);
}

int sg_dither_fill_launch(int B, int m, uint64_t seed, uint64_t pass, float* out, cudaStream_t st) {
  dim3 grid((m + 15) / 16, B);
  dither_fill2_kernel<<<grid, 256, 0, st>>>(m, make_dither(SG_DITHER_PHILOX, nullptr, seed, pass), out);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

// EOT samples as batch rows: out[e * B + b] = in[b] for e < copies (float4 rows when N % 4 == 0), labels alike
__global__ void tile_rows_kernel(const float* __restrict__ in, float* __restrict__ out, size_t row_floats, int B, int copies,
                                 const long long* __restrict__ yin, long long* __restrict__ yout) {
  const size_t total = row_floats * B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float v = in[i];
    for (int e = 0; e < copies; ++e) out[(size_t)e * total + i] = v;
  }
  if (yin != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < B * copies; i += blockDim.x) yout[i] = yin[i % B];
}
// acc[b] (+)= sum over copies e of rows[e * B + b], summed in copy order
__global__ void reduce_rows_kernel(const float* __restrict__ rows, float* __restrict__ acc, size_t row_floats, int B, int copies,
                                   int accumulate) {
  const size_t total = row_floats * B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float a = accumulate ? acc[i] : 0.f;
    for (int e = 0; e < copies; ++e) a += rows[(size_t)e * total + i];
    acc[i] = a;
  }
}
int sg_tile_rows_launch(const float* in, float* out, size_t row_floats, int B, int copies, const long long* yin, long long* yout,
                        cudaStream_t st) {
  tile_rows_kernel<<<148 * 8, 256, 0, st>>>(in, out, row_floats, B, copies, yin, yout);
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_reduce_rows_launch(const float* rows, float* acc, size_t row_floats, int B, int copies, int accumulate, cudaStream_t st) {
  reduce_rows_kernel<<<148 * 8, 256, 0, st>>>(rows, acc, row_floats, B, copies, accumulate);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

int sg_step_linf_launch(float* x, const float* x0, const float* grad, size_t n, float step, float eps,
                        cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  step_linf_kernel<<<blocks, 256, 0, st>>>(x, x0, grad, n, step, eps);
  SG_LAUNCH_CHECK();
  return SG_OK;
}

#define CMN_PREFIX_MAX_T 850
static int cmvn_dispatch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int ncol, int backward, dim3 grid,
                         cudaStream_t st) {
  if (T > CMN_WIN && T <= CMN_PREFIX_MAX_T) {
    static std::atomic<unsigned long long> configured{0};
    if (sg_first_on_device(&configured))
      SG_CUDA_CHECK(cudaFuncSetAttribute(cmvn_prefix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (CMN_PREFIX_MAX_T + 1) * 32 * (int)sizeof(double)));
    cmvn_prefix_kernel<<<grid, dim3(32, 8), (size_t)(T + 1) * 32 * sizeof(double), st>>>(in, ld_in, out, ld_out, T, ncol, backward);
  } else {
    cmvn_kernel<<<grid, dim3(32, 8), 0, st>>>(in, ld_in, out, ld_out, T, ncol, backward);
  }
  SG_LAUNCH_CHECK();
  return SG_OK;
}
int sg_cmvn_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int backward, cudaStream_t st) {
  return cmvn_dispatch(in, ld_in, out, ld_out, B, T, SG_NCEP, backward, dim3(B), st);
}
// any number of columns (the 72-dim MFCC+delta features of the i-vector system)
int sg_cmvn_cols_launch(const float* in, int ld_in, float* out, int ld_out, int B, int T, int ncol, int backward, cudaStream_t st) {
  return cmvn_dispatch(in, ld_in, out, ld_out, B, T, ncol, backward, dim3(B, (max(ncol, ld_out) + 31) / 32), st);
}

size_t sg_feat_stash_floats(int B, int m) { return (size_t)B * m * SG_STASH_FLOATS; }
