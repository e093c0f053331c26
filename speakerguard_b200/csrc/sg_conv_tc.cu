// T1 / T2: the TDNN contractions on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same contract as sg_conv_simt.cu: out[p, n] = epi(sum_{tap,c} A[p + tap*step, c] * W[n, tap*cin + c])
// with channels-last fp32 activations, so a dilated tap is a TMA box shifted by `step` rows (rows
// outside the tensor are zero-filled by TMA: no im2col, no halo copies).  Both operands are K-major,
// SWIZZLE_128B: A box = 128 rows x 32 fp32, B box = BN rows x 32 fp32 per pipeline stage.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (warp-convergent loop, one elected
// lane issues) and TMEM owner, warps 2..9 = two epilogue groups of four (tcgen05.ld -> bias / ReLU + 1-bit ReLU mask, or
// mask + row validity -> swizzled smem staging box -> TMA store).  The fp32 accumulator (128 lanes x BN columns) is
// double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Template variants of the one kernel: KIND_BF16 (kind::f16 on bf16 operands, else kind::tf32 on fp32), OUT_BF16,
// XFORM (the statistics-pooling adjoint applied to the staged A tiles of the layer-5 dgrad by 8 extra warps), and PAIR:
// a 2-CTA cluster computes a 256 x BN tile with tcgen05.mma.cta_group::2, each CTA staging its 128 rows of A and half
// of B (32 KB instead of 48 KB of L2 -> SM traffic per k-block and CTA, 6 pipeline stages instead of 4).  PAIR is the
// default for every bf16 contraction with a bf16 output (+5 % on the PGD step, mostly through the lower power draw:
// the step runs under the 1 kW cap).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "sg_common.cuh"

#define TC_BM 128
#define TC_BK 32                     // fp32 elements per k-block = one 128-byte swizzle span
#define TC_STAGES 4
#define TC_MAX_BN 256
#define TC_A_BYTES (TC_BM * TC_BK * 4)          // 16 KB
#define TC_B_BYTES (TC_MAX_BN * TC_BK * 4)      // 32 KB
#define TC_STAGE_BYTES (TC_A_BYTES + TC_B_BYTES)
#define TC_STG_BYTES (TC_BM * 32 * 4)            // 16 KB epilogue staging box (128 rows x 32 fp32), x2
#define TC_MAX_STAGES 12
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + 2 * TC_STG_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 1024 /*bias*/)
#define TC_MAIN_THREADS 320           // main kernel: TMA warp, MMA warp, 2 x 4 epilogue warps
#define TC_XF_THREADS (TC_MAIN_THREADS + 256)   // XFORM variant: + 8 warps that rewrite the staged A tile in place

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (it > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 3-D load (channel, frame, utterance): frames outside [0, T) - negative ones included - are zero-filled, which is exactly the
// 'same' padding of a convolution that must not read across utterances
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* smem_src, const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// 3-D store: elements whose coordinates fall beyond the tensor's extents are not written.  (Only the UPPER bound clips: a
// store with a negative coordinate raises an illegal-instruction fault on sm_100, unlike a load, which zero-fills.)
__device__ __forceinline__ void tma_store_3d(const void* smem_src, const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND_BF16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND_BF16) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// bf16x2 {lo, hi} = {max(lo, 0), max(hi, 0)} rounded to nearest even: ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// split form: several loads in flight, one wait (the registers must not be read before tc_ld_wait)
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
        "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
        "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B
// (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

struct TcArgs {
  uint32_t* bits_out; const uint32_t* bits_in; int ldbits;
  const float* bias;
  float* out; int ldo;
  const float* mask; int ldmask;
  int rows, N, bn;                 // bn: N-tile (multiple of 32, <= 256)
  int kchunks, taps, tap_step;     // kchunks = cin / 32
  int epilogue, T, t_valid;
  int m_tiles, n_tiles;
  // XFORM variant (statistics-pooling adjoint fused into the layer-5 dgrad): A is the stored post-ReLU activation r5;
  // the transform warps turn each staged tile into dA5 = (t < xf_tv && r > 0) ? alpha' + beta * r : 0 before the MMA
  // reads it.  xf_ab: [rows / T][xf_ld / 2] x {alpha' pair, beta pair} (bf16x2 each) per utterance and channel pair.
  const uint4* xf_ab; int xf_ld, xf_tv;
  int pf_dist;                     // L2 prefetch distance of the A operand in k-blocks (0 = off)
  int issue_mode;                  // MMA issuer: 0 single-lane region, 1 warp-convergent loop with an elected lane
  int nst, stb;                    // pipeline ring: stages and bytes per stage (main kernel)
  // split-K (small problems): tile index = (m-tile, n-tile, k-slice); slice ks contracts k-blocks [ks * kb_per, (ks+1) * kb_per)
  // and stores its fp32 partial at row offset ks * rows_pad of the partial buffer (the output map then covers that buffer)
  int ksplit, kb_per, rows_pad;
  // re-strided output: the tile rows are frames of utterances with a.T rows each; the output tensor is [utterance][st_T frames]
  // (st_T <= a.T when compacting to the valid frames, or a.T when expanding a compact tensor into a wider row stride) and a
  // staged box is stored once per utterance it touches (3-D map, out-of-range frames clipped).  bits_T: row stride of the
  // ReLU-bit words written by the forward epilogue (0: a.T)
  int st_dual, bits_T, st_nutt, st_stride;
  // utterance-tiled mode ('same' padding, AudioNet): an M tile is 128 frames of ONE utterance (utt_tpu tiles per utterance of
  // utt_T frames); A and the output are addressed through 3-D maps (channel, frame, utterance), taps read frame
  // f + tap_base + tap * tap_step and frames outside the utterance read zero
  int utt_T, utt_tpu, tap_base;
  int wb_rows;                     // rows of the B operand per utterance (0: one operand shared by all tiles)
};

// ---- cluster / cta_group::2 helpers ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// remote arrive with the default (.release.cta) semantics, as CUTLASS's ClusterBarrier::arrive(cta_id) does: the TMEM
// hand-over is ordered by the tcgen05 fences around it; a .release.cluster arrive costs a GPU-scope membar per call.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int KIND_BF16>
__device__ __forceinline__ void tc_mma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND_BF16) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}


// one 16-byte chunk (8 bf16 channels) of the pooling adjoint in packed bf16x2 arithmetic: P holds, per channel pair,
// {alpha' pair, beta pair}; one fma.rn.bf16x2 + one compare mask + one lop3 per pair (the first version did this in fp32,
// ~9 instructions per pair, and made the transform warps - not the tensor pipe - the limiter of the contraction).
// r5 is post-ReLU (>= +0), so r > 0 <=> r != 0.  okm: all ones for valid frames, 0 otherwise.
__device__ __forceinline__ void xf8(uint4& w, const uint4 (&P)[2], uint32_t okm) {
  uint32_t u[4] = {w.x, w.y, w.z, w.w};
  const uint32_t al[4] = {P[0].x, P[0].z, P[1].x, P[1].z}, be[4] = {P[0].y, P[0].w, P[1].y, P[1].w};
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 r = *reinterpret_cast<const __nv_bfloat162*>(&u[k]);
    const __nv_bfloat162 f = __hfma2(*reinterpret_cast<const __nv_bfloat162*>(&be[k]), r, *reinterpret_cast<const __nv_bfloat162*>(&al[k]));
    u[k] = *reinterpret_cast<const uint32_t*>(&f) & __hne2_mask(r, zero) & okm;
  }
  w = make_uint4(u[0], u[1], u[2], u[3]);
}

template <int KIND_BF16, int OUT_BF16, int XFORM = 0, int PAIR = 0>
__global__ void __launch_bounds__(XFORM ? TC_XF_THREADS : TC_MAIN_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapO2, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // PAIR: a CTA pair (cluster of 2 on one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA stages
  // its own 128 rows of A and HALF of the B tile, so a k-block costs 32 KB of L2 -> SM traffic per CTA instead of 48 KB
  // for the same 4.2 MFLOP, and the smaller stage leaves room for 6 pipeline stages.  Measured: the cycles per k-block
  // stay at ~650 (512 for the tensor pipe alone; cuBLAS ~590), but the pair draws less power and the step runs under
  // the 1 kW cap: +10 % SM clock, +5 % on the PGD step.  Rank 0 issues the MMAs; both CTAs run TMA producers
  // (signalling rank 0's full barriers) and epilogues (own 128 TMEM lanes); commits are multicast.
  // The ring geometry is chosen by the host from the B box: stage = 16 KB of A + this CTA's B rows x 128 B; 4 stages
  // for a 256-row B box, 6 for a pair's 128 rows (SGB200_TC_DEEP_RING=1 allows up to 12 for small boxes: no gain measured).
  const int NST = a.nst;
  const uint32_t STB = (uint32_t)a.stb;
  // layout: [2 epilogue staging boxes][pipeline ring][barriers, bias].  The staging boxes come first so that the shifted
  // source window of a re-strided store (st_dual, below) stays inside the CTA's shared memory.
  uint8_t* stg = smem;                                                 // 2 x 16 KB, 1024-byte aligned
  smem += 2 * TC_STG_BYTES;                                            // from here on `smem` is the ring
  uint64_t* bars = (uint64_t*)(smem + TC_STAGES * TC_STAGE_BYTES);
  uint64_t* full = bars;                                // [TC_MAX_STAGES]   (PAIR: rank 0's copy is the live one)
  uint64_t* empty = bars + TC_MAX_STAGES;               // [TC_MAX_STAGES]   per CTA
  uint64_t* tfull = bars + 2 * TC_MAX_STAGES;           // [2]     per CTA
  uint64_t* tempty = bars + 2 * TC_MAX_STAGES + 2;      // [2]     (PAIR: rank 0's copy is the live one)
  uint64_t* xfull = bars + 2 * TC_MAX_STAGES + 4;       // [TC_MAX_STAGES] (XFORM: tile transformed, ready for the MMA)
  uint64_t* xpeer = bars + 3 * TC_MAX_STAGES + 4;       // [TC_MAX_STAGES] (XFORM + PAIR, rank 0's copy: the peer's tile is transformed)
  uint32_t* tmem_slot = (uint32_t*)(bars + 4 * TC_MAX_STAGES + 4);
  float* bias_s = (float*)(bars + 64);                  // [2 epilogue groups][128]: the bias of the current tile's columns

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = a.taps * a.kchunks;
  const int ntiles = a.m_tiles * a.n_tiles * a.ksplit;
  constexpr int KB_ELEMS = KIND_BF16 ? 64 : 32;     // elements per 128-byte k-block
  const uint32_t rank = PAIR ? cluster_rank() : 0u;
  const int cta0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first tile of this CTA (pair)
  const int tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_ROWS = PAIR ? 2 * TC_BM : TC_BM;
  const int rbase = (int)rank * TC_BM;                                   // this CTA's rows inside a pair tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&xfull[s], 4); mbar_init(&xpeer[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();                  // barriers of both CTAs initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int bh = PAIR ? (a.bn >> 1) : a.bn;                       // B rows staged by this CTA
      // XFORM + PAIR: every CTA's boxes signal its OWN full barrier (its transform warps wait on it); plain PAIR: the leader's
      const uint32_t tx = ((PAIR && !XFORM) ? 2u : 1u) * (TC_A_BYTES + (uint32_t)bh * TC_BK * 4);
      // L2 prefetch cursor for the A operand, a.pf_dist k-blocks ahead of the load cursor.  A single-tap layer streams A
      // from HBM exactly once, and with only TC_STAGES - 1 boxes in flight per SM the ring cannot cover the HBM latency
      // (Little: 3 x 16 KB x 148 SMs / ~2 us = 3.5 TB/s); the prefetch moves that wait out of the ring, so the ring
      // only sees L2 latency.  B (weights) is L2-resident anyway.
      int ptile = cta0, pkb = 0;
      auto prefetch_next = [&]() {
        if (ptile < ntiles) {
          const int ptap = pkb / a.kchunks, pkc = pkb - ptap * a.kchunks;
          tma_prefetch_2d(&mapA, pkc * KB_ELEMS, (ptile / a.n_tiles) * TILE_ROWS + rbase + ptap * a.tap_step);
          if (++pkb == nkb) { pkb = 0; ptile += tstep; }
        }
      };
      for (int i = 0; i < a.pf_dist; ++i) prefetch_next();
      for (int tile = cta0; tile < ntiles; tile += tstep) {
        const int t2 = tile / a.ksplit, ks = tile - t2 * a.ksplit;
        const int mt = t2 / a.n_tiles, nt = t2 - mt * a.n_tiles;
        const int p0 = mt * TILE_ROWS + rbase, n0 = nt * a.bn + (int)rank * (PAIR ? bh : 0);
        const int kb0 = ks * a.kb_per, kb1 = min(nkb, kb0 + a.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int tap = kb / a.kchunks, kc = kb - tap * a.kchunks;
          if (a.pf_dist > 0) prefetch_next();
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STB;
          if (PAIR && !XFORM) {
            const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
            // only the leader arms the barrier (count 1, transaction bytes of BOTH CTAs' boxes).  The peer's bytes may land
            // first: the transaction count then goes negative while the phase's single arrival is still pending, so the
            // phase cannot complete early, and the peer cannot run a whole ring phase ahead because its empty[stage] is
            // released by the commit that follows the leader's MMAs.  (A per-k-block remote mbarrier.arrive.release.cluster
            // from the peer costs a GPU-scope membar each time and throttled the pair to ~1700 clk per k-block.)
            if (rank == 0) mbar_expect_tx(&full[stage], tx);
            tma_load_2d_2sm(sa, &mapA, lead_full, kc * KB_ELEMS, p0 + tap * a.tap_step);
            tma_load_2d_2sm(sa + TC_A_BYTES, &mapB, lead_full, kb * KB_ELEMS, n0);
          } else if (!PAIR && a.utt_T > 0) {
            const int ub = mt / a.utt_tpu, f0 = (mt - ub * a.utt_tpu) * TC_BM;
            mbar_expect_tx(&full[stage], tx);
            tma_load_3d(sa, &mapA, &full[stage], kc * KB_ELEMS, f0 + a.tap_base + tap * a.tap_step, ub);
            tma_load_2d(sa + TC_A_BYTES, &mapB, &full[stage], kb * KB_ELEMS, n0 + ub * a.wb_rows);
          } else {
            mbar_expect_tx(&full[stage], tx);
            tma_load_2d(sa, &mapA, &full[stage], kc * KB_ELEMS, p0 + tap * a.tap_step);
            tma_load_2d(sa + TC_A_BYTES, &mapB, &full[stage], kb * KB_ELEMS, n0);
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // issue_mode 1: the whole warp runs the loop convergently (descriptors and counters stay in uniform registers, the
    // four tcgen05.mma of a k-block issue back to back) and one elected lane issues; issue_mode 0: a single-lane region, where
    // ptxas re-derives uniformity for every tcgen05.mma operand (ELECT + R2UR.BROADCAST per instruction, ~70 dependent
    // instructions per k-block).
    if (PAIR && XFORM && rank != 0) {
      // the peer's otherwise idle MMA warp forwards "my A tile is transformed" to the leader, one remote arrive per k-block
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = cta0; tile < ntiles; tile += tstep)
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&xfull[stage], phase);
            mbar_arrive_cluster(mapa_u32(smem_u32(&xpeer[stage]), 0));
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
      }
    } else if (a.issue_mode != 0 && rank == 0) {
      const uint32_t fmt = KIND_BF16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24);
      const uint32_t smem_base = smem_u32(smem);
      const bool leader = elect_one();
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cta0; tile < ntiles; tile += tstep) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * TC_MAX_BN;
        const int nk = min(nkb, (tile % a.ksplit) * a.kb_per + a.kb_per) - (tile % a.ksplit) * a.kb_per;   // k-blocks of this slice
        for (int kb = 0; kb < nk; ++kb) {
          const uint32_t sa = smem_base + (uint32_t)stage * STB;
          const uint64_t da = make_desc(sa), db = make_desc(sa + TC_A_BYTES);
          mbar_wait(XFORM ? &xfull[stage] : &full[stage], phase);
          if (PAIR && XFORM) mbar_wait(&xpeer[stage], phase);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {      // (two N/2 MMAs per k-step on alternating column halves were measured: no gain)
              if (PAIR) tc_mma_2sm<KIND_BF16>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              else tc_mma<KIND_BF16>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            }
            if (PAIR) tc_commit_2sm(&empty[stage]); else tc_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (leader) { if (PAIR) tc_commit_2sm(&tfull[acc]); else tc_commit(&tfull[acc]); }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D=f32 [4,6)=1, A/B format [7,10)/[10,13) (tf32=2, bf16=1), K-major both,
      // N>>3 in [17,23), M>>4 in [24,29)
      const uint32_t fmt = KIND_BF16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(a.bn >> 3) << 17) | ((uint32_t)(TILE_ROWS >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cta0; tile < ntiles; tile += tstep) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * TC_MAX_BN;
        const int nk = min(nkb, (tile % a.ksplit) * a.kb_per + a.kb_per) - (tile % a.ksplit) * a.kb_per;   // k-blocks of this slice
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(XFORM ? &xfull[stage] : &full[stage], phase);
          if (PAIR && XFORM) mbar_wait(&xpeer[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STB);
          const uint64_t da = make_desc(sa), db = make_desc(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 32 bytes) per 128-byte swizzle span: +2 in 16-byte units
            if (PAIR) tc_mma_2sm<KIND_BF16>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else tc_mma<KIND_BF16>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if (PAIR) tc_commit_2sm(&empty[stage]); else tc_commit(&empty[stage]);   // frees the smem slot when these MMAs retire
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_2sm(&tfull[acc]); else tc_commit(&tfull[acc]);          // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (XFORM && warp >= TC_MAIN_THREADS / 32) {
    // ===== A-operand transform: warps 10..17 (bf16 tiles: 128 rows x 64 channels, SWIZZLE_128B) =====
    // Two groups of four warps take ALTERNATE k-blocks of the ring, so that one group's latencies (barrier wait, shared-memory
    // round trip, the generic->async proxy fence: 16 % of all samples as `membar` in the one-group form, and the parameter
    // loads: another 16 % as long-scoreboard) overlap the other group's arithmetic; the parameters of a group's next k-block
    // are requested before it waits for the current one.
    // thread -> logical 16-byte chunk j (8 channels) of rows g, g+16, ..., g+112; (g + 16 i) & 7 == g & 7, so the physical
    // chunk position j ^ (row & 7) is the same for all eight rows.  A tile spans at most two utterances (the host only
    // selects this variant for T >= 128): rows before `isplit` use the parameters of utterance b0, the others those of b0 + 1.
    const int t = (int)threadIdx.x - TC_MAIN_THREADS;
    const int xg = t >> 7, tg = t & 127;                            // transform group, thread within the group
    const int j = tg & 7, g = tg >> 3;                              // g in [0, 16)
    const uint32_t off = (uint32_t)g * 128u + (uint32_t)((j ^ (g & 7)) << 4);
    const int nutt = a.rows / a.T;
    struct XfTile { const uint4* q0; const uint4* q1; int isplit; uint32_t okbits; };
    auto tile_info = [&](int tile) {
      XfTile ti;
      const int mt = tile / a.n_tiles;
      const int row0 = mt * TILE_ROWS + rbase + g;
      const int b0 = row0 / a.T;
      const int tt0 = row0 - b0 * a.T;
      int isplit = (a.T - tt0 + 15) >> 4;
      if (isplit > 8) isplit = 8;
      uint32_t ok = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int tt = tt0 + 16 * i - (i >= isplit ? a.T : 0);
        ok |= ((row0 + 16 * i < a.rows) && (tt < a.xf_tv)) ? (1u << i) : 0u;
      }
      const bool second = (isplit < 8) && (b0 + 1 < nutt);
      ti.q0 = a.xf_ab + (((size_t)(b0 < nutt ? b0 : 0) * a.xf_ld + j * 8) >> 2);   // one uint4 = 4 channels
      ti.q1 = ti.q0 + (second ? (a.xf_ld >> 2) : 0);
      ti.isplit = isplit; ti.okbits = ok;
      return ti;
    };
    // ring position of this group's first k-block: s = xg
    int stage = xg % NST; uint32_t phase = (uint32_t)((xg / NST) & 1);
    int tile = cta0, kb = xg;
    while (kb >= nkb) { kb -= nkb; tile += tstep; }
    XfTile cur = tile_info(tile < ntiles ? tile : 0);
    uint4 P[2], Q[2];
    if (tile < ntiles) {
#pragma unroll
      for (int k = 0; k < 2; ++k) { P[k] = __ldg(cur.q0 + kb * 16 + k); Q[k] = __ldg(cur.q1 + kb * 16 + k); }
    }
    while (tile < ntiles) {
      // this group's next k-block: two ring positions on
      int ntile = tile, nkbk = kb + 2;
      while (nkbk >= nkb) { nkbk -= nkb; ntile += tstep; }
      XfTile nxt = cur;
      if (ntile != tile && ntile < ntiles) nxt = tile_info(ntile);
      uint4 Pn[2] = {P[0], P[1]}, Qn[2] = {Q[0], Q[1]};
      if (ntile < ntiles) {
#pragma unroll
        for (int k = 0; k < 2; ++k) { Pn[k] = __ldg(nxt.q0 + nkbk * 16 + k); Qn[k] = __ldg(nxt.q1 + nkbk * 16 + k); }
      }
      mbar_wait(&full[stage], phase);
      uint8_t* sa = smem + stage * STB + off;
      uint4 w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = *reinterpret_cast<const uint4*>(sa + i * 2048);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t okm = 0u - ((cur.okbits >> i) & 1u);
        if (i < cur.isplit) xf8(w[i], P, okm); else xf8(w[i], Q, okm);
        *reinterpret_cast<uint4*>(sa + i * 2048) = w[i];
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfull[stage]);
      stage += 2;
      if (stage >= NST) { stage -= NST; phase ^= 1; }
      cur = nxt; tile = ntile; kb = nkbk;
#pragma unroll
      for (int k = 0; k < 2; ++k) { P[k] = Pn[k]; Q[k] = Qn[k]; }
    }
  } else {
    // ===== epilogue: warps 2..9 = two groups of four; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
    // TMEM -> registers -> bias / ReLU / mask -> swizzled smem staging box (128 rows x 128 B) -> TMA store.
    // (direct st.global from one-row-per-thread registers ran the short-K layers at ~2 TB/s of output; with a
    // single group of four warps the short-K layers were still epilogue-bound, so two groups take alternate
    // column chunks, each with its own staging box and named barrier.)
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                      // 0: warps 2..5, 1: warps 6..9
    const int r_in = q * 32 + lane;                       // row within the tile
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    uint8_t* buf = stg + grp * TC_STG_BYTES;
    constexpr int CW = OUT_BF16 ? 64 : 32;                // columns per 128-byte staging row
    const bool has_bias = (a.epilogue == SG_EPI_BIAS || a.epilogue == SG_EPI_BIAS_RELU);
    float* bias_g = bias_s + grp * 128;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t nstore = 0;
    for (int tile = cta0; tile < ntiles; tile += tstep) {
      const int t2 = tile / a.ksplit, ks = tile - t2 * a.ksplit;
      const int mt = t2 / a.n_tiles, nt = t2 - mt * a.n_tiles;
      // utterance-tiled mode: the tile is frames f0 .. f0 + 127 of utterance ub; rows beyond the utterance are not stored
      const int u_ub = a.utt_T > 0 ? mt / a.utt_tpu : 0, u_f0 = a.utt_T > 0 ? (mt - u_ub * a.utt_tpu) * TC_BM : 0;
      const int row = a.utt_T > 0 ? u_ub * a.utt_T + u_f0 + r_in : mt * TILE_ROWS + rbase + r_in;
      const bool row_in = a.utt_T > 0 ? (u_f0 + r_in < a.utt_T) : (row < a.rows);
      const int n0 = nt * a.bn;
      if (has_bias) {
        // this group's 128 columns of the bias (chunk k, column i -> bias_g[k * CW + i]); the last named barrier of the
        // previous tile ordered every read of bias_g before this write
        const int t = ((warp - 2) & 3) * 32 + lane;
        const int cc = grp * CW + (t / CW) * 2 * CW + (t % CW);
        bias_g[t] = cc < a.bn ? __ldg(a.bias + n0 + cc) : 0.f;
        epi_bar(1 + grp);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      bool row_ok = row_in;
      if (a.epilogue == SG_EPI_MASK) row_ok = row_ok && ((row % a.T) < a.t_valid);
      // One chunk = NG groups of 32 columns.  The operands of the epilogue op (bias / ReLU bits) are requested first,
      // then all TMEM loads of the chunk are issued with a single wait, so that their latencies overlap instead of
      // adding up (the short-K layers are bound by this warp's critical path, not by issue slots).
      constexpr int NG = OUT_BF16 ? 2 : 1;
      const bool use_bits = (a.epilogue == SG_EPI_MASK) && a.bits_in != nullptr;
      auto apply = [&](int col, float (&v)[32], const float* bsm, uint32_t wbits) {
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(bsm + 4 * j);      // uniform address: one broadcast wavefront
            v[4 * j + 0] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
          if (a.epilogue == SG_EPI_BIAS_RELU) {
            // The epilogue warps are bound by the 16-lane integer/ALU pipe (2 warps per scheduler), so every instruction per
            // element counts: the ReLU of a bf16 output is folded into the conversion (cvt.rn.relu.bf16x2.f32) and the
            // ReLU bit is one set + one lop3.
            if (!OUT_BF16) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (a.bits_out != nullptr && row_in) {
              // bit j = (x_j > 0): 0 - x has its sign bit set exactly then (0 - (+-0) = +0), and a funnel shift moves that
              // sign into the word: two FADDs on the wide FMA pipe + one SHF per element instead of FSETP + SEL + IADD3
              uint32_t ob = 0;
#pragma unroll
              for (int j = 31; j >= 0; --j) ob = __funnelshift_l(__float_as_uint(0.f - v[j]), ob, 1);
              if (a.bits_T == 0) {
                a.bits_out[(size_t)row * a.ldbits + (col >> 5)] = ob;
              } else {                                              // compacted output: the bit rows follow it
                const int ub = row / a.T, ut = row - ub * a.T;
                if (ut < a.bits_T) a.bits_out[((size_t)ub * a.bits_T + ut) * a.ldbits + (col >> 5)] = ob;
              }
            }
          }
        } else if (a.epilogue == SG_EPI_MASK) {
          if (use_bits) {                                     // wbits is 0 for rows outside the valid range
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((wbits >> j) & 1u) ? v[j] : 0.f;
          } else if (row_ok) {
            const float4* mp = reinterpret_cast<const float4*>(a.mask + (size_t)row * a.ldmask + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 m4 = __ldg(mp + j);
              v[4 * j + 0] = m4.x > 0.f ? v[4 * j + 0] : 0.f;
              v[4 * j + 1] = m4.y > 0.f ? v[4 * j + 1] : 0.f;
              v[4 * j + 2] = m4.z > 0.f ? v[4 * j + 2] : 0.f;
              v[4 * j + 3] = m4.w > 0.f ? v[4 * j + 3] : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
        }
      };
      for (int c = grp * CW; c < a.bn; c += 2 * CW, ++nstore) {
        float v[NG][32];
        uint32_t wb[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const int col = n0 + c + 32 * g;
          wb[g] = 0u;
          if (use_bits && row_ok) {
            wb[g] = __ldg(a.bits_in + (size_t)row * a.ldbits + (col >> 5));
          }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g)
          tc_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_MAX_BN + c + 32 * g), v[g]);
        tc_ld_wait();
#pragma unroll
        for (int g = 0; g < NG; ++g) apply(n0 + c + 32 * g, v[g], bias_g + (c - grp * CW) / 2 + 32 * g, wb[g]);
        uint4 packed[8];
        if (OUT_BF16 && a.epilogue == SG_EPI_BIAS_RELU) {
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              packed[4 * g + j] = make_uint4(pack_relu_bf16x2(v[g][8 * j + 0], v[g][8 * j + 1]), pack_relu_bf16x2(v[g][8 * j + 2], v[g][8 * j + 3]),
                                             pack_relu_bf16x2(v[g][8 * j + 4], v[g][8 * j + 5]), pack_relu_bf16x2(v[g][8 * j + 6], v[g][8 * j + 7]));
        } else if (OUT_BF16) {
          const float (&v1)[32] = v[0];
          const float (&v2)[32] = v[NG - 1];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v1[8 * j + 0], v1[8 * j + 1]), p1 = __floats2bfloat162_rn(v1[8 * j + 2], v1[8 * j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v1[8 * j + 4], v1[8 * j + 5]), p3 = __floats2bfloat162_rn(v1[8 * j + 6], v1[8 * j + 7]);
            packed[j] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                   *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
            __nv_bfloat162 s0 = __floats2bfloat162_rn(v2[8 * j + 0], v2[8 * j + 1]), s1 = __floats2bfloat162_rn(v2[8 * j + 2], v2[8 * j + 3]);
            __nv_bfloat162 s2 = __floats2bfloat162_rn(v2[8 * j + 4], v2[8 * j + 5]), s3 = __floats2bfloat162_rn(v2[8 * j + 6], v2[8 * j + 7]);
            packed[4 + j] = make_uint4(*reinterpret_cast<uint32_t*>(&s0), *reinterpret_cast<uint32_t*>(&s1),
                                       *reinterpret_cast<uint32_t*>(&s2), *reinterpret_cast<uint32_t*>(&s3));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            packed[j] = make_uint4(__float_as_uint(v[0][4 * j]), __float_as_uint(v[0][4 * j + 1]), __float_as_uint(v[0][4 * j + 2]),
                                   __float_as_uint(v[0][4 * j + 3]));
        }
        if (nstore >= 1) {                                  // this group's previous store must have drained the box
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          epi_bar(1 + grp);
        }
        uint4* srow = reinterpret_cast<uint4*>(buf + r_in * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) srow[j ^ (r_in & 7)] = packed[j];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        epi_bar(1 + grp);
        if (issuer) {
          if (a.utt_T > 0) {
            tma_store_3d(buf, &mapO, n0 + c, u_f0, u_ub);                                   // frames >= T are clipped
          } else if (!a.st_dual) {
            tma_store_2d(buf, &mapO, n0 + c, ks * a.rows_pad + mt * TILE_ROWS + rbase);   // rows / columns beyond the tensor are clipped by TMA
          } else {
            const int R0 = mt * TILE_ROWS + rbase, ub = R0 / a.T, ut = R0 - ub * a.T;      // first row of the box: utterance, frame
            tma_store_3d(buf, &mapO, n0 + c, ut, ub);                                       // frames >= the map's extent are dropped
            // The box's rows [k, 128) belong to the next utterance (frames 0 ...).  They are stored through the "sliding" map
            // (element (c, i, s) = row s + i, i < 128): source window shifted down by k rows - the 128-byte swizzle is a function
            // of the shared-memory address, so a row-shifted window reads what was staged - and i-coordinate k, so that the
            // window's rows beyond the box (i >= 128) are clipped.
            const int k = a.T - ut;
            if (k < TC_BM && ub + 1 < a.st_nutt) tma_store_3d(buf + k * 128, &mapO2, n0 + c, k, (ub + 1) * a.st_stride - k);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();                  // the peer may still be reading its TMEM / signalling our barriers
  else __syncthreads();
  if (warp == 1) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}


// (Two earlier experimental kernels lived here and were removed after the CTA-pair template variant above superseded them:
// a TF32-only cta_group::2 kernel whose per-k-block remote mbarrier arrive cost a GPU-scope membar each time (-40 %), and a
// single-CTA 256 x 256 tile kernel whose two accumulators filled TMEM so that the epilogue could not overlap the next
// tile's MMAs (+-0 %).  DESIGN.md section 3 keeps the measurements.)

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;
static int g_num_sms = 0;
static int g_pf_dist = 0;        // L2 prefetch distance (k-blocks) of the A operand; SGB200_TC_PREFETCH overrides, 0 = off
static int g_pf_all = 0;         // SGB200_TC_PREFETCH_ALL=1: also for multi-tap layers (their A re-reads hit L2 anyway)
static int g_pair_xf = 1;        // SGB200_TC_PAIR_XF=0: keep the fused layer-5 dgrad on the single-CTA kernel
static int g_max_bn = TC_MAX_BN;  // SGB200_TC_BN: cap of the N-tile width (A/B switch)
static int g_small_bn = 1;       // SGB200_TC_SMALL_BN=0: keep 256-column tiles for problems that do not fill the GPU
static int g_deep_ring = 0;      // SGB200_TC_DEEP_RING=1: as many stages as fit when the B box is small (measured: no gain)
static int g_issue_mode = 1;     // SGB200_TC_ISSUE: see TcArgs::issue_mode
static int g_pair_bf16 = 2;      // SGB200_TC_PAIR_BF16: contractions on CTA pairs (cta_group::2): 1 bf16 long-K only, 2 all bf16 -> bf16, 3 / 4 see sg_conv_tc()

static std::atomic<unsigned long long> g_tc_dev_init{0};
static int tc_init() {
  if (!sg_first_on_device(&g_tc_dev_init)) return SG_OK;   // shared-memory opt-ins are per device
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    sg_set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
    return SG_ECUDA;
  }
  int dev = 0;
  SG_CUDA_CHECK(cudaGetDevice(&dev));
  SG_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<0, 1, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<1, 1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  SG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<0, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  if (const char* e = getenv("SGB200_TC_PAIR_BF16")) g_pair_bf16 = atoi(e);
  if (const char* e = getenv("SGB200_TC_ISSUE")) g_issue_mode = atoi(e);
  if (const char* e = getenv("SGB200_TC_DEEP_RING")) g_deep_ring = atoi(e);
  if (const char* e = getenv("SGB200_TC_SMALL_BN")) g_small_bn = atoi(e) != 0;
  if (const char* e = getenv("SGB200_TC_BN")) { const int v = atoi(e); if (v >= 32 && v <= TC_MAX_BN && v % 32 == 0) g_max_bn = v; }
  if (const char* e = getenv("SGB200_TC_PAIR_XF")) g_pair_xf = atoi(e);
  if (const char* e = getenv("SGB200_TC_PREFETCH")) { g_pf_dist = atoi(e); if (g_pf_dist < 0 || g_pf_dist > 64) g_pf_dist = 0; }
  if (const char* e = getenv("SGB200_TC_PREFETCH_ALL")) g_pf_all = atoi(e) != 0;
  g_encode = (EncodeFn)fn;
  return SG_OK;
}

int sg_conv_tc_warm() { return tc_init(); }

static int make_map(CUtensorMap* m, const void* base, int bf16, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 64 : 32), box_rows};       // 128-byte inner box
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { sg_set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu ld=%llu box=%u", (int)r,
                                        (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows); return SG_ECUDA; }
  return SG_OK;
}

// sliding view of a [rows][cols] tensor: element (c, i, s) = row s + i for i < 128 (overlapping strides), box = 128 x 128 bytes
static int make_map_slide(CUtensorMap* m, const void* base, int bf16, uint64_t rows, uint64_t cols, uint64_t ld) {
  const uint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {cols, TC_BM, rows};
  cuuint64_t strides[2] = {ld * es, ld * es};
  cuuint32_t box[3] = {(cuuint32_t)(bf16 ? 64 : 32), TC_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { sg_set_error("cuTensorMapEncodeTiled (sliding) failed (%d): rows=%llu cols=%llu", (int)r, (unsigned long long)rows, (unsigned long long)cols); return SG_ECUDA; }
  return SG_OK;
}

// [utterance][frames][cols] output with `frames` <= `tstride` rows per utterance in memory; box = 128 frames x 128 bytes
static int make_map3(CUtensorMap* m, const void* base, int bf16, uint64_t nutt, uint64_t frames, uint64_t tstride, uint64_t cols, uint64_t ld) {
  const uint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[3] = {cols, frames, nutt};
  cuuint64_t strides[2] = {ld * es, tstride * ld * es};
  cuuint32_t box[3] = {(cuuint32_t)(bf16 ? 64 : 32), TC_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { sg_set_error("cuTensorMapEncodeTiled (3-D) failed (%d): utt=%llu frames=%llu stride=%llu cols=%llu", (int)r,
                                        (unsigned long long)nutt, (unsigned long long)frames, (unsigned long long)tstride, (unsigned long long)cols); return SG_ECUDA; }
  return SG_OK;
}

// split-K epilogue: out[r][c] = bias[c] + sum over slices s (in order) of part[s][r][c]; float4 per thread
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int S, size_t slice_floats, int rows, int N, const float* __restrict__ bias,
                                     float* __restrict__ out, int ldo) {
  const int n4 = N >> 2;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * n4) return;
  const int r = (int)(i / n4), c = (int)(i - (size_t)r * n4) * 4;
  float4 a = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = part + (size_t)r * N + c;
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p + (size_t)s * slice_floats));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  *reinterpret_cast<float4*>(out + (size_t)r * ldo + c) = a;
}

static thread_local int g_tc_extra = 0;
int sg_conv_tc_extra_launches() { return g_tc_extra; }

// uses a.Wk, the K-major copy of the weights: [N][taps*cin]
int sg_conv_tc(const SgConvArgs& a, int precision, cudaStream_t st) {
  g_tc_extra = 0;
  if (precision == SG_PREC_FP32) { sg_set_error("sg_conv_tc called in fp32 mode"); return SG_EINVAL; }
  const int kbe = a.op_bf16 ? 64 : 32;
  const bool utt = a.same_utt != 0;
  if (utt && (a.op_bf16 || a.out_bf16 || a.xf_ab || a.out_T > 0 || a.bits_out || a.bits_in || a.T < 1 || a.rows % a.T != 0)) {
    sg_set_error("sg_conv_tc: 'same' padding is built for fp32 tensors with TF32 operands (rows %d, T %d)", a.rows, a.T); return SG_EUNSUPPORTED;
  }
  if (!utt && a.tap_base != 0) { sg_set_error("sg_conv_tc: tap_base needs the utterance-tiled ('same' padding) mode"); return SG_EUNSUPPORTED; }
  int r = tc_init();
  if (r != SG_OK) return r;
  if (a.cin % kbe != 0 || a.N % 32 != 0 || a.lda % 8 != 0 || a.ldo % 8 != 0 || (a.out_bf16 && a.N % 64 != 0) ||
      (a.epilogue == SG_EPI_MASK && !a.bits_in && (a.ldmask % 4 != 0 || a.op_bf16))) {
    sg_set_error("sg_conv_tc: cin %% 32, N %% 32, lda/ldo %% 4 required (cin=%d N=%d)", a.cin, a.N);
    return SG_EINVAL;
  }
  int bn = g_max_bn;
  while (bn > 32 && a.N % bn != 0) bn -= 32;     // largest N-tile (multiple of 32, <= 256) dividing N
  // (Narrower tiles against wave quantisation - B = 128 x 300 frames is 300 pair tiles on 74 pairs = 4.05 -> 5 rounds - were
  // measured and rejected: 128-column tiles run the 512-channel layers 55 % slower (A is re-read once more per output column
  // and the tensor pipe waits on shared memory), 112.8 k vs 143.2 k utt-iter/s at B = 128.  SGB200_TC_BN caps the width for A/B runs.
  // Narrow tiles for the LAST wave only - the four full waves as one launch, the 4 leftover tiles as a second launch of sixteen
  // 64-column tiles - were measured too: 150.4 k vs 152.0 k (a k-block costs ~650 cycles whatever its N, so a quarter-width
  // tile is not a quarter of the time).)
  if (a.N % bn != 0) { sg_set_error("sg_conv_tc: N=%d is not a multiple of 32", a.N); return SG_EINVAL; }
  // Small problems (the head's fc1 / LDA contractions: B rows): with 256-column tiles only m_tiles x N/256 SMs work, and each
  // pulls its whole A + B stream through one SM's L2 port (the B = 1024 fc1 forward ran 16 CTAs for 41 us).  Narrower tiles
  // spread the same stream over more SMs; the tensor pipe is nowhere near busy in this regime.
  // Split-K first: the head's fc1 forward (K = 3000: 94 k-blocks in sequence on 16 CTAs, each k-block a full L2 round trip
  // behind a 4-deep ring) becomes 8 x fewer k-blocks per CTA on 8 x more CTAs; the fp32 partials are summed in slice order by
  // splitk_reduce_kernel (deterministic), which also applies the bias.
  int ksplit = 1;
  const int nkb_all = a.taps * (a.cin / kbe);
  // The slice count is a function of the contraction's K and N only, never of the row count: the summation order of an
  // output element must not depend on the batch size, or a sharded run would differ from the unsharded one in the last bit
  // (tests/test_gpu_shard.py).
  if (g_small_bn && !utt && !a.xf_ab && a.splitk_ws && !a.out_bf16 && !a.bits_out && a.N % 4 == 0 && a.ldo % 4 == 0 && a.N <= 512 &&
      (a.epilogue == SG_EPI_BIAS || a.epilogue == SG_EPI_NONE)) {
    // 32 slices: the i-vector linear term, K = 3 F C (13 824 k-blocks at C = 2048; 1 728 at the C = 256 system of tests/test_gpu_iv.py)
    const int S = nkb_all >= 1024 ? 32 : (nkb_all >= 64 ? 8 : (nkb_all >= 16 ? 4 : 1));
    const size_t need = (size_t)S * ((a.rows + TC_BM - 1) / TC_BM) * TC_BM * a.N;
    if (S > 1 && (S - 1) * ((nkb_all + S - 1) / S) < nkb_all) {
      if (need > a.splitk_floats) { sg_set_error("sg_conv_tc: split-K scratch too small (%zu floats needed, %zu given)", need, a.splitk_floats); return SG_EINVAL; }
      ksplit = S;
    }
  }
  if (g_small_bn && !a.xf_ab) {
    const int mt = utt ? (a.rows / a.T) * ((a.T + TC_BM - 1) / TC_BM) : ((a.rows + TC_BM - 1) / TC_BM) * ksplit;
    while (bn >= 128 && (bn / 2) % 32 == 0 && a.N % (bn / 2) == 0 && (!a.out_bf16 || (bn / 2) % 64 == 0) &&
           2 * mt * (a.N / bn) <= g_num_sms)
      bn /= 2;
  }
  CUtensorMap mapA, mapB;
  if (utt) r = make_map3(&mapA, a.A, 0, (uint64_t)(a.rows / a.T), (uint64_t)a.T, (uint64_t)a.T, (uint64_t)a.cin, (uint64_t)a.lda);
  else r = make_map(&mapA, a.A, a.op_bf16, (uint64_t)a.rows, (uint64_t)a.cin, (uint64_t)a.lda, TC_BM);
  if (r != SG_OK) return r;
  if (!a.Wk) { sg_set_error("sg_conv_tc: K-major weights missing"); return SG_EINVAL; }
  if (a.w_per_utt && !utt) { sg_set_error("sg_conv_tc: per-utterance weights need the utterance-tiled mode"); return SG_EUNSUPPORTED; }
  r = make_map(&mapB, a.Wk, a.op_bf16, (uint64_t)a.N * (a.w_per_utt ? (uint64_t)(a.rows / a.T) : 1), (uint64_t)a.taps * a.cin,
               (uint64_t)a.taps * a.cin, (uint32_t)bn);
  if (r != SG_OK) return r;
  TcArgs t;
  t.bits_out = a.bits_out; t.bits_in = a.bits_in; t.ldbits = a.ldbits;
  t.bias = a.bias; t.out = a.out; t.ldo = a.ldo; t.mask = a.mask; t.ldmask = a.ldmask;
  t.rows = a.rows; t.N = a.N; t.bn = bn; t.kchunks = a.cin / kbe; t.taps = a.taps; t.tap_step = a.tap_step;
  t.epilogue = a.epilogue; t.T = a.T > 0 ? a.T : 1; t.t_valid = a.t_valid;
  t.m_tiles = (a.rows + TC_BM - 1) / TC_BM; t.n_tiles = a.N / bn;
  t.utt_T = 0; t.utt_tpu = 1; t.tap_base = 0; t.wb_rows = a.w_per_utt ? a.N : 0;
  if (utt) { t.utt_T = a.T; t.utt_tpu = (a.T + TC_BM - 1) / TC_BM; t.m_tiles = (a.rows / a.T) * t.utt_tpu; t.tap_base = a.tap_base; }
  t.xf_ab = reinterpret_cast<const uint4*>(a.xf_ab); t.xf_ld = a.xf_ld; t.xf_tv = a.xf_tv;
  t.pf_dist = (!utt && (a.taps == 1 || g_pf_all)) ? g_pf_dist : 0;
  t.issue_mode = g_issue_mode;
  t.stb = TC_A_BYTES + bn * TC_BK * 4; t.nst = (TC_STAGES * TC_STAGE_BYTES) / t.stb;
  if (t.nst > TC_MAX_STAGES) t.nst = TC_MAX_STAGES;
  // a problem that does not fill the GPU is latency-bound per CTA: use every stage that fits (a narrow B box leaves room for 8)
  const bool small_problem = g_small_bn && t.m_tiles * t.n_tiles * ksplit < g_num_sms;
  if (!g_deep_ring && !small_problem && t.nst > TC_STAGES) t.nst = TC_STAGES;
  t.ksplit = ksplit; t.kb_per = (nkb_all + ksplit - 1) / ksplit; t.rows_pad = t.m_tiles * TC_BM;
  if (ksplit > 1) { t.bias = nullptr; t.epilogue = SG_EPI_NONE; t.out = a.splitk_ws; t.ldo = a.N; }
  if (a.xf_ab && !(a.op_bf16 && a.out_bf16 && a.taps == 1 && a.T >= TC_BM && a.rows % a.T == 0 && a.xf_ld % 8 == 0 && a.cin <= a.xf_ld)) {
    sg_set_error("sg_conv_tc: the fused pooling adjoint needs bf16 operands/output, one tap, T >= 128 (T=%d taps=%d)", a.T, a.taps);
    return SG_EINVAL;
  }
  if ((a.epilogue == SG_EPI_BIAS || a.epilogue == SG_EPI_BIAS_RELU) && !a.bias) { sg_set_error("sg_conv_tc: bias epilogue without bias"); return SG_EINVAL; }
  CUtensorMap mapO, mapO2;
  memset(&mapO2, 0, sizeof(mapO2));
  t.st_dual = 0; t.bits_T = a.bits_T; t.st_nutt = 0; t.st_stride = 0;
  if (utt) {
    r = make_map3(&mapO, a.out, 0, (uint64_t)(a.rows / a.T), (uint64_t)a.T, (uint64_t)a.T, (uint64_t)a.N, (uint64_t)a.ldo);
  } else
  if (a.out_T > 0) {
    if (ksplit > 1 || a.T < TC_BM || a.rows % a.T != 0 || a.out_T > a.out_Tstride || a.out_T > a.T) {
      sg_set_error("sg_conv_tc: re-strided output needs T >= 128 rows per utterance and rows %% T == 0 (T=%d rows=%d)", a.T, a.rows);
      return SG_EINVAL;
    }
    t.st_dual = 1; t.st_nutt = a.rows / a.T; t.st_stride = a.out_Tstride;
    r = make_map_slide(&mapO2, a.out, a.out_bf16, (uint64_t)(a.rows / a.T) * a.out_Tstride, (uint64_t)a.N, (uint64_t)a.ldo);
    if (r != SG_OK) return r;
    r = make_map3(&mapO, a.out, a.out_bf16, (uint64_t)(a.rows / a.T), (uint64_t)a.out_T, (uint64_t)a.out_Tstride, (uint64_t)a.N, (uint64_t)a.ldo);
  } else
  if (ksplit > 1) r = make_map(&mapO, a.splitk_ws, 0, (uint64_t)ksplit * t.rows_pad, (uint64_t)a.N, (uint64_t)a.N, TC_BM);
  else r = make_map(&mapO, a.out, a.out_bf16, (uint64_t)a.rows, (uint64_t)a.N, (uint64_t)a.ldo, TC_BM);
  if (r != SG_OK) return r;
  // bf16 CTA-pair variant (cta_group::2): mode 1 = contractions with >= 16 k-blocks, 2 = every eligible one
  // (modes 3 and 4 are not validated defaults: 3 adds the tf32-operand / bf16-output layer-1 forward (measured: no gain),
  //  4 adds fp32-output contractions - the tf32 mode and the i-vector UBM contraction: tests/test_gpu_tc.py and
  //  tests/test_gpu_iv.py pass with it, its throughput has not been measured yet)
  if (!utt && ksplit == 1 && g_pair_bf16 && (a.op_bf16 || g_pair_bf16 >= 3) && (a.out_bf16 || (g_pair_bf16 >= 4 && !a.op_bf16)) && (!a.xf_ab || g_pair_xf) && bn % 32 == 0 && bn >= 64 && a.rows > 256 &&
      (g_pair_bf16 >= 2 || a.taps * t.kchunks >= 16)) {
    CUtensorMap mapBh;
    r = make_map(&mapBh, a.Wk, a.op_bf16, (uint64_t)a.N, (uint64_t)a.taps * a.cin, (uint64_t)a.taps * a.cin, (uint32_t)(bn / 2));
    if (r != SG_OK) return r;
    t.m_tiles = (a.rows + 2 * TC_BM - 1) / (2 * TC_BM);
    t.stb = TC_A_BYTES + (bn / 2) * TC_BK * 4; t.nst = (TC_STAGES * TC_STAGE_BYTES) / t.stb;
    if (t.nst > TC_MAX_STAGES) t.nst = TC_MAX_STAGES;
    if (!g_deep_ring && t.nst > 6) t.nst = 6;
    int pairs = t.m_tiles * t.n_tiles;
    if (pairs > g_num_sms / 2) pairs = g_num_sms / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC_MAIN_THREADS); cfg.dynamicSmemBytes = TC_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (a.xf_ab) { cfg.blockDim = dim3(TC_XF_THREADS); SG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 1, 1, 1>, mapA, mapBh, mapO, mapO2, t)); }
    else if (a.op_bf16) SG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 1, 0, 1>, mapA, mapBh, mapO, mapO2, t));
    else if (a.out_bf16) SG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 1, 0, 1>, mapA, mapBh, mapO, mapO2, t));
    else SG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 0, 0, 1>, mapA, mapBh, mapO, mapO2, t));
    return SG_OK;
  }
  int grid = t.m_tiles * t.n_tiles * ksplit;
  if (grid > g_num_sms) grid = g_num_sms;
  if (a.xf_ab) conv_tc_kernel<1, 1, 1><<<grid, TC_XF_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapO, mapO2, t);
  else if (a.op_bf16 && a.out_bf16) conv_tc_kernel<1, 1><<<grid, TC_MAIN_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapO, mapO2, t);
  else if (a.op_bf16) conv_tc_kernel<1, 0><<<grid, TC_MAIN_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapO, mapO2, t);
  else if (a.out_bf16) conv_tc_kernel<0, 1><<<grid, TC_MAIN_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapO, mapO2, t);
  else conv_tc_kernel<0, 0><<<grid, TC_MAIN_THREADS, TC_SMEM_BYTES, st>>>(mapA, mapB, mapO, mapO2, t);
  SG_LAUNCH_CHECK();
  if (ksplit > 1) {
    const size_t n = (size_t)a.rows * (a.N / 4);
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a.splitk_ws, ksplit, (size_t)t.rows_pad * a.N, a.rows, a.N,
                                                                    a.epilogue == SG_EPI_BIAS ? a.bias : nullptr, a.out, a.ldo);
    SG_LAUNCH_CHECK();
    g_tc_extra = 1;
  }
  return SG_OK;
}
