// Handle layout and small helpers shared by the API translation units (sg_api.cu, sg_api_audionet.cu).
#pragma once
#include <vector>

#include "sg_audionet.cuh"
#include "sg_common.cuh"
#include "sg_head.cuh"

// per-category device timing (CUDA events on the launching stream); off by default
struct SgProf {
  bool on = false;
  std::vector<cudaEvent_t> ev;      // pairs (start, stop)
  std::vector<int> cat;             // category of pair i
  std::vector<int> tag;             // caller-defined label of pair i (TDNN layer 1..5; 0 = none)
  int next_tag = 0;                 // label of the next recorded launch (reset to 0 after each one)
  size_t used = 0;                  // pairs recorded since the last reset
  cudaEvent_t* begin(int c, cudaStream_t st) {
    if (!on) return nullptr;
    if (used == cat.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return nullptr;
      ev.push_back(a); ev.push_back(b); cat.push_back(c); tag.push_back(0);
    }
    cat[used] = c; tag[used] = next_tag; next_tag = 0;
    cudaEventRecord(ev[2 * used], st);
    return &ev[2 * used + 1];
  }
  void end(cudaEvent_t* stop, cudaStream_t st) {
    if (!stop) return;
    cudaEventRecord(*stop, st);
    ++used;
  }
};

// One captured PGD iteration per ping-pong parity (sg_pgd_run): replayed for every iteration of every attack whose shapes,
// workspace and parameters match `key`.
struct SgPgdGraph {
  cudaGraphExec_t exec[2] = {nullptr, nullptr};
  unsigned long long key[16] = {0};
  cudaStream_t cap_stream = nullptr;   // private capture stream (the caller's may be the legacy default stream)
  int kernels = 0;                     // kernels per captured iteration (launch accounting)
  bool valid = false;
};

struct sg_handle {
  SgProf prof;
  int device = 0;
  int precision = SG_PREC_FP32;
  long long launches = 0;
  long long cw2_iters = 0;          // gradient iterations executed by the last sg_cw2_audionet_run
  int l1_tap_form = 1;              // SG_OPT_L1_TAP_FORM: bf16 mode computes the layer-1 dgrad per tap (K = 512) + a shifted sum
  int feat_stash = 1;               // SG_OPT_FEAT_STASH: the fused attack loop hands the per-frame forward state to the MFCC adjoint
  int utt_offset = 0;               // SG_OPT_UTT_OFFSET: global index of utterance 0 (philox dither key)
  int use_graph = 1;                // SG_OPT_CUDA_GRAPH: sg_pgd_run replays one captured iteration instead of ~26 launches per pass
  SgPgdGraph pgd_graph;
  float* cm_part = nullptr; unsigned int* cm_count = nullptr; int cm_cap = 0;   // fused-CMVN scratch (chunk sums, counters) for cm_cap utterances
  int row_compaction = 1;           // SG_OPT_ROW_COMPACTION: layers 4 / 5, pooling and their adjoints on the valid frames only (tensor-core modes)
  int cmvn_fusion = 0;              // SG_OPT_CMVN_FUSION: utterances of <= 300 frames run CMVN (and its adjoint) inside the MFCC kernels
                                    // (off: measured 2 ms per PGD-100 step SLOWER than the two 23 us cmvn launches, see sgb200.h)
  int pool_fusion = 1;              // SG_OPT_POOL_FUSION: bf16 mode contracts the pooling adjoint inside the layer-5 dgrad
  SgFeatTables* d_tables = nullptr;
  bool xv_loaded = false;
  bool backend_loaded = false;      // PLDA back-end + enrolled speakers (set by sg_load_xv / sg_load_iv)
  int L = 0, Lp = 0, S = 0;
  // packed TDNN weights
  float* Wf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [taps*cinP, coutP]
  float* Wb[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [taps*coutP, cinP]
  float* Wfk[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // K-major copies for the tensor-core path: [coutP, taps*cinP]
  float* Wbk[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [cinP, taps*coutP]
  void* Wfk_h[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // bf16 copies of Wfk / Wbk (SG_PREC_BF16)
  void* Wbk_h[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void* Wg1_h = nullptr;            // layer-1 dgrad in per-tap form: bf16 [taps * 32, 512], row k*32 + c = w[:, c, k]
  float* bias[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // [coutP] (BN of the previous layer folded)
  float* bn5_mean = nullptr; float* bn5_istd = nullptr;           // [C5P]
  float* Wfc = nullptr; float* Wfc_b = nullptr; float* bfc = nullptr;     // fc1: [3072,512], [512,3072], [512]
  float* Wlda = nullptr; float* Wlda_b = nullptr; float* blda = nullptr;  // LDA: [512,Lp], [Lp,512], [Lp]
  // K-major copies for the tensor-core path ([N, K]): transposes of the four matrices above
  float* Wfc_k = nullptr; float* Wfc_bk = nullptr; float* Wlda_k = nullptr; float* Wlda_bk = nullptr;
  float* plda_mean = nullptr; float* plda_T = nullptr; float* plda_Tt = nullptr;
  float* inv_psi1 = nullptr; float* psi_ratio = nullptr; float* inv_var_given = nullptr;
  float* enroll = nullptr;
  SgHeadConst H;
  std::vector<void*> allocs;
  void* comm = nullptr; int comm_rank = 0, comm_world = 0;   // NCCL communicator of the metric all-reduce (sg_comm.cu)
  struct SgAudioNet* an = nullptr;   // AudioNet state (sg_api_audionet.cu)
  struct SgIv* iv = nullptr;         // i-vector system state (sg_api_iv.cu)
};


#define SG_TRY(expr) do { int _r = (expr); if (_r != SG_OK) return _r; } while (0)
#define PROF(h, c, st, call) do { cudaEvent_t* _pe = (h)->prof.begin((c), (st)); int _pr = (call); (h)->prof.end(_pe, (st)); if (_pr != SG_OK) return _pr; } while (0)

int sg_dev_upload(sg_handle* h, float** dst, const std::vector<float>& src);
int sg_check_handle(sg_handle* h, bool need_xv);
int sg_run_conv(sg_handle* h, const SgConvArgs& a, bool tensor_ok, int cat, cudaStream_t st);
void sg_audionet_free(sg_handle* h);
void sg_iv_free(sg_handle* h);
int sg_load_backend(sg_handle* h, const float* plda_mean, const float* plda_transform, const float* plda_psi,
                    const float* enroll, int L, int S);
extern "C" int sg_comm_destroy(sg_handle* h);
