// tcgen05 / TMEM tensor-core path for the TDNN contractions (placeholder until the kernel lands).
#include "sg_common.cuh"

int sg_conv_tc(const SgConvArgs& a, int precision, cudaStream_t st) {
  (void)a; (void)precision; (void)st;
  sg_set_error("tensor-core path not built yet: use SG_PREC_FP32");
  return SG_EUNSUPPORTED;
}
