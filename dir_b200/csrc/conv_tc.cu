// tcgen05 / TMA implicit-GEMM convolution (bf16 operands, fp32 accumulate in TMEM).
#include "engine.h"

namespace dirb200 {

bool conv_tc_supported(const ConvLayer&, int, int, int) { return false; }
int conv_tc_prepare_weights(ConvLayer&) { return 0; }
int launch_conv_tc(const ConvLayer&, const __nv_bfloat16*, __nv_bfloat16*, const __nv_bfloat16*, int, int, int,
                   cudaStream_t) {
  return DIRB200_E_STATE;
}

}  // namespace dirb200
