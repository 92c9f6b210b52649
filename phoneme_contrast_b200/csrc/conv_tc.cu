// tcgen05 tensor-core implicit-GEMM convolution (sm_100a). Placeholder dispatcher until the kernel lands:
// reports PC_EUNSUPPORTED so pc_conv_fwd falls through to the exact-fp32 SIMT kernel.
#include "common.cuh"

extern "C" int pc_conv_fwd_tc(const float*, const float*, const float*, const PcConvGeom*, const PcInXform*, float*, double*, int,
                              pc_stream_t) {
  return PC_EUNSUPPORTED;
}
