// tcgen05 (3xTF32) implementation of the fused GCN layer -- placeholder until the kernel lands.
#include "common.cuh"

namespace gmeta {
bool gcn_layer_fwd_tc_supported(const GatherSrc&, int, int, int, const float*, int) { return false; }
int gcn_layer_fwd_tc(const GatherSrc&, const int32_t*, const int32_t*, const int32_t*, int, const float*,
                     int64_t, int, int, const float*, int64_t, int, int, const float*, float*, int,
                     cudaStream_t) {
  return GMETA_ERR_UNSUPPORTED;
}
}  // namespace gmeta
