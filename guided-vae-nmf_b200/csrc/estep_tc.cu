// E-step, tensor-core variant (tcgen05 / TMEM) -- placeholder until the kernel lands.
#include "gvn_common.cuh"

namespace gvn {

int32_t launch_pack_tc(const float*, const float*, const float*, const float*, int, int, int, unsigned char*,
                       cudaStream_t) {
  return GVN_OK;
}

int32_t launch_estep_tc(const gvn_batch*, const void*, int, int, float, const gvn_noise*, const gvn_trace*, int,
                        cudaStream_t) {
  return fail(GVN_E_UNSUPPORTED_SHAPE, "tensor-core E-step not built yet");
}

}  // namespace gvn
