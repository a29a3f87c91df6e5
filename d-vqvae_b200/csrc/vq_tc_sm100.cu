// tcgen05 filter kernel — placeholder until the tensor-core path lands.
#include "dvq_common.cuh"

namespace dvq {
bool vq_tc_supported(int64_t, int, int) { return false; }
size_t vq_tc_operand_bytes(int, int) { return 0; }
int launch_vq_tc(const float*, const float*, const float*, int64_t, int, int, int, float*, int64_t*,
                 unsigned long long*, double*, void*, int*, int*, cudaStream_t) {
  return fail(DVQ_ERR_BAD_SHAPE, "tcgen05 path not built");
}
}  // namespace dvq
