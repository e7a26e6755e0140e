// bf16 tcgen05 path — placeholder until the tensor-core kernel lands (next commit).
#include "dpi_common.cuh"
#include "dpi_kernels.h"

namespace dpi {
int tc_pack_network(const dpi_network&, const ProblemDev&, const MlpDev&, std::vector<void*>&, TcNetDev*, bool* ready,
                    int64_t*, cudaStream_t) {
  *ready = false;
  return DPI_OK;
}
int launch_cha_tc_path(const ProblemDev&, const TcNetDev&, const float*, int64_t, int, uint64_t, int64_t, const float*,
                       const float*, const float*, const float*, float, float*, cudaStream_t) {
  set_error("bf16 tensor-core path not built");
  return DPI_ERR_UNSUPPORTED;
}
}  // namespace dpi
