// tcgen05 tensor-core gather-GEMM (placeholder until the kernel lands; the dispatcher in capi.cu
// only routes here when tc_available() says so).
#include "common.cuh"
namespace scn {
int tc_available() { return 0; }
int launch_conv_plan_tc(const float *, float *, const float *, const int *, const int *, int, int, int, int, const float *, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return -4;
}
} // namespace scn
