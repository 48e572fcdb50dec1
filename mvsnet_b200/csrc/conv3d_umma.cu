// placeholder (replaced below in the same commit series)
#include "common.cuh"
namespace mvsb200 {
size_t conv3d_umma_scratch_bytes(int, int, int) { return 0; }
int launch_conv3d_umma(const void*, int, const float*, const float*, const void*, const float*, const float*,
                       const float*, int, int, int, int, int, int, int, void*, int, double*, void*, cudaStream_t) {
  set_error("conv3d bf16/tcgen05 path not built yet");
  return MVSB200_ERR_UNSUPPORTED;
}
}  // namespace mvsb200
