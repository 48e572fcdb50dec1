// Error reporting, version and device queries of the C ABI (include/mvsnet_b200.h).
#include "common.cuh"
#include <string.h>

namespace mvsb200 {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" const char* mvsb200_last_error(void) { return t_error; }

extern "C" int mvsb200_version(void) { return 100; }

extern "C" uint64_t mvsb200_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

extern "C" int mvsb200_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  MVS_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MVS_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (prop.major != 10) {
    set_error("mvsnet_b200 is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return MVSB200_ERR_UNSUPPORTED;
  }
  return MVSB200_OK;
}
