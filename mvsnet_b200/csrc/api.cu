// Error reporting, version and device queries of the C ABI (include/mvsnet_b200.h).
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <mutex>

namespace mvsb200 {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

// ---- tuning snapshot ---------------------------------------------------------------------------------------------
namespace {
struct TuningField { const char* name; int Tuning::*field; };
const TuningField kTuningInts[] = {
    {"NO_FUSED_REGRESS", &Tuning::no_fused_regress}, {"CV_FP32_TAPS", &Tuning::cv_fp32_taps},
    {"CV_KERNEL", &Tuning::cv_kernel}, {"CV_FP32_BLEND", &Tuning::cv_fp32_blend}, {"CV_MINB", &Tuning::cv_minb},
    {"CV_REC16", &Tuning::cv_rec16}, {"CV_KDC", &Tuning::cv_kdc}, {"CV_PLANES", &Tuning::cv_planes},
    {"CV_STATS", &Tuning::cv_stats}, {"CV_DBG", &Tuning::cv_dbg}, {"TC_ZF", &Tuning::tc_zf}, {"TC_FUSE01", &Tuning::tc_fuse01}, {"TC_RANK", &Tuning::tc_rank}, {"INFER_SIDE", &Tuning::infer_side}, {"TC_TRIM", &Tuning::tc_trim}, {"TC_XF_GROUPS", &Tuning::tc_xf_groups}, {"TC_XFOLD", &Tuning::tc_xfold},
    {"TC_ZSPLIT", &Tuning::tc_zsplit}, {"TC_DBG", &Tuning::tc_dbg}, {"TC_VERBOSE", &Tuning::tc_verbose},
    {"TC_PROF", &Tuning::tc_prof}, {"TC_EXACT_SMEM", &Tuning::tc_exact_smem}, {"TC_NO_PDL", &Tuning::tc_no_pdl},
    {"REGNET_PROFILE", &Tuning::regnet_profile}, {"UNET_NO_TILE", &Tuning::unet_no_tile},
    {"UNET_PROFILE", &Tuning::unet_profile}, {"UNET_FP32", &Tuning::unet_fp32}, {"UNET_MB", &Tuning::unet_mb}, {"UNET_DBG", &Tuning::unet_dbg}, {"UNET_OBUF", &Tuning::unet_obuf}, {"UNET_HOT", &Tuning::unet_hot}, {"UNET_INPLACE", &Tuning::unet_inplace},
};
std::atomic<const Tuning*> g_tuning{nullptr};
std::mutex g_tuning_mutex;

// value == NULL restores the default of the field
bool apply_tuning(Tuning* t, const char* name, const char* value) {
  const Tuning defaults;
  for (const TuningField& f : kTuningInts)
    if (!strcmp(name, f.name)) {
      // flags given without a number ("MVSB200_TC_PROF=" or "=yes") count as 1
      t->*(f.field) = value ? ((*value == '-' || (*value >= '0' && *value <= '9')) ? atoi(value) : 1) : defaults.*(f.field);
      return true;
    }
  if (!strcmp(name, "TC_TILE")) {
    t->tc_tile_x = t->tc_tile_y = 0;
    if (value) sscanf(value, "%dx%d", &t->tc_tile_x, &t->tc_tile_y);
    return true;
  }
  if (!strcmp(name, "TC_LAYER")) {
    t->tc_layer_set = 0;
    if (value && sscanf(value, "%d,%d,%d", &t->tc_layer[0], &t->tc_layer[1], &t->tc_layer[2]) == 3) t->tc_layer_set = 1;
    return true;
  }
  return false;
}

const Tuning* tuning_init_locked() {
  const Tuning* cur = g_tuning.load(std::memory_order_acquire);
  if (cur) return cur;
  Tuning* t = new Tuning();
  char env[64];
  for (const TuningField& f : kTuningInts) {
    snprintf(env, sizeof(env), "MVSB200_%s", f.name);
    if (const char* v = getenv(env)) apply_tuning(t, f.name, v);
  }
  for (const char* name : {"TC_TILE", "TC_LAYER"}) {
    snprintf(env, sizeof(env), "MVSB200_%s", name);
    if (const char* v = getenv(env)) apply_tuning(t, name, v);
  }
  g_tuning.store(t, std::memory_order_release);
  return t;
}
}  // namespace

const Tuning& tuning() {
  const Tuning* t = g_tuning.load(std::memory_order_acquire);
  if (t) return *t;
  std::lock_guard<std::mutex> lock(g_tuning_mutex);
  return *tuning_init_locked();
}

int sm_count_current() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = cache[dev & 63].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  cache[dev & 63].store(v, std::memory_order_relaxed);
  return v;
}

}  // namespace mvsb200

using namespace mvsb200;

// Development / tuning switch by name (the MVSB200_<NAME> environment variables are only its initial values, read
// once): value NULL restores the default.  Old snapshots are kept alive (a call in flight may still read one).
extern "C" int mvsb200_set_tuning(const char* name, const char* value) {
  MVS_CHECK_ARG(name != nullptr, "set_tuning: NULL name");
  std::lock_guard<std::mutex> lock(g_tuning_mutex);
  Tuning* t = new Tuning(*tuning_init_locked());
  if (!apply_tuning(t, name, value)) {
    delete t;
    set_error("set_tuning: unknown switch '%s'", name);
    return MVSB200_ERR_INVALID;
  }
  g_tuning.store(t, std::memory_order_release);
  return MVSB200_OK;
}

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
