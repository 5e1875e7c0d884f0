// ABI bookkeeping: version, thread-local error text, device query.
#include "vu_common.cuh"

namespace vu {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int fail_arg(const char* fn, const char* what) {
  g_err = std::string(fn) + ": " + what;
  return VU_ERR_ARG;
}
int check_launch(const char* fn) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_err = std::string(fn) + ": CUDA error: " + cudaGetErrorString(e);
    return VU_ERR_CUDA;
  }
  return VU_OK;
}
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount; cached_dev = dev;
  }
  return cached;
}
}  // namespace vu

extern "C" int vu_version(void) { return 9; }
extern "C" const char* vu_last_error(void) { return vu::g_err.c_str(); }
extern "C" int vu_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaGetLastError();
    vu::set_error("vu_device_sm_count: no such CUDA device");
    return -1;
  }
  return n;
}
