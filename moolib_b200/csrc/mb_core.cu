// Error reporting and device queries shared by the C-ABI entry points.
#include "mb_common.cuh"

#include <mutex>

namespace mb {

namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? MB_ENOMEM : MB_ECUDA;
}

int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) return -1;
  return d;
}

int sm_count(int device) {
  static std::mutex m;
  static int cache[64];
  if (device < 0 || device >= 64) {
    set_error("sm_count: bad device %d", device);
    return -1;
  }
  std::lock_guard<std::mutex> l(m);
  if (cache[device] == 0) {
    int n = 0;
    cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
      cuda_fail(e, "cudaDeviceGetAttribute(MultiProcessorCount)", __FILE__, __LINE__);
      return -1;
    }
    cache[device] = n;
  }
  return cache[device];
}

}  // namespace mb

extern "C" {
int mb_version(void) { return MB_VERSION; }
const char* mb_last_error(void) { return mb::g_err; }
int mb_sm_count(int device) { return mb::sm_count(device); }
}
