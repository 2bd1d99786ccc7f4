// Error reporting and device queries shared by every entry point of libdpl_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "dpl_common.cuh"

namespace dpl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace dpl

extern "C" int dpl_version(void) { return DPL_VERSION; }
extern "C" const char* dpl_last_error(void) { return dpl::g_err; }
