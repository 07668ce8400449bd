// Host-side plumbing shared by all entry points: thread-local error text, launch checks.
#include <stdarg.h>
#include <string.h>

#include "ub_common.cuh"

namespace ub {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch error so later calls are not poisoned
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return UB_ERR_LAUNCH;
  }
  return UB_OK;
}

int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

}  // namespace ub

extern "C" {

int ub_abi_version(void) { return UB_ABI_VERSION; }

const char* ub_last_error(void) { return ub::g_last_error; }

int ub_sm_count(void) { return ub::sm_count(); }

}  // extern "C"
