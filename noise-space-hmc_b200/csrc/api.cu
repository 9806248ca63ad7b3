// Error plumbing and trivial queries of the C ABI (include/nhmc.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace nhmc {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NHMC_OK;
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return NHMC_ERR_CUDA;
}
}  // namespace nhmc

extern "C" {
int nhmc_version(void) { return 200; }
const char* nhmc_last_error(void) { return nhmc::g_err; }
size_t nhmc_sizeof_chain(void) { return sizeof(nhmc_chain); }
size_t nhmc_sizeof_op(void) { return sizeof(nhmc_op); }
}
