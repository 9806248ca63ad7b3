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
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) { const char* e = getenv("NHMC_PDL"); g_pdl = (e && e[0] == '0') ? 0 : 1; }
  return g_pdl != 0;
}
void pdl_set(int on) { g_pdl = on ? 1 : 0; }
}  // namespace nhmc

extern "C" {
int nhmc_set_pdl(int on) { const int prev = nhmc::pdl_enabled() ? 1 : 0; nhmc::pdl_set(on); return prev; }
int nhmc_version(void) { return 100; }
const char* nhmc_last_error(void) { return nhmc::g_err; }
size_t nhmc_sizeof_chain(void) { return sizeof(nhmc_chain); }
size_t nhmc_sizeof_op(void) { return sizeof(nhmc_op); }
}
