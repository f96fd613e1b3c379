// C-ABI glue: error message storage, version, device check.
#include "common.cuh"
#include "../../include/diffusynth_b200.h"
#include <cstdarg>

namespace ds {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
}  // namespace ds

extern "C" {
const char* ds_last_error(void) { return ds::get_error(); }
int ds_version(void) { return 100; }
int ds_operand_dtype(void) { return ds::kOperandIsFp16; }
int ds_check_device(int dev) {
  cudaDeviceProp p;
  DS_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  DS_REQUIRE(p.major == 10, "device %d is sm_%d%d; this library contains sm_100a code only", dev, p.major, p.minor);
  return ds::DS_OK;
}
}
