// Version / error plumbing of the C ABI (include/ast_sm100.h).
#include <stdarg.h>
#include <string.h>

#include "ast_common.cuh"

namespace ast {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace ast

extern "C" int ast_version(void) { return AST_ABI_VERSION; }

extern "C" const char* ast_last_error(void) { return ast::g_last_error; }

extern "C" int ast_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    ast::set_error("ast_device_check: cudaGetDevice: %s", cudaGetErrorString(e));
    return AST_ERR_CUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    ast::set_error("ast_device_check: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev,
                   major, minor);
    return AST_ERR_UNSUPPORTED;
  }
  return AST_OK;
}
