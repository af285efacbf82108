// Thread-local error string of the C ABI.
#include "common.cuh"

namespace xlbn {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(error_buffer(), 512, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

}  // namespace xlbn
