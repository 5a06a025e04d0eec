// capi_util.cu -- error plumbing and device checks shared by every C-ABI entry point.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace hyp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HYP_ERR_CUDA;
  }
  return HYP_OK;
}

}  // namespace hyp

extern "C" int hyp_abi_version(void) { return HYP_ABI_VERSION; }

extern "C" const char *hyp_last_error(void) { return hyp::g_err; }

extern "C" int hyp_check_device(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    hyp::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return HYP_ERR_CUDA;
  }
  if (p.major != 10) {
    hyp::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, p.major, p.minor);
    return HYP_ERR_UNSUPPORTED;
  }
  return HYP_OK;
}
