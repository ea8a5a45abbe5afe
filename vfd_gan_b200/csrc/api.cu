// Error state + version entry points of the C-ABI (include/vfd_b200.h).
#include <cstdio>
#include <cstring>
#include "vfd_internal.h"

namespace vfd {
static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  (void)cudaGetLastError();  // clear the (non-sticky) error so later launches are judged on their own
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return VFD_ERR_CUDA;
}
int check_launch(const char* kernel_name) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, kernel_name);
  return VFD_OK;
}
}  // namespace vfd

VFD_API const char* vfd_last_error(void) { return vfd::g_err; }
VFD_API int vfd_abi_version(void) { return 3; }
