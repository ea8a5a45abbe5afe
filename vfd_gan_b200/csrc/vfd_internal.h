// Error plumbing shared by every translation unit of libvfd_b200.so.
#pragma once
#include <cuda_runtime.h>
#include "../../include/vfd_b200.h"

#define VFD_OK 0
#define VFD_ERR_ARG 1
#define VFD_ERR_CUDA 2
#define VFD_ERR_DRIVER 3

namespace vfd {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* kernel_name);
}  // namespace vfd
