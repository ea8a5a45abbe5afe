/* Test / tool-only entry points of libvfd_b200_debug.so (the product sources built with -DVFD_DEBUG plus
 * csrc/conv_direct.cu). NOT part of the drop-in boundary: libvfd_b200.so exports none of them. */
#ifndef VFD_B200_DEBUG_H
#define VFD_B200_DEBUG_H
#include "vfd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Stage-isolation timing of the conv pipelines (tools/gpu_stage_probe.py): bit 0 skips the TMA loads, bit 1 the
 * tcgen05.mma issue, bit 2 the epilogue arithmetic / stores. Results are garbage while any bit is set. */
VFD_API int vfd_set_debug(int flags);

/* CUDA-core convs on the same packed operands as vfd_conv3d_fwd / vfd_conv3d_wgrad; the tests cross-check the
 * tcgen05 kernels against them on the device (tests/test_parity_gpu.py::test_conv_fwd_dgrad_wgrad). */
VFD_API int vfd_conv3d_fwd_direct(const void* x, long long x_ld, int cin, const void* w_packed,
                                  int w_rows, int cin_k, const float* bias, void* out,
                                  long long out_ld, int out_cols, int out_fp32, int N, int D, int H,
                                  int W, int kd, int kh, int kw, void* stream);
VFD_API int vfd_conv3d_wgrad_direct(const void* dy, long long dy_ld, int cout, const void* x,
                                    long long x_ld, int cin, float* acc, int co_pad, int ci_pad, int N,
                                    int D, int H, int W, int kd, int kh, int kw, void* stream);

#ifdef __cplusplus
}
#endif
#endif
