// Direct (CUDA-core) conv3d on the same packed operands as the tcgen05 path. One thread per
// (voxel, output channel), fp32 accumulation. It exists to cross-check the tensor-core kernels
// on the device (tests/test_conv_gpu.py) and as the selectable path VFD_CONV_IMPL=direct; it is
// far too slow to be a production path and bench.py never selects it.
#include <cuda_bf16.h>
#include <cstdint>
#include "vfd_internal.h"

namespace vfd {
typedef __nv_bfloat16 bf16;

__global__ void conv_direct_fwd_kernel(const bf16* __restrict__ x, long long x_ld, int cin,
                                       const bf16* __restrict__ wp, int w_rows, int cin_k,
                                       const float* __restrict__ bias, void* __restrict__ out,
                                       long long out_ld, int out_cols, int out_fp32, int N, int D,
                                       int H, int W, int kd, int kh, int kw) {
  const long long V = (long long)N * D * H * W;
  const long long total = V * out_cols;
  const int taps = kd * kh * kw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % out_cols);
    long long t = i / out_cols;
    const long long vox = t;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    t /= H;
    const int d = (int)(t % D);
    const int n = (int)(t / D);
    float acc = 0.f;
    if (co < w_rows) {
      if (bias != nullptr) acc = bias[co];
      int tap = 0;
      for (int a = 0; a < kd; ++a)
        for (int b = 0; b < kh; ++b)
          for (int c = 0; c < kw; ++c, ++tap) {
            const int dd = d + a - kd / 2, hh = h + b - kh / 2, ww = w + c - kw / 2;
            if (dd < 0 || dd >= D || hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const bf16* xp = x + ((((long long)n * D + dd) * H + hh) * W + ww) * x_ld;
            const bf16* wr = wp + ((long long)co * taps + tap) * cin_k;
            for (int ci = 0; ci < cin; ++ci)
              acc = fmaf(__bfloat162float(xp[ci]), __bfloat162float(wr[ci]), acc);
          }
    }
    if (out_fp32)
      reinterpret_cast<float*>(out)[vox * out_ld + co] = acc;
    else
      reinterpret_cast<bf16*>(out)[vox * out_ld + co] = __float2bfloat16(acc);
  }
}

// acc[tap][ci][co] += sum_vox dy[vox][co] * x[vox+tap][ci]; one thread per (tap, ci, co)
__global__ void conv_direct_wgrad_kernel(const bf16* __restrict__ dy, long long dy_ld, int cout,
                                         const bf16* __restrict__ x, long long x_ld, int cin,
                                         float* __restrict__ acc, int co_pad, int ci_pad, int N,
                                         int D, int H, int W, int kd, int kh, int kw) {
  const int taps = kd * kh * kw;
  const long long total = (long long)taps * cin * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    long long t = i / cout;
    const int ci = (int)(t % cin);
    const int tap = (int)(t / cin);
    const int c = tap % kw, b = (tap / kw) % kh, a = tap / (kw * kh);
    float s = 0.f;
    for (int n = 0; n < N; ++n)
      for (int d = 0; d < D; ++d) {
        const int dd = d + a - kd / 2;
        if (dd < 0 || dd >= D) continue;
        for (int h = 0; h < H; ++h) {
          const int hh = h + b - kh / 2;
          if (hh < 0 || hh >= H) continue;
          for (int w = 0; w < W; ++w) {
            const int ww = w + c - kw / 2;
            if (ww < 0 || ww >= W) continue;
            const float g = __bfloat162float(dy[((((long long)n * D + d) * H + h) * W + w) * dy_ld + co]);
            const float v = __bfloat162float(x[((((long long)n * D + dd) * H + hh) * W + ww) * x_ld + ci]);
            s = fmaf(g, v, s);
          }
        }
      }
    acc[((long long)tap * ci_pad + ci) * co_pad + co] += s;
  }
}
}  // namespace vfd

using namespace vfd;

VFD_API int vfd_conv3d_fwd_direct(const void* x, long long x_ld, int cin, const void* w_packed,
                                     int w_rows, int cin_k, const float* bias, void* out,
                                     long long out_ld, int out_cols, int out_fp32, int N, int D,
                                     int H, int W, int kd, int kh, int kw, void* stream_) {
  const long long total = (long long)N * D * H * W * out_cols;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  conv_direct_fwd_kernel<<<(int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      (const bf16*)x, x_ld, cin, (const bf16*)w_packed, w_rows, cin_k, bias, out, out_ld, out_cols,
      out_fp32, N, D, H, W, kd, kh, kw);
  return check_launch("conv_direct_fwd");
}

VFD_API int vfd_conv3d_wgrad_direct(const void* dy, long long dy_ld, int cout, const void* x,
                                       long long x_ld, int cin, float* acc, int co_pad, int ci_pad,
                                       int N, int D, int H, int W, int kd, int kh, int kw,
                                       void* stream_) {
  const long long total = (long long)kd * kh * kw * cin * cout;
  if (total == 0 || N == 0) return 0;
  long long blocks = (total + 127) / 128;
  if (blocks > 148 * 32) blocks = 148 * 32;
  conv_direct_wgrad_kernel<<<(int)blocks, 128, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(
      (const bf16*)dy, dy_ld, cout, (const bf16*)x, x_ld, cin, acc, co_pad, ci_pad, N, D, H, W, kd,
      kh, kw);
  return check_launch("conv_direct_wgrad");
}
