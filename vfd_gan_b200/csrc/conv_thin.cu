// Weight gradient of THIN 1x1x1 GEMMs (<= 32 channels on both sides) -- the tap-folded first / last layers
// and the discriminator's 1x1x1 convs (autograd of nn.Conv3d, models/mygannet.py:311,344 through
// models/spatiotempconv.py:49-50,59-60):
//     acc[ci][co] += sum_v x[v][ci] * dy[v][co]
// These layers are pure HBM streams (6.4 M voxels x <= 128 bytes, ~1 kFLOP per voxel): a 128-row tcgen05 tile
// would be > 90 % padding and its single-thread TMA / MMA issue cannot keep up with the memory system. Here
// every warp streams 16-voxel groups with coalesced 128-bit loads, stages them in padded shared memory and
// feeds warp-level mma.sync (m16n8k16, bf16 -> fp32) through ldmatrix.trans; the 32 x 32 fp32 accumulator
// lives in registers for the whole kernel and is reduced block-wide before one atomic per element.
#include <cuda_bf16.h>
#include <cstdint>
#include "vfd_internal.h"

namespace vfd {

typedef __nv_bfloat16 bf16;
constexpr int kThinThreads = 256;
constexpr int kThinRowBytes = 80;   // 16 voxels x (64 B data + 16 B pad): conflict-free ldmatrix rows

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MT: 16-row tiles of input channels (1 or 2), NT: 8-column tiles of output channels (1..4)
template <int MT, int NT>
__global__ void __launch_bounds__(kThinThreads)
thin_wgrad_kernel(const bf16* __restrict__ dy, long long dy_ld, int cout, const bf16* __restrict__ x,
                  long long x_ld, int cin, float* __restrict__ acc, int co_pad, long long V) {
  const int xchunks = (cin + 7) >> 3;   // 16-byte chunks that exist in a row of x (dy has exactly NT)
  __shared__ __align__(16) uint8_t tiles[kThinThreads / 32][2][16 * kThinRowBytes];
  __shared__ float red[32 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 32; i += kThinThreads) red[i] = 0.f;
  __syncthreads();

  uint8_t* tx = tiles[warp][0];
  uint8_t* ty = tiles[warp][1];
  const uint32_t tx_s = static_cast<uint32_t>(__cvta_generic_to_shared(tx));
  const uint32_t ty_s = static_cast<uint32_t>(__cvta_generic_to_shared(ty));
  float c[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) c[m][n][i] = 0.f;

  // lane -> (voxel row, 16-byte chunk) of the staged tiles: two chunks per lane and tensor
  const int lrow = lane >> 2, lchunk = lane & 3;
  // ldmatrix row addresses: matrix = lane / 8, row = lane % 8
  const int mat = lane >> 3, mrow = lane & 7;
  const uint32_t a_off = ((mat >> 1) * 8 + mrow) * kThinRowBytes + (mat & 1) * 16;   // + mtile * 32 bytes
  const uint32_t b_off = ((mat & 1) * 8 + mrow) * kThinRowBytes + (mat >> 1) * 16;   // + ntile pair * 32 bytes

  const long long groups = (V + 15) >> 4;
  const long long gstride = static_cast<long long>(gridDim.x) * (kThinThreads / 32);
  for (long long grp = static_cast<long long>(blockIdx.x) * (kThinThreads / 32) + warp; grp < groups; grp += gstride) {
    const long long v0 = grp << 4;
    uint4 rx[2], ry[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long v = v0 + lrow + 8 * i;
      rx[i] = make_uint4(0u, 0u, 0u, 0u);
      ry[i] = make_uint4(0u, 0u, 0u, 0u);
      if (v < V) {
        if (lchunk < xchunks) rx[i] = __ldg(reinterpret_cast<const uint4*>(x + v * x_ld) + lchunk);
        if (lchunk < NT) ry[i] = __ldg(reinterpret_cast<const uint4*>(dy + v * dy_ld) + lchunk);
      }
    }
    __syncwarp();   // the previous group's ldmatrix reads are done
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<uint4*>(tx + (lrow + 8 * i) * kThinRowBytes + lchunk * 16) = rx[i];
      *reinterpret_cast<uint4*>(ty + (lrow + 8 * i) * kThinRowBytes + lchunk * 16) = ry[i];
    }
    __syncwarp();
    uint32_t a[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) ldmatrix_x4_trans(tx_s + a_off + m * 32, a[m][0], a[m][1], a[m][2], a[m][3]);
#pragma unroll
    for (int np = 0; np < (NT + 1) / 2; ++np) {
      uint32_t b0, b1, b2, b3;   // n-tile 2np: (b0, b1); n-tile 2np+1: (b2, b3)
      ldmatrix_x4_trans(ty_s + b_off + np * 32, b0, b1, b2, b3);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        mma_bf16_16816(c[m][2 * np], a[m][0], a[m][1], a[m][2], a[m][3], b0, b1);
        if (2 * np + 1 < NT) mma_bf16_16816(c[m][2 * np + 1], a[m][0], a[m][1], a[m][2], a[m][3], b2, b3);
      }
    }
  }

  // block reduction, then one atomic per valid element
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      const int ci = 16 * m + g, co = 8 * n + 2 * t;
      atomicAdd(&red[ci * 32 + co], c[m][n][0]);
      atomicAdd(&red[ci * 32 + co + 1], c[m][n][1]);
      atomicAdd(&red[(ci + 8) * 32 + co], c[m][n][2]);
      atomicAdd(&red[(ci + 8) * 32 + co + 1], c[m][n][3]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += kThinThreads) {
    const int ci = i >> 5, co = i & 31;
    if (ci < cin && co < cout) atomicAdd(acc + static_cast<size_t>(ci) * co_pad + co, red[i]);
  }
}

}  // namespace vfd

using namespace vfd;

VFD_API int vfd_conv3d_wgrad_thin(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                                     int cin, float* acc, int co_pad, int ci_pad, long long V, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (V <= 0) return 0;
  if (cout < 1 || cout > 32 || cin < 1 || cin > 32 || co_pad < cout || ci_pad < cin)
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: needs 1 <= cin, cout <= 32 and a large enough accumulator");
  if ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) || (dy_ld % 8) || (x_ld % 8) ||
      dy_ld < ((cout + 7) & ~7) || x_ld < ((cin + 7) & ~7))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: tensors must be 16-byte aligned channels-last bf16");
  const int mt = (cin + 15) / 16, nt = (cout + 7) / 8;
  // the staged tiles read whole 16-byte chunks: chunks beyond the tensor's padded width are not loaded
  const int grid = 148 * 4;
#define VFD_THIN(MT, NT)                                                                                     \
  if (mt == MT && nt == NT) {                                                                                \
    thin_wgrad_kernel<MT, NT><<<grid, kThinThreads, 0, stream>>>((const bf16*)dy, dy_ld, cout, (const bf16*)x, \
                                                                 x_ld, cin, acc, co_pad, V);                 \
    return check_launch("thin_wgrad");                                                                       \
  }
  VFD_THIN(1, 1) VFD_THIN(1, 2) VFD_THIN(1, 3) VFD_THIN(1, 4)
  VFD_THIN(2, 1) VFD_THIN(2, 2) VFD_THIN(2, 3) VFD_THIN(2, 4)
#undef VFD_THIN
  return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: unsupported tile shape");
}
