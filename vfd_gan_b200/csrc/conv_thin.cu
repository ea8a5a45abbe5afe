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

// Tap folding fused into the stream (see vfd_tap_gather): one side of the GEMM is not read as stored but
// gathered on the fly, element e of the folded row = channel e % CS of the source voxel shifted by tap e / CS.
struct FastDivT {
  uint32_t m, s, d;
};
__device__ __forceinline__ uint32_t fdivt(uint32_t n, const FastDivT& f) { return (__umulhi(n, f.m) + n) >> f.s; }
struct FoldGeom {
  int D, H, W;
  FastDivT fW, fH, fD;
};
struct NoFold {
  static constexpr bool kFold = false;
};
template <int KD, int KH, int KW, int CS_>
struct Fold {
  static constexpr bool kFold = true;
  static constexpr int TAPS = KD * KH * KW, CS = CS_;
  // 16-byte chunk `chunk` of the folded row of voxel (n, d, h, w); sign = +1 gathers x, -1 gathers dy
  static __device__ __forceinline__ uint4 chunk_of(const bf16* __restrict__ src, long long ld, const FoldGeom& g,
                                                   unsigned n, int d, int h, int w, int chunk, int sign) {
    uint32_t words[4] = {0u, 0u, 0u, 0u};
    const unsigned short* s16 = reinterpret_cast<const unsigned short*>(src);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = chunk * 8 + j;
      if (e < TAPS * CS) {
        const int tap = e / CS, ch = e - tap * CS;
        const int a = tap / (KH * KW), r = tap - a * (KH * KW), b = r / KW, c = r - b * KW;
        const int dd = d + sign * (a - KD / 2), hh = h + sign * (b - KH / 2), ww = w + sign * (c - KW / 2);
        if (dd >= 0 && dd < g.D && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W) {
          const uint32_t val = __ldg(s16 + (size_t)(((n * g.D + dd) * g.H + hh) * g.W + ww) * ld + ch);
          words[j >> 1] |= val << ((j & 1) * 16);
        }
      }
    }
    return make_uint4(words[0], words[1], words[2], words[3]);
  }
};

// MT: 16-row tiles of input channels (1 or 2), NT: 8-column tiles of output channels (1..4).
// FX / FY: tap folding of the x / dy side (at most one of them).
template <int MT, int NT, typename FX = NoFold, typename FY = NoFold>
__global__ void __launch_bounds__(kThinThreads)
thin_wgrad_kernel(const bf16* __restrict__ dy, long long dy_ld, int cout, const bf16* __restrict__ x,
                  long long x_ld, int cin, float* __restrict__ acc, int co_pad, long long V, FoldGeom fg,
                  float* __restrict__ partial) {
  const int xchunks = (cin + 7) >> 3;   // 16-byte chunks that exist in a (folded) row of x (dy has exactly NT)
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
  // GU groups per iteration: all their loads are issued before the first one is consumed, so a warp keeps
  // GU x 4 independent 128-bit loads in flight (a single 16-voxel group of a 3 -> 2 channel layer is 512 bytes --
  // far too little to cover the DRAM latency). The on-the-fly tap gathers are instruction bound: GU = 1 there.
  constexpr int GU = (FX::kFold || FY::kFold) ? 1 : 4;
  for (long long grp0 = static_cast<long long>(blockIdx.x) * (kThinThreads / 32) + warp; grp0 < groups;
       grp0 += gstride * GU) {
    uint4 rx[GU][2], ry[GU][2];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const long long v0 = (grp0 + u * gstride) << 4;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const long long v = v0 + lrow + 8 * i;
        rx[u][i] = make_uint4(0u, 0u, 0u, 0u);
        ry[u][i] = make_uint4(0u, 0u, 0u, 0u);
        if (v < V) {
          unsigned n = 0;
          int d = 0, h = 0, w = 0;
          if (FX::kFold || FY::kFold) {
            const unsigned vv = static_cast<unsigned>(v);
            const unsigned t1 = fdivt(vv, fg.fW);
            w = vv - t1 * fg.W;
            const unsigned t2 = fdivt(t1, fg.fH);
            h = t1 - t2 * fg.H;
            n = fdivt(t2, fg.fD);
            d = t2 - n * fg.D;
          }
          if (lchunk < xchunks) {
            if constexpr (FX::kFold) rx[u][i] = FX::chunk_of(x, x_ld, fg, n, d, h, w, lchunk, 1);
            else rx[u][i] = __ldg(reinterpret_cast<const uint4*>(x + v * x_ld) + lchunk);
          }
          if (lchunk < NT) {
            if constexpr (FY::kFold) ry[u][i] = FY::chunk_of(dy, dy_ld, fg, n, d, h, w, lchunk, -1);
            else ry[u][i] = __ldg(reinterpret_cast<const uint4*>(dy + v * dy_ld) + lchunk);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      if (grp0 + u * gstride >= groups) break;   // warp-uniform
      __syncwarp();   // the previous group's ldmatrix reads are done
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        *reinterpret_cast<uint4*>(tx + (lrow + 8 * i) * kThinRowBytes + lchunk * 16) = rx[u][i];
        *reinterpret_cast<uint4*>(ty + (lrow + 8 * i) * kThinRowBytes + lchunk * 16) = ry[u][i];
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
  }

  // block reduction in warp order (every lane of a warp owns distinct elements, so the sum does not depend on
  // timing), then one atomic per valid element -- or, in deterministic mode, the block's own partial slot
  const int g = lane >> 2, t = lane & 3;
  for (int w = 0; w < kThinThreads / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const int ci = 16 * m + g, co = 8 * n + 2 * t;
          red[ci * 32 + co] += c[m][n][0];
          red[ci * 32 + co + 1] += c[m][n][1];
          red[(ci + 8) * 32 + co] += c[m][n][2];
          red[(ci + 8) * 32 + co + 1] += c[m][n][3];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 32 * 32; i += kThinThreads) {
    const int ci = i >> 5, co = i & 31;
    if (partial != nullptr) partial[static_cast<size_t>(blockIdx.x) * 1024 + i] = red[i];
    else if (ci < cin && co < cout) atomicAdd(acc + static_cast<size_t>(ci) * co_pad + co, red[i]);
  }
}

// ------------------------------------------------------------------------------------------ tiny 1x1x1 forward
// 1x1x1 conv with at most 8 input and 8 output channels (the temporal discriminator's first "spatial" conv,
// 3 -> 2 channels, models/spatiotempconv.py:49-50 with kernel (3,1,1)): 32 bytes of traffic per voxel and 64
// FMAs -- a pure HBM stream that the 128-row tcgen05 tile serves at a fifth of the bandwidth. One thread per
// voxel: one 128-bit load, the 8 x 8 product from registers, one 128-bit store; the BatchNorm statistics of the
// stored (bf16-rounded) values are kept in registers and reduced once per block.
__global__ void __launch_bounds__(256)
tiny_pointwise_kernel(const bf16* __restrict__ x, long long x_ld, const bf16* __restrict__ w_packed, int cin_k,
                      const float* __restrict__ bias, bf16* __restrict__ out, long long out_ld, long long V,
                      double* __restrict__ stats, int stats_ld) {
  __shared__ double s_stat[16];   // exact sums of the warps' fp32 partials: order-independent
  float w[8][8], b[8], ssum[8], ssq[8];
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    b[co] = bias != nullptr ? __ldg(bias + co) : 0.f;
    ssum[co] = ssq[co] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) w[co][ci] = __bfloat162float(w_packed[co * cin_k + ci]);
  }
  if (threadIdx.x < 16) s_stat[threadIdx.x] = 0.0;
  __syncthreads();
  // 32 bytes per voxel: four voxels per iteration with their loads issued together, or a thread has a single
  // 16-byte load in flight and the kernel is latency bound (2.7 TB/s measured with one voxel per iteration)
  constexpr int U = 4;
  const long long tstride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; v0 < V; v0 += tstride * U) {
    uint4 raw[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long v = v0 + k * tstride;
      if (v < V) raw[k] = __ldg(reinterpret_cast<const uint4*>(x + v * x_ld));
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long v = v0 + k * tstride;
      if (v >= V) break;
      const uint32_t uu[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
      float xi[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xi[2 * i] = __uint_as_float(uu[i] << 16);
        xi[2 * i + 1] = __uint_as_float(uu[i] & 0xFFFF0000u);
      }
      uint32_t o[4];
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        float a0 = b[2 * cp], a1 = b[2 * cp + 1];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
          a0 = fmaf(xi[ci], w[2 * cp][ci], a0);
          a1 = fmaf(xi[ci], w[2 * cp + 1][ci], a1);
        }
        const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
        o[cp] = *reinterpret_cast<const uint32_t*>(&h);
        const float r0 = __uint_as_float(o[cp] << 16), r1 = __uint_as_float(o[cp] & 0xFFFF0000u);
        ssum[2 * cp] += r0;
        ssum[2 * cp + 1] += r1;
        ssq[2 * cp] = fmaf(r0, r0, ssq[2 * cp]);
        ssq[2 * cp + 1] = fmaf(r1, r1, ssq[2 * cp + 1]);
      }
      *reinterpret_cast<uint4*>(out + v * out_ld) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a = ssum[c], q = ssq[c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
      }
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_stat[c], static_cast<double>(a));
        atomicAdd(&s_stat[8 + c], static_cast<double>(q));
      }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      const int c = threadIdx.x & 7;
      const double val = s_stat[threadIdx.x];
      if (c < stats_ld && val != 0.0) atomicAdd(stats + (threadIdx.x < 8 ? 0 : stats_ld) + c, val);
    }
  }
}

// 1x1x1 weight gradient with at most 8 channels on both sides (TDisc's 3 -> 2 "spatial" conv): 32 bytes and 64 MACs
// per voxel. One thread per voxel (four per iteration, loads issued together), the 8 x 8 products accumulate in
// registers, one block reduction at the end. The 16-voxel mma.sync groups of thin_wgrad_kernel keep only 8 of 32
// lanes loading on such rows and ran at 1.5 TB/s.
__global__ void __launch_bounds__(256)
tiny_wgrad_kernel(const bf16* __restrict__ dy, long long dy_ld, int cout, const bf16* __restrict__ x, long long x_ld,
                  int cin, float* __restrict__ acc, int co_pad, long long V, float* __restrict__ partial) {
  __shared__ float wred[8][64];   // per-warp partials, summed in warp order
  float p[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) p[i][j] = 0.f;
  constexpr int U = 4;
  const long long tstride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; v0 < V; v0 += tstride * U) {
    uint4 rx[U], ry[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long v = v0 + k * tstride;
      rx[k] = ry[k] = make_uint4(0u, 0u, 0u, 0u);
      if (v < V) {
        rx[k] = __ldg(reinterpret_cast<const uint4*>(x + v * x_ld));
        ry[k] = __ldg(reinterpret_cast<const uint4*>(dy + v * dy_ld));
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const uint32_t xs[4] = {rx[k].x, rx[k].y, rx[k].z, rx[k].w}, ys[4] = {ry[k].x, ry[k].y, ry[k].z, ry[k].w};
      float xf[8], yf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xf[2 * i] = __uint_as_float(xs[i] << 16);
        xf[2 * i + 1] = __uint_as_float(xs[i] & 0xFFFF0000u);
        yf[2 * i] = __uint_as_float(ys[i] << 16);
        yf[2 * i + 1] = __uint_as_float(ys[i] & 0xFFFF0000u);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) p[i][j] = fmaf(xf[i], yf[j], p[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = p[i][j];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5][i * 8 + j] = a;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += wred[w][threadIdx.x];      // warp order: independent of timing
    const int ci = threadIdx.x >> 3, co = threadIdx.x & 7;
    if (partial != nullptr) partial[static_cast<size_t>(blockIdx.x) * 64 + threadIdx.x] = a;
    else if (ci < cin && co < cout) atomicAdd(acc + (size_t)ci * co_pad + co, a);
  }
}

}  // namespace vfd

using namespace vfd;

static FastDivT make_fastdivt(uint32_t d) {
  FastDivT f;
  f.d = d;
  if (d <= 1) {
    f.m = 0;
    f.s = 0;
  } else {
    uint32_t sft = 0;
    while ((1u << sft) < d) ++sft;
    f.m = (uint32_t)((((1ull << sft) - d) << 32) / d + 1);
    f.s = sft;
  }
  return f;
}

// fold: 0 = none; 1 = x is folded on the fly (x holds cs channels per voxel, cin = taps * cs);
//       2 = dy is folded on the fly (dy holds cs channels per voxel, cout = taps * cs)
// partial == nullptr: blocks add into acc with atomics. Otherwise every block writes its 32 x 32 (tiny kernel: 8 x 8)
// result into its own slot of `partial` and *blocks_out / *slot_out tell the caller how many slots of which size.
static int thin_dispatch(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, int cin, float* acc,
                         int co_pad, int ci_pad, int fold, int cs, int N, int D, int H, int W, int kd, int kh, int kw,
                         cudaStream_t stream, float* partial, int* blocks_out, int* slot_out, bool query) {
  const long long V = (long long)N * D * H * W;
  if (blocks_out != nullptr) *blocks_out = 0;
  if (V <= 0) return 0;
  if (V >= (1LL << 31)) return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: more than 2^31 voxels");
  if (cout < 1 || cout > 32 || cin < 1 || cin > 32 || co_pad < cout || ci_pad < cin)
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: needs 1 <= cin, cout <= 32 and a large enough accumulator");
  const int x_need = fold == 1 ? ((cs + 7) & ~7) : ((cin + 7) & ~7), y_need = fold == 2 ? ((cs + 7) & ~7) : ((cout + 7) & ~7);
  if (!query && ((reinterpret_cast<uintptr_t>(dy) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) || (dy_ld % 8) || (x_ld % 8) ||
      dy_ld < y_need || x_ld < x_need))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: tensors must be 16-byte aligned channels-last bf16");
  const int taps = kd * kh * kw;
  if ((fold == 1 && cin != taps * cs) || (fold == 2 && cout != taps * cs) || fold < 0 || fold > 2)
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: folded channel count must be taps * cs");
  FoldGeom fg;
  fg.D = D; fg.H = H; fg.W = W;
  fg.fW = make_fastdivt(W); fg.fH = make_fastdivt(H); fg.fD = make_fastdivt(D);
  if (fold == 0 && cin <= 8 && cout <= 8) {   // 32 bytes per voxel: the one-thread-per-voxel kernel
    long long blocks = (V + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks_out != nullptr) { *blocks_out = (int)blocks; *slot_out = 64; }
    if (query) return 0;
    tiny_wgrad_kernel<<<(int)blocks, 256, 0, stream>>>((const bf16*)dy, dy_ld, cout, (const bf16*)x, x_ld, cin, acc,
                                                       co_pad, V, partial);
    return check_launch("tiny_wgrad");
  }
  const int mt = (cin + 15) / 16, nt = (cout + 7) / 8;
  const int grid = 148 * 4;
  if (blocks_out != nullptr) { *blocks_out = grid; *slot_out = 1024; }
  if (query) return 0;
#define VFD_THIN_ARGS (const bf16*)dy, dy_ld, cout, (const bf16*)x, x_ld, cin, acc, co_pad, V, fg, partial
#define VFD_THIN(MT, NT, FX, FY)                                                                  \
  if (mt == MT && nt == NT) {                                                                     \
    thin_wgrad_kernel<MT, NT, FX, FY><<<grid, kThinThreads, 0, stream>>>(VFD_THIN_ARGS);           \
    return check_launch("thin_wgrad");                                                            \
  }
#define VFD_THIN_NT(MT, FX, FY) VFD_THIN(MT, 1, FX, FY) VFD_THIN(MT, 2, FX, FY) VFD_THIN(MT, 3, FX, FY) VFD_THIN(MT, 4, FX, FY)
  using F133_3 = Fold<1, 3, 3, 3>;
  using F311_2 = Fold<3, 1, 1, 2>;
  using F333_1 = Fold<3, 3, 3, 1>;
  if (fold == 0) {
    VFD_THIN_NT(1, NoFold, NoFold)
    VFD_THIN_NT(2, NoFold, NoFold)
  } else if (fold == 1 && kd == 1 && kh == 3 && kw == 3 && cs == 3) {
    VFD_THIN_NT(2, F133_3, NoFold)
  } else if (fold == 1 && kd == 3 && kh == 1 && kw == 1 && cs == 2) {
    VFD_THIN_NT(1, F311_2, NoFold)
  } else if (fold == 2 && kd == 3 && kh == 3 && kw == 3 && cs == 1) {
    VFD_THIN(1, 4, NoFold, F333_1)
    VFD_THIN(2, 4, NoFold, F333_1)
  }
#undef VFD_THIN_NT
#undef VFD_THIN
#undef VFD_THIN_ARGS
  return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin: unsupported tile shape / fold combination");
}

VFD_API int vfd_conv3d_wgrad_thin(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                                     int cin, float* acc, int co_pad, int ci_pad, int fold, int cs, int N, int D,
                                     int H, int W, int kd, int kh, int kw, void* stream_) {
  return thin_dispatch(dy, dy_ld, cout, x, x_ld, cin, acc, co_pad, ci_pad, fold, cs, N, D, H, W, kd, kh, kw,
                       reinterpret_cast<cudaStream_t>(stream_), nullptr, nullptr, nullptr, false);
}

namespace vfd {
namespace {
// acc[ci][co] += slot 0 + slot 1 + ... in block order; slot = `side` x `side` row-major
__global__ void thin_ordered_reduce_kernel(const float* __restrict__ partial, int blocks, int side, int cin, int cout,
                                           float* __restrict__ acc, int co_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= side * side) return;
  const int ci = i / side, co = i % side;
  if (ci >= cin || co >= cout) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[static_cast<size_t>(b) * side * side + i];
  acc[static_cast<size_t>(ci) * co_pad + co] += s;
}
}  // namespace
}  // namespace vfd

VFD_API long long vfd_conv3d_wgrad_thin_det_workspace(int cout, int cin, int fold, int cs, int N, int D, int H, int W,
                                                      int kd, int kh, int kw) {
  int blocks = 0, slot = 0;
  if (thin_dispatch(nullptr, 0, cout, nullptr, 0, cin, nullptr, 32, 32, fold, cs, N, D, H, W, kd, kh, kw, nullptr,
                    nullptr, &blocks, &slot, true))
    return -1;
  return 4ll * blocks * slot;
}

VFD_API int vfd_conv3d_wgrad_thin_det(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                                      int cin, float* acc, int co_pad, int ci_pad, int fold, int cs, int N, int D,
                                      int H, int W, int kd, int kh, int kw, void* workspace, long long ws_bytes,
                                      void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int blocks = 0, slot = 0;
  if (int e = thin_dispatch(dy, dy_ld, cout, x, x_ld, cin, acc, co_pad, ci_pad, fold, cs, N, D, H, W, kd, kh, kw,
                            stream, nullptr, &blocks, &slot, true))
    return e;
  if (blocks == 0) return 0;
  if (workspace == nullptr || ws_bytes < 4ll * blocks * slot || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_thin_det: workspace too small (vfd_conv3d_wgrad_thin_det_workspace)");
  if (int e = thin_dispatch(dy, dy_ld, cout, x, x_ld, cin, acc, co_pad, ci_pad, fold, cs, N, D, H, W, kd, kh, kw,
                            stream, static_cast<float*>(workspace), &blocks, &slot, false))
    return e;
  const int side = slot == 64 ? 8 : 32;
  thin_ordered_reduce_kernel<<<(side * side + 255) / 256, 256, 0, stream>>>(static_cast<const float*>(workspace), blocks,
                                                                          side, cin, cout, acc, co_pad);
  return check_launch("thin_ordered_reduce");
}

// Internal (called from vfd_conv3d_fwd): 1x1x1, <= 8 input and <= 8 output channels, bf16 output.
namespace vfd {
int launch_tiny_pointwise(const void* x, long long x_ld, const void* w_packed, int cin_k, const float* bias, void* out,
                          long long out_ld, long long V, double* stats, int stats_ld, cudaStream_t stream) {
  long long blocks = (V + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  tiny_pointwise_kernel<<<(int)blocks, 256, 0, stream>>>((const bf16*)x, x_ld, (const bf16*)w_packed, cin_k, bias,
                                                         (bf16*)out, out_ld, V, stats, stats_ld);
  return check_launch("tiny_pointwise");
}
}  // namespace vfd
