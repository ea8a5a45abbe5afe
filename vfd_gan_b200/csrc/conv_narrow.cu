// Forward of a 3x3x3 conv with 32 input channels and ONE output channel -- NetG's conv_last
// (models/mygannet.py:52,97: nn.Conv3d(ngf, 1, 3, padding=1) before the sigmoid), fp32 logits out.
//
// On the 128-row tcgen05 tile this layer is 27 taps x 2 K16 steps, each re-reading its A tile from shared memory for
// 16 output columns of which one is used: MMA-issue bound at 0.20 of the HBM rate (profiles/r2_thin_epilogue.txt).
// Here the taps become the GEMM's N dimension instead:
//     P[u][tap] = sum_c x[u][c] * w[tap][c]          one m16n8k16 mma.sync chain per 16 INPUT voxels (N = 27 -> 32)
//     out[v]    = bias + sum_tap P[v + off(tap)][tap]  27 shared-memory reads per output voxel
// i.e. every input voxel is multiplied with all 27 filter rows once (8 MMAs per 16 voxels instead of 54 per 16
// outputs), and the 3x3x3 neighbourhood sum runs over the fp32 partial products. A CTA owns an 8 x 16 (h, w) window and
// walks along d: input plane j's partial products go into slot j % 3 of a three-plane ring, so each input plane is
// loaded and multiplied once per window (halo overhead 180 / 128 in h, w only).
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "vfd_internal.h"

// stage-isolation switches (debug library only; tools/gpu_time_conv_last.py): 1 = no global loads, 2 = no MMAs / partial
// product stores, 4 = no neighbourhood sum, 8 = no output store
#ifdef VFD_DEBUG
#define VFD_NDBG(p, bit) ((p).dbg & (bit))
#else
#define VFD_NDBG(p, bit) (0)
#endif

namespace vfd {
namespace {

typedef __nv_bfloat16 bf16;

constexpr int kNarrowThreads = 128;
constexpr int kTH = 8, kTW = 16;                 // output window of one CTA (one d-plane at a time)
constexpr int kPH = kTH + 2, kPW = kTW + 2;      // haloed input window
constexpr int kPV = kPH * kPW;                   // 180 input voxels per plane
constexpr int kPVpad = 192;                      // 12 m16 tiles
constexpr int kXRow = 80;                        // bytes per staged voxel row: 64 B data + 16 B pad (conflict-free ldmatrix)
constexpr int kTaps = 27;
constexpr int kXBytes = kPVpad * kXRow;          // 15360
constexpr int kPSlot = kTaps * kPV;              // floats per ring slot
constexpr int kNarrowSmem = kXBytes + 3 * kPSlot * 4;   // 73680

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// first K step: D = A * B (+ 0), no accumulator registers to clear
__device__ __forceinline__ void mma16816_zero(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                              uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%10, %10, %10, %10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

struct NarrowParams {
  const bf16* x;        // channels-last [N][D][H][W][x_ld], 32 valid channels
  long long x_ld;
  const bf16* w;        // packed forward weights, row 0: [27 taps][cin_k = 32]
  int cin_k;
  const float* bias;    // 1 entry or nullptr
  float* out;           // fp32 [N][D][H][W][out_ld], out_cols columns written (column 0 = logit, the rest 0)
  long long out_ld;
  int out_cols;
  int N, D, H, W, tilesH, tilesW;
  int dbg;
};

constexpr int kChunks = kPV * 4;                                   // 16-byte chunks of one input plane window
constexpr int kChunksPerThread = (kChunks + kNarrowThreads - 1) / kNarrowThreads;   // 6

__global__ void __launch_bounds__(kNarrowThreads)
conv_narrow_fwd_kernel(const NarrowParams p) {
  extern __shared__ __align__(16) uint8_t nsm[];
  uint8_t* xs = nsm;
  float* P = reinterpret_cast<float*>(nsm + kXBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // B fragments (the 27 x 32 filter, zero rows for taps 27..31) stay in registers: [n-tile][k-step][2]
  uint32_t bfr[4][2][2];
  {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int tap = 8 * nt + g;
        uint32_t b0 = 0u, b1 = 0u;
        if (tap < kTaps) {
          const bf16* row = p.w + static_cast<size_t>(tap) * p.cin_k + ks * 16 + 2 * tq;
          b0 = *reinterpret_cast<const uint32_t*>(row);
          b1 = *reinterpret_cast<const uint32_t*>(row + 8);
        }
        bfr[nt][ks][0] = b0;
        bfr[nt][ks][1] = b1;
      }
  }
  // rows 180..191 of the staged window are never loaded: keep them zero
  for (int i = tid; i < (kPVpad - kPV) * kXRow / 4; i += kNarrowThreads)
    reinterpret_cast<uint32_t*>(xs + kPV * kXRow)[i] = 0u;
  const float bias = p.bias != nullptr ? __ldg(p.bias) : 0.f;

  // two input planes in flight per thread (registers): one plane ahead left the memory system idle for most of a step
  // persistent: a CTA takes (n, h-window, w-window) columns round-robin; the filter fragments are loaded once
  const int ncols = p.N * p.tilesH * p.tilesW;
#pragma unroll 1
  for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
  int t = col;
  const int w0 = (t % p.tilesW) * kTW;
  t /= p.tilesW;
  const int h0 = (t % p.tilesH) * kTH;
  const int n = t / p.tilesH;
  uint4 pre_a[kChunksPerThread], pre_b[kChunksPerThread];
  // the kernel is instruction-issue bound, so everything that does not depend on the plane is computed once: element
  // offsets of this thread's 16-byte chunks inside a plane (-1 = outside the image or past the window)
  int coff[kChunksPerThread];
#pragma unroll
  for (int k = 0; k < kChunksPerThread; ++k) {
    const int c = tid + k * kNarrowThreads;
    const int pv = c >> 2, part = c & 3;
    const int hh = h0 - 1 + pv / kPW, ww = w0 - 1 + pv % kPW;
    coff[k] = (c < kChunks && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
                  ? static_cast<int>((static_cast<long long>(hh) * p.W + ww) * p.x_ld + part * 8) : -1;
  }
  const long long plane_elems = static_cast<long long>(p.H) * p.W * p.x_ld;
  auto load_plane = [&](int d, uint4 (&pre)[kChunksPerThread]) {
    const bool plane_ok = d >= 0 && d < p.D && !VFD_NDBG(p, 1);
    const bf16* base = p.x + (static_cast<long long>(n) * p.D + (plane_ok ? d : 0)) * plane_elems;
#pragma unroll
    for (int k = 0; k < kChunksPerThread; ++k) {
      pre[k] = make_uint4(0u, 0u, 0u, 0u);
      if (plane_ok && coff[k] >= 0) pre[k] = __ldg(reinterpret_cast<const uint4*>(base + coff[k]));
    }
  };
  // float offsets of this lane's accumulator columns inside a ring slot (-1: tap 27..31, not stored)
  int poff[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tap = 8 * nt + 2 * (lane & 3) + j;
      poff[nt][j] = tap < kTaps ? tap * kPV : -1;
    }

  const uint32_t xs_s = static_cast<uint32_t>(__cvta_generic_to_shared(xs));
  const int oh = tid / kTW, ow = tid % kTW;
  const bool out_ok = h0 + oh < p.H && w0 + ow < p.W;
  // one step: input plane d_in = s - 1 (planes -1 and D are the zero padding) arrives in `pre`
  auto step = [&](int s, uint4 (&pre)[kChunksPerThread]) {
    const int d_in = s - 1;
#pragma unroll
    for (int k = 0; k < kChunksPerThread; ++k) {
      const int c = tid + k * kNarrowThreads;
      if (c < kChunks) *reinterpret_cast<uint4*>(xs + (c >> 2) * kXRow + (c & 3) * 16) = pre[k];
    }
    __syncthreads();
    if (s + 2 <= p.D + 1) load_plane(d_in + 2, pre);   // this buffer is free again: fetch the plane after next
    float* Ps = P + (s % 3) * kPSlot;
    if (d_in >= 0 && d_in < p.D && !VFD_NDBG(p, 2)) {
      const int g = lane >> 2;
#pragma unroll 1
      for (int mt = warp; mt < kPVpad / 16; mt += kNarrowThreads / 32) {
        float acc[4][4];
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4(xs_s + (mt * 16 + (lane & 15)) * kXRow + (lane >> 4) * 16, a0, a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816_zero(acc[nt], a0, a1, a2, a3, bfr[nt][0][0], bfr[nt][0][1]);
        ldmatrix_x4(xs_s + (mt * 16 + (lane & 15)) * kXRow + 32 + (lane >> 4) * 16, a0, a1, a2, a3);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(acc[nt], a0, a1, a2, a3, bfr[nt][1][0], bfr[nt][1][1]);
        const int v0 = mt * 16 + g;
        const bool ok0 = v0 < kPV, ok1 = v0 + 8 < kPV;
        float* pv0 = Ps + v0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (poff[nt][j] >= 0) {
              if (ok0) pv0[poff[nt][j]] = acc[nt][j];
              if (ok1) pv0[poff[nt][j] + 8] = acc[nt][2 + j];
            }
      }
    } else if (!VFD_NDBG(p, 2)) {
      for (int i = tid; i < kPSlot; i += kNarrowThreads) Ps[i] = 0.f;
    }
    __syncthreads();
    const int d_out = d_in - 1;                  // planes d_out - 1, d_out, d_out + 1 are in the ring now
    if (d_out >= 0 && d_out < p.D && out_ok) {
      float sum[3] = {bias, 0.f, 0.f};           // three independent chains
#pragma unroll
      for (int a = 0; a < (VFD_NDBG(p, 4) ? 0 : 3); ++a) {
        const float* Pa = P + ((d_out + a) % 3) * kPSlot;   // plane d_out - 1 + a sits in slot (d_out + a) % 3
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            sum[a] += Pa[((a * 3 + b) * 3 + c) * kPV + (oh + b) * kPW + (ow + c)];
      }
      const long long vox = ((static_cast<long long>(n) * p.D + d_out) * p.H + (h0 + oh)) * p.W + (w0 + ow);
      float* o = p.out + vox * p.out_ld;
      if (!VFD_NDBG(p, 8) || sum[0] == 12345.f)
      *reinterpret_cast<float4*>(o) = make_float4((sum[0] + sum[1]) + sum[2], 0.f, 0.f, 0.f);
      for (int c4 = 4; c4 < p.out_cols; c4 += 4) *reinterpret_cast<float4*>(o + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_plane(-1, pre_a);
  load_plane(0, pre_b);
  for (int s = 0; s <= p.D + 1; s += 2) {
    step(s, pre_a);
    if (s + 1 <= p.D + 1) step(s + 1, pre_b);
  }
  }
}

// ---- dgrad of the same layer: dx[u][c] = sum_tap g[u + off(tap)] * Wd[c][tap], one gradient channel in, 32 out ------
// Here the taps are the GEMM's K dimension: A[u][tap] = g[u + off(tap)] is an im2col of a SCALAR field, gathered
// straight into mma.sync A fragments from a three-plane shared-memory ring of the gradient (16 two-byte reads per 16
// voxels), B[tap][c] = the dgrad-packed filter (32 x 32, in registers). 8 MMAs per 16 voxels instead of 27 K16 MMAs per
// 128 voxels of which one of 16 K columns is used.
struct NarrowDgradParams {
  const bf16* g;        // channels-last [N][D][H][W][g_ld], column 0 = the gradient of the logit
  long long g_ld;
  const bf16* w;        // dgrad-packed weights [32 rows = ci][27 taps (mirrored)][w_ck], column 0
  int w_ck;
  bf16* dx;             // channels-last [N][D][H][W][dx_ld], 32 channels
  long long dx_ld;
  int N, D, H, W, tilesH, tilesW;
};

__global__ void __launch_bounds__(kNarrowThreads)
conv_narrow_dgrad_kernel(const NarrowDgradParams p) {
  __shared__ __align__(16) unsigned short ring[3][kPVpad];   // gradient planes d-1, d, d+1 of the haloed window (bf16 bits)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const unsigned short* wbits = reinterpret_cast<const unsigned short*>(p.w);
  const unsigned short* gbits = reinterpret_cast<const unsigned short*>(p.g);
  // B fragments: B[k = tap][n = c] = w[c][tap][0]; [n-tile][k-step][2]
  uint32_t bfr[4][2][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int c = 8 * nt + gq;
        uint32_t v = 0u;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int tap = 16 * ks + 8 * h2 + 2 * tq + j;
          if (tap < kTaps) v |= static_cast<uint32_t>(wbits[(static_cast<size_t>(c) * kTaps + tap) * p.w_ck]) << (16 * j);
        }
        bfr[nt][ks][h2] = v;
      }
  // this lane's eight A columns (taps): plane index a (-1 for taps 27..31) and offset inside a haloed plane
  int ta[2][2][2], trel[2][2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int tap = 16 * ks + 8 * h2 + 2 * tq + j;
        ta[ks][h2][j] = tap < kTaps ? tap / 9 : -1;
        trel[ks][h2][j] = ((tap / 3) % 3) * kPW + tap % 3;
      }
  for (int i = tid; i < 3 * kPVpad; i += kNarrowThreads) (&ring[0][0])[i] = 0;
  __syncthreads();

  const int ncols = p.N * p.tilesH * p.tilesW;
#pragma unroll 1
  for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
    int t = col;
    const int w0 = (t % p.tilesW) * kTW;
    t /= p.tilesW;
    const int h0 = (t % p.tilesH) * kTH;
    const int n = t / p.tilesH;
    // this thread's (at most two) scalars of a plane window
    long long soff[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pv = tid + k * kNarrowThreads;
      const int hh = h0 - 1 + pv / kPW, ww = w0 - 1 + pv % kPW;
      soff[k] = (pv < kPV && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
                    ? (static_cast<long long>(hh) * p.W + ww) * p.g_ld : -1;
    }
    const long long plane_elems = static_cast<long long>(p.H) * p.W * p.g_ld;
    auto load_scalars = [&](int d, unsigned short (&v)[2]) {
      const bool ok = d >= 0 && d < p.D;
      const unsigned short* base = gbits + (static_cast<long long>(n) * p.D + (ok ? d : 0)) * plane_elems;
#pragma unroll
      for (int k = 0; k < 2; ++k) v[k] = (ok && soff[k] >= 0) ? __ldg(base + soff[k]) : static_cast<unsigned short>(0);
    };
    auto store_scalars = [&](int slot, const unsigned short (&v)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (tid + k * kNarrowThreads < kPV) ring[slot][tid + k * kNarrowThreads] = v[k];
    };
    unsigned short nxt[2];
    // ring slot of plane j is (j + 1) % 3; planes -1 and D are the zero padding
    load_scalars(-1, nxt);
    store_scalars(0, nxt);
    load_scalars(0, nxt);
    store_scalars(1, nxt);
    load_scalars(1, nxt);
    for (int d = 0; d < p.D; ++d) {
      // plane d + 1 into its slot (that of plane d - 2, which nobody reads any more), then fetch plane d + 2
      store_scalars((d + 2) % 3, nxt);
      __syncthreads();
      load_scalars(d + 2, nxt);
      const unsigned short* s0 = ring[d % 3];          // plane d - 1
      const unsigned short* s1 = ring[(d + 1) % 3];    // plane d
      const unsigned short* s2 = ring[(d + 2) % 3];    // plane d + 1
#pragma unroll 1
      for (int mt = warp; mt < kTH; mt += kNarrowThreads / 32) {     // one m-tile = the 16 voxels of window row mt
        float acc[4][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t a[4];   // a0: rows gq, cols 2tq..; a1: rows gq + 8; a2 / a3: cols + 8
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              uint32_t v = 0u;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int aa = ta[ks][h2][j];
                if (aa >= 0) {
                  const unsigned short* sp = aa == 0 ? s0 : (aa == 1 ? s1 : s2);
                  v |= static_cast<uint32_t>(sp[mt * kPW + gq + 8 * r2 + trel[ks][h2][j]]) << (16 * j);
                }
              }
              a[2 * h2 + r2] = v;
            }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (ks == 0) mma16816_zero(acc[nt], a[0], a[1], a[2], a[3], bfr[nt][0][0], bfr[nt][0][1]);
            else mma16816(acc[nt], a[0], a[1], a[2], a[3], bfr[nt][1][0], bfr[nt][1][1]);
          }
        }
        const int hh = h0 + mt;
        if (hh < p.H) {
          const long long row = ((static_cast<long long>(n) * p.D + d) * p.H + hh) * p.W + w0;
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2) {
            const int ww = gq + 8 * r2;
            if (w0 + ww < p.W) {
              bf16* o = p.dx + (row + ww) * p.dx_ld + 2 * tq;
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<__nv_bfloat162*>(o + 8 * nt) = __floats2bfloat162_rn(acc[nt][2 * r2], acc[nt][2 * r2 + 1]);
            }
          }
        }
      }
      __syncthreads();   // every warp is done with plane d - 1's slot before the next step overwrites it
    }
  }
}

// ---- weight gradient of the same layer: dW[tap][c] = sum_u g[u - off(tap)] * x[u][c] ---------------------------------
// M = taps (32 rows, 27 used), N = the 32 input channels, K = voxels. The A operand (taps x voxels) is gathered from the
// three-plane ring of the scalar gradient like in the dgrad kernel, the B operand (voxels x channels) comes from the
// staged x rows through ldmatrix.trans; the 32 x 32 fp32 accumulator lives in registers for the whole kernel and is
// reduced block-wide in warp order before one atomic per element. x and g are each read once.
struct NarrowWgradParams {
  const bf16* g;        // channels-last [N][D][H][W][g_ld], column 0 = the gradient of the logit
  long long g_ld;
  const bf16* x;        // channels-last [N][D][H][W][x_ld], 32 channels
  long long x_ld;
  float* acc;           // fp32 [32 channels][acc_ld]: acc[c][tap] += dW[tap][c]
  int acc_ld;
  int N, D, H, W, tilesH, tilesW;
};

__global__ void __launch_bounds__(kNarrowThreads)
conv_narrow_wgrad_kernel(const NarrowWgradParams p) {
  __shared__ __align__(16) unsigned short ring[3][kPVpad];
  __shared__ __align__(16) uint8_t xs[kTH * kTW * kXRow];        // 128 voxels x (64 B + 16 B pad)
  __shared__ float red[32 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const unsigned short* gbits = reinterpret_cast<const unsigned short*>(p.g);
  // A rows of this lane: taps gq, gq + 8 (m-tile 0) and 16 + gq, 24 + gq (m-tile 1): plane selector and window offset
  int ta[2][2], trel[2][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const int tap = 16 * m + 8 * r2 + gq;
      const int a = tap / 9, b = (tap / 3) % 3, c = tap % 3;
      ta[m][r2] = tap < kTaps ? a : -1;
      trel[m][r2] = (2 - b) * kPW + (2 - c);        // g[u - off(tap)] relative to x voxel (row, col) of the window
    }
  float acc[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][nt][j] = 0.f;
  for (int i = tid; i < 3 * kPVpad; i += kNarrowThreads) (&ring[0][0])[i] = 0;
  for (int i = tid; i < 32 * 32; i += kNarrowThreads) red[i] = 0.f;
  __syncthreads();
  const uint32_t xs_s = static_cast<uint32_t>(__cvta_generic_to_shared(xs));
  const int mat = lane >> 3, mrow = lane & 7;
  const uint32_t b_off = ((mat & 1) * 8 + mrow) * kXRow + (mat >> 1) * 16;   // + n-tile pair * 32 bytes

  const int ncols = p.N * p.tilesH * p.tilesW;
#pragma unroll 1
  for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
    int t = col;
    const int w0 = (t % p.tilesW) * kTW;
    t /= p.tilesW;
    const int h0 = (t % p.tilesH) * kTH;
    const int n = t / p.tilesH;
    long long soff[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pv = tid + k * kNarrowThreads;
      const int hh = h0 - 1 + pv / kPW, ww = w0 - 1 + pv % kPW;
      soff[k] = (pv < kPV && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
                    ? (static_cast<long long>(hh) * p.W + ww) * p.g_ld : -1;
    }
    // this thread's four 16-byte chunks of the 8 x 16 x-window (no halo): chunk c -> voxel c >> 2, part c & 3
    int xoff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = tid + k * kNarrowThreads;
      const int v = c >> 2, hh = h0 + v / kTW, ww = w0 + v % kTW;
      xoff[k] = (hh < p.H && ww < p.W) ? static_cast<int>((static_cast<long long>(hh) * p.W + ww) * p.x_ld + (c & 3) * 8) : -1;
    }
    const long long gplane = static_cast<long long>(p.H) * p.W * p.g_ld;
    const long long xplane = static_cast<long long>(p.H) * p.W * p.x_ld;
    auto load_scalars = [&](int d, unsigned short (&v)[2]) {
      const bool ok = d >= 0 && d < p.D;
      const unsigned short* base = gbits + (static_cast<long long>(n) * p.D + (ok ? d : 0)) * gplane;
#pragma unroll
      for (int k = 0; k < 2; ++k) v[k] = (ok && soff[k] >= 0) ? __ldg(base + soff[k]) : static_cast<unsigned short>(0);
    };
    auto store_scalars = [&](int slot, const unsigned short (&v)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (tid + k * kNarrowThreads < kPV) ring[slot][tid + k * kNarrowThreads] = v[k];
    };
    auto load_x = [&](int d, uint4 (&v)[4]) {
      const bf16* base = p.x + (static_cast<long long>(n) * p.D + d) * xplane;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        v[k] = (d < p.D && xoff[k] >= 0) ? __ldg(reinterpret_cast<const uint4*>(base + xoff[k])) : make_uint4(0u, 0u, 0u, 0u);
    };
    unsigned short nxt[2];
    uint4 xn[4];
    load_scalars(-1, nxt);
    store_scalars(0, nxt);
    load_scalars(0, nxt);
    store_scalars(1, nxt);
    load_scalars(1, nxt);
    load_x(0, xn);
    for (int d = 0; d < p.D; ++d) {
      store_scalars((d + 2) % 3, nxt);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = tid + k * kNarrowThreads;
        *reinterpret_cast<uint4*>(xs + (c >> 2) * kXRow + (c & 3) * 16) = xn[k];
      }
      __syncthreads();
      load_scalars(d + 2, nxt);
      load_x(d + 1, xn);
      const unsigned short* s0 = ring[d % 3];          // plane d - 1
      const unsigned short* s1 = ring[(d + 1) % 3];    // plane d
      const unsigned short* s2 = ring[(d + 2) % 3];    // plane d + 1
#pragma unroll 1
      for (int mt = warp; mt < kTH; mt += kNarrowThreads / 32) {     // K step = the 16 voxels of window row mt
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {     // a[.][r2 + 2 * h2]: row gq + 8 r2, voxel columns 2tq + 8 h2, + 1
              uint32_t v = 0u;
              const int aa = ta[m][r2];
              if (aa >= 0) {
                const unsigned short* sp = aa == 0 ? s2 : (aa == 1 ? s1 : s0);   // tap plane a reads g plane d - a + 1
                const unsigned short* q = sp + mt * kPW + trel[m][r2] + 2 * tq + 8 * h2;
                v = static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 16);
              }
              a[m][r2 + 2 * h2] = v;
            }
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          uint32_t b0, b1, b2, b3;   // n-tile 2np: (b0, b1); n-tile 2np + 1: (b2, b3)
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(xs_s + mt * kTW * kXRow + b_off + np * 32));
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            mma16816(acc[m][2 * np], a[m][0], a[m][1], a[m][2], a[m][3], b0, b1);
            mma16816(acc[m][2 * np + 1], a[m][0], a[m][1], a[m][2], a[m][3], b2, b3);
          }
        }
      }
      __syncthreads();
    }
  }
  // block reduction in warp order, then one atomic per valid (tap, channel)
  for (int w = 0; w < kNarrowThreads / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int tap = 16 * m + gq, c = 8 * nt + 2 * tq;
          red[tap * 32 + c] += acc[m][nt][0];
          red[tap * 32 + c + 1] += acc[m][nt][1];
          red[(tap + 8) * 32 + c] += acc[m][nt][2];
          red[(tap + 8) * 32 + c + 1] += acc[m][nt][3];
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < 32 * 32; i += kNarrowThreads) {
    const int tap = i >> 5, c = i & 31;
    if (tap < kTaps) atomicAdd(p.acc + static_cast<size_t>(c) * p.acc_ld + tap, red[i]);
  }
}

// ---- weight gradient of the FIRST layers: 1x3x3, three input channels (NetG dconv1 / SDisc dconv1 spatial convs,
// models/mygannet.py:37,130 through models/spatiotempconv.py:49-50) --------------------------------------------------
//   dW[co][c][tap] = sum_v dy[v][co] * x[v + off(tap)][c]        M = (tap, c) = 27 rows, N = cout <= 32, K = voxels
// The tap-folded path writes x as a 27-channel tensor (411 MB at the bench size) and reads it back; here the A operand
// is gathered from a haloed three-channel window in shared memory straight into mma.sync fragments, the B operand is the
// staged dy rows through ldmatrix.trans, and x (16 bytes per voxel) and dy are read once.
struct FirstWgradParams {
  const bf16* x;        // channels-last [N][D][H][W][x_ld], 3 valid channels
  long long x_ld;
  const bf16* dy;       // channels-last [N][D][H][W][dy_ld], cout valid channels
  long long dy_ld;
  int cout;
  float* acc;           // fp32 [32 rows = tap * 3 + c][acc_ld]: acc[r][co] += dW
  int acc_ld;
  int N, D, H, W, tilesH, tilesW;
};

__global__ void __launch_bounds__(kNarrowThreads)
conv_first_wgrad_kernel(const FirstWgradParams p) {
  __shared__ __align__(16) unsigned short xw[3][kPVpad];           // haloed window, one plane per input channel
  __shared__ __align__(16) uint8_t ys[kTH * kTW * kXRow];          // 128 dy rows x (64 B + 16 B pad)
  __shared__ float red[32 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int ychunks = (p.cout + 7) >> 3;                            // 16-byte chunks of a dy row
  // A rows of this lane: r = 16 m + 8 r2 + gq = tap * 3 + c (27 used)
  int rc[2][2], rrel[2][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      const int r = 16 * m + 8 * r2 + gq;
      const int tap = r / 3;
      rc[m][r2] = r < 27 ? r % 3 : -1;
      rrel[m][r2] = (tap / 3) * kPW + tap % 3;
    }
  float acc[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[m][nt][j] = 0.f;
  for (int i = tid; i < 3 * kPVpad; i += kNarrowThreads) (&xw[0][0])[i] = 0;
  for (int i = tid; i < kTH * kTW * kXRow / 4; i += kNarrowThreads) reinterpret_cast<uint32_t*>(ys)[i] = 0u;
  for (int i = tid; i < 32 * 32; i += kNarrowThreads) red[i] = 0.f;
  __syncthreads();
  const uint32_t ys_s = static_cast<uint32_t>(__cvta_generic_to_shared(ys));
  const int mat = lane >> 3, mrow = lane & 7;
  const uint32_t b_off = ((mat & 1) * 8 + mrow) * kXRow + (mat >> 1) * 16;   // + n-tile pair * 32 bytes

  const long long ncols = static_cast<long long>(p.N) * p.D * p.tilesH * p.tilesW;   // one (n, d, window) per step
  // this thread's two voxels of the haloed x window and its (at most) four 16-byte chunks of the dy window
  uint2 xn[2];
  uint4 yn[4];
  auto load = [&](long long col, uint2 (&xv)[2], uint4 (&yv)[4]) {
    long long t = col;
    const int w0 = static_cast<int>(t % p.tilesW) * kTW;
    t /= p.tilesW;
    const int h0 = static_cast<int>(t % p.tilesH) * kTH;
    t /= p.tilesH;                                                   // t = n * D + d
    const bf16* xb = p.x + t * p.H * p.W * p.x_ld;
    const bf16* yb = p.dy + t * p.H * p.W * p.dy_ld;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pv = tid + k * kNarrowThreads;
      const int hh = h0 - 1 + pv / kPW, ww = w0 - 1 + pv % kPW;
      xv[k] = make_uint2(0u, 0u);
      if (col < ncols && pv < kPV && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
        xv[k] = __ldg(reinterpret_cast<const uint2*>(xb + (static_cast<long long>(hh) * p.W + ww) * p.x_ld));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = tid + k * kNarrowThreads;
      const int v = c >> 2, part = c & 3;
      const int hh = h0 + v / kTW, ww = w0 + v % kTW;
      yv[k] = make_uint4(0u, 0u, 0u, 0u);
      if (col < ncols && part < ychunks && hh < p.H && ww < p.W)
        yv[k] = __ldg(reinterpret_cast<const uint4*>(yb + (static_cast<long long>(hh) * p.W + ww) * p.dy_ld + part * 8));
    }
  };
  load(blockIdx.x, xn, yn);
#pragma unroll 1
  for (long long col = blockIdx.x; col < ncols; col += gridDim.x) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pv = tid + k * kNarrowThreads;
      if (pv < kPV) {
        xw[0][pv] = static_cast<unsigned short>(xn[k].x & 0xFFFFu);
        xw[1][pv] = static_cast<unsigned short>(xn[k].x >> 16);
        xw[2][pv] = static_cast<unsigned short>(xn[k].y & 0xFFFFu);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = tid + k * kNarrowThreads;
      if ((c & 3) < ychunks) *reinterpret_cast<uint4*>(ys + (c >> 2) * kXRow + (c & 3) * 16) = yn[k];
    }
    __syncthreads();
    load(col + gridDim.x, xn, yn);                                   // the next window while this one is multiplied
#pragma unroll 1
    for (int mt = warp; mt < kTH; mt += kNarrowThreads / 32) {       // K step = the 16 voxels of window row mt
      uint32_t a[2][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {     // a[.][r2 + 2 h2]: row gq + 8 r2, voxel columns 2tq + 8 h2, + 1
            uint32_t v = 0u;
            const int cc = rc[m][r2];
            if (cc >= 0) {
              const unsigned short* q = (cc == 0 ? xw[0] : (cc == 1 ? xw[1] : xw[2])) + mt * kPW + rrel[m][r2] + 2 * tq + 8 * h2;
              v = static_cast<uint32_t>(q[0]) | (static_cast<uint32_t>(q[1]) << 16);
            }
            a[m][r2 + 2 * h2] = v;
          }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        if (np * 16 >= p.cout) break;
        uint32_t b0, b1, b2, b3;   // n-tile 2np: (b0, b1); n-tile 2np + 1: (b2, b3)
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(ys_s + mt * kTW * kXRow + b_off + np * 32));
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          mma16816(acc[m][2 * np], a[m][0], a[m][1], a[m][2], a[m][3], b0, b1);
          mma16816(acc[m][2 * np + 1], a[m][0], a[m][1], a[m][2], a[m][3], b2, b3);
        }
      }
    }
    __syncthreads();
  }
  for (int w = 0; w < kNarrowThreads / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int r = 16 * m + gq, c = 8 * nt + 2 * tq;
          red[r * 32 + c] += acc[m][nt][0];
          red[r * 32 + c + 1] += acc[m][nt][1];
          red[(r + 8) * 32 + c] += acc[m][nt][2];
          red[(r + 8) * 32 + c + 1] += acc[m][nt][3];
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < 32 * 32; i += kNarrowThreads) {
    const int r = i >> 5, c = i & 31;
    if (r < 27 && c < p.cout) atomicAdd(p.acc + static_cast<size_t>(r) * p.acc_ld + c, red[i]);
  }
}

}  // namespace

}  // namespace vfd

using namespace vfd;

VFD_API int vfd_conv3d_fwd_narrow(const void* x, long long x_ld, int cin, const void* w_packed, int cin_k,
                                  const float* bias, float* out, long long out_ld, int out_cols, int N, int D, int H,
                                  int W, void* stream_) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  if (x == nullptr || w_packed == nullptr || out == nullptr) return set_error(VFD_ERR_ARG, "conv3d_fwd_narrow: null pointer");
  if (cin != 32 || cin_k != 32)
    return set_error(VFD_ERR_ARG, "conv3d_fwd_narrow: serves 32 input channels (packed K = 32) only");
  if (out_cols < 4 || out_cols % 4 || out_ld < out_cols || out_ld % 4 || (reinterpret_cast<uintptr_t>(out) & 15) ||
      x_ld < 32 || x_ld % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w_packed) & 3))
    return set_error(VFD_ERR_ARG, "conv3d_fwd_narrow: tensors must be 16-byte aligned channels-last");
  if (static_cast<long long>(N) * D * H * W >= (1LL << 31))
    return set_error(VFD_ERR_ARG, "conv3d_fwd_narrow: more than 2^31 voxels");
  NarrowParams p;
  p.x = static_cast<const bf16*>(x); p.x_ld = x_ld; p.w = static_cast<const bf16*>(w_packed); p.cin_k = cin_k;
  p.bias = bias; p.out = out; p.out_ld = out_ld; p.out_cols = out_cols;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.tilesH = (H + kTH - 1) / kTH; p.tilesW = (W + kTW - 1) / kTW;
  p.dbg = 0;
#ifdef VFD_DEBUG
  if (const char* e = getenv("VFD_NARROW_DBG")) p.dbg = atoi(e);
#endif
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_narrow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kNarrowSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_narrow_fwd)");
    // three CTAs per SM need 3 x 73 KB of shared memory: ask for the largest carveout instead of the driver's default
    e = cudaFuncSetAttribute(conv_narrow_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_narrow_fwd, carveout)");
    attr = true;
#ifdef VFD_DEBUG
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_narrow_fwd_kernel, kNarrowThreads, kNarrowSmem);
    fprintf(stderr, "conv_narrow_fwd: %d resident CTAs per SM\n", occ);
#endif
  }
  long long grid = static_cast<long long>(N) * p.tilesH * p.tilesW;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (grid > 3LL * sms) grid = 3LL * sms;       // three resident CTAs per SM, each walks several columns
  conv_narrow_fwd_kernel<<<static_cast<unsigned>(grid), kNarrowThreads, kNarrowSmem, static_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("conv_narrow_fwd");
}

VFD_API int vfd_conv3d_dgrad_narrow(const void* g, long long g_ld, const void* w_dgrad_packed, int w_rows, int w_ck,
                                    void* dx, long long dx_ld, int N, int D, int H, int W, void* stream_) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  if (g == nullptr || w_dgrad_packed == nullptr || dx == nullptr)
    return set_error(VFD_ERR_ARG, "conv3d_dgrad_narrow: null pointer");
  if (w_rows != 32 || w_ck < 1) return set_error(VFD_ERR_ARG, "conv3d_dgrad_narrow: serves 32 input channels only");
  if (g_ld < 1 || dx_ld < 32 || dx_ld % 2 || (reinterpret_cast<uintptr_t>(dx) & 3) || (reinterpret_cast<uintptr_t>(g) & 1))
    return set_error(VFD_ERR_ARG, "conv3d_dgrad_narrow: bad tensor layout");
  if (static_cast<long long>(N) * D * H * W >= (1LL << 31))
    return set_error(VFD_ERR_ARG, "conv3d_dgrad_narrow: more than 2^31 voxels");
  NarrowDgradParams p;
  p.g = static_cast<const bf16*>(g); p.g_ld = g_ld; p.w = static_cast<const bf16*>(w_dgrad_packed); p.w_ck = w_ck;
  p.dx = static_cast<bf16*>(dx); p.dx_ld = dx_ld;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.tilesH = (H + kTH - 1) / kTH; p.tilesW = (W + kTW - 1) / kTW;
  long long grid = static_cast<long long>(N) * p.tilesH * p.tilesW;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (grid > 8LL * sms) grid = 8LL * sms;       // eight 128-thread CTAs per SM, each walks several columns
  conv_narrow_dgrad_kernel<<<static_cast<unsigned>(grid), kNarrowThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("conv_narrow_dgrad");
}

VFD_API int vfd_conv3d_wgrad_narrow(const void* g, long long g_ld, const void* x, long long x_ld, float* acc, int acc_ld,
                                    int N, int D, int H, int W, void* stream_) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  if (g == nullptr || x == nullptr || acc == nullptr) return set_error(VFD_ERR_ARG, "conv3d_wgrad_narrow: null pointer");
  if (g_ld < 1 || x_ld < 32 || x_ld % 8 || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(g) & 1) ||
      acc_ld < 27)
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_narrow: bad tensor layout");
  if (static_cast<long long>(N) * D * H * W >= (1LL << 31))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_narrow: more than 2^31 voxels");
  NarrowWgradParams p;
  p.g = static_cast<const bf16*>(g); p.g_ld = g_ld; p.x = static_cast<const bf16*>(x); p.x_ld = x_ld;
  p.acc = acc; p.acc_ld = acc_ld;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.tilesH = (H + kTH - 1) / kTH; p.tilesW = (W + kTW - 1) / kTW;
  long long grid = static_cast<long long>(N) * p.tilesH * p.tilesW;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (grid > 4LL * sms) grid = 4LL * sms;
  conv_narrow_wgrad_kernel<<<static_cast<unsigned>(grid), kNarrowThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("conv_narrow_wgrad");
}

VFD_API int vfd_conv3d_wgrad_first(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, float* acc,
                                   int acc_ld, int N, int D, int H, int W, void* stream_) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  if (dy == nullptr || x == nullptr || acc == nullptr) return set_error(VFD_ERR_ARG, "conv3d_wgrad_first: null pointer");
  if (cout < 1 || cout > 32 || acc_ld < cout) return set_error(VFD_ERR_ARG, "conv3d_wgrad_first: needs 1 <= cout <= 32");
  if (x_ld < 8 || x_ld % 4 || (reinterpret_cast<uintptr_t>(x) & 7) || dy_ld % 8 || dy_ld < ((cout + 7) & ~7) ||
      (reinterpret_cast<uintptr_t>(dy) & 15))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_first: tensors must be aligned channels-last bf16");
  if (static_cast<long long>(N) * D * H * W >= (1LL << 31))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_first: more than 2^31 voxels");
  FirstWgradParams p;
  p.x = static_cast<const bf16*>(x); p.x_ld = x_ld; p.dy = static_cast<const bf16*>(dy); p.dy_ld = dy_ld; p.cout = cout;
  p.acc = acc; p.acc_ld = acc_ld;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.tilesH = (H + kTH - 1) / kTH; p.tilesW = (W + kTW - 1) / kTW;
  long long grid = static_cast<long long>(N) * D * p.tilesH * p.tilesW;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (grid > 4LL * sms) grid = 4LL * sms;
  conv_first_wgrad_kernel<<<static_cast<unsigned>(grid), kNarrowThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  return check_launch("conv_first_wgrad");
}
