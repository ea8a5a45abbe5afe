// Device side of the reference's clip pipeline (lib/data.py:14-161), from decoded uint8 frames onwards:
//   * Resize((isize, isize)) of the test transform (test.py:150-153, lib/data.py:143-146). videotransforms/
//     functional.py:54-58 maps the default 'nearest' to PIL.Image.BILINEAR, so this is Pillow's ImagingResample
//     (src/libImaging/Resample.c) with the bilinear filter on 8-bit channels: antialiased (the support grows with
//     the down-scale factor), two passes -- horizontal into a uint8 intermediate, then vertical -- each a dot
//     product with 22-bit fixed-point coefficients, rounded (+ 1 << 21) and clipped to [0, 255]. Integer work:
//     bit-exact against Pillow.
//   * ClipToTensor + the `*2-1` of MdfDataLoader.__getitem__ (videotransforms/volume_transforms.py:17-58,
//     lib/data.py:78): uint8 (T, H, W, C) frames -> float32 (C, T, H, W), x / 255 in float32 (a correctly rounded
//     division, like torch's), optionally 2x - 1 (exact in float32 after the division's rounding).
// Moving the frames as bytes cuts the host->device traffic of a step 4x; both kernels are HBM-bound byte streams.
#include <cstdint>
#include "vfd_internal.h"

namespace vfd {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c: PRECISION_BITS

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), box = the whole
// axis. Double arithmetic in the reference's order with explicit roundings (no FMA contraction).
__global__ void resample_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds,
                                       int* __restrict__ kk) {
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;        // 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int xx = blockIdx.x * blockDim.x + threadIdx.x; xx < out_size; xx += gridDim.x * blockDim.x) {
    const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);   // in0 = 0
    int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      ww = __dadd_rn(ww, w);
    }
    int* k = kk + static_cast<size_t>(xx) * ksize;
    for (int x = 0; x < ksize; ++x) {
      double w = 0.0;
      if (x < xmax) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        if (ww != 0.0) w = __ddiv_rn(w, ww);
      }
      const double scaled = __dmul_rn(w, static_cast<double>(1 << kPrecisionBits));
      k[x] = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, scaled)) : static_cast<int>(__dadd_rn(0.5, scaled));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// one thread = one output pixel (all C channels); src rows [row0, row0 + rows) of every frame -> tmp [n][rows][Wout][C]
template <int C>
__global__ void __launch_bounds__(256)
resample_horizontal_kernel(const uint8_t* __restrict__ src, int n, int Hin, int Win, int row0, int rows, int Wout,
                           int ksize, const int* __restrict__ bounds, const int* __restrict__ kk,
                           uint8_t* __restrict__ tmp) {
  const long long total = static_cast<long long>(n) * rows * Wout;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += gridDim.x * 256ll) {
    const int xx = static_cast<int>(t % Wout);
    const long long r = t / Wout;
    const int y = static_cast<int>(r % rows);
    const long long f = r / rows;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int* k = kk + static_cast<size_t>(xx) * ksize;
    const uint8_t* line = src + ((f * Hin + row0 + y) * Win + xmin) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < xmax; ++x) {
      const int w = k[x];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(line[x * C + c]) * w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) tmp[t * C + c] = clip8(acc[c]);
  }
}

// tmp [n][rows][Wout][C] (row 0 = source row row0) -> dst [n][Hout][Wout][C]
template <int C>
__global__ void __launch_bounds__(256)
resample_vertical_kernel(const uint8_t* __restrict__ tmp, int n, int rows, int row0, int Wout, int Hout, int ksize,
                         const int* __restrict__ bounds, const int* __restrict__ kk, uint8_t* __restrict__ dst) {
  const long long total = static_cast<long long>(n) * Hout * Wout;
  for (long long t = blockIdx.x * 256ll + threadIdx.x; t < total; t += gridDim.x * 256ll) {
    const int xx = static_cast<int>(t % Wout);
    const long long r = t / Wout;
    const int yy = static_cast<int>(r % Hout);
    const long long f = r / Hout;
    const int ymin = bounds[2 * yy] - row0, ymax = bounds[2 * yy + 1];
    const int* k = kk + static_cast<size_t>(yy) * ksize;
    const uint8_t* col = tmp + ((f * rows + ymin) * Wout + xx) * C;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < ymax; ++y) {
      const int w = k[y];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(col[static_cast<size_t>(y) * Wout * C + c]) * w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dst[t * C + c] = clip8(acc[c]);
  }
}

__global__ void __launch_bounds__(256)
copy_bytes_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long nbytes) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nbytes; i += gridDim.x * 256ll) dst[i] = src[i];
}

// frames uint8 [B][T][H][W][C] -> out float32 [B][Cout][T][H][W]; Cout == C, or C == 1 broadcast to Cout channels
// (ClipToTensor(channel_nb=3) broadcasts an 'L' image over the three channels, volume_transforms.py:46).
// one thread = one (b, t, h, w) pixel: coalesced byte reads, Cout coalesced float writes
__global__ void __launch_bounds__(256)
frames_to_clip_kernel(const uint8_t* __restrict__ frames, long long B, int T, long long HW, int C, int Cout,
                      int pm1, float* __restrict__ out) {
  const long long total = B * T * HW;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const long long hw = i % HW;
    const long long bt = i / HW;
    const int t = static_cast<int>(bt % T);
    const long long b = bt / T;
    for (int c = 0; c < Cout; ++c) {
      const uint8_t u = frames[i * C + (C == 1 ? 0 : c)];
      float v = __fdiv_rn(static_cast<float>(u), 255.0f);
      if (pm1) v = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
      out[((b * Cout + c) * T + t) * HW + hw] = v;
    }
  }
}

struct ResizeLayout {
  int ksize_h, ksize_v;
  long long bounds_h, kk_h, bounds_v, kk_v, tmp, bytes;
};

inline long long align256(long long v) { return (v + 255) & ~255ll; }

inline int ksize_for(int in_size, int out_size) {
  double scale = static_cast<double>(in_size) / out_size;
  if (scale < 1.0) scale = 1.0;
  long long c = static_cast<long long>(scale);
  if (static_cast<double>(c) < scale) ++c;           // ceil(support)
  return static_cast<int>(c) * 2 + 1;
}

ResizeLayout resize_layout(long long n, int Hin, int Win, int C, int Hout, int Wout) {
  ResizeLayout L;
  L.ksize_h = ksize_for(Win, Wout);
  L.ksize_v = ksize_for(Hin, Hout);
  long long off = 0;
  L.bounds_h = off; off = align256(off + 8ll * Wout);
  L.kk_h = off; off = align256(off + 4ll * Wout * L.ksize_h);
  L.bounds_v = off; off = align256(off + 8ll * Hout);
  L.kk_v = off; off = align256(off + 4ll * Hout * L.ksize_v);
  L.tmp = off; off = align256(off + n * Hin * Wout * C);   // upper bound: every source row used
  L.bytes = off;
  return L;
}

inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace vfd

using namespace vfd;

VFD_API long long vfd_resize_frames_u8_workspace(long long n, int Hin, int Win, int C, int Hout, int Wout) {
  if (n < 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || (C != 1 && C != 3)) return -1;
  return resize_layout(n, Hin, Win, C, Hout, Wout).bytes;
}

VFD_API int vfd_resize_frames_u8(const void* src_, long long n, int Hin, int Win, int C, void* dst_, int Hout,
                                 int Wout, void* workspace, long long ws_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n < 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0 || (C != 1 && C != 3))
    return set_error(VFD_ERR_ARG, "resize_frames_u8: bad geometry (C must be 1 or 3)");
  if (n == 0) return VFD_OK;
  if (src_ == nullptr || dst_ == nullptr) return set_error(VFD_ERR_ARG, "resize_frames_u8: null pointer");
  const ResizeLayout L = resize_layout(n, Hin, Win, C, Hout, Wout);
  if (workspace == nullptr || ws_bytes < L.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(VFD_ERR_ARG, "resize_frames_u8: workspace too small or not 256-byte aligned "
                                  "(vfd_resize_frames_u8_workspace)");
  const uint8_t* src = static_cast<const uint8_t*>(src_);
  uint8_t* dst = static_cast<uint8_t*>(dst_);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* bounds_h = reinterpret_cast<int*>(ws + L.bounds_h);
  int* kk_h = reinterpret_cast<int*>(ws + L.kk_h);
  int* bounds_v = reinterpret_cast<int*>(ws + L.bounds_v);
  int* kk_v = reinterpret_cast<int*>(ws + L.kk_v);
  uint8_t* tmp = ws + L.tmp;
  const bool need_h = Wout != Win, need_v = Hout != Hin;   // Resample.c: a pass whose size does not change is skipped
  if (!need_h && !need_v) {
    const long long nbytes = n * Hin * Win * C;
    copy_bytes_kernel<<<grid_for(nbytes), 256, 0, st>>>(src, dst, nbytes);
    return check_launch("resize_frames_u8(copy)");
  }
  if (need_h) resample_coeffs_kernel<<<(Wout + 127) / 128, 128, 0, st>>>(Win, Wout, L.ksize_h, bounds_h, kk_h);
  if (need_v) resample_coeffs_kernel<<<(Hout + 127) / 128, 128, 0, st>>>(Hin, Hout, L.ksize_v, bounds_v, kk_v);
  // Without a box the vertical pass of a whole-axis resize touches every source row (first bound 0, last bound
  // ends at Hin), so the horizontal pass covers rows [0, Hin).
  const uint8_t* vin = src;
  uint8_t* hout = need_v ? tmp : dst;
  if (need_h) {
    const long long total = n * Hin * Wout;
    if (C == 3)
      resample_horizontal_kernel<3><<<grid_for(total), 256, 0, st>>>(src, (int)n, Hin, Win, 0, Hin, Wout, L.ksize_h,
                                                                     bounds_h, kk_h, hout);
    else
      resample_horizontal_kernel<1><<<grid_for(total), 256, 0, st>>>(src, (int)n, Hin, Win, 0, Hin, Wout, L.ksize_h,
                                                                     bounds_h, kk_h, hout);
    vin = hout;
  }
  if (need_v) {
    const long long total = n * Hout * Wout;
    if (C == 3)
      resample_vertical_kernel<3><<<grid_for(total), 256, 0, st>>>(vin, (int)n, Hin, 0, Wout, Hout, L.ksize_v,
                                                                   bounds_v, kk_v, dst);
    else
      resample_vertical_kernel<1><<<grid_for(total), 256, 0, st>>>(vin, (int)n, Hin, 0, Wout, Hout, L.ksize_v,
                                                                   bounds_v, kk_v, dst);
  }
  return check_launch("resize_frames_u8");
}

VFD_API int vfd_frames_to_clip(const void* frames, long long B, int T, int H, int W, int C, int Cout, int pm1,
                               float* out, void* stream_) {
  if (B < 0 || T <= 0 || H <= 0 || W <= 0 || C <= 0 || Cout <= 0 || (C != Cout && C != 1))
    return set_error(VFD_ERR_ARG, "frames_to_clip: bad geometry (C must equal Cout or be 1)");
  if (B == 0) return VFD_OK;
  if (frames == nullptr || out == nullptr) return set_error(VFD_ERR_ARG, "frames_to_clip: null pointer");
  const long long total = B * T * H * W;
  frames_to_clip_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint8_t*>(frames), B, T, static_cast<long long>(H) * W, C, Cout, pm1, out);
  return check_launch("frames_to_clip");
}
