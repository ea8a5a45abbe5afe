// video_to_flow (lib/utils.py:94-129) on the device: the reference moves the generator's fresh output to the
// host twice per training step, runs cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0) on
// every frame pair on one core and encodes the field as an RGB video (cartToPolar -> HSV with S = 255 ->
// cvtColor on float32 -> np.uint8 wrap). This file restates that pipeline as data-parallel kernels over all
// (clip, frame) images / (clip, frame pair) fields at once; the algorithm and its arithmetic order follow
// oracle/flow_oracle.py (OpenCV's optflowgf.cpp, getGaussianKernel, resize, fastAtan2, normalize, HSV2RGB).
//
// Arithmetic mirrors the CPU code where it matters for byte parity of the encoded video: float steps use
// __fmul_rn / __fadd_rn (no FMA contraction, OpenCV's generic paths are not contracted), the horizontal pass of
// the polynomial expansion and the box sums accumulate in double like the reference.
//
// All kernels are tiny HBM/L2 streams (B*D images of <= 128 x 128 pixels); one thread per pixel.
#include <cstdint>
#include <cmath>
#include "vfd_internal.h"

namespace vfd {
namespace {

constexpr int kPolyN = 5;          // poly_n
constexpr int kWin = 15;           // winsize
constexpr int kIters = 3;          // iterations
constexpr int kMaxLevels = 3;

struct PolyConst {
  float g[kPolyN + 1], xg[kPolyN + 1], xxg[kPolyN + 1];   // index k = 0..n (symmetric / antisymmetric halves)
  double ig11, ig03, ig33, ig55;
};
struct BlurKernel {
  int ksize;
  float k[9];
};

__device__ __forceinline__ float mulf(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subf(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// warp results -> one atomic pair per block
__device__ __forceinline__ void block_minmax(float lo, float hi, float* dst) {
  __shared__ float s_lo[32], s_hi[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) {
    s_lo[wid] = lo;
    s_hi[wid] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i) {
      lo = fminf(lo, s_lo[i]);
      hi = fmaxf(hi, s_hi[i]);
    }
    atomic_min_f(dst, lo);
    atomic_max_f(dst + 1, hi);
  }
}

// mm[2k] = +inf, mm[2k+1] = -inf
__global__ void fill_minmax_kernel(float* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = (i & 1) ? -INFINITY : INFINITY;
}

// normalize() of lib/utils.py:81-89 works per frame index over the whole batch: min / max over (B, C, H, W).
// grid = (blocks, D)
__global__ void __launch_bounds__(256)
frame_minmax_kernel(const float* __restrict__ video, int B, int D, int HW, float* __restrict__ mm) {
  const int d = blockIdx.y;
  const long long per = (long long)B * 3 * HW;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const long long bc = i / HW;
    const int p = (int)(i - bc * HW);
    const float v = video[(bc * D + d) * HW + p];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  block_minmax(lo, hi, mm + 2 * d);
}

// grey[b][d][p] = cv2.cvtColor(RGB2GRAY) of (x - min_d) / (max_d - min_d + 1e-5)
__global__ void __launch_bounds__(256)
gray_kernel(const float* __restrict__ video, int B, int D, int HW, const float* __restrict__ mm,
            float* __restrict__ grey) {
  const long long total = (long long)B * D * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long bd = i / HW;
    const int p = (int)(i - bd * HW);
    const int b = (int)(bd / D), d = (int)(bd - (long long)b * D);
    const float mn = mm[2 * d], mx = mm[2 * d + 1];
    const float den = (float)((double)mx - (double)mn + 1e-5);
    const float nmn = -mn;
    const float* src = video + ((long long)b * 3 * D + d) * HW + p;
    const float r = __fdiv_rn(addf(src[0], nmn), den);
    const float g = __fdiv_rn(addf(src[(long long)D * HW], nmn), den);
    const float bl = __fdiv_rn(addf(src[2LL * D * HW], nmn), den);
    grey[i] = addf(addf(mulf(r, 0.299f), mulf(g, 0.587f)), mulf(bl, 0.114f));
  }
}

// One pyramid level of one image: GaussianBlur(ksize, reflect-101) at full resolution, then cv::resize
// (INTER_LINEAR) to (w, h). Evaluated directly per output pixel in the oracle's order (rows, then columns).
__device__ __forceinline__ float blurred_at(const float* __restrict__ img, int H, int W, int y, int x, const BlurKernel& bk) {
  const int r = bk.ksize >> 1;
  float out = 0.f;
  for (int jy = 0; jy < bk.ksize; ++jy) {
    const float* row = img + (long long)reflect101(y + jy - r, H) * W;
    float t = 0.f;
    for (int jx = 0; jx < bk.ksize; ++jx) t = addf(t, mulf(bk.k[jx], row[reflect101(x + jx - r, W)]));
    out = addf(out, mulf(bk.k[jy], t));
  }
  return out;
}
__device__ __forceinline__ void lin_coord(int dst, int n_dst, int n_src, int& i0, int& i1, float& a) {
  const double f = ((double)dst + 0.5) * ((double)n_src / (double)n_dst) - 0.5;
  const int fl = (int)floor(f);
  a = (float)(f - (double)fl);
  if (fl < 0 || fl >= n_src - 1) a = 0.f;
  i0 = min(max(fl, 0), n_src - 1);
  i1 = min(max(fl + 1, 0), n_src - 1);
}
__global__ void __launch_bounds__(256)
level_image_kernel(const float* __restrict__ grey, int frames, int H, int W, int h, int w, BlurKernel bk,
                   float* __restrict__ out) {
  const long long total = (long long)frames * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const long long f = i / ((long long)w * h);
    const float* img = grey + f * (long long)H * W;
    if (h == H && w == W) {
      out[i] = blurred_at(img, H, W, y, x, bk);
      continue;
    }
    int x0, x1, y0, y1;
    float ax, ay;
    lin_coord(x, w, W, x0, x1, ax);
    lin_coord(y, h, H, y0, y1, ay);
    const float v00 = blurred_at(img, H, W, y0, x0, bk), v01 = blurred_at(img, H, W, y0, x1, bk);
    const float v10 = blurred_at(img, H, W, y1, x0, bk), v11 = blurred_at(img, H, W, y1, x1, bk);
    const float top = addf(mulf(v00, subf(1.f, ax)), mulf(v01, ax));
    const float bot = addf(mulf(v10, subf(1.f, ax)), mulf(v11, ax));
    out[i] = addf(mulf(top, subf(1.f, ay)), mulf(bot, ay));
  }
}

// FarnebackPolyExp, vertical pass (float, rows clamped): rows[f][y][x][3]
__global__ void __launch_bounds__(256)
polyexp_v_kernel(const float* __restrict__ I, int frames, int h, int w, PolyConst pc, float* __restrict__ rows) {
  const long long total = (long long)frames * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const float* img = I + (i / ((long long)w * h)) * (long long)h * w;
    float r0 = mulf(img[(long long)y * w + x], pc.g[0]), r1 = 0.f, r2 = 0.f;
#pragma unroll
    for (int k = 1; k <= kPolyN; ++k) {
      const float s0 = img[(long long)max(y - k, 0) * w + x], s1 = img[(long long)min(y + k, h - 1) * w + x];
      const float p = addf(s0, s1);
      r0 = addf(r0, mulf(pc.g[k], p));
      r1 = addf(r1, mulf(pc.xg[k], subf(s1, s0)));
      r2 = addf(r2, mulf(pc.xxg[k], p));
    }
    rows[i * 3] = r0;
    rows[i * 3 + 1] = r1;
    rows[i * 3 + 2] = r2;
  }
}
// horizontal pass (double accumulators, columns clamped): R[f][y][x][5]
__global__ void __launch_bounds__(256)
polyexp_h_kernel(const float* __restrict__ rows, int frames, int h, int w, PolyConst pc, float* __restrict__ R) {
  const long long total = (long long)frames * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const float* row = rows + (i - x) * 3;     // start of this image row
    double b1 = (double)mulf(row[x * 3], pc.g[0]), b2 = 0.0, b3 = (double)mulf(row[x * 3 + 1], pc.g[0]), b4 = 0.0,
           b5 = (double)mulf(row[x * 3 + 2], pc.g[0]), b6 = 0.0;
#pragma unroll
    for (int k = 1; k <= kPolyN; ++k) {
      const int xp = min(x + k, w - 1) * 3, xm = max(x - k, 0) * 3;
      const double tg = (double)addf(row[xp], row[xm]);
      b1 += tg * (double)pc.g[k];
      b4 += tg * (double)pc.xxg[k];
      b2 += (double)mulf(subf(row[xp], row[xm]), pc.xg[k]);
      b3 += (double)mulf(addf(row[xp + 1], row[xm + 1]), pc.g[k]);
      b6 += (double)mulf(subf(row[xp + 1], row[xm + 1]), pc.xg[k]);
      b5 += (double)mulf(addf(row[xp + 2], row[xm + 2]), pc.g[k]);
    }
    float* o = R + i * 5;
    o[0] = (float)(b3 * pc.ig11);
    o[1] = (float)(b2 * pc.ig11);
    o[2] = (float)(b1 * pc.ig03 + b5 * pc.ig33);
    o[3] = (float)(b1 * pc.ig03 + b4 * pc.ig33);
    o[4] = (float)(b6 * pc.ig55);
  }
}

// pair p = b * (D - 1) + j uses frames b * D + j (prev) and b * D + j + 1 (next)
__device__ __forceinline__ long long pair_frame(long long pair, int D) {
  const long long b = pair / (D - 1);
  return b * D + (pair - b * (D - 1));
}

// FarnebackUpdateMatrices
__global__ void __launch_bounds__(256)
update_matrices_kernel(const float* __restrict__ R, const float* __restrict__ flow, int pairs, int D, int h, int w,
                       float* __restrict__ M) {
  const long long total = (long long)pairs * h * w;
  const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const long long pair = i / ((long long)w * h);
    const long long f0 = pair_frame(pair, D);
    const float* R0 = R + ((f0 * h + y) * w + x) * 5;
    const float* R1 = R + (f0 + 1) * (long long)h * w * 5;
    const float dx = flow ? flow[i * 2] : 0.f, dy = flow ? flow[i * 2 + 1] : 0.f;
    float fx = addf((float)x, dx), fy = addf((float)y, dy);
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx = subf(fx, (float)x1);
    fy = subf(fy, (float)y1);
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
      const float a00 = mulf(subf(1.f, fx), subf(1.f, fy)), a01 = mulf(fx, subf(1.f, fy));
      const float a10 = mulf(subf(1.f, fx), fy), a11 = mulf(fx, fy);
      const float* p = R1 + ((long long)y1 * w + x1) * 5;
      const float* q = p + (long long)w * 5;
      r2 = addf(addf(addf(mulf(a00, p[0]), mulf(a01, p[5])), mulf(a10, q[0])), mulf(a11, q[5]));
      r3 = addf(addf(addf(mulf(a00, p[1]), mulf(a01, p[6])), mulf(a10, q[1])), mulf(a11, q[6]));
      r4 = addf(addf(addf(mulf(a00, p[2]), mulf(a01, p[7])), mulf(a10, q[2])), mulf(a11, q[7]));
      r5 = addf(addf(addf(mulf(a00, p[3]), mulf(a01, p[8])), mulf(a10, q[3])), mulf(a11, q[8]));
      r6 = addf(addf(addf(mulf(a00, p[4]), mulf(a01, p[9])), mulf(a10, q[4])), mulf(a11, q[9]));
      r4 = mulf(addf(R0[2], r4), 0.5f);
      r5 = mulf(addf(R0[3], r5), 0.5f);
      r6 = mulf(addf(R0[4], r6), 0.25f);
    } else {
      r2 = r3 = 0.f;
      r4 = R0[2];
      r5 = R0[3];
      r6 = mulf(R0[4], 0.5f);
    }
    r2 = mulf(subf(R0[0], r2), 0.5f);
    r3 = mulf(subf(R0[1], r3), 0.5f);
    r2 = addf(addf(r2, mulf(r4, dy)), mulf(r6, dx));
    r3 = addf(addf(r3, mulf(r6, dy)), mulf(r5, dx));
    float sx = 1.f, sy = 1.f;
    if (x < 5) sx = mulf(sx, border[x]);
    if (x >= w - 5) sx = mulf(sx, border[w - x - 1]);
    if (y < 5) sy = mulf(sy, border[y]);
    if (y >= h - 5) sy = mulf(sy, border[h - y - 1]);
    const float scale = mulf(sx, sy);
    r2 = mulf(r2, scale);
    r3 = mulf(r3, scale);
    r4 = mulf(r4, scale);
    r5 = mulf(r5, scale);
    r6 = mulf(r6, scale);
    float* o = M + i * 5;
    o[0] = addf(mulf(r4, r4), mulf(r6, r6));
    o[1] = mulf(addf(r4, r5), r6);
    o[2] = addf(mulf(r5, r5), mulf(r6, r6));
    o[3] = addf(mulf(r4, r2), mulf(r6, r3));
    o[4] = addf(mulf(r6, r2), mulf(r5, r3));
  }
}

// FarnebackUpdateFlow_Blur: 15 x 15 box sums of the five matrix channels with replicated borders, accumulated in
// double like the reference, and the 2 x 2 solve with the 1e-3 regulariser. One block = one 16 x 16 tile of one
// field: the (16+14)^2 x 5 halo is staged in shared memory (clamped loads are the replicated border), column sums
// go to a second shared array in double, row sums and the solve finish per pixel.
constexpr int kBoxT = 16;
constexpr int kBoxS = kBoxT + kWin - 1;   // 30
__global__ void __launch_bounds__(kBoxT * kBoxT)
box_blur_solve_kernel(const float* __restrict__ M, int h, int w, int tiles_x, int tiles_y, float* __restrict__ flow) {
  __shared__ float s_in[kBoxS * kBoxS * 5];
  __shared__ double s_v[kBoxT * kBoxS * 5];
  const int tiles = tiles_x * tiles_y;
  const long long pair = blockIdx.x / tiles;
  const int t = blockIdx.x - (int)(pair * tiles);
  const int y0 = (t / tiles_x) * kBoxT, x0 = (t % tiles_x) * kBoxT;
  const float* img = M + pair * (long long)h * w * 5;
  const int tid = threadIdx.x;
  constexpr int m = kWin / 2;
  for (int i = tid; i < kBoxS * kBoxS * 5; i += kBoxT * kBoxT) {
    const int c = i % 5, col = (i / 5) % kBoxS, row = i / (5 * kBoxS);
    const int gy = min(max(y0 + row - m, 0), h - 1), gx = min(max(x0 + col - m, 0), w - 1);
    s_in[i] = img[((long long)gy * w + gx) * 5 + c];
  }
  __syncthreads();
  // column sums: one thread per (column, channel) slides the 15-row window down the tile (running sums in
  // double, like the reference's vsum)
  for (int cc = tid; cc < kBoxS * 5; cc += kBoxT * kBoxT) {
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < kWin; ++k) sum += (double)s_in[k * kBoxS * 5 + cc];
    s_v[cc] = sum;
#pragma unroll
    for (int r = 1; r < kBoxT; ++r) {
      sum += (double)s_in[(r + kWin - 1) * kBoxS * 5 + cc] - (double)s_in[(r - 1) * kBoxS * 5 + cc];
      s_v[r * kBoxS * 5 + cc] = sum;
    }
  }
  __syncthreads();
  const int py = tid / kBoxT, px = tid % kBoxT;
  const int y = y0 + py, x = x0 + px;
  if (y >= h || x >= w) return;
  double sacc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < kWin; ++k) {
    const double* p = s_v + (py * kBoxS + px + k) * 5;
#pragma unroll
    for (int c = 0; c < 5; ++c) sacc[c] += p[c];
  }
  const double scale = 1.0 / (double)(kWin * kWin);
  const double g11 = sacc[0] * scale, g12 = sacc[1] * scale, g22 = sacc[2] * scale, h1 = sacc[3] * scale, h2 = sacc[4] * scale;
  const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
  float* o = flow + ((pair * h + y) * (long long)w + x) * 2;
  o[0] = (float)((g11 * h2 - g12 * h1) * idet);
  o[1] = (float)((g22 * h1 - g12 * h2) * idet);
}

// resize(prevFlow, (w, h), INTER_LINEAR) * (1 / pyr_scale)
__global__ void __launch_bounds__(256)
flow_upsample_kernel(const float* __restrict__ src, int pairs, int hs, int ws, int h, int w, float gain,
                     float* __restrict__ dst) {
  const long long total = (long long)pairs * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)((i / w) % h);
    const float* img = src + (i / ((long long)w * h)) * (long long)hs * ws * 2;
    int x0, x1, y0, y1;
    float ax, ay;
    lin_coord(x, w, ws, x0, x1, ax);
    lin_coord(y, h, hs, y0, y1, ay);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float v00 = img[((long long)y0 * ws + x0) * 2 + c], v01 = img[((long long)y0 * ws + x1) * 2 + c];
      const float v10 = img[((long long)y1 * ws + x0) * 2 + c], v11 = img[((long long)y1 * ws + x1) * 2 + c];
      const float top = addf(mulf(v00, subf(1.f, ax)), mulf(v01, ax));
      const float bot = addf(mulf(v10, subf(1.f, ax)), mulf(v11, ax));
      dst[i * 2 + c] = mulf(addf(mulf(top, subf(1.f, ay)), mulf(bot, ay)), gain);
    }
  }
}

// cv::fastAtan2 in degrees
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = (float)(0.9997878412794807 * 57.29577951308232), p3 = (float)(-0.3258083974640975 * 57.29577951308232),
              p5 = (float)(0.1555786518463281 * 57.29577951308232), p7 = (float)(-0.04432655554792128 * 57.29577951308232);
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = 2.220446049250313e-16f;
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, addf(ax, eps)), c2 = mulf(c, c);
    a = mulf(addf(mulf(addf(mulf(addf(mulf(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    const float c = __fdiv_rn(ax, addf(ay, eps)), c2 = mulf(c, c);
    a = subf(90.f, mulf(addf(mulf(addf(mulf(addf(mulf(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = subf(180.f, a);
  if (y < 0.f) a = subf(360.f, a);
  return a;
}

// cartToPolar: magnitude + per-pair min / max (cv2.normalize NORM_MINMAX works per image). grid = (blocks, pairs)
__global__ void __launch_bounds__(256)
mag_minmax_kernel(const float* __restrict__ flow, int hw, float* __restrict__ mm) {
  const long long base = (long long)blockIdx.y * hw;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    const float fx = flow[(base + i) * 2], fy = flow[(base + i) * 2 + 1];
    const float mag = __fsqrt_rn(addf(mulf(fx, fx), mulf(fy, fy)));
    lo = fminf(lo, mag);
    hi = fmaxf(hi, mag);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  block_minmax(lo, hi, mm + 2 * blockIdx.y);
}

// HSV (H = ang / 2, S = 255, V = minmax-normalised magnitude) -> cv::cvtColor(HSV2RGB) on float -> np.uint8 wrap ->
// / 255 * 2 - 1, written into frame j of out [B][3][D][H][W]; frame D-1 repeats frame D-2 (lib/utils.py:125).
__global__ void __launch_bounds__(256)
encode_kernel(const float* __restrict__ flow, const float* __restrict__ mm, int pairs, int D, int hw,
              float* __restrict__ out) {
  const long long total = (long long)pairs * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pair = i / hw;
    const int p = (int)(i - pair * hw);
    const long long b = pair / (D - 1);
    const int j = (int)(pair - b * (D - 1));
    const float fx = flow[i * 2], fy = flow[i * 2 + 1];
    const float mag = __fsqrt_rn(addf(mulf(fx, fx), mulf(fy, fy)));
    const float ang = fast_atan2_deg(fy, fx);
    const float mn = mm[2 * pair], mx = mm[2 * pair + 1];
    const float nscale = mx > mn ? __fdiv_rn(255.f, subf(mx, mn)) : 0.f;
    const float v = mulf(subf(mag, mn), nscale);
    const float s = 255.f;
    float hh = mulf(mulf(ang, 0.5f), (float)(6.0 / 360.0));
    if (hh < 0.f) hh = addf(hh, mulf(6.f, ceilf(__fdiv_rn(-hh, 6.f))));
    if (hh >= 6.f) hh = subf(hh, mulf(6.f, floorf(__fdiv_rn(hh, 6.f))));
    int sector = (int)floorf(hh);
    float f = subf(hh, (float)sector);
    if ((unsigned)sector >= 6u) {
      sector = 0;
      f = 0.f;
    }
    float tab[4];
    tab[0] = v;
    tab[1] = mulf(v, subf(1.f, s));
    tab[2] = mulf(v, subf(1.f, mulf(s, f)));
    tab[3] = mulf(v, subf(1.f, mulf(s, subf(1.f, f))));
    const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    const float rgb[3] = {tab[sd[sector][2]], tab[sd[sector][1]], tab[sd[sector][0]]};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int byte = __float2int_rz(rgb[c]) & 0xFF;
      const float val = subf(mulf(__fdiv_rn((float)byte, 255.f), 2.f), 1.f);
      float* dst = out + ((b * 3 + c) * D + j) * hw + p;
      *dst = val;
      if (j == D - 2) dst[hw] = val;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
void make_poly_const(PolyConst& pc) {
  // FarnebackPrepareGaussian(n = 5, sigma = 1.2)
  const int n = kPolyN;
  const double sigma = 1.2;
  float g[2 * kPolyN + 1];
  double s = 0.0;
  for (int x = -n; x <= n; ++x) {
    g[x + n] = (float)std::exp(-x * x / (2 * sigma * sigma));
    s += g[x + n];
  }
  s = 1.0 / s;
  float xg[2 * kPolyN + 1], xxg[2 * kPolyN + 1];
  for (int x = -n; x <= n; ++x) {
    g[x + n] = (float)(g[x + n] * s);
    xg[x + n] = (float)(x * g[x + n]);
    xxg[x + n] = (float)(x * x * g[x + n]);
  }
  double G[6][6] = {};
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      const double gg = (double)g[y + n] * (double)g[x + n];
      G[0][0] += gg;
      G[1][1] += gg * x * x;
      G[3][3] += gg * x * x * x * x;
      G[5][5] += gg * x * x * y * y;
    }
  G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
  G[4][4] = G[3][3];
  G[3][4] = G[4][3] = G[5][5];
  // Gauss-Jordan inverse of the 6 x 6 symmetric positive-definite matrix
  double A[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 12; ++j) A[i][j] = j < 6 ? G[i][j] : (j - 6 == i ? 1.0 : 0.0);
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r)
      if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
    for (int j = 0; j < 12; ++j) std::swap(A[c][j], A[piv][j]);
    const double d = 1.0 / A[c][c];
    for (int j = 0; j < 12; ++j) A[c][j] *= d;
    for (int r = 0; r < 6; ++r)
      if (r != c) {
        const double f = A[r][c];
        for (int j = 0; j < 12; ++j) A[r][j] -= f * A[c][j];
      }
  }
  pc.ig11 = A[1][7];
  pc.ig03 = A[0][9];
  pc.ig33 = A[3][9];
  pc.ig55 = A[5][11];
  for (int k = 0; k <= n; ++k) {
    pc.g[k] = g[n + k];
    pc.xg[k] = xg[n + k];
    pc.xxg[k] = xxg[n + k];
  }
}

void make_blur_kernel(BlurKernel& bk, int ksize, double sigma) {
  // cv::getGaussianKernel(ksize, sigma, CV_32F)
  bk.ksize = ksize;
  if (sigma <= 0 && ksize == 3) {
    bk.k[0] = 0.25f; bk.k[1] = 0.5f; bk.k[2] = 0.25f;
    return;
  }
  const double s = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
  double t[9], sum = 0.0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    t[i] = std::exp(-(x * x) / (2 * s * s));
    sum += t[i];
  }
  for (int i = 0; i < ksize; ++i) bk.k[i] = (float)(t[i] / sum);
}

struct LevelGeom {
  int h, w, ksize;
  double sigma;
};
int plan_levels(int H, int W, LevelGeom* lv) {
  int k = 0;
  double scale = 1.0;
  while (k < kMaxLevels) {
    scale *= 0.5;
    if (W * scale < 32 || H * scale < 32) break;
    ++k;
  }
  for (int i = 0; i <= k; ++i) {
    const double sc = std::pow(0.5, i), sigma = (1.0 / sc - 1) * 0.5;
    int sm = (int)std::nearbyint(sigma * 5) | 1;
    if (sm < 3) sm = 3;
    lv[i].h = (int)std::nearbyint(H * sc);
    lv[i].w = (int)std::nearbyint(W * sc);
    lv[i].ksize = sm;
    lv[i].sigma = sigma;
  }
  return k;
}

inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}
inline long long align256(long long x) { return (x + 255) & ~255LL; }

}  // namespace
}  // namespace vfd

using namespace vfd;

// Workspace layout (bytes, each region 256-byte aligned):
//   frame min/max [2*D] f32 | pair min/max [2*pairs] f32 | grey [B*D*H*W] f32 | level image [frames*H*W] f32 |
//   rows [frames*H*W*3] f32 | R per level [frames*h*w*5] f32 | M [pairs*H*W*5] f32 |
//   flow A, flow B [pairs*H*W*2] f32
VFD_API long long vfd_video_to_flow_workspace(int B, int D, int H, int W) {
  if (B <= 0 || D < 2 || H <= 0 || W <= 0) return 0;
  const long long frames = (long long)B * D, pairs = (long long)B * (D - 1), hw = (long long)H * W;
  LevelGeom lv[kMaxLevels + 1];
  const int levels = plan_levels(H, W, lv);
  long long total = align256(2LL * D * 4) + align256(2 * pairs * 4) + align256(frames * hw * 4) * 2 +
                    align256(frames * hw * 3 * 4);
  for (int k = 0; k <= levels; ++k) total += align256(frames * lv[k].h * lv[k].w * 5 * 4);
  total += align256(pairs * hw * 5 * 4) + 2 * align256(pairs * hw * 2 * 4);
  return total;
}

VFD_API int vfd_video_to_flow(const float* video, int B, int D, int H, int W, float* out, float* raw_flow,
                              void* workspace, long long ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B < 0 || D < 2 || H < 12 || W < 12) return set_error(VFD_ERR_ARG, "video_to_flow: needs D >= 2 and frames of at least 12 x 12");
  if (B == 0) return VFD_OK;   // empty batch: nothing to do
  if (video == nullptr || out == nullptr || workspace == nullptr) return set_error(VFD_ERR_ARG, "video_to_flow: null pointer");
  if (ws_bytes < vfd_video_to_flow_workspace(B, D, H, W)) return set_error(VFD_ERR_ARG, "video_to_flow: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return set_error(VFD_ERR_ARG, "video_to_flow: workspace must be 256-byte aligned");
  const long long frames = (long long)B * D, pairs = (long long)B * (D - 1), hw = (long long)H * W;
  if (pairs > 65535) return set_error(VFD_ERR_ARG, "video_to_flow: at most 65535 frame pairs per call");
  LevelGeom lv[kMaxLevels + 1];
  const int levels = plan_levels(H, W, lv);
  char* p = static_cast<char*>(workspace);
  auto take = [&](long long bytes) { char* r = p; p += align256(bytes); return r; };
  float* fmm = reinterpret_cast<float*>(take(2LL * D * 4));
  float* pmm = reinterpret_cast<float*>(take(2 * pairs * 4));
  float* grey = reinterpret_cast<float*>(take(frames * hw * 4));
  float* limg = reinterpret_cast<float*>(take(frames * hw * 4));
  float* rows = reinterpret_cast<float*>(take(frames * hw * 3 * 4));
  float* R[kMaxLevels + 1];
  for (int k = 0; k <= levels; ++k) R[k] = reinterpret_cast<float*>(take(frames * lv[k].h * lv[k].w * 5 * 4));
  float* M = reinterpret_cast<float*>(take(pairs * hw * 5 * 4));
  float* flowA = reinterpret_cast<float*>(take(pairs * hw * 2 * 4));
  float* flowB = reinterpret_cast<float*>(take(pairs * hw * 2 * 4));

  PolyConst pc;
  make_poly_const(pc);

  fill_minmax_kernel<<<(2 * D + 2 * (int)pairs + 255) / 256, 256, 0, stream>>>(fmm, 2 * D);
  fill_minmax_kernel<<<(2 * (int)pairs + 255) / 256, 256, 0, stream>>>(pmm, 2 * (int)pairs);
  {
    int bx = (int)((((long long)B * 3 * hw) + 256 * 8 - 1) / (256 * 8));
    if (bx < 1) bx = 1;
    if (bx > 148 * 4) bx = 148 * 4;
    if (D > 65535) return set_error(VFD_ERR_ARG, "video_to_flow: too many frames per clip");
    frame_minmax_kernel<<<dim3(bx, D), 256, 0, stream>>>(video, B, D, (int)hw, fmm);
  }
  gray_kernel<<<grid_for(frames * hw), 256, 0, stream>>>(video, B, D, (int)hw, fmm, grey);
  for (int k = 0; k <= levels; ++k) {
    BlurKernel bk;
    make_blur_kernel(bk, lv[k].ksize, lv[k].sigma);
    const long long n = frames * lv[k].h * lv[k].w;
    level_image_kernel<<<grid_for(n), 256, 0, stream>>>(grey, (int)frames, H, W, lv[k].h, lv[k].w, bk, limg);
    polyexp_v_kernel<<<grid_for(n), 256, 0, stream>>>(limg, (int)frames, lv[k].h, lv[k].w, pc, rows);
    polyexp_h_kernel<<<grid_for(n), 256, 0, stream>>>(rows, (int)frames, lv[k].h, lv[k].w, pc, R[k]);
  }
  if (int e = check_launch("video_to_flow: polynomial expansion")) return e;

  float* cur = flowA;
  float* other = flowB;
  for (int k = levels; k >= 0; --k) {
    const int h = lv[k].h, w = lv[k].w;
    const long long n = pairs * h * w;
    const int tiles_x = (w + kBoxT - 1) / kBoxT, tiles_y = (h + kBoxT - 1) / kBoxT;
    const float* init = nullptr;
    if (k < levels) {
      flow_upsample_kernel<<<grid_for(n), 256, 0, stream>>>(cur, (int)pairs, lv[k + 1].h, lv[k + 1].w, h, w, 2.0f, other);
      std::swap(cur, other);
      init = cur;
    }
    update_matrices_kernel<<<grid_for(n), 256, 0, stream>>>(R[k], init, (int)pairs, D, h, w, M);
    for (int it = 0; it < kIters; ++it) {
      box_blur_solve_kernel<<<(unsigned)(pairs * tiles_x * tiles_y), kBoxT * kBoxT, 0, stream>>>(M, h, w, tiles_x, tiles_y, cur);
      if (it < kIters - 1) update_matrices_kernel<<<grid_for(n), 256, 0, stream>>>(R[k], cur, (int)pairs, D, h, w, M);
    }
  }
  if (int e = check_launch("video_to_flow: flow iterations")) return e;
  if (raw_flow != nullptr) {
    cudaError_t e = cudaMemcpyAsync(raw_flow, cur, pairs * hw * 2 * 4, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return set_cuda_error(e, "video_to_flow: raw flow copy");
  }
  {
    int bx = (int)((hw + 255) / 256);
    if (bx > 64) bx = 64;
    mag_minmax_kernel<<<dim3(bx, (int)pairs), 256, 0, stream>>>(cur, (int)hw, pmm);
  }
  encode_kernel<<<grid_for(pairs * hw), 256, 0, stream>>>(cur, pmm, (int)pairs, D, (int)hw, out);
  return check_launch("video_to_flow: encode");
}
