// Scoring-side kernels of the hot path: the per-clip anomaly score and the latent (L2) / contextual (L1)
// reductions of the enc-dec-enc composition (definitions: models/ganomaly.py:372,396,437-440,475-480, the
// only place the reference states them), and the device versions of the host detours MyGAN.test takes on
// every batch: threshold + 5x5 morphological opening (lib/utils.py:139-152) and the confusion counts / ROC
// area that lib/evaluate.py:14-91 derives with sklearn.
//
// All HBM-bound: 128-bit loads, warp-shuffle + one atomic per block reductions, fp64 accumulators.
#include <cuda_bf16.h>
#include <cstdint>
#include "vfd_internal.h"

namespace vfd {

typedef __nv_bfloat16 bf16;

namespace {

__device__ __forceinline__ void ld8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void st8(bf16* p, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block; valid in thread 0
__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum(v);
  return v;
}

// ---------------------------------------------------------------------------------- anomaly score
// per_clip[n] += sum over the clip's voxels and channels of (a - b)^2, channels-last bf16 latents.
// models/ganomaly.py:372 takes mean(pow(latent_i - latent_o, 2), dim=1) of a (B, nz, 1, 1) latent, i.e. the
// mean over everything but the batch; the 3-D latent is (B, 512, D/16, H/16, W/16).
// grid = (blocks per clip, clips)
__global__ void __launch_bounds__(256)
latent_score_kernel(const bf16* __restrict__ a, long long a_ld, const bf16* __restrict__ b, long long b_ld,
                    int C, long long rows, double* __restrict__ per_clip) {
  const int CG = C / 8;
  const long long total = rows * CG;
  const long long base = (long long)blockIdx.y * rows;
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const long long r = base + i / CG;
    float x[8], y[8];
    ld8(a + r * a_ld + cg * 8, x);
    ld8(b + r * b_ld + cg * 8, y);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = x[j] - y[j];
      s = fmaf(d, d, s);
    }
    local += (double)s;
  }
  local = block_sum_d(local);
  if (threadIdx.x == 0) atomicAdd(per_clip + blockIdx.y, local);
}

// gradient of scale * sum((a - b)^2): ga = 2 * scale * (a - b), gb = -ga (either may be null)
__global__ void __launch_bounds__(256)
sqdiff_bwd_kernel(const bf16* __restrict__ a, long long a_ld, const bf16* __restrict__ b, long long b_ld, int C,
                  long long V, const float* __restrict__ gscale, float scale, bf16* __restrict__ ga,
                  long long ga_ld, bf16* __restrict__ gb, long long gb_ld) {
  const int CG = C / 8;
  const long long total = V * CG;
  const float s2 = 2.f * scale * (gscale != nullptr ? __ldg(gscale) : 1.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const long long r = i / CG;
    float x[8], y[8], g[8];
    ld8(a + r * a_ld + cg * 8, x);
    ld8(b + r * b_ld + cg * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = s2 * (x[j] - y[j]);
    if (ga != nullptr) st8(ga + r * ga_ld + cg * 8, g);
    if (gb != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = -g[j];
      st8(gb + r * gb_ld + cg * 8, g);
    }
  }
}

// nn.L1Loss numerator (models/ganomaly.py:438,476) on fp32 tensors; optional gradient
// d/da = grad_scale * sign(a - b) (sign(0) = 0, like torch)
__global__ void __launch_bounds__(256)
l1_kernel(const float* __restrict__ a, const float* __restrict__ b, long long V, float grad_scale,
          double* __restrict__ sum, float* __restrict__ ga) {
  double local = 0.0;
  const long long V4 = V / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    local += (double)(fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3));
    if (ga != nullptr) {
      float4 g;
      g.x = d0 > 0.f ? grad_scale : (d0 < 0.f ? -grad_scale : 0.f);
      g.y = d1 > 0.f ? grad_scale : (d1 < 0.f ? -grad_scale : 0.f);
      g.z = d2 > 0.f ? grad_scale : (d2 < 0.f ? -grad_scale : 0.f);
      g.w = d3 > 0.f ? grad_scale : (d3 < 0.f ? -grad_scale : 0.f);
      reinterpret_cast<float4*>(ga)[i] = g;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = V4 * 4 + threadIdx.x; i < V; i += blockDim.x) {
      const float d = a[i] - b[i];
      local += (double)fabsf(d);
      if (ga != nullptr) ga[i] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);
    }
  }
  local = block_sum_d(local);
  if (threadIdx.x == 0) atomicAdd(sum, local);
}

// nn.BCELoss (models/mygannet.py:267, lib/train_stcnn.py:90) numerator: -(t * max(log p, -100) +
// (1 - t) * max(log(1 - p), -100)); optional gradient grad_scale * (p - t) / max(p * (1 - p), 1e-12), the
// formula torch's binary_cross_entropy_backward uses.
__global__ void __launch_bounds__(256)
bce_kernel(const float* __restrict__ p, const float* __restrict__ t, long long V, float grad_scale,
           double* __restrict__ sum, float* __restrict__ gp) {
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i], ti = t[i];
    const float lp = fmaxf(logf(pi), -100.f), lq = fmaxf(logf(1.f - pi), -100.f);
    local -= (double)(ti * lp + (1.f - ti) * lq);
    if (gp != nullptr) gp[i] = grad_scale * (pi - ti) / fmaxf(pi * (1.f - pi), 1e-12f);
  }
  local = block_sum_d(local);
  if (threadIdx.x == 0) atomicAdd(sum, local);
}

// per_clip (fp64 sums) -> fp32 means, plus the running min / max of the sweep (mm[0] = min, mm[1] = max,
// initialised by the caller to +inf / -inf). One block.
__global__ void score_finalize_kernel(const double* __restrict__ per_clip, int n, double inv_count,
                                      float* __restrict__ scores, float* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = (float)(per_clip[i] * inv_count);
    scores[i] = s;
    lo = fminf(lo, s);
    hi = fmaxf(hi, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[32], shi[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    slo[wid] = lo;
    shi[wid] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0 && mm != nullptr) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      lo = fminf(lo, slo[w]);
      hi = fmaxf(hi, shi[w]);
    }
    mm[0] = fminf(mm[0], lo);
    mm[1] = fmaxf(mm[1], hi);
  }
}

// models/ganomaly.py:396: (s - min) / (max - min) over the whole sweep
__global__ void score_scale_kernel(const float* __restrict__ s, long long n, const float* __restrict__ mm,
                                   float* __restrict__ out) {
  const float lo = mm[0], span = mm[1] - mm[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (s[i] - lo) / span;
}

// ---------------------------------------------------------------------------------- threshold + opening
// threshold (lib/utils.py:149-152): t = (p > thr) as float. morphology_proc (lib/utils.py:139-147) hands each
// clip's (D, H, W) array to cv2.morphologyEx(MORPH_OPEN, ones(5,5)): OpenCV reads a 3-D array as an image with
// D rows, H columns and W channels, so the 5x5 opening runs in the (D, H) plane independently for every w,
// with the default border (pixels outside the image never win the min / max). The rectangle is separable:
// min over h then over d, max over h then over d. One block = one clip x TW consecutive w; the plane lives
// in shared memory as bytes, w fastest (coalesced global access, conflict-free shared access).
template <bool IS_MAX>
__device__ __forceinline__ void pass_h(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int D, int H,
                                       int TW) {
  const int total = D * H * TW;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = i % TW, h = (i / TW) % H, d = i / (TW * H);
    const int h0 = max(h - 2, 0), h1 = min(h + 2, H - 1);
    uint8_t v = src[(d * H + h0) * TW + w];
    for (int hh = h0 + 1; hh <= h1; ++hh) {
      const uint8_t u = src[(d * H + hh) * TW + w];
      v = IS_MAX ? max(v, u) : min(v, u);
    }
    dst[i] = v;
  }
}
template <bool IS_MAX>
__device__ __forceinline__ void pass_d(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int D, int H,
                                       int TW) {
  const int total = D * H * TW;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = i % TW, h = (i / TW) % H, d = i / (TW * H);
    const int d0 = max(d - 2, 0), d1 = min(d + 2, D - 1);
    uint8_t v = src[(d0 * H + h) * TW + w];
    for (int dd = d0 + 1; dd <= d1; ++dd) {
      const uint8_t u = src[(dd * H + h) * TW + w];
      v = IS_MAX ? max(v, u) : min(v, u);
    }
    dst[i] = v;
  }
}

__global__ void __launch_bounds__(512)
threshold_open_kernel(const float* __restrict__ p, int D, int H, int W, int TW, float thr,
                      float* __restrict__ t_out, float* __restrict__ m_out) {
  extern __shared__ uint8_t plane[];
  const int total = D * H * TW;
  uint8_t* s0 = plane;
  uint8_t* s1 = plane + total;
  const int w0 = blockIdx.x * TW;
  const long long base = (long long)blockIdx.y * D * H * W;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = w0 + i % TW;
    const long long row = i / TW;   // d * H + h
    uint8_t t = 0;
    if (w < W) {
      t = p[base + row * W + w] > thr ? 1 : 0;
      if (t_out != nullptr) t_out[base + row * W + w] = (float)t;
    }
    s0[i] = t;
  }
  __syncthreads();
  pass_h<false>(s0, s1, D, H, TW);
  __syncthreads();
  pass_d<false>(s1, s0, D, H, TW);
  __syncthreads();
  pass_h<true>(s0, s1, D, H, TW);
  __syncthreads();
  pass_d<true>(s1, s0, D, H, TW);
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = w0 + i % TW;
    const long long row = i / TW;
    if (w < W) m_out[base + row * W + w] = (float)s0[i];
  }
}

// ---------------------------------------------------------------------------------- evaluation counts
// counts[0..3] += TP, FP, FN, TN with prediction = (score >= thr) (lib/evaluate.py:22-25 binarises at 0.20 in
// place before f1_score) and label = (label > 0.5) (models/mygannet.py:444 casts the {0,1} mask to int32).
__global__ void __launch_bounds__(256)
confusion_kernel(const float* __restrict__ labels, const float* __restrict__ scores, long long n, float thr,
                 unsigned long long* __restrict__ counts) {
  unsigned int tp = 0, fp = 0, fn = 0, tn = 0;   // a thread sees < 2^32 elements
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 l = reinterpret_cast<const float4*>(labels)[i], s = reinterpret_cast<const float4*>(scores)[i];
    const float lv[4] = {l.x, l.y, l.z, l.w}, sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool pos = lv[j] > 0.5f, pred = sv[j] >= thr;
      tp += pos && pred;
      fp += !pos && pred;
      fn += pos && !pred;
      tn += !pos && !pred;
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const bool pos = labels[i] > 0.5f, pred = scores[i] >= thr;
      tp += pos && pred;
      fp += !pos && pred;
      fn += pos && !pred;
      tn += !pos && !pred;
    }
  }
  unsigned int v[4] = {tp, fp, fn, tn};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[k] = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0 && v[k]) atomicAdd(counts + k, (unsigned long long)v[k]);
  }
}

// Exact ROC area of up to kAucMax (score, label) pairs in one block: bitonic sort of the scores in shared
// memory, an inclusive scan of the negatives, and the tie-aware Mann-Whitney count
//   AUC = sum over positives of (#negatives below + 0.5 * #negatives tied) / (P * N),
// which is what sklearn's auc(roc_curve(labels, scores)) (lib/evaluate.py:37-38) integrates.
// out[0] = AUC (NaN when a class is empty), out[1] = P, out[2] = N, out[3] = area under the precision-recall curve
// as lib/evaluate.py:67-68 computes it (auc(recall, precision) of precision_recall_curve).
constexpr int kAucMax = 16384;
constexpr int kAucThreads = 1024;

__device__ __forceinline__ uint32_t float_order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kAucThreads)
auc_kernel(const float* __restrict__ scores, const float* __restrict__ labels, int n, int npad,
           double* __restrict__ out) {
  extern __shared__ uint8_t auc_smem[];
  uint32_t* key = reinterpret_cast<uint32_t*>(auc_smem);            // npad sortable keys
  uint32_t* val = key + npad;                                        // npad: label, later the negative prefix count
  const int tid = threadIdx.x;
  for (int i = tid; i < npad; i += kAucThreads) {
    if (i < n) {
      float s = scores[i];
      s += 0.f;                // -0 and +0 are one score (-0 + 0 = +0)
      key[i] = float_order_key(s);
      val[i] = labels[i] > 0.5f ? 1u : 0u;
    } else {
      key[i] = 0xFFFFFFFFu;    // padding sorts last (a NaN score would tie with it; scores are finite)
      val[i] = 2u;
    }
  }
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += kAucThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const uint32_t a = key[i], b = key[ixj];
          if ((a > b) == up) {
            key[i] = b;
            key[ixj] = a;
            const uint32_t t = val[i];
            val[i] = val[ixj];
            val[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  // inclusive prefix count of negatives over the sorted order; every thread owns a contiguous span
  const int per = npad / kAucThreads > 0 ? npad / kAucThreads : 1;
  const int begin = tid * per;
  uint32_t negs = 0, poss = 0;
  uint32_t lab[kAucMax / kAucThreads];
#pragma unroll
  for (int e = 0; e < kAucMax / kAucThreads; ++e) {
    lab[e] = 2u;
    if (e < per && begin + e < npad) {
      lab[e] = val[begin + e];
      negs += lab[e] == 0u;
      poss += lab[e] == 1u;
    }
  }
  __shared__ uint32_t wtot[32];
  __shared__ uint32_t ptot[32];
  uint32_t incl = negs, pw = poss;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  pw = warp_sum(pw);
  if ((tid & 31) == 31) wtot[tid >> 5] = incl;
  if ((tid & 31) == 0) ptot[tid >> 5] = pw;
  __syncthreads();
  uint32_t woff = 0, P = 0, N = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < (tid >> 5)) woff += wtot[w];
    N += wtot[w];
    P += ptot[w];
  }
  __syncthreads();
  uint32_t run = woff + incl - negs;   // negatives before this thread's span
#pragma unroll
  for (int e = 0; e < kAucMax / kAucThreads; ++e) {
    if (e < per && begin + e < npad) {
      run += lab[e] == 0u;
      val[begin + e] = (run << 2) | lab[e];   // inclusive negative count, label kept in the low bits
    }
  }
  __syncthreads();
  double local = 0.0, local_pr = 0.0;
  for (int i = tid; i < n; i += kAucThreads) {
    const uint32_t kk = key[i];
    const bool group_start = i == 0 || key[i - 1] != kk;
    const bool is_pos = (val[i] & 3u) == 1u;
    if (!is_pos && !group_start) continue;
    int lo = i, hi = n - 1;              // last index whose key == kk
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (key[mid] > kk) hi = mid - 1; else lo = mid;
    }
    const int last = lo;
    if (is_pos) {
      lo = 0;
      hi = i;                            // first index whose key == kk
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (key[mid] < kk) lo = mid + 1; else hi = mid;
      }
      const int first = lo;
      const uint32_t below = first > 0 ? (val[first - 1] >> 2) : 0u;
      const uint32_t upto = val[last] >> 2;
      local += (double)below + 0.5 * (double)(upto - below);
    }
    if (group_start && P > 0) {
      // precision-recall point of the threshold "score >= this value" and of the next higher distinct value
      // (or the end point (recall 0, precision 1)): one trapezoid of sklearn's auc(recall, precision)
      const uint32_t nb = i > 0 ? (val[i - 1] >> 2) : 0u;          // negatives below the threshold
      const double tp = (double)P - (double)((uint32_t)i - nb), fp = (double)N - (double)nb;
      const double r0 = tp / (double)P, p0 = tp / (tp + fp);
      double r1 = 0.0, p1 = 1.0;
      const int j = last + 1;
      if (j < n) {
        const uint32_t nbj = val[j - 1] >> 2;
        const double tpj = (double)P - (double)((uint32_t)j - nbj), fpj = (double)N - (double)nbj;
        r1 = tpj / (double)P;
        p1 = tpj / (tpj + fpj);
      }
      local_pr += (r0 - r1) * (p0 + p1) * 0.5;
    }
  }
  local = block_sum_d(local);
  __syncthreads();
  local_pr = block_sum_d(local_pr);
  if (tid == 0) {
    out[0] = (P > 0 && N > 0) ? local / ((double)P * (double)N) : nan("");
    out[1] = (double)P;
    out[2] = (double)N;
    out[3] = P > 0 ? local_pr : nan("");
  }
}

}  // namespace
}  // namespace vfd

using namespace vfd;
#define STREAM static_cast<cudaStream_t>(stream_)

static inline int blocks_for(long long total, int block, int max_blocks) {
  long long b = (total + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}
static inline int bad_cl(const void* p, long long ld, int C) {
  return p == nullptr || (C % 8) != 0 || ld < C || (ld % 8) != 0 || (reinterpret_cast<uintptr_t>(p) % 16) != 0;
}

VFD_API int vfd_latent_score(const void* a, long long a_ld, const void* b, long long b_ld, int C,
                             long long rows_per_clip, int N, double* per_clip, void* stream_) {
  if (rows_per_clip < 0 || N < 0) return set_error(VFD_ERR_ARG, "latent_score: bad arguments");
  if (N == 0 || rows_per_clip == 0) return VFD_OK;   // empty batch: nothing to do
  if (bad_cl(a, a_ld, C) || bad_cl(b, b_ld, C)) return set_error(VFD_ERR_ARG, "latent_score: bad latent tensor");
  if (per_clip == nullptr) return set_error(VFD_ERR_ARG, "latent_score: bad arguments");
  if (N > 65535) return set_error(VFD_ERR_ARG, "latent_score: at most 65535 clips per call");
  // whole waves: about 148 * 8 blocks in total
  int per = blocks_for(rows_per_clip * (C / 8), 256, (148 * 8 + N - 1) / N);
  latent_score_kernel<<<dim3(per, N), 256, 0, STREAM>>>((const bf16*)a, a_ld, (const bf16*)b, b_ld, C, rows_per_clip,
                                                       per_clip);
  return check_launch("latent_score");
}

VFD_API int vfd_sqdiff_bwd(const void* a, long long a_ld, const void* b, long long b_ld, int C, long long V,
                           const float* gscale, float scale, void* ga, long long ga_ld, void* gb, long long gb_ld,
                           void* stream_) {
  if (V <= 0) return VFD_OK;
  if (bad_cl(a, a_ld, C) || bad_cl(b, b_ld, C)) return set_error(VFD_ERR_ARG, "sqdiff_bwd: bad input tensor");
  if ((ga != nullptr && bad_cl(ga, ga_ld, C)) || (gb != nullptr && bad_cl(gb, gb_ld, C)))
    return set_error(VFD_ERR_ARG, "sqdiff_bwd: bad gradient tensor");
  if (V <= 0) return VFD_OK;
  sqdiff_bwd_kernel<<<blocks_for(V * (C / 8), 256, 148 * 8), 256, 0, STREAM>>>(
      (const bf16*)a, a_ld, (const bf16*)b, b_ld, C, V, gscale, scale, (bf16*)ga, ga_ld, (bf16*)gb, gb_ld);
  return check_launch("sqdiff_bwd");
}

VFD_API int vfd_l1_loss(const float* a, const float* b, long long V, float grad_scale, double* sum, float* ga,
                        void* stream_) {
  if (V <= 0) return VFD_OK;
  if (a == nullptr || b == nullptr || sum == nullptr) return set_error(VFD_ERR_ARG, "l1_loss: null pointer");
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(ga)) % 16)
    return set_error(VFD_ERR_ARG, "l1_loss: tensors must be 16-byte aligned");
  if (V <= 0) return VFD_OK;
  l1_kernel<<<blocks_for(V / 4 + 1, 256, 148 * 4), 256, 0, STREAM>>>(a, b, V, grad_scale, sum, ga);
  return check_launch("l1_loss");
}

VFD_API int vfd_bce_loss(const float* p, const float* t, long long V, float grad_scale, double* sum, float* gp,
                         void* stream_) {
  if (V <= 0) return VFD_OK;
  if (p == nullptr || t == nullptr || sum == nullptr) return set_error(VFD_ERR_ARG, "bce_loss: null pointer");
  if (V <= 0) return VFD_OK;
  bce_kernel<<<blocks_for(V, 256, 148 * 4), 256, 0, STREAM>>>(p, t, V, grad_scale, sum, gp);
  return check_launch("bce_loss");
}

VFD_API int vfd_score_finalize(const double* per_clip, int n, double inv_count, float* scores, float* minmax,
                               void* stream_) {
  if (n == 0) return VFD_OK;
  if (per_clip == nullptr || scores == nullptr || n < 0) return set_error(VFD_ERR_ARG, "score_finalize: bad arguments");
  score_finalize_kernel<<<1, 256, 0, STREAM>>>(per_clip, n, inv_count, scores, minmax);
  return check_launch("score_finalize");
}

VFD_API int vfd_score_scale(const float* scores, long long n, const float* minmax, float* out, void* stream_) {
  if (n == 0) return VFD_OK;
  if (scores == nullptr || minmax == nullptr || out == nullptr || n < 0)
    return set_error(VFD_ERR_ARG, "score_scale: bad arguments");
  score_scale_kernel<<<blocks_for(n, 256, 148 * 4), 256, 0, STREAM>>>(scores, n, minmax, out);
  return check_launch("score_scale");
}

VFD_API int vfd_threshold_open(const float* predict, int N, int D, int H, int W, float thr, float* t_out,
                               float* m_out, void* stream_) {
  if (N == 0) return VFD_OK;
  if (predict == nullptr || m_out == nullptr || N < 0 || D <= 0 || H <= 0 || W <= 0)
    return set_error(VFD_ERR_ARG, "threshold_open: bad arguments");
  if (W > 512) return set_error(VFD_ERR_ARG, "threshold_open: W > 512 (OpenCV's channel limit; the reference fails too)");
  if (N == 0) return VFD_OK;
  if (N > 65535) return set_error(VFD_ERR_ARG, "threshold_open: at most 65535 clips per call");
  int tw = 32;
  while (tw > 1 && 2ll * D * H * tw > 200 * 1024) tw >>= 1;
  const long long smem = 2ll * D * H * tw;
  if (smem > 200 * 1024) return set_error(VFD_ERR_ARG, "threshold_open: a (D, H) plane does not fit shared memory");
  cudaError_t e = cudaFuncSetAttribute(threshold_open_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_cuda_error(e, "threshold_open: cudaFuncSetAttribute");
  threshold_open_kernel<<<dim3((W + tw - 1) / tw, N), 512, smem, STREAM>>>(predict, D, H, W, tw, thr, t_out, m_out);
  return check_launch("threshold_open");
}

VFD_API int vfd_confusion_counts(const float* labels, const float* scores, long long n, float thr,
                                 unsigned long long* counts, void* stream_) {
  if (n == 0) return VFD_OK;
  if (labels == nullptr || scores == nullptr || counts == nullptr || n < 0)
    return set_error(VFD_ERR_ARG, "confusion_counts: bad arguments");
  if ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(scores)) % 16)
    return set_error(VFD_ERR_ARG, "confusion_counts: tensors must be 16-byte aligned");
  if (n == 0) return VFD_OK;
  confusion_kernel<<<blocks_for(n / 4 + 1, 256, 148 * 8), 256, 0, STREAM>>>(labels, scores, n, thr, counts);
  return check_launch("confusion_counts");
}

VFD_API int vfd_roc_auc(const float* scores, const float* labels, int n, double* out, void* stream_) {
  if (out == nullptr || n < 0 || (n > 0 && (scores == nullptr || labels == nullptr)))
    return set_error(VFD_ERR_ARG, "roc_auc: bad arguments");
  if (n > kAucMax) return set_error(VFD_ERR_ARG, "roc_auc: at most 16384 scores per call");
  int npad = kAucThreads;
  while (npad < n) npad <<= 1;
  const int smem = npad * 8;
  cudaError_t e = cudaFuncSetAttribute(auc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return set_cuda_error(e, "roc_auc: cudaFuncSetAttribute");
  auc_kernel<<<1, kAucThreads, smem, STREAM>>>(scores, labels, n, npad, out);
  return check_launch("roc_auc");
}
