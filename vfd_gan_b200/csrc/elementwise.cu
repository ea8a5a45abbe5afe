// HBM-bound kernels of the GAN step: layout packing, BatchNorm3d statistics / fused
// BatchNorm + (Leaky)ReLU (+ AvgPool3d, + Dropout) forward and backward, trilinear x2 upsample
// into a concat buffer, loss reductions and the ConvLSTM cell update.
//
// All activations are channels-last bf16 [N][D][H][W][ld] with the logical channel count C a
// multiple of 8, so every thread moves one 128-bit vector (8 channels) per access and a warp
// touches consecutive 16-byte chunks of a voxel row.
//
// Reference call sites replaced: nn.BatchNorm3d/ReLU (models/spatiotempconv.py:51-52,63),
// BatchNorm3d/LeakyReLU (models/mygannet.py:19-20,25-26,109-110,114-115), AvgPool3d (:41,132,174),
// Dropout (:49,76), Upsample+cat (:50,77-94), l2_loss / weighted_bce (lib/utils.py:59-71),
// ConvLSTM cell update (models/convlstm.py:49-58).
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdlib>
#include "vfd_internal.h"

namespace vfd {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void load8(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// Inverted-dropout multipliers (0 or 1/(1-p)) for the 8 channels of group cg at voxel vox.
__device__ __forceinline__ void dropout8(unsigned long long seed, long long vox, int cg, float p,
                                         float* m) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint4 a = philox4x32_10(make_uint4((uint32_t)vox, (uint32_t)(vox >> 32), (uint32_t)cg, 0u), key);
  const uint4 b = philox4x32_10(make_uint4((uint32_t)vox, (uint32_t)(vox >> 32), (uint32_t)cg, 1u), key);
  const uint32_t r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const float keep = 1.0f / (1.0f - p);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = ((r[i] >> 8) * (1.0f / 16777216.0f) >= p) ? keep : 0.0f;
}

// ------------------------------------------------------------------------------ layout packing
// fp32 NCDHW [N][Csrc][S] -> bf16 channels-last [N][S][ld]; channel c of the output reads source
// channel (c % Csrc) when `replicate` (gray2rgb, lib/utils.py:91-92), zero beyond C.
__global__ void pack_ncdhw_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N,
                                  int Csrc, long long S, int C, long long ld, int Cp,
                                  int replicate) {
  const long long total = (long long)N * S * (Cp / 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const long long t = i / S;
    const int cg = (int)(t % (Cp / 8));
    const long long n = t / (Cp / 8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      float x = 0.f;
      if (c < C) {
        const int cs = replicate ? (c % Csrc) : c;
        x = src[((long long)n * Csrc + cs) * S + s];
      }
      v[j] = x;
    }
    store8(dst + ((long long)n * S + s) * ld + cg * 8, v);
  }
}

// channels-last (bf16 or fp32) [N][S][ld] -> fp32 NCDHW [N][C][S]
template <typename T>
__global__ void unpack_ncdhw_kernel(const T* __restrict__ src, float* __restrict__ dst, int N,
                                    int C, long long S, long long ld) {
  const long long total = (long long)N * C * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    const long long t = i / S;
    const int c = (int)(t % C);
    const long long n = t / C;
    dst[i] = (float)src[((long long)n * S + s) * ld + c];
  }
}

// fp32 conv weight [Cout][Cin][taps] -> bf16 packed GEMM operand.
//  mode 0 (forward): Wp[row=co][tap][c=ci]
//  mode 1 (dgrad)  : Wp[row=ci][tap][c=co] with the taps mirrored (tap -> taps-1-tap)
__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int Cout,
                                   int Cin, int taps, int rows, int ck, int mode) {
  const long long total = (long long)rows * taps * ck;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ck);
    const long long t = i / ck;
    const int tap = (int)(t % taps);
    const int row = (int)(t / taps);
    float v = 0.f;
    if (mode == 0) {
      if (row < Cout && c < Cin) v = w[((long long)row * Cin + c) * taps + tap];
    } else {
      if (row < Cin && c < Cout) v = w[((long long)c * Cin + row) * taps + (taps - 1 - tap)];
    }
    wp[i] = __float2bfloat16(v);
  }
}

// All conv weights of a network in ONE launch: job j packs weight j into one of its two GEMM layouts
// (see pack_weight_kernel). begin[] is the running element offset, so a thread finds its job by a short scan.
struct PackJob {
  const float* w;
  bf16* dst;
  int cout, cin, taps, rows, ck, mode;
  long long begin;   // first flattened element of this job
};
constexpr int kPackPerBlock = 256 * 8;   // flattened output elements per block
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const PackJob* __restrict__ jobs, int njobs, long long total) {
  // one job lookup per block (binary search for the block's first element); a block that straddles job
  // boundaries walks forward from there
  __shared__ int s_job;
  const long long base = (long long)blockIdx.x * kPackPerBlock;
  if (threadIdx.x == 0) {
    int lo = 0, hi = njobs - 1;   // last job with begin <= base
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].begin <= base) lo = mid; else hi = mid - 1;
    }
    s_job = lo;
  }
  __syncthreads();
  int j = s_job;
  PackJob jb = jobs[j];
  long long next_begin = j + 1 < njobs ? jobs[j + 1].begin : total;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const long long i = base + k * 256 + threadIdx.x;
    if (i >= total) break;
    while (i >= next_begin) {
      ++j;
      jb = jobs[j];
      next_begin = j + 1 < njobs ? jobs[j + 1].begin : total;
    }
    const unsigned e = (unsigned)(i - jb.begin);          // a single packed weight has < 2^32 elements
    const unsigned t = e / (unsigned)jb.ck;
    const int c = (int)(e - t * (unsigned)jb.ck);
    const int row = (int)(t / (unsigned)jb.taps);
    const int tap = (int)(t - (unsigned)row * (unsigned)jb.taps);
    float v = 0.f;
    if (jb.mode == 0) {
      if (row < jb.cout && c < jb.cin) v = jb.w[((long long)row * jb.cin + c) * jb.taps + tap];
    } else {
      if (row < jb.cin && c < jb.cout) v = jb.w[((long long)c * jb.cin + row) * jb.taps + (jb.taps - 1 - tap)];
    }
    jb.dst[e] = __float2bfloat16(v);
  }
}

// wgrad accumulator [taps][ci_pad][co_pad] fp32 -> weight gradient [Cout][Cin][taps] fp32
__global__ void unpack_wgrad_kernel(const float* __restrict__ acc, float* __restrict__ gw, int Cout,
                                    int Cin, int taps, int co_pad, int ci_pad, int accumulate) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const long long t = i / taps;
    const int ci = (int)(t % Cin);
    const int co = (int)(t / Cin);
    const float v = acc[((long long)tap * ci_pad + ci) * co_pad + co];
    gw[i] = accumulate ? gw[i] + v : v;
  }
}

// ------------------------------------------------------------------------------ BN statistics
// Per-channel sum / sum of squares over V voxels; fp32 per-thread partials, double across blocks.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const bf16* __restrict__ x, long long ld, int C, long long V,
                double* __restrict__ sums) {
  extern __shared__ double sh[];  // [2][C]; double: exact sums of the threads' fp32 partials, order-independent
  const int CG = C / 8;
  const int rows_per_iter = 256 / CG;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * C; i += 256) sh[i] = 0.0;
  __syncthreads();
  const bool active = tid < rows_per_iter * CG;
  const int cg = tid % CG;
  const int rl = tid / CG;
  const long long per_block = (V + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per_block;
  const long long r1 = (r0 + per_block < V) ? (r0 + per_block) : V;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  if (active) {
    for (long long r = r0 + rl; r < r1; r += 4 * rows_per_iter) {
      uint4 raw[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + (long long)u * rows_per_iter;
        ok[u] = rr < r1;
        if (ok[u]) raw[u] = __ldg(reinterpret_cast<const uint4*>(x + rr * ld + cg * 8));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h[j]);
          s[2 * j] += f.x;
          s[2 * j + 1] += f.y;
          ss[2 * j] = fmaf(f.x, f.x, ss[2 * j]);
          ss[2 * j + 1] = fmaf(f.y, f.y, ss[2 * j + 1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[cg * 8 + j], static_cast<double>(s[j]));
      atomicAdd(&sh[C + cg * 8 + j], static_cast<double>(ss[j]));
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += 256) atomicAdd(&sums[i], sh[i]);
}

// sums -> (mean, invstd, scale, shift), running-stat update, and clears the accumulator.
// Matches nn.BatchNorm3d: biased variance for normalisation, unbiased for running_var,
// momentum 0.1, eps 1e-5. In eval mode (train == 0) the running statistics are used instead.
// pre_bias (optional): bias of the producing conv that was NOT added to the stored tensor -- a
// per-channel constant cancels in the normalisation, so it only enters the running mean (train)
// or the shift (eval). Keeping it out of the bf16 tensor keeps the stored values centred.
__global__ void bn_finalize_kernel(double* __restrict__ sums, int C, int Cvalid, long long V,
                                   const float* __restrict__ pre_bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int train, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, float* __restrict__ scale_out,
                                   float* __restrict__ shift_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean = 0.f, invstd = 0.f, sc = 0.f, sh = 0.f;
  if (c < Cvalid) {
    const float pb = pre_bias != nullptr ? pre_bias[c] : 0.f;
    if (train) {
      const double m = sums[c] / (double)V;
      double var = sums[C + c] / (double)V - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean != nullptr) {
        const double unbiased = V > 1 ? var * (double)V / (double)(V - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mean + pb);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    } else {
      mean = running_mean[c] - pb;
      invstd = rsqrtf(running_var[c] + eps);
    }
    sc = gamma[c] * invstd;
    sh = beta[c] - mean * sc;
  }
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  scale_out[c] = sc;
  shift_out[c] = sh;
  sums[c] = 0.0;
  sums[C + c] = 0.0;
}

// Division by a run-time constant without the integer divide sequence: q = (umulhi(n, m) + n) >> s,
// exact for n < 2^31 (all voxel / window counts here are checked to be below that).
struct FastDiv {
  uint32_t m, s, d;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.m) + n) >> f.s; }

struct ActGeom {
  int N, D, H, W;     // full-resolution voxel grid
  int pd, ph, pw;     // pooling window (1 or 2 per axis); floor semantics like nn.AvgPool3d
  int C;              // padded channel count (multiple of 8)
  int WD, WH, WW;     // windows per axis (ceil: partial windows still carry full-resolution voxels)
  int QD, QH, QW;     // pooled extent (floor)
  unsigned nwin;
  FastDiv fWW, fWH, fWD;
  int pool_bcast;     // backward only: g_pool is [N][QD][C], broadcast over the H and W axes
  // backward "gather" mode: windows are single voxels (pd = ph = pw = 1 above, so every load / store is contiguous
  // across the warp whatever the pooling), and each voxel fetches the pooled gradient of the (gpd, gph, gpw) window
  // it belongs to (served by L1 / L2 for the window's other voxels). 0 = off.
  int gpd, gph, gpw;
  float inv_win;
};

// The BN/activation kernels map one thread to (pool window, 8-channel group). A block owns a
// contiguous range of windows; a thread keeps its channel group for the whole kernel (so the
// per-channel constants live in registers) and strides over windows. All loads of an iteration are
// issued before any arithmetic (U windows x NV voxels independent 128-bit loads per tensor). These
// kernels are instruction-issue bound before they are HBM bound, so the index math avoids integer
// division (FastDiv; none at all for un-pooled tensors) and the per-element arithmetic is folded into
// the fewest FFMA / FMNMX it takes.
__device__ __forceinline__ void decode_win(unsigned win, const ActGeom& g, int& n, int& wd, int& wh, int& ww) {
  const unsigned t = fdiv(win, g.fWW);
  ww = win - t * g.WW;
  const unsigned t2 = fdiv(t, g.fWH);
  wh = t - t2 * g.WH;
  const unsigned t3 = fdiv(t2, g.fWD);
  wd = t2 - t3 * g.WD;
  n = t3;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  // bf16 -> fp32 is a 16-bit shift: low halves by shift, high halves by mask
  v[0] = __uint_as_float(u.x << 16);
  v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16);
  v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  v[4] = __uint_as_float(u.z << 16);
  v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  v[6] = __uint_as_float(u.w << 16);
  v[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack4(const uint2& u, float* v) {
  v[0] = __uint_as_float(u.x << 16);
  v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16);
  v[3] = __uint_as_float(u.y & 0xFFFF0000u);
}
__device__ __forceinline__ uint2 ldg8(const bf16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint4 ldg16(const bf16* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// Voxels of one window (template pool shape): indices, validity, pooled index.
template <int PD, int PH, int PW>
struct Window {
  static constexpr int NV = PD * PH * PW;
  unsigned vox[NV];
  bool ok[NV];
  bool pool_ok;
  unsigned pvox;
  __device__ __forceinline__ void locate(unsigned win, bool wv, const ActGeom& g) {
    if (NV == 1 && !g.pool_bcast && !g.gpd) {
      vox[0] = win;
      ok[0] = wv;
      pool_ok = wv;
      pvox = win;
      return;
    }
    int n, wd, wh, ww;
    decode_win(wv ? win : 0u, g, n, wd, wh, ww);
    if (NV == 1 && g.gpd) {
      const int qd = wd >> (g.gpd - 1), qh = wh >> (g.gph - 1), qw = ww >> (g.gpw - 1);
      vox[0] = win;
      ok[0] = wv;
      pool_ok = wv && qd < g.QD && qh < g.QH && qw < g.QW;
      pvox = g.pool_bcast ? (n * g.QD + qd) : (((n * g.QD + qd) * g.QH + qh) * g.QW + qw);
      return;
    }
    pool_ok = wv && wd < g.QD && wh < g.QH && ww < g.QW;
    pvox = g.pool_bcast ? (n * g.QD + wd) : (((n * g.QD + wd) * g.QH + wh) * g.QW + ww);
    const unsigned base = ((n * g.D + wd * PD) * g.H + wh * PH) * g.W + ww * PW;
#pragma unroll
    for (int a = 0; a < PD; ++a)
#pragma unroll
      for (int b = 0; b < PH; ++b)
#pragma unroll
        for (int c = 0; c < PW; ++c) {
          const int v = (a * PH + b) * PW + c;
          ok[v] = wv && (PD == 1 || wd * PD + a < g.D) && (PH == 1 || wh * PH + b < g.H) &&
                  (PW == 1 || ww * PW + c < g.W);
          vox[v] = base + (a * g.H + b) * g.W + c;
        }
  }
};

// out = dropout(act(y * scale + shift)); optionally also the average-pooled tensor.
// act(z) = max(z, slope * z) for 0 <= slope <= 1 (ReLU, LeakyReLU).
template <int PD, int PH, int PW, bool DROP>
__global__ void __launch_bounds__(256, 2)
bn_act_fwd_kernel(const bf16* __restrict__ y, long long y_ld, ActGeom g,
                  const float* __restrict__ scale, const float* __restrict__ shift, float slope,
                  bf16* __restrict__ out_full, long long full_ld, bf16* __restrict__ out_pool,
                  long long pool_ld, float drop_p, unsigned long long seed,
                  const unsigned long long* __restrict__ seed_dev) {
  constexpr int NV = PD * PH * PW;
  constexpr int U = NV >= 8 ? 1 : 8 / NV;
  if (DROP && seed_dev != nullptr) seed += *seed_dev;  // per-step offset kept on the device (CUDA-graph replay)
  const int CG = g.C >> 3;
  const int rpi = 256 / CG;
  const int tid = threadIdx.x;
  if (tid >= rpi * CG) return;
  const int cg = tid % CG, wl = tid / CG;
  const unsigned per_block = (g.nwin + gridDim.x - 1) / gridDim.x;
  const unsigned w_begin = blockIdx.x * per_block;
  const unsigned w_end = min(g.nwin, w_begin + per_block);
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + cg * 8 + j);
    sf[j] = __ldg(shift + cg * 8 + j);
  }
  y += cg * 8;
  if (out_full != nullptr) out_full += cg * 8;
  if (out_pool != nullptr) out_pool += cg * 8;
  const float inv_win = 1.0f / (float)NV;
  for (unsigned wb = w_begin + wl; wb < w_end; wb += rpi * U) {
    Window<PD, PH, PW> wn[U];
    uint4 raw[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned win = wb + u * rpi;
      wn[u].locate(win, win < w_end, g);
#pragma unroll
      for (int v = 0; v < NV; ++v)
        if (wn[u].ok[v]) raw[u][v] = ldg16(y + (size_t)wn[u].vox[v] * y_ld);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!wn[u].ok[v]) continue;
        float x[8];
        unpack8(raw[u][v], x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = fmaf(x[j], sc[j], sf[j]);
          x[j] = fmaxf(z, z * slope);
        }
        if (DROP) {
          float m[8];
          dropout8(seed, (long long)wn[u].vox[v], cg, drop_p, m);
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] *= m[j];
        }
        if (out_full != nullptr) store8(out_full + (size_t)wn[u].vox[v] * full_ld, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += x[j];
      }
      if (out_pool != nullptr && wn[u].pool_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv_win;
        store8(out_pool + (size_t)wn[u].pvox * pool_ld, acc);
      }
    }
  }
}

// ---- backward, 8-channel window-per-thread mapping (un-pooled and depth-only pooled tensors)
// Shared front end of the two backward passes: loads y / g_full for the window's voxels and the
// window's g_pool, and turns them into the gradient w.r.t. the BatchNorm output.
template <int PD, int PH, int PW, bool DROP>
struct BwdWindow8 {
  static constexpr int NV = PD * PH * PW;
  Window<PD, PH, PW> w;
  uint4 ry[NV], rg[NV], rp;
  float inv_win;

  __device__ __forceinline__ void load(unsigned win, bool wv, const ActGeom& g, const bf16* __restrict__ y,
                                       long long y_ld, const bf16* __restrict__ g_full, long long gf_ld,
                                       const bf16* __restrict__ g_pool, long long gp_ld) {
    w.locate(win, wv, g);
    inv_win = g.inv_win;
    w.pool_ok = w.pool_ok && g_pool != nullptr;
    if (w.pool_ok) rp = ldg16(g_pool + (size_t)w.pvox * gp_ld);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (w.ok[v]) {
        ry[v] = ldg16(y + (size_t)w.vox[v] * y_ld);
        if (g_full != nullptr) rg[v] = ldg16(g_full + (size_t)w.vox[v] * gf_ld);
      }
  }
  // gradient w.r.t. z = y*scale+shift for voxel v (yv receives the unpacked y)
  __device__ __forceinline__ void grad(int v, bool has_full, const float* sc, const float* sf,
                                       float slope, float drop_p, unsigned long long seed, int cg,
                                       float* yv, float* gv) const {
    unpack8(ry[v], yv);
    if (has_full) {
      unpack8(rg[v], gv);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = 0.f;
    }
    if (w.pool_ok) {
      float t[8];
      unpack8(rp, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] = fmaf(t[j], inv_win, gv[j]);
    }
    if (DROP) {
      float m[8];
      dropout8(seed, (long long)w.vox[v], cg, drop_p, m);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv[j] *= m[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(yv[j], sc[j], sf[j]);
      gv[j] = z > 0.f ? gv[j] : gv[j] * slope;
    }
  }
};

// Tail of both reduce kernels when a ticket counter is supplied: the block that finishes last (every block adds its
// partials to `sums` with atomics, fences, then takes a ticket) turns the complete sums into dgamma / dbeta and the two
// per-channel means of pass 2 and clears the accumulator and the ticket -- what bn_bwd_finalize_kernel does as a separate
// launch otherwise (56 launches of ~3 us per train step on the critical path).
struct BwdFinalize {
  unsigned int* ticket;   // nullptr: the caller launches bn_bwd_finalize_kernel
  int Cvalid, train, accumulate;
  long long V;
  float *dgamma, *dbeta, *c1, *c2;
};
__device__ __forceinline__ void bwd_finalize_by_last_block(const BwdFinalize& f, double* sums, int C) {
  if (f.ticket == nullptr) return;
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(f.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double sg = __ldcg(sums + c), sgx = __ldcg(sums + C + c);
    if (c < f.Cvalid) {
      f.dbeta[c] = f.accumulate ? f.dbeta[c] + (float)sg : (float)sg;
      f.dgamma[c] = f.accumulate ? f.dgamma[c] + (float)sgx : (float)sgx;
    }
    f.c1[c] = f.train ? (float)(sg / (double)f.V) : 0.f;
    f.c2[c] = f.train ? (float)(sgx / (double)f.V) : 0.f;
    sums[c] = 0.0;
    sums[C + c] = 0.0;
  }
  if (threadIdx.x == 0) *f.ticket = 0u;
}

// Pass 1 of BN backward: per-channel sum(g) and sum(g * xhat).
template <int PD, int PH, int PW, bool DROP>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd8_reduce_kernel(const bf16* __restrict__ y, long long y_ld, ActGeom g,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         float slope, const bf16* __restrict__ g_full, long long gf_ld,
                         const bf16* __restrict__ g_pool, long long gp_ld, float drop_p,
                         unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                         double* __restrict__ sums, BwdFinalize fin) {
  __shared__ float part[16 * 256];  // [sum | sum*xhat][channel in group][thread]
  constexpr int NV = PD * PH * PW;
  constexpr int U = NV >= 4 ? 1 : 4 / NV;
  if (DROP && seed_dev != nullptr) seed += *seed_dev;
  const int C = g.C;
  const int CG = C >> 3;
  const int rpi = 256 / CG;
  const int tid = threadIdx.x;
  if (tid < rpi * CG) {
    const int cg = tid % CG, wl = tid / CG;
    const unsigned per_block = (g.nwin + gridDim.x - 1) / gridDim.x;
    const unsigned w_begin = blockIdx.x * per_block;
    const unsigned w_end = min(g.nwin, w_begin + per_block);
    float sc[8], sf[8], is[8], nm[8], s1[8], s2[8];   // xhat = y * is + nm, nm = -mean * invstd
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + cg * 8 + j);
      sf[j] = __ldg(shift + cg * 8 + j);
      is[j] = __ldg(invstd + cg * 8 + j);
      nm[j] = -__ldg(mean + cg * 8 + j) * is[j];
      s1[j] = s2[j] = 0.f;
    }
    y += cg * 8;
    if (g_full != nullptr) g_full += cg * 8;
    if (g_pool != nullptr) g_pool += cg * 8;
    const bool has_full = g_full != nullptr;
    for (unsigned wb = w_begin + wl; wb < w_end; wb += rpi * U) {
      BwdWindow8<PD, PH, PW, DROP> bw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        bw[u].load(wb + u * rpi, wb + u * rpi < w_end, g, y, y_ld, g_full, gf_ld, g_pool, gp_ld);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (!bw[u].w.ok[v]) continue;
          float yv[8], gv[8];
          bw[u].grad(v, has_full, sc, sf, slope, drop_p, seed, cg, yv, gv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += gv[j];
            s2[j] = fmaf(gv[j], fmaf(yv[j], is[j], nm[j]), s2[j]);
          }
        }
    }
    // block reduction without shared-memory atomics (with few channel groups every thread of the block
    // would hit the same handful of addresses): stage the partials, then one thread per channel sums them
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part[j * 256 + tid] = s1[j];
      part[(8 + j) * 256 + tid] = s2[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += 256) {
    const int which = i >= C, c = which ? i - C : i;
    const int cgi = c >> 3, j = c & 7;
    const float* src = part + (which * 8 + j) * 256 + cgi;
    float acc = 0.f;
    for (int wl = 0; wl < rpi; ++wl) acc += src[wl * CG];
    atomicAdd(&sums[i], (double)acc);
  }
  bwd_finalize_by_last_block(fin, sums, C);
}

// Pass 2 of BN backward: dy = scale * (g - c1 - xhat * c2) = scale * g + (kb + kc * xhat)
template <int PD, int PH, int PW, bool DROP>
__global__ void __launch_bounds__(256, 2)
bn_act_bwd8_apply_kernel(const bf16* __restrict__ y, long long y_ld, ActGeom g,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        float slope, const bf16* __restrict__ g_full, long long gf_ld,
                        const bf16* __restrict__ g_pool, long long gp_ld, float drop_p,
                        unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                        const float* __restrict__ c1, const float* __restrict__ c2, bf16* __restrict__ dy,
                        long long dy_ld) {
  constexpr int NV = PD * PH * PW;
  constexpr int U = NV >= 4 ? 1 : 4 / NV;
  if (DROP && seed_dev != nullptr) seed += *seed_dev;
  const int CG = g.C >> 3;
  const int rpi = 256 / CG;
  const int tid = threadIdx.x;
  if (tid >= rpi * CG) return;
  const int cg = tid % CG, wl = tid / CG;
  const unsigned per_block = (g.nwin + gridDim.x - 1) / gridDim.x;
  const unsigned w_begin = blockIdx.x * per_block;
  const unsigned w_end = min(g.nwin, w_begin + per_block);
  float sc[8], sf[8], is[8], nm[8], kb[8], kc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + cg * 8 + j);
    sf[j] = __ldg(shift + cg * 8 + j);
    is[j] = __ldg(invstd + cg * 8 + j);
    nm[j] = -__ldg(mean + cg * 8 + j) * is[j];
    kb[j] = -sc[j] * __ldg(c1 + cg * 8 + j);
    kc[j] = -sc[j] * __ldg(c2 + cg * 8 + j);
  }
  y += cg * 8;
  dy += cg * 8;
  if (g_full != nullptr) g_full += cg * 8;
  if (g_pool != nullptr) g_pool += cg * 8;
  const bool has_full = g_full != nullptr;
  for (unsigned wb = w_begin + wl; wb < w_end; wb += rpi * U) {
    BwdWindow8<PD, PH, PW, DROP> bw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      bw[u].load(wb + u * rpi, wb + u * rpi < w_end, g, y, y_ld, g_full, gf_ld, g_pool, gp_ld);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!bw[u].w.ok[v]) continue;
        float yv[8], gv[8], o[8];
        bw[u].grad(v, has_full, sc, sf, slope, drop_p, seed, cg, yv, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = fmaf(sc[j], gv[j], fmaf(kc[j], fmaf(yv[j], is[j], nm[j]), kb[j]));
        store8(dy + (size_t)bw[u].w.vox[v] * dy_ld, o);
      }
  }
}

// ---- backward for spatially pooled tensors (ph = pw = 2). One thread = (plane window of PH x PW voxels of
// one d-plane, 4-channel group): half the
// per-channel state of the 8-channel forward mapping and at most four voxels per window, which keeps the
// kernels under 85 registers without spills (24 resident warps per SM) -- they are latency bound, not issue bound. Depth
// pooling (pd = 2) only changes which pooled-gradient element a plane reads.
struct BwdGeom {
  int N, D, H, W, C;
  int WH, WW;        // windows per plane axis (ceil)
  int QD, QH, QW;    // pooled extents (floor)
  int pd;            // depth pooling factor (1 or 2)
  unsigned nwin;     // N * D * WH * WW
  FastDiv fWW, fWH, fD;
  float inv_win;     // 1 / (pd * PH * PW)
  int pool_bcast;    // g_pool is [N][QD][C], broadcast over the H and W axes (PH == PW == 1 only)
};

__device__ __forceinline__ void store4(bf16* p, const float* v) {
  uint2 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <int PH, int PW, bool DROP>
struct BwdWindow {
  static constexpr int NV = PH * PW;
  unsigned vox[NV];
  bool ok[NV], pool_ok;
  uint2 ry[NV], rg[NV], rp;

  __device__ __forceinline__ void load(unsigned win, bool wv, const BwdGeom& g, const bf16* __restrict__ y,
                                       long long y_ld, const bf16* __restrict__ g_full, long long gf_ld,
                                       const bf16* __restrict__ g_pool, long long gp_ld) {
    unsigned pvox;
    if (NV == 1 && g.pd == 1 && !g.pool_bcast) {
      vox[0] = win;
      ok[0] = wv;
      pool_ok = wv && g_pool != nullptr;
      pvox = win;
    } else {
      const unsigned w0 = wv ? win : 0u;
      const unsigned t = fdiv(w0, g.fWW);
      const int ww = w0 - t * g.WW;
      const unsigned t2 = fdiv(t, g.fWH);
      const int wh = t - t2 * g.WH;
      const unsigned n = fdiv(t2, g.fD);
      const int d = t2 - n * g.D;
      const int wd = g.pd == 2 ? (d >> 1) : d;
      pool_ok = wv && g_pool != nullptr && wd < g.QD && wh < g.QH && ww < g.QW;
      pvox = g.pool_bcast ? (n * g.QD + wd) : (((n * g.QD + wd) * g.QH + wh) * g.QW + ww);
      const unsigned base = ((n * g.D + d) * g.H + wh * PH) * g.W + ww * PW;
#pragma unroll
      for (int b = 0; b < PH; ++b)
#pragma unroll
        for (int c = 0; c < PW; ++c) {
          ok[b * PW + c] = wv && (PH == 1 || wh * PH + b < g.H) && (PW == 1 || ww * PW + c < g.W);
          vox[b * PW + c] = base + b * g.W + c;
        }
    }
    if (pool_ok) rp = ldg8(g_pool + (size_t)pvox * gp_ld);
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (ok[v]) {
        ry[v] = ldg8(y + (size_t)vox[v] * y_ld);
        if (g_full != nullptr) rg[v] = ldg8(g_full + (size_t)vox[v] * gf_ld);
      }
  }
  // gradient w.r.t. z = y*scale+shift for voxel v (yv receives the unpacked y)
  __device__ __forceinline__ void grad(int v, bool has_full, const float* sc, const float* sf, float slope,
                                       float inv_win, float drop_p, unsigned long long seed, int cg4, float* yv,
                                       float* gv) const {
    unpack4(ry[v], yv);
    if (has_full) {
      unpack4(rg[v], gv);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = 0.f;
    }
    if (pool_ok) {
      float t[4];
      unpack4(rp, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] = fmaf(t[j], inv_win, gv[j]);
    }
    if (DROP) {
      float m[8];
      dropout8(seed, (long long)vox[v], cg4 >> 1, drop_p, m);
#pragma unroll
      for (int j = 0; j < 4; ++j) gv[j] *= m[(cg4 & 1) * 4 + j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = fmaf(yv[j], sc[j], sf[j]);
      gv[j] = z > 0.f ? gv[j] : gv[j] * slope;
    }
  }
};

// Pass 1 of BN backward: per-channel sum(g) and sum(g * xhat). The loop accumulates sum(g) and sum(g * y);
// sum(g * xhat) = invstd * (sum(g * y) - mean * sum(g)) is formed once per thread at the end.
template <int PH, int PW, bool DROP>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_reduce_kernel(const bf16* __restrict__ y, long long y_ld, BwdGeom g,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         float slope, const bf16* __restrict__ g_full, long long gf_ld,
                         const bf16* __restrict__ g_pool, long long gp_ld, float drop_p,
                         unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                         double* __restrict__ sums, BwdFinalize fin) {
  __shared__ float part[8 * 256];  // [sum | sum*xhat][channel in group][thread]
  constexpr int NV = PH * PW;
  constexpr int U = 8 / NV;
  if (DROP && seed_dev != nullptr) seed += *seed_dev;
  const int C = g.C;
  const int CG = C >> 2;
  const int rpi = 256 / CG;
  const int tid = threadIdx.x;
  if (tid < rpi * CG) {
    const int cg = tid % CG, wl = tid / CG;
    const unsigned per_block = (g.nwin + gridDim.x - 1) / gridDim.x;
    const unsigned w_begin = blockIdx.x * per_block;
    const unsigned w_end = min(g.nwin, w_begin + per_block);
    float sc[4], sf[4], s1[4], sy[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc[j] = __ldg(scale + cg * 4 + j);
      sf[j] = __ldg(shift + cg * 4 + j);
      s1[j] = sy[j] = 0.f;
    }
    y += cg * 4;
    if (g_full != nullptr) g_full += cg * 4;
    if (g_pool != nullptr) g_pool += cg * 4;
    const bool has_full = g_full != nullptr;
    for (unsigned wb = w_begin + wl; wb < w_end; wb += rpi * U) {
      BwdWindow<PH, PW, DROP> bw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        bw[u].load(wb + u * rpi, wb + u * rpi < w_end, g, y, y_ld, g_full, gf_ld, g_pool, gp_ld);
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (!bw[u].ok[v]) continue;
          float yv[4], gv[4];
          bw[u].grad(v, has_full, sc, sf, slope, g.inv_win, drop_p, seed, cg, yv, gv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s1[j] += gv[j];
            sy[j] = fmaf(gv[j], yv[j], sy[j]);
          }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float is = __ldg(invstd + cg * 4 + j), mu = __ldg(mean + cg * 4 + j);
      part[j * 256 + tid] = s1[j];
      part[(4 + j) * 256 + tid] = is * (sy[j] - mu * s1[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < 2 * C; i += 256) {
    const int which = i >= C, c = which ? i - C : i;
    const int cgi = c >> 2, j = c & 3;
    const float* src = part + (which * 4 + j) * 256 + cgi;
    float acc = 0.f;
    for (int wl = 0; wl < rpi; ++wl) acc += src[wl * CG];
    atomicAdd(&sums[i], (double)acc);
  }
  bwd_finalize_by_last_block(fin, sums, C);
}

// sums -> dgamma, dbeta and the two per-channel means used by pass 2; clears the accumulator.
__global__ void bn_bwd_finalize_kernel(double* __restrict__ sums, int C, int Cvalid, long long V,
                                       int train, int accumulate, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ c1,
                                       float* __restrict__ c2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sg = sums[c], sgx = sums[C + c];
  if (c < Cvalid) {
    dbeta[c] = accumulate ? dbeta[c] + (float)sg : (float)sg;
    dgamma[c] = accumulate ? dgamma[c] + (float)sgx : (float)sgx;
  }
  c1[c] = train ? (float)(sg / (double)V) : 0.f;
  c2[c] = train ? (float)(sgx / (double)V) : 0.f;
  sums[c] = 0.0;
  sums[C + c] = 0.0;
}

// Pass 2 of BN backward: dy = scale * (g - c1 - xhat * c2) = scale * g + a2 * y + b2 with
// a2 = -scale * c2 * invstd, b2 = -scale * (c1 - c2 * mean * invstd).
template <int PH, int PW, bool DROP>
__global__ void __launch_bounds__(256, 3)
bn_act_bwd_apply_kernel(const bf16* __restrict__ y, long long y_ld, BwdGeom g,
                        const float* __restrict__ mean, const float* __restrict__ invstd,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        float slope, const bf16* __restrict__ g_full, long long gf_ld,
                        const bf16* __restrict__ g_pool, long long gp_ld, float drop_p,
                        unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                        const float* __restrict__ c1, const float* __restrict__ c2, bf16* __restrict__ dy,
                        long long dy_ld) {
  constexpr int NV = PH * PW;
  constexpr int U = 8 / NV;
  if (DROP && seed_dev != nullptr) seed += *seed_dev;
  const int CG = g.C >> 2;
  const int rpi = 256 / CG;
  const int tid = threadIdx.x;
  if (tid >= rpi * CG) return;
  const int cg = tid % CG, wl = tid / CG;
  const unsigned per_block = (g.nwin + gridDim.x - 1) / gridDim.x;
  const unsigned w_begin = blockIdx.x * per_block;
  const unsigned w_end = min(g.nwin, w_begin + per_block);
  float sc[4], sf[4], a2[4], b2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = __ldg(scale + cg * 4 + j);
    sf[j] = __ldg(shift + cg * 4 + j);
    const float is = __ldg(invstd + cg * 4 + j), mu = __ldg(mean + cg * 4 + j);
    const float k1 = __ldg(c1 + cg * 4 + j), k2 = __ldg(c2 + cg * 4 + j);
    a2[j] = -sc[j] * k2 * is;
    b2[j] = -sc[j] * (k1 - k2 * mu * is);
  }
  y += cg * 4;
  dy += cg * 4;
  if (g_full != nullptr) g_full += cg * 4;
  if (g_pool != nullptr) g_pool += cg * 4;
  const bool has_full = g_full != nullptr;
  for (unsigned wb = w_begin + wl; wb < w_end; wb += rpi * U) {
    BwdWindow<PH, PW, DROP> bw[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      bw[u].load(wb + u * rpi, wb + u * rpi < w_end, g, y, y_ld, g_full, gf_ld, g_pool, gp_ld);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!bw[u].ok[v]) continue;
        float yv[4], gv[4], o[4];
        bw[u].grad(v, has_full, sc, sf, slope, g.inv_win, drop_p, seed, cg, yv, gv);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(sc[j], gv[j], fmaf(a2[j], yv[j], b2[j]));
        store4(dy + (size_t)bw[u].vox[v] * dy_ld, o);
      }
  }
}

// Per-channel sum over voxels of a channels-last bf16 tensor (conv bias gradient).
__global__ void __launch_bounds__(256)
channel_sum_kernel(const bf16* __restrict__ x, long long ld, int C, long long V,
                   double* __restrict__ out) {
  extern __shared__ double sh[];  // [C]; double (shared and global): exact, order-independent sums of fp32 partials
  const int CG = C / 8;
  const int rows_per_iter = 256 / CG;
  const int tid = threadIdx.x;
  for (int i = tid; i < C; i += 256) sh[i] = 0.0;
  __syncthreads();
  const int cg = tid % CG;
  const int rl = tid / CG;
  const long long per_block = (V + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * per_block;
  const long long r1 = (r0 + per_block < V) ? (r0 + per_block) : V;
  if (tid < rows_per_iter * CG) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (long long r = r0 + rl; r < r1; r += rows_per_iter) {
      float v[8];
      load8(x + r * ld + cg * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sh[cg * 8 + j], static_cast<double>(s[j]));
  }
  __syncthreads();
  for (int i = tid; i < C; i += 256) atomicAdd(&out[i], sh[i]);
}

// ------------------------------------------------------------------------------ upsample x2
// nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True) (models/mygannet.py:50):
// source index = dst * (in-1)/(out-1), computed in fp32 like ATen's upsample kernels.
__device__ __forceinline__ void src_index(int o, int in_size, int out_size, int& i0, int& i1,
                                          float& l0, float& l1) {
  const float r = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  const float s = r * (float)o;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

struct UpGeom {
  int N, D, H, W, CG;          // low-resolution grid, channel groups of 8
  unsigned total;              // work items (voxels x channel groups) of the kernel's index space
  FastDiv fCG, fX, fY, fZ;     // divisors: CG, then the W / H / D extents of the index space
};

__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const bf16* __restrict__ x, long long x_ld, UpGeom g, bf16* __restrict__ out,
                      long long out_ld) {
  const int D = g.D, H = g.H, W = g.W;
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < g.total; i += gridDim.x * blockDim.x) {
    const unsigned ov = fdiv(i, g.fCG);
    const int cg = i - ov * g.CG;
    unsigned t = fdiv(ov, g.fX);
    const int ow = ov - t * OW;
    unsigned t2 = fdiv(t, g.fY);
    const int oh = t - t2 * OH;
    const unsigned n = fdiv(t2, g.fZ);
    const int od = t2 - n * OD;
    int d0, d1, h0, h1, w0, w1;
    float ld0, ld1, lh0, lh1, lw0, lw1;
    src_index(od, D, OD, d0, d1, ld0, ld1);
    src_index(oh, H, OH, h0, h1, lh0, lh1);
    src_index(ow, W, OW, w0, w1, lw0, lw1);
    const bf16* xb = x + cg * 8;
    uint4 raw[8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int d = a ? d1 : d0, h = b ? h1 : h0, w = c ? w1 : w0;
          raw[(a * 2 + b) * 2 + c] = ldg16(xb + (size_t)(((n * D + d) * H + h) * W + w) * x_ld);
        }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float wt = (a ? ld1 : ld0) * (b ? lh1 : lh0) * (c ? lw1 : lw0);
          float v[8];
          unpack8(raw[(a * 2 + b) * 2 + c], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, v[j], acc[j]);
        }
    store8(out + (size_t)ov * out_ld + cg * 8, acc);
  }
}

// ---- register-blocked forward: one thread = one low-resolution cell (n, jd, jh, jw) x 4 channels. The 2 x 2 x 2
// outputs (2j, 2j+1)^3 of a cell only read the 3 x 3 x 3 inputs (j-1 .. j+1)^3: with align_corners the source
// coordinate of output 2j lies in (j - 1/2, j] and that of 2j+1 in [j, j + 1/2), so along every axis the even output
// interpolates slots (j-1, j) and the odd one slots (j, j+1). The interpolation is done separably in registers
// (W, then H, then D: 76 FMA-class operations per channel for 8 outputs instead of 8 x 8 in the gather form), the
// per-axis weights are computed once per thread with the same fp32 arithmetic as src_index() (bit-identical
// weights), and 27 loads feed 8 stores. Cells whose weights do not follow the two-slot pattern (possible only
// where the fp32 source coordinate rounds across an integer, i.e. at the far border) take the general three-slot
// path. ~13 thread-instructions per output element against ~33 for the gather kernel, which was issue-bound at
// 0.23 of the HBM rate.
__device__ __forceinline__ void up_axis_weights(int j, int L, float (&w)[2][3]) {
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    int i0, i1;
    float l0, l1;
    src_index(2 * j + e, L, 2 * L, i0, i1, l0, l1);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int idx = j - 1 + k;
      w[e][k] = (i0 == idx ? l0 : 0.f) + (i1 == idx ? l1 : 0.f);
    }
  }
}

struct UpCellGeom {
  int N, D, H, W, CG;          // low-resolution grid, channel groups of 4
  unsigned total;              // cells x channel groups
  FastDiv fCG, fW, fH, fD;
};

__device__ __forceinline__ void store4v(bf16* p, const float* v) {
  uint2 u;
  const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// per-axis weight table entry: w[2][3], then 1.0 when the two-slot pattern holds (padded to 8 floats)
__device__ __forceinline__ void up_fill_table(float* tab, int L) {
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float w[2][3];
    up_axis_weights(j, L, w);
    float* e = tab + 8 * j;
    e[0] = w[0][0]; e[1] = w[0][1]; e[2] = w[0][2]; e[3] = w[1][0]; e[4] = w[1][1]; e[5] = w[1][2];
    e[6] = (w[0][2] == 0.f && w[1][0] == 0.f) ? 1.f : 0.f;
    e[7] = 0.f;
  }
}

__global__ void __launch_bounds__(256)
upsample2x_fwd_cell_kernel(const bf16* __restrict__ x, long long x_ld, UpCellGeom g, bf16* __restrict__ out,
                           long long out_ld) {
  extern __shared__ float4 up_tab4[];     // [D + H + W] entries of 8 floats, filled once per block
  float* tab = reinterpret_cast<float*>(up_tab4);
  const int D = g.D, H = g.H, W = g.W;
  const int OH = 2 * H, OW = 2 * W;
  up_fill_table(tab, D);
  up_fill_table(tab + 8 * D, H);
  up_fill_table(tab + 8 * (D + H), W);
  __syncthreads();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < g.total; i += gridDim.x * blockDim.x) {
    const unsigned cell = fdiv(i, g.fCG);
    const int cg = i - cell * g.CG;
    unsigned t = fdiv(cell, g.fW);
    const int jw = cell - t * W;
    unsigned t2 = fdiv(t, g.fH);
    const int jh = t - t2 * H;
    const unsigned n = fdiv(t2, g.fD);
    const int jd = t2 - n * D;
    float wd[2][3], wh[2][3], ww[2][3];
    float ok = 1.f;
    {
      const float4 a0 = up_tab4[2 * jd], a1 = up_tab4[2 * jd + 1];
      wd[0][0] = a0.x; wd[0][1] = a0.y; wd[0][2] = a0.z; wd[1][0] = a0.w; wd[1][1] = a1.x; wd[1][2] = a1.y;
      const float4 b0 = up_tab4[2 * (D + jh)], b1 = up_tab4[2 * (D + jh) + 1];
      wh[0][0] = b0.x; wh[0][1] = b0.y; wh[0][2] = b0.z; wh[1][0] = b0.w; wh[1][1] = b1.x; wh[1][2] = b1.y;
      const float4 c0 = up_tab4[2 * (D + H + jw)], c1 = up_tab4[2 * (D + H + jw) + 1];
      ww[0][0] = c0.x; ww[0][1] = c0.y; ww[0][2] = c0.z; ww[1][0] = c0.w; ww[1][1] = c1.x; ww[1][2] = c1.y;
      ok = a1.z * b1.z * c1.z;
    }
    const bool fast = ok != 0.f;
    const int dz[3] = {max(jd - 1, 0), jd, min(jd + 1, D - 1)};
    const int hy[3] = {max(jh - 1, 0), jh, min(jh + 1, H - 1)};
    const int wx[3] = {max(jw - 1, 0), jw, min(jw + 1, W - 1)};
    const bf16* xb = x + cg * 4;
    bf16* ob = out + cg * 4;
    const unsigned obase = ((n * 2 * D + 2 * jd) * OH + 2 * jh) * OW + 2 * jw;   // output voxel (2jd, 2jh, 2jw)
    if (!fast) {
      // general three-slot weights (rare: far-border cells only): plain gather of the eight outputs
#pragma unroll 1
      for (int e = 0; e < 8; ++e) {
        const int ed = e >> 2, eh = (e >> 1) & 1, ew = e & 1;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int a = 0; a < 3; ++a)
#pragma unroll 1
          for (int b = 0; b < 3; ++b)
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
              const float wt = wd[ed][a] * wh[eh][b] * ww[ew][c];
              if (wt == 0.f) continue;
              float v[4];
              const int sd = min(max(jd - 1 + a, 0), D - 1), sh = min(max(jh - 1 + b, 0), H - 1);
              const int sw = min(max(jw - 1 + c, 0), W - 1);
              unpack4(ldg8(xb + (size_t)(((n * D + sd) * H + sh) * W + sw) * x_ld), v);
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = fmaf(wt, v[j], o[j]);
            }
        store4v(ob + (size_t)(obase + (ed * OH + eh) * OW + ew) * out_ld, o);
      }
      continue;
    }
    float prev[2][2][4];   // H/W-interpolated plane of the previous input depth slot
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      uint2 raw[3][3];
      const unsigned pbase = (n * D + dz[a]) * H;
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          raw[b][c] = ldg8(xb + (size_t)((pbase + hy[b]) * W + wx[c]) * x_ld);
      float r[3][2][4];    // W-interpolated rows: even outputs read slots (0, 1), odd ones slots (1, 2)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float v0[4], v1[4], v2[4];
        unpack4(raw[b][0], v0);
        unpack4(raw[b][1], v1);
        unpack4(raw[b][2], v2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r[b][0][j] = fmaf(ww[0][0], v0[j], ww[0][1] * v1[j]);
          r[b][1][j] = fmaf(ww[1][1], v1[j], ww[1][2] * v2[j]);
        }
      }
      float cur[2][2][4];  // [h parity][w parity]
#pragma unroll
      for (int ew = 0; ew < 2; ++ew)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cur[0][ew][j] = fmaf(wh[0][0], r[0][ew][j], wh[0][1] * r[1][ew][j]);
          cur[1][ew][j] = fmaf(wh[1][1], r[1][ew][j], wh[1][2] * r[2][ew][j]);
        }
      if (a >= 1) {        // output depth 2jd + (a - 1) interpolates depth slots (a - 1, a)
        const int e = a - 1;
#pragma unroll
        for (int eh = 0; eh < 2; ++eh)
#pragma unroll
          for (int ew = 0; ew < 2; ++ew) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaf(wd[e][e], prev[eh][ew][j], wd[e][e + 1] * cur[eh][ew][j]);
            store4v(ob + (size_t)(obase + (e * OH + eh) * OW + ew) * out_ld, o);
          }
      }
#pragma unroll
      for (int eh = 0; eh < 2; ++eh)
#pragma unroll
        for (int ew = 0; ew < 2; ++ew)
#pragma unroll
          for (int j = 0; j < 4; ++j) prev[eh][ew][j] = cur[eh][ew][j];
    }
  }
}

// weights with which input index i contributes to outputs o in [2i-2, 2i+4] along one axis
__device__ __forceinline__ void bwd_weights(int i, int in_size, float* wts) {
  const int out_size = 2 * in_size;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int o = 2 * i - 2 + k;
    float wt = 0.f;
    if (o >= 0 && o < out_size) {
      int i0, i1;
      float l0, l1;
      src_index(o, in_size, out_size, i0, i1, l0, l1);
      if (i0 == i) wt += l0;
      if (i1 == i) wt += l1;
    }
    wts[k] = wt;
  }
}

// gradient of the x2 trilinear upsample w.r.t. its low-resolution input (gather form)
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const bf16* __restrict__ go, long long go_ld, UpGeom g, bf16* __restrict__ gx,
                      long long gx_ld) {
  const int D = g.D, H = g.H, W = g.W;
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < g.total; i += gridDim.x * blockDim.x) {
    const unsigned iv = fdiv(i, g.fCG);
    const int cg = i - iv * g.CG;
    unsigned t = fdiv(iv, g.fX);
    const int w = iv - t * W;
    unsigned t2 = fdiv(t, g.fY);
    const int h = t - t2 * H;
    const unsigned n = fdiv(t2, g.fZ);
    const int d = t2 - n * D;
    float wd[7], wh[7], ww[7];
    bwd_weights(d, D, wd);
    bwd_weights(h, H, wh);
    bwd_weights(w, W, ww);
    const bf16* gb = go + cg * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int a = 0; a < 7; ++a) {
      if (wd[a] == 0.f) continue;
      const int od = 2 * d - 2 + a;
      for (int b = 0; b < 7; ++b) {
        if (wh[b] == 0.f) continue;
        const int oh = 2 * h - 2 + b;
        const unsigned rowv = ((n * OD + od) * OH + oh) * OW + (2 * w - 2);
        // at most four of the seven w-taps are non-zero: issue their loads together
        uint4 raw[7];
#pragma unroll
        for (int c = 0; c < 7; ++c)
          if (ww[c] != 0.f) raw[c] = ldg16(gb + (size_t)(rowv + c) * go_ld);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          if (ww[c] == 0.f) continue;
          const float wt = wd[a] * wh[b] * ww[c];
          float v[8];
          unpack8(raw[c], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt, v[j], acc[j]);
        }
      }
    }
    store8(gx + (size_t)iv * gx_ld + cg * 8, acc);
  }
}


// One axis of the adjoint of the separable x2 trilinear upsample: src [outer][2L][inner][C] -> dst [outer][L][inner][C],
// dst[l] = sum_k wt_k(l) * src[2l - 2 + k]. Three such passes (W, H, D) read 1 + 1/2 + 1/4 of the gradient
// tensor instead of gathering a 5 x 5 x 5 neighbourhood per low-resolution voxel.
struct UpAxisGeom {
  unsigned total;     // outer * L * inner * CG
  int L, inner, CG;
  FastDiv fCG, fInner, fL;
};

__global__ void __launch_bounds__(256)
upsample_bwd_axis_kernel(const bf16* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long dst_ld,
                         UpAxisGeom g) {
  const int L = g.L, OL = 2 * g.L;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < g.total; i += gridDim.x * blockDim.x) {
    const unsigned r = fdiv(i, g.fCG);
    const int cg = i - r * g.CG;
    const unsigned r2 = fdiv(r, g.fInner);
    const int in = r - r2 * g.inner;
    const unsigned outer = fdiv(r2, g.fL);
    const int l = r2 - outer * L;
    float wts[7];
    bwd_weights(l, L, wts);
    uint4 raw[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int o = 2 * l - 2 + k;
      if (wts[k] != 0.f) raw[k] = ldg16(src + (size_t)((outer * OL + o) * g.inner + in) * src_ld + cg * 8);
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      if (wts[k] == 0.f) continue;
      float v[8];
      unpack8(raw[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(wts[k], v[j], acc[j]);
    }
    store8(dst + (size_t)r * dst_ld + cg * 8, acc);
  }
}

// ---- lean adjoint of one axis: dst[l] = sum_{t<4} wt[l][t] * src[2l - 1 + t] (the outputs 2l-1 .. 2l+2 are the only
// ones that read input l, see above); the per-position weights are tabulated once per block in shared memory with
// the arithmetic of src_index(). ~100 thread-instructions per 8 channels against ~380 for the 7-tap kernel.
__global__ void __launch_bounds__(256)
upsample_bwd_axis4_kernel(const bf16* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long dst_ld,
                          UpAxisGeom g) {
  extern __shared__ float wtab[];   // [L][4]
  const int L = g.L, OL = 2 * g.L;
  for (int e = threadIdx.x; e < 4 * L; e += blockDim.x) {
    const int l = e >> 2, o = 2 * l - 1 + (e & 3);
    float wt = 0.f;
    if (o >= 0 && o < OL) {
      int i0, i1;
      float l0, l1;
      src_index(o, L, OL, i0, i1, l0, l1);
      wt = (i0 == l ? l0 : 0.f) + (i1 == l ? l1 : 0.f);
    }
    wtab[e] = wt;
  }
  __syncthreads();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < g.total; i += gridDim.x * blockDim.x) {
    const unsigned r = fdiv(i, g.fCG);
    const int cg = i - r * g.CG;
    const unsigned r2 = fdiv(r, g.fInner);
    const int in = r - r2 * g.inner;
    const unsigned outer = fdiv(r2, g.fL);
    const int l = r2 - outer * L;
    const float4 wt = *reinterpret_cast<const float4*>(wtab + 4 * l);
    const float w4[4] = {wt.x, wt.y, wt.z, wt.w};
    // tap t reads output position 2l - 1 + t; addressed relative to tap 1 (always inside the tensor)
    const bf16* sp = src + (size_t)((outer * OL + 2 * l) * g.inner + in) * src_ld + cg * 8;
    const ptrdiff_t step = (ptrdiff_t)g.inner * (ptrdiff_t)src_ld;
    uint4 raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (w4[k] != 0.f) raw[k] = ldg16(sp + (k - 1) * step);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (w4[k] == 0.f) continue;
      float v[8];
      unpack8(raw[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(w4[k], v[j], acc[j]);
    }
    store8(dst + (size_t)r * dst_ld + cg * 8, acc);
  }
}

// ------------------------------------------------------------------------------ heads / losses
// predict = sigmoid(logit[:,0]) from the fp32 conv_last output [V][ld]
__global__ void sigmoid_head_fwd_kernel(const float* __restrict__ logits, long long ld, long long V,
                                        float* __restrict__ predict) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x)
    predict[i] = 1.f / (1.f + expf(-logits[i * ld]));
}
// dlogit (bf16 channels-last, 8 channels, only channel 0 non-zero) = g * p * (1 - p)
__global__ void sigmoid_head_bwd_kernel(const float* __restrict__ gpred,
                                        const float* __restrict__ predict, long long V,
                                        bf16* __restrict__ dlogit) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x) {
    const float p = predict[i];
    float v[8] = {gpred[i] * p * (1.f - p), 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8(dlogit + i * 8, v);
  }
}

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;
}

// weighted_bce (lib/utils.py:65-71): p clamped to [1e-8, 1-1e-8] in fp32 (the upper bound rounds
// to 1.0f exactly like the reference), loss = -mean(t*log p + pos_weight*(1-t)*log(1-p)).
// Also emits d loss / d predict scaled by grad_scale (0 where the clamp is active).
__global__ void __launch_bounds__(256)
wbce_kernel(const float* __restrict__ predict, const float* __restrict__ target, long long V,
            float pos_weight, float grad_scale, double* __restrict__ loss_sum,
            float* __restrict__ gpred) {
  double local = 0.0;
  const float lo = 1e-8f, hi = 1.0f - 1e-8f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V;
       i += (long long)gridDim.x * blockDim.x) {
    const float p0 = predict[i], t = target[i];
    const float p = fminf(fmaxf(p0, lo), hi);
    const float l = t * logf(p) + pos_weight * (1.f - t) * logf(1.f - p);
    local += (double)l;
    if (gpred != nullptr) {
      const bool pass = (p0 >= lo) && (p0 <= hi);
      const float gl = t / p - pos_weight * (1.f - t) / (1.f - p);
      gpred[i] = pass ? (-grad_scale * gl) : 0.f;
    }
  }
  local = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(loss_sum, local);
}

// l2_loss numerator (lib/utils.py:59-63) between two channels-last bf16 tensors
__global__ void __launch_bounds__(256)
sqdiff_kernel(const bf16* __restrict__ a, long long a_ld, const bf16* __restrict__ b, long long b_ld,
              int C, long long V, double* __restrict__ out) {
  const int CG = C / 8;
  const long long total = V * CG;
  double local = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const long long r = i / CG;
    float x[8], y[8];
    load8(a + r * a_ld + cg * 8, x);
    load8(b + r * b_ld + cg * 8, y);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = x[j] - y[j];
      s = fmaf(d, d, s);
    }
    local += (double)s;
  }
  local = block_sum(local);
  if (threadIdx.x == 0) atomicAdd(out, local);
}

// ------------------------------------------------------------------------------ ConvLSTM cell
// gates: fp32 channels-last [V][4*hid] in the reference's split order i,f,o,g
// (models/convlstm.py:49-58); c, h: fp32 [V][hid].
__global__ void convlstm_cell_fwd_kernel(const float* __restrict__ gates, long long g_ld,
                                         const float* __restrict__ c_cur, int hid, long long V,
                                         float* __restrict__ h_next, float* __restrict__ c_next,
                                         float* __restrict__ act) {
  const long long total = V * hid;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % hid);
    const long long v = i / hid;
    const float* gp = gates + v * g_ld;
    const float gi = 1.f / (1.f + expf(-gp[k]));
    const float gf = 1.f / (1.f + expf(-gp[hid + k]));
    const float go = 1.f / (1.f + expf(-gp[2 * hid + k]));
    const float gg = tanhf(gp[3 * hid + k]);
    const float c = gf * c_cur[i] + gi * gg;
    c_next[i] = c;
    h_next[i] = go * tanhf(c);
    if (act != nullptr) {  // saved activations for backward: [V][4*hid]
      float* ap = act + v * 4 * hid;
      ap[k] = gi;
      ap[hid + k] = gf;
      ap[2 * hid + k] = go;
      ap[3 * hid + k] = gg;
    }
  }
}

// backward of the cell update: given dh_next, dc_next -> dgates (bf16 channels-last, for the gate
// conv's dgrad/wgrad) and dc_cur.
__global__ void convlstm_cell_bwd_kernel(const float* __restrict__ act, const float* __restrict__ c_cur,
                                         const float* __restrict__ c_next,
                                         const float* __restrict__ dh, const float* __restrict__ dc_in,
                                         int hid, long long V, bf16* __restrict__ dgates,
                                         long long dg_ld, float* __restrict__ dc_cur) {
  const long long total = V * hid;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % hid);
    const long long v = i / hid;
    const float* ap = act + v * 4 * hid;
    const float gi = ap[k], gf = ap[hid + k], go = ap[2 * hid + k], gg = ap[3 * hid + k];
    const float tc = tanhf(c_next[i]);
    const float dhn = dh != nullptr ? dh[i] : 0.f;
    const float dcn = (dc_in != nullptr ? dc_in[i] : 0.f) + dhn * go * (1.f - tc * tc);
    bf16* dg = dgates + v * dg_ld;
    dg[k] = __float2bfloat16(dcn * gg * gi * (1.f - gi));
    dg[hid + k] = __float2bfloat16(dcn * c_cur[i] * gf * (1.f - gf));
    dg[2 * hid + k] = __float2bfloat16(dhn * tc * go * (1.f - go));
    dg[3 * hid + k] = __float2bfloat16(dcn * gi * (1.f - gg * gg));
    dc_cur[i] = dcn * gf;
  }
}


// ------------------------------------------------------------------------------ tap gather (im2col into channels)
// dst[v][t*CS + c] = src[v + sign*off(t)][c] for the KD*KH*KW taps t (offsets centred, zero outside the
// volume) and c < CS; the remaining destination columns are zero. Folding the taps of a thin conv into
// the channel dimension turns its weight gradient into a 1x1x1 wgrad with a dense K = taps*CS (<= 32)
// instead of taps x 8 nearly empty 128 x 16 x 16 MMAs per 128 voxels. One thread per voxel; lanes are
// consecutive voxels, so the 16-byte source loads of a warp are contiguous.
struct GatherGeom {
  int N, D, H, W;
  unsigned V;
  FastDiv fW, fH, fD;
};

template <int KD, int KH, int KW, int CS>
__global__ void __launch_bounds__(256)
tap_gather_kernel(const bf16* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long dst_ld,
                  GatherGeom g, int sign) {
  constexpr int TAPS = KD * KH * KW;
  constexpr int COLS = (TAPS * CS + 7) & ~7;
  static_assert(COLS <= 32 && CS <= 4, "tap gather folds at most 32 columns");
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < g.V; v += gridDim.x * blockDim.x) {
    unsigned t = fdiv(v, g.fW);
    const int w = v - t * g.W;
    unsigned t2 = fdiv(t, g.fH);
    const int h = t - t2 * g.H;
    const unsigned n = fdiv(t2, g.fD);
    const int d = t2 - n * g.D;
    uint32_t raw[TAPS][2];
#pragma unroll
    for (int a = 0; a < KD; ++a)
#pragma unroll
      for (int b = 0; b < KH; ++b)
#pragma unroll
        for (int c = 0; c < KW; ++c) {
          const int tp = (a * KH + b) * KW + c;
          const int dd = d + sign * (a - KD / 2), hh = h + sign * (b - KH / 2), ww = w + sign * (c - KW / 2);
          const bool ok = dd >= 0 && dd < g.D && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
          uint2 r = make_uint2(0u, 0u);
          if (ok)
            r = __ldg(reinterpret_cast<const uint2*>(src + (size_t)(((n * g.D + dd) * g.H + hh) * g.W + ww) * src_ld));
          raw[tp][0] = r.x;
          raw[tp][1] = r.y;
        }
    // element e of the destination row = channel (e % CS) of tap (e / CS)
    uint32_t outw[COLS / 2];
#pragma unroll
    for (int e2 = 0; e2 < COLS / 2; ++e2) {
      uint32_t word = 0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int e = 2 * e2 + half;
        if (e < TAPS * CS) {
          const int tp = e / CS, ch = e % CS;
          const uint32_t val = (raw[tp][ch >> 1] >> ((ch & 1) * 16)) & 0xFFFFu;
          word |= val << (half * 16);
        }
      }
      outw[e2] = word;
    }
    uint4* o = reinterpret_cast<uint4*>(dst + (size_t)v * dst_ld);
#pragma unroll
    for (int i = 0; i < COLS / 8; ++i) o[i] = make_uint4(outw[4 * i], outw[4 * i + 1], outw[4 * i + 2], outw[4 * i + 3]);
  }
}

static inline int grid_for(long long total, int block = 256, int max_blocks = 148 * 16) {
  long long b = (total + block - 1) / block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

static inline int check_cl(const void* p, long long ld, int C, const char* what) {
  if ((reinterpret_cast<uintptr_t>(p) & 15) || (ld % 8) || (C % 8) || C <= 0 || C > 2048) return set_error(VFD_ERR_ARG, what);
  return 0;
}

}  // namespace vfd

using namespace vfd;
#define STREAM reinterpret_cast<cudaStream_t>(stream_)

VFD_API int vfd_pack_ncdhw(const float* src, void* dst, int N, int Csrc, long long S, int C,
                              long long ld, int Cp, int replicate, void* stream_) {
  if (int e = check_cl(dst, ld, Cp, "pack_ncdhw: destination must be 16-byte aligned, C%8==0")) return e;
  const long long total = (long long)N * S * (Cp / 8);
  if (total == 0) return 0;
  pack_ncdhw_kernel<<<grid_for(total), 256, 0, STREAM>>>(src, (bf16*)dst, N, Csrc, S, C, ld, Cp, replicate);
  return check_launch("pack_ncdhw");
}

VFD_API int vfd_unpack_ncdhw(const void* src, int src_fp32, float* dst, int N, int C, long long S,
                                long long ld, void* stream_) {
  const long long total = (long long)N * C * S;
  if (total == 0) return 0;
  if (src_fp32)
    unpack_ncdhw_kernel<float><<<grid_for(total), 256, 0, STREAM>>>((const float*)src, dst, N, C, S, ld);
  else
    unpack_ncdhw_kernel<bf16><<<grid_for(total), 256, 0, STREAM>>>((const bf16*)src, dst, N, C, S, ld);
  return check_launch("unpack_ncdhw");
}

VFD_API int vfd_pack_weight(const float* w, void* wp, int Cout, int Cin, int taps, int rows, int ck,
                               int mode, void* stream_) {
  const long long total = (long long)rows * taps * ck;
  if (total == 0) return 0;
  pack_weight_kernel<<<grid_for(total), 256, 0, STREAM>>>(w, (bf16*)wp, Cout, Cin, taps, rows, ck, mode);
  return check_launch("pack_weight");
}

VFD_API int vfd_pack_weights_batched(const void* jobs, int njobs, long long total, void* stream_) {
  if (njobs <= 0 || total <= 0) return 0;
  const long long blocks = (total + kPackPerBlock - 1) / kPackPerBlock;
  if (blocks >= (1LL << 31)) return set_error(VFD_ERR_ARG, "pack_weights_batched: too many elements");
  pack_weights_batched_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>((const PackJob*)jobs, njobs, total);
  return check_launch("pack_weights_batched");
}

VFD_API int vfd_unpack_wgrad(const float* acc, float* gw, int Cout, int Cin, int taps, int co_pad,
                                int ci_pad, int accumulate, void* stream_) {
  const long long total = (long long)Cout * Cin * taps;
  if (total == 0) return 0;
  unpack_wgrad_kernel<<<grid_for(total), 256, 0, STREAM>>>(acc, gw, Cout, Cin, taps, co_pad, ci_pad, accumulate);
  return check_launch("unpack_wgrad");
}

static int stats_grid(long long V, int C) {
  const int rows_per_iter = 256 / (C / 8);
  long long b = (V + (long long)rows_per_iter * 8 - 1) / ((long long)rows_per_iter * 8);
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

VFD_API int vfd_bn_stats(const void* x, long long ld, int C, long long V, double* sums, void* stream_) {
  if (int e = check_cl(x, ld, C, "bn_stats: bad tensor")) return e;
  if (V <= 0) return 0;
  bn_stats_kernel<<<stats_grid(V, C), 256, 2 * C * sizeof(double), STREAM>>>((const bf16*)x, ld, C, V, sums);
  return check_launch("bn_stats");
}

VFD_API int vfd_bn_finalize(double* sums, int C, int Cvalid, long long V, const float* pre_bias,
                               const float* gamma, const float* beta, float* running_mean, float* running_var,
                               float momentum, float eps, int train, float* mean, float* invstd,
                               float* scale, float* shift, void* stream_) {
  if (!train && (running_mean == nullptr || running_var == nullptr))
    return set_error(VFD_ERR_ARG, "bn_finalize: eval mode needs running statistics");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums, C, Cvalid, V, pre_bias, gamma, beta, running_mean,
                                                          running_var, momentum, eps, train, mean,
                                                          invstd, scale, shift);
  return check_launch("bn_finalize");
}

static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) {
    f.m = 0;
    f.s = 0;
  } else {
    uint32_t sft = 0;
    while ((1u << sft) < d) ++sft;  // ceil(log2 d)
    f.m = (uint32_t)((((1ull << sft) - d) << 32) / d + 1);
    f.s = sft;
  }
  return f;
}

static int fill_act_geom(ActGeom& g, int N, int D, int H, int W, int C, int pd, int ph, int pw) {
  if ((pd != 1 && pd != 2) || (ph != 1 && ph != 2) || (pw != 1 && pw != 2))
    return set_error(VFD_ERR_ARG, "pool window must be 1 or 2 per axis");
  g.N = N; g.D = D; g.H = H; g.W = W; g.C = C; g.pd = pd; g.ph = ph; g.pw = pw;
  g.WD = (D + pd - 1) / pd; g.WH = (H + ph - 1) / ph; g.WW = (W + pw - 1) / pw;
  g.QD = D / pd; g.QH = H / ph; g.QW = W / pw;
  g.nwin = (unsigned)((long long)N * g.WD * g.WH * g.WW);
  g.fWW = make_fastdiv(g.WW); g.fWH = make_fastdiv(g.WH); g.fWD = make_fastdiv(g.WD);
  g.pool_bcast = 0;
  g.gpd = g.gph = g.gpw = 0;
  g.inv_win = 1.0f / (float)(pd * ph * pw);
  return 0;
}

static int bn_waves() {
  static int w = 0;
  if (w == 0) {
    const char* e = getenv("VFD_BN_WAVES");
    w = e ? atoi(e) : 16;
    if (w < 1) w = 16;
  }
  return w;
}

static int win_grid(const ActGeom& g, int pd, int ph, int pw, int nv_per_iter, int waves = 0) {
  const long long nwin = (long long)g.N * ((g.D + pd - 1) / pd) * ((g.H + ph - 1) / ph) * ((g.W + pw - 1) / pw);
  const int rpi = 256 / (g.C / 8);
  long long b = (nwin + (long long)rpi * nv_per_iter * 4 - 1) / ((long long)rpi * nv_per_iter * 4);
  if (b < 1) b = 1;
  // blocks per SM in the grid (2 are resident). Measured on B200: the forward kernels like many small
  // blocks (16), the 8-channel backward kernels exactly one resident wave (2); VFD_BN_WAVES overrides both
  const int cap = 148 * (getenv("VFD_BN_WAVES") ? bn_waves() : (waves ? waves : 16));
  if (b > cap) b = cap;
  return (int)b;
}

#define VFD_POOL_DISPATCH(KERNEL, DROPFLAG, LAUNCH)                                               \
  do {                                                                                            \
    if (pd == 1 && ph == 1 && pw == 1) {                                                          \
      if (DROPFLAG) { auto kfn = KERNEL<1, 1, 1, true>; LAUNCH; }                                 \
      else { auto kfn = KERNEL<1, 1, 1, false>; LAUNCH; }                                         \
    } else if (DROPFLAG) {                                                                        \
      return set_error(VFD_ERR_ARG, "dropout is only fused into un-pooled BN+activation");        \
    } else if (pd == 2 && ph == 2 && pw == 2) { auto kfn = KERNEL<2, 2, 2, false>; LAUNCH; }      \
    else if (pd == 1 && ph == 2 && pw == 2) { auto kfn = KERNEL<1, 2, 2, false>; LAUNCH; }        \
    else if (pd == 2 && ph == 1 && pw == 1) { auto kfn = KERNEL<2, 1, 1, false>; LAUNCH; }        \
    else return set_error(VFD_ERR_ARG, "pool window must be (1,1,1), (2,2,2), (1,2,2) or (2,1,1)"); \
  } while (0)

VFD_API int vfd_bn_act_fwd(const void* y, long long y_ld, int N, int D, int H, int W, int C,
                              const float* scale, const float* shift, float slope, void* out_full,
                              long long full_ld, void* out_pool, long long pool_ld, int pd, int ph,
                              int pw, float drop_p, unsigned long long seed,
                              const unsigned long long* seed_dev, void* stream_) {
  if (int e = check_cl(y, y_ld, C, "bn_act_fwd: bad input")) return e;
  if (out_full && check_cl(out_full, full_ld, C, "bn_act_fwd: bad full output")) return VFD_ERR_ARG;
  if (out_pool && check_cl(out_pool, pool_ld, C, "bn_act_fwd: bad pooled output")) return VFD_ERR_ARG;
  ActGeom g;
  if (int e = fill_act_geom(g, N, D, H, W, C, pd, ph, pw)) return e;
  if ((long long)N * D * H * W >= (1LL << 31)) return set_error(VFD_ERR_ARG, "bn_act_fwd: more than 2^31 voxels");
  if ((long long)N * D * H * W == 0) return 0;
  if (!(slope >= 0.f && slope <= 1.f)) return set_error(VFD_ERR_ARG, "bn_act_fwd: slope must be in [0, 1]");
  const bool drop = drop_p > 0.f;
  const int grid = win_grid(g, pd, ph, pw, 8 / (pd * ph * pw) > 0 ? 8 / (pd * ph * pw) : 1);
  VFD_POOL_DISPATCH(bn_act_fwd_kernel, drop,
                    (kfn<<<grid, 256, 0, STREAM>>>((const bf16*)y, y_ld, g, scale, shift, slope, (bf16*)out_full,
                                                   full_ld, (bf16*)out_pool, pool_ld, drop_p, seed, seed_dev)));
  return check_launch("bn_act_fwd");
}

VFD_API int vfd_bn_act_bwd(const void* y, long long y_ld, int N, int D, int H, int W, int C, int Cvalid,
                              const float* mean, const float* invstd, const float* scale,
                              const float* shift, float slope, const void* g_full, long long gf_ld,
                              const void* g_pool, long long gp_ld, int pd, int ph, int pw, float drop_p,
                              unsigned long long seed, const unsigned long long* seed_dev, int train,
                              double* sums, float* c1, float* c2, float* dgamma, float* dbeta, void* dy,
                              long long dy_ld, unsigned int* ticket, void* stream_) {
  if (int e = check_cl(y, y_ld, C, "bn_act_bwd: bad input")) return e;
  if (int e = check_cl(dy, dy_ld, C, "bn_act_bwd: bad output")) return e;
  if (g_full && check_cl(g_full, gf_ld, C, "bn_act_bwd: bad full-resolution gradient")) return VFD_ERR_ARG;
  if (g_pool && check_cl(g_pool, gp_ld, C, "bn_act_bwd: bad pooled gradient")) return VFD_ERR_ARG;
  if ((pd != 1 && pd != 2) || (ph != 1 && ph != 2) || ph != pw)
    return set_error(VFD_ERR_ARG, "bn_act_bwd: pool window must be (1,1,1), (2,2,2), (1,2,2) or (2,1,1)");
  if (C > 1024) return set_error(VFD_ERR_ARG, "bn_act_bwd: at most 1024 channels");
  const long long V = (long long)N * D * H * W;
  if (V >= (1LL << 31)) return set_error(VFD_ERR_ARG, "bn_act_bwd: more than 2^31 voxels");
  if (V == 0) return 0;
  const bool drop = drop_p > 0.f;
  if (drop && (pd != 1 || ph != 1)) return set_error(VFD_ERR_ARG, "dropout is only fused into un-pooled BN+activation");
  const int pool_bcast = (train >> 1) & 1;   // bit 1: g_pool is [N][D/pd][C], broadcast over H and W
  const int accumulate = (train >> 2) & 1;   // bit 2: add to dgamma / dbeta instead of overwriting them
  train &= 1;
  if (pool_bcast && (ph != 1 || g_pool == nullptr))
    return set_error(VFD_ERR_ARG, "bn_act_bwd: a broadcast pooled gradient needs a (1,1,1) or (2,1,1) window");
  static const int force4 = getenv("VFD_BN_BWD4") ? atoi(getenv("VFD_BN_BWD4")) : 0;
  static const int gather = getenv("VFD_BN_GATHER") ? atoi(getenv("VFD_BN_GATHER")) : 0;
  const bool use_gather = gather && pd * ph * pw > 1 && !drop;
  // (1,2,2) windows (SDisc) also take the 8-channel kernels: 0.435 -> 0.347 ms on the first SDisc block, step -0.5 ms
  // (profiles/r2_negative_overlap_and_bn_mapping.txt); (2,2,2) windows measured equal and stay on the 4-channel ones
  // (8 voxels x 16 bytes per thread spill). VFD_BN_BWD8_POOL = 0 / 2 forces neither / both.
  static const int pool8 = getenv("VFD_BN_BWD8_POOL") ? atoi(getenv("VFD_BN_BWD8_POOL")) : 1;
  const bool pooled8 = ph == 2 && !drop && !pool_bcast && ((pool8 == 1 && pd == 1) || pool8 == 2);
  BwdFinalize fin;
  fin.ticket = ticket; fin.Cvalid = Cvalid; fin.train = train; fin.accumulate = accumulate; fin.V = V;
  fin.dgamma = dgamma; fin.dbeta = dbeta; fin.c1 = c1; fin.c2 = c2;
  if ((ph == 1 && !force4) || use_gather || pooled8) {
    // un-pooled / depth-pooled: 8-channel window-per-thread kernels
    ActGeom g;
    if (use_gather) {
      if (int e = fill_act_geom(g, N, D, H, W, C, 1, 1, 1)) return e;
      g.gpd = pd; g.gph = ph; g.gpw = pw;
      g.QD = D / pd; g.QH = H / ph; g.QW = W / pw;
      g.inv_win = 1.0f / (float)(pd * ph * pw);
      pd = ph = pw = 1;
    } else if (int e = fill_act_geom(g, N, D, H, W, C, pd, ph, pw)) return e;
    g.pool_bcast = pool_bcast;
    const int nvi = 4 / (pd * ph * pw) > 0 ? 4 / (pd * ph * pw) : 1;
    const int grid = win_grid(g, pd, ph, pw, nvi, 2);
#define VFD_BWD8_LAUNCH(KERNEL, SMEM, ...)                                                             \
    do {                                                                                                \
      if (ph == 2 && pd == 1) {                                                                         \
        KERNEL<1, 2, 2, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                               \
      } else if (ph == 2) {                                                                             \
        KERNEL<2, 2, 2, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                               \
      } else if (pd == 1) {                                                                             \
        if (drop) KERNEL<1, 1, 1, true><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                      \
        else KERNEL<1, 1, 1, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                          \
      } else {                                                                                          \
        KERNEL<2, 1, 1, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                               \
      }                                                                                                 \
    } while (0)
    VFD_BWD8_LAUNCH(bn_act_bwd8_reduce_kernel, 2 * C * sizeof(float), (const bf16*)y, y_ld, g, mean, invstd, scale,
                    shift, slope, (const bf16*)g_full, gf_ld, (const bf16*)g_pool, gp_ld, drop_p, seed, seed_dev, sums, fin);
    if (int e = check_launch("bn_act_bwd_reduce")) return e;
    if (ticket == nullptr) {
      bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums, C, Cvalid, V, train, accumulate, dgamma, dbeta, c1, c2);
      if (int e = check_launch("bn_bwd_finalize")) return e;
    }
    VFD_BWD8_LAUNCH(bn_act_bwd8_apply_kernel, 0, (const bf16*)y, y_ld, g, mean, invstd, scale, shift, slope,
                    (const bf16*)g_full, gf_ld, (const bf16*)g_pool, gp_ld, drop_p, seed, seed_dev, c1, c2, (bf16*)dy,
                    dy_ld);
#undef VFD_BWD8_LAUNCH
    return check_launch("bn_act_bwd_apply");
  }
  BwdGeom g;
  g.N = N; g.D = D; g.H = H; g.W = W; g.C = C; g.pd = pd;
  g.WH = (H + ph - 1) / ph; g.WW = (W + pw - 1) / pw;
  g.QD = D / pd; g.QH = H / ph; g.QW = W / pw;
  g.nwin = (unsigned)((long long)N * D * g.WH * g.WW);
  g.fWW = make_fastdiv(g.WW); g.fWH = make_fastdiv(g.WH); g.fD = make_fastdiv(D);
  g.inv_win = 1.0f / (float)(pd * ph * pw);
  g.pool_bcast = pool_bcast;
  {
    const int rpi = 256 / (C / 4);
    const int per_iter = rpi * (8 / (ph * pw));
    long long b = ((long long)g.nwin + (long long)per_iter * 4 - 1) / ((long long)per_iter * 4);
    if (b < 1) b = 1;
    if (b > 148 * 6) b = 148 * 6;   // 3 resident blocks per SM, two waves
    const int grid = (int)b;
#define VFD_BWD_LAUNCH(KERNEL, SMEM, ...)                                                              \
    do {                                                                                                \
      if (ph == 1) {                                                                                    \
        if (drop) KERNEL<1, 1, true><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                         \
        else KERNEL<1, 1, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                             \
      } else {                                                                                          \
        KERNEL<2, 2, false><<<grid, 256, SMEM, STREAM>>>(__VA_ARGS__);                                  \
      }                                                                                                 \
    } while (0)
    VFD_BWD_LAUNCH(bn_act_bwd_reduce_kernel, 2 * C * sizeof(float), (const bf16*)y, y_ld, g, mean, invstd, scale,
                   shift, slope, (const bf16*)g_full, gf_ld, (const bf16*)g_pool, gp_ld, drop_p, seed, seed_dev, sums, fin);
    if (int e = check_launch("bn_act_bwd_reduce")) return e;
    if (ticket == nullptr) {
      bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(sums, C, Cvalid, V, train, accumulate, dgamma, dbeta, c1, c2);
      if (int e = check_launch("bn_bwd_finalize")) return e;
    }
    VFD_BWD_LAUNCH(bn_act_bwd_apply_kernel, 0, (const bf16*)y, y_ld, g, mean, invstd, scale, shift, slope,
                   (const bf16*)g_full, gf_ld, (const bf16*)g_pool, gp_ld, drop_p, seed, seed_dev, c1, c2, (bf16*)dy,
                   dy_ld);
#undef VFD_BWD_LAUNCH
  }
  return check_launch("bn_act_bwd_apply");
}

VFD_API int vfd_tap_gather(const void* src, long long src_ld, int cs, void* dst, long long dst_ld, int dst_cols,
                              int N, int D, int H, int W, int kd, int kh, int kw, int sign, void* stream_) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (src_ld % 8) || (reinterpret_cast<uintptr_t>(dst) & 15) || (dst_ld % 8))
    return set_error(VFD_ERR_ARG, "tap_gather: tensors must be 16-byte aligned channels-last bf16");
  const long long V = (long long)N * D * H * W;
  if (V >= (1LL << 31)) return set_error(VFD_ERR_ARG, "tap_gather: more than 2^31 voxels");
  if (V == 0) return 0;
  if (sign != 1 && sign != -1) return set_error(VFD_ERR_ARG, "tap_gather: sign must be +1 or -1");
  GatherGeom g;
  g.N = N; g.D = D; g.H = H; g.W = W; g.V = (unsigned)V;
  g.fW = make_fastdiv(W); g.fH = make_fastdiv(H); g.fD = make_fastdiv(D);
  const int taps = kd * kh * kw;
  if (dst_cols != ((taps * cs + 7) & ~7) || dst_ld < dst_cols)
    return set_error(VFD_ERR_ARG, "tap_gather: dst_cols must be taps*cs rounded up to 8");
#define VFD_GATHER(KD, KH, KW, CS)                                                                       \
  if (kd == KD && kh == KH && kw == KW && cs == CS) {                                                    \
    tap_gather_kernel<KD, KH, KW, CS><<<grid_for(V), 256, 0, STREAM>>>((const bf16*)src, src_ld, (bf16*)dst, \
                                                                       dst_ld, g, sign);                \
    return check_launch("tap_gather");                                                                   \
  }
  VFD_GATHER(3, 3, 3, 1)
  VFD_GATHER(1, 3, 3, 3)
  VFD_GATHER(1, 3, 3, 1)
  VFD_GATHER(1, 3, 3, 2)
  VFD_GATHER(3, 1, 1, 1)
  VFD_GATHER(3, 1, 1, 2)
  VFD_GATHER(3, 1, 1, 3)
  VFD_GATHER(3, 1, 1, 4)
#undef VFD_GATHER
  return set_error(VFD_ERR_ARG, "tap_gather: unsupported (kernel, channels-per-tap) combination");
}

VFD_API int vfd_channel_sum(const void* x, long long ld, int C, long long V, double* out, void* stream_) {
  if (int e = check_cl(x, ld, C, "channel_sum: bad tensor")) return e;
  if (V <= 0) return 0;
  channel_sum_kernel<<<stats_grid(V, C), 256, C * sizeof(double), STREAM>>>((const bf16*)x, ld, C, V, out);
  return check_launch("channel_sum");
}

static int fill_up_geom(UpGeom& g, int N, int D, int H, int W, int C, int scale, const char* what) {
  const long long total = (long long)N * D * H * W * scale * scale * scale * (C / 8);
  if (total >= (1LL << 31)) return set_error(VFD_ERR_ARG, what);
  g.N = N; g.D = D; g.H = H; g.W = W; g.CG = C / 8;
  g.total = (unsigned)total;
  g.fCG = make_fastdiv(C / 8);
  g.fX = make_fastdiv(W * scale); g.fY = make_fastdiv(H * scale); g.fZ = make_fastdiv(D * scale);
  return 0;
}

VFD_API int vfd_upsample2x_fwd(const void* x, long long x_ld, int N, int D, int H, int W, int C,
                                  void* out, long long out_ld, void* stream_) {
  if (int e = check_cl(x, x_ld, C, "upsample2x_fwd: bad input")) return e;
  if (int e = check_cl(out, out_ld, C, "upsample2x_fwd: bad output")) return e;
  UpGeom g;
  if (int e = fill_up_geom(g, N, D, H, W, C, 2, "upsample2x_fwd: more than 2^31 output vectors")) return e;
  if (g.total == 0) return 0;
  static const int gather = getenv("VFD_UPSAMPLE_GATHER") ? atoi(getenv("VFD_UPSAMPLE_GATHER")) : 0;
  const long long cells = (long long)N * D * H * W * (C / 4);
  if (!gather && cells < (1LL << 31) && (long long)N * 8 * D * H * W < (1LL << 31) && D + H + W <= 1400) {
    UpCellGeom c;
    c.N = N; c.D = D; c.H = H; c.W = W; c.CG = C / 4; c.total = (unsigned)cells;
    c.fCG = make_fastdiv(C / 4); c.fW = make_fastdiv(W); c.fH = make_fastdiv(H); c.fD = make_fastdiv(D);
    upsample2x_fwd_cell_kernel<<<grid_for(cells), 256, 32 * (size_t)(D + H + W), STREAM>>>((const bf16*)x, x_ld, c, (bf16*)out, out_ld);
    return check_launch("upsample2x_fwd_cell");
  }
  upsample2x_fwd_kernel<<<grid_for(g.total), 256, 0, STREAM>>>((const bf16*)x, x_ld, g, (bf16*)out, out_ld);
  return check_launch("upsample2x_fwd");
}

static int launch_up_axis(const bf16* src, long long src_ld, bf16* dst, long long dst_ld, long long outer, int L,
                          long long inner, int C, cudaStream_t stream) {
  const long long total = outer * L * inner * (C / 8);
  if (total >= (1LL << 31) || inner >= (1LL << 31)) return set_error(VFD_ERR_ARG, "upsample2x_bwd: more than 2^31 vectors");
  UpAxisGeom g;
  g.total = (unsigned)total; g.L = L; g.inner = (int)inner; g.CG = C / 8;
  g.fCG = make_fastdiv(C / 8); g.fInner = make_fastdiv((uint32_t)inner); g.fL = make_fastdiv(L);
  static const int seven = getenv("VFD_UPSAMPLE_GATHER") ? atoi(getenv("VFD_UPSAMPLE_GATHER")) : 0;
  if (!seven && L <= 2048) {
    upsample_bwd_axis4_kernel<<<grid_for(total), 256, 4 * L * sizeof(float), stream>>>(src, src_ld, dst, dst_ld, g);
    return check_launch("upsample2x_bwd_axis4");
  }
  upsample_bwd_axis_kernel<<<grid_for(total), 256, 0, stream>>>(src, src_ld, dst, dst_ld, g);
  return check_launch("upsample2x_bwd_axis");
}

VFD_API int vfd_upsample2x_bwd(const void* go, long long go_ld, int N, int D, int H, int W, int C,
                                  void* gx, long long gx_ld, void* workspace, long long ws_bytes, void* stream_) {
  if (int e = check_cl(go, go_ld, C, "upsample2x_bwd: bad input")) return e;
  if (int e = check_cl(gx, gx_ld, C, "upsample2x_bwd: bad output")) return e;
  if ((long long)N * D * H * W == 0) return 0;
  // separable path: W, H, D adjoint passes through two bf16 temporaries in the workspace
  const long long t1 = (long long)N * 2 * D * 2 * H * W * C * 2, t2 = (long long)N * 2 * D * H * W * C * 2;
  if (workspace != nullptr && ws_bytes >= t1 + t2 && !(reinterpret_cast<uintptr_t>(workspace) & 15)) {
    bf16* tmp1 = reinterpret_cast<bf16*>(workspace);
    bf16* tmp2 = reinterpret_cast<bf16*>(reinterpret_cast<char*>(workspace) + t1);
    if (int e = launch_up_axis((const bf16*)go, go_ld, tmp1, C, (long long)N * 2 * D * 2 * H, W, 1, C, STREAM)) return e;
    if (int e = launch_up_axis(tmp1, C, tmp2, C, (long long)N * 2 * D, H, W, C, STREAM)) return e;
    return launch_up_axis(tmp2, C, (bf16*)gx, gx_ld, N, D, (long long)H * W, C, STREAM);
  }
  UpGeom g;
  if (int e = fill_up_geom(g, N, D, H, W, C, 1, "upsample2x_bwd: more than 2^31 vectors")) return e;
  upsample2x_bwd_kernel<<<grid_for(g.total), 256, 0, STREAM>>>((const bf16*)go, go_ld, g, (bf16*)gx, gx_ld);
  return check_launch("upsample2x_bwd");
}

VFD_API int vfd_sigmoid_head_fwd(const float* logits, long long ld, long long V, float* predict, void* stream_) {
  if (V <= 0) return 0;
  sigmoid_head_fwd_kernel<<<grid_for(V), 256, 0, STREAM>>>(logits, ld, V, predict);
  return check_launch("sigmoid_head_fwd");
}

VFD_API int vfd_sigmoid_head_bwd(const float* gpred, const float* predict, long long V, void* dlogit, void* stream_) {
  if (V <= 0) return 0;
  sigmoid_head_bwd_kernel<<<grid_for(V), 256, 0, STREAM>>>(gpred, predict, V, (bf16*)dlogit);
  return check_launch("sigmoid_head_bwd");
}

VFD_API int vfd_weighted_bce(const float* predict, const float* target, long long V, float pos_weight,
                                float grad_scale, double* loss_sum, float* gpred, void* stream_) {
  if (V <= 0) return 0;
  wbce_kernel<<<grid_for(V, 256, 148 * 4), 256, 0, STREAM>>>(predict, target, V, pos_weight, grad_scale, loss_sum, gpred);
  return check_launch("weighted_bce");
}

VFD_API int vfd_sqdiff(const void* a, long long a_ld, const void* b, long long b_ld, int C, long long V,
                          double* out, void* stream_) {
  if (int e = check_cl(a, a_ld, C, "sqdiff: bad a")) return e;
  if (int e = check_cl(b, b_ld, C, "sqdiff: bad b")) return e;
  if (V <= 0) return 0;
  sqdiff_kernel<<<grid_for(V * (C / 8), 256, 148 * 4), 256, 0, STREAM>>>((const bf16*)a, a_ld, (const bf16*)b, b_ld, C, V, out);
  return check_launch("sqdiff");
}

VFD_API int vfd_convlstm_cell_fwd(const float* gates, long long g_ld, const float* c_cur, int hid,
                                     long long V, float* h_next, float* c_next, float* act, void* stream_) {
  if (V <= 0) return 0;
  convlstm_cell_fwd_kernel<<<grid_for(V * hid), 256, 0, STREAM>>>(gates, g_ld, c_cur, hid, V, h_next, c_next, act);
  return check_launch("convlstm_cell_fwd");
}

VFD_API int vfd_convlstm_cell_bwd(const float* act, const float* c_cur, const float* c_next,
                                     const float* dh, const float* dc_in, int hid, long long V,
                                     void* dgates, long long dg_ld, float* dc_cur, void* stream_) {
  if (V <= 0) return 0;
  convlstm_cell_bwd_kernel<<<grid_for(V * hid), 256, 0, STREAM>>>(act, c_cur, c_next, dh, dc_in, hid, V, (bf16*)dgates, dg_ld, dc_cur);
  return check_launch("convlstm_cell_bwd");
}
