// Voxel-level ROC / precision-recall areas on the device for any number of (score, label) pairs up to 2^31 - 1:
// what lib/evaluate.py:27-38 (`roc`: auc(roc_curve(labels, scores))) and :66-68 (`pr`: auc(recall, precision) of
// precision_recall_curve) compute with sklearn on the flattened voxel arrays that test.py:175-202 and
// models/mygannet.py:444-470 collect over the whole test set (B*D*H*W per batch, 1e8 and more per sweep).
//
// Pipeline (HBM-bound integer work, no host synchronisation, deterministic):
//   1. keys      key64 = order-preserving 32-bit image of the score << 1 | label      (read 8 B, write 8 B per pair)
//   2. sort      stable LSD radix sort of the 33 significant bits, 4 passes of 9 bits. Every pass = per-block digit
//                histogram over the block's contiguous chunk, one single-block exclusive scan of the [512][blocks]
//                table, and a scatter in which a warp ranks its keys with match.any (no shared-memory sort)
//   3. prefix    negx[i] = number of negatives before sorted position i (chunk totals, scan, per-chunk block scans)
//   4. areas     one thread per position; the last position of every run of equal scores ("group") owns the group:
//                  ROC   2U += pos_g * (2 * neg_below_g + neg_g)      exact in 64-bit integers (the tie-aware
//                        Mann-Whitney count; AUC = U / (P * N) is the trapezoid area under sklearn's roc_curve)
//                  PR    the trapezoid between this group's (recall, precision) point and the next higher threshold's
//                        (or the end point (0, 1)), summed in double per block, blocks added in index order
//   5. finalize  out = {AUC, P, N, PR area}
// The group's first position is found by galloping backwards from its last one (1, 2, 4, ... then bisection), so
// continuous scores (groups of one) cost one neighbouring read.
#include <cstdint>
#include "vfd_internal.h"

namespace vfd {
namespace {

constexpr int kRadixBits = 9;
constexpr int kBins = 1 << kRadixBits;          // 512
constexpr int kPasses = 4;                      // 36 >= 33 significant key bits
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 8;
constexpr int kTile = kSortThreads * kItems;    // 2048 keys per tile
constexpr int kMaxBlocks = 148 * 4;

__device__ __forceinline__ uint32_t order_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(256)
roc_keys_kernel(const float* __restrict__ scores, const float* __restrict__ labels, long long n,
                uint64_t* __restrict__ keys) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    float s = scores[i];
    s += 0.f;   // -0 and +0 are one score
    keys[i] = (static_cast<uint64_t>(order_key(s)) << 1) | (labels[i] > 0.5f ? 1ull : 0ull);
  }
}

// counts[d * nblocks + b] = number of keys of block b's chunk whose digit is d
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const uint64_t* __restrict__ keys, long long n, long long chunk, int shift,
                  uint32_t* __restrict__ counts) {
  __shared__ uint32_t h[kBins];
  for (int d = threadIdx.x; d < kBins; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const long long begin = blockIdx.x * chunk;
  long long end = begin + chunk;
  if (end > n) end = n;
  const int lane = threadIdx.x & 31;
  for (long long base = begin; base < end; base += kSortThreads) {   // block-uniform trip count
    const long long i = base + threadIdx.x;
    const bool valid = i < end;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = static_cast<uint32_t>(keys[i] >> shift) & (kBins - 1);
      const unsigned peers = __match_any_sync(vm, d);          // one shared atomic per distinct digit of the warp
      if ((peers & ((1u << lane) - 1)) == 0) atomicAdd(&h[d], __popc(peers));
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kBins; d += kSortThreads)
    counts[static_cast<size_t>(d) * gridDim.x + blockIdx.x] = h[d];
}

// in-place exclusive scan of m 32-bit counters by one block; total (optional) receives the sum
__global__ void __launch_bounds__(1024)
scan_u32_kernel(uint32_t* __restrict__ v, int m, uint32_t* __restrict__ total) {
  __shared__ uint32_t wsum[32];
  const int per = (m + 1023) / 1024;
  const int begin = threadIdx.x * per;
  uint32_t s = 0;
  for (int e = 0; e < per; ++e)
    if (begin + e < m) s += v[begin + e];
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  uint32_t off = 0, all = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < (threadIdx.x >> 5)) off += wsum[w];
    all += wsum[w];
  }
  uint32_t run = off + incl - s;
  for (int e = 0; e < per; ++e)
    if (begin + e < m) {
      const uint32_t c = v[begin + e];
      v[begin + e] = run;
      run += c;
    }
  if (total != nullptr && threadIdx.x == 0) *total = all;
}

// stable scatter of block b's chunk to the positions the scanned table assigns to (digit, block)
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, long long n, long long chunk,
                     int shift, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t run[kBins];                     // next free output slot of every digit of this block
  __shared__ uint32_t tot[kBins];                     // digit totals of the current tile
  __shared__ uint32_t cnt[kSortWarps][kBins];         // per-warp digit counts, then exclusive prefix over the warps
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < kBins; d += kSortThreads)
    run[d] = offsets[static_cast<size_t>(d) * gridDim.x + blockIdx.x];
  const long long begin = blockIdx.x * chunk;
  long long end = begin + chunk;
  if (end > n) end = n;
  for (long long base = begin; base < end; base += kTile) {
    for (int d = threadIdx.x; d < kSortWarps * kBins; d += kSortThreads) (&cnt[0][0])[d] = 0;
    __syncthreads();
    uint64_t key[kItems];
    uint32_t rank[kItems];
    // order inside the tile: warp, item, lane -- ranks follow it, so equal digits keep their input order
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const long long i = base + w * (32 * kItems) + k * 32 + lane;
      const bool valid = i < end;
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      key[k] = 0;
      rank[k] = 0;
      if (valid) {
        key[k] = in[i];
        const uint32_t d = static_cast<uint32_t>(key[k] >> shift) & (kBins - 1);
        const unsigned peers = __match_any_sync(vm, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader) {
          old = cnt[w][d];
          cnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(vm, old, leader);
        rank[k] = old + __popc(peers & ((1u << lane) - 1));
      }
      __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kBins; d += kSortThreads) {
      uint32_t acc = 0;
#pragma unroll
      for (int ww = 0; ww < kSortWarps; ++ww) {
        const uint32_t c = cnt[ww][d];
        cnt[ww][d] = acc;
        acc += c;
      }
      tot[d] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const long long i = base + w * (32 * kItems) + k * 32 + lane;
      if (i < end) {
        const uint32_t d = static_cast<uint32_t>(key[k] >> shift) & (kBins - 1);
        out[run[d] + cnt[w][d] + rank[k]] = key[k];
      }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kBins; d += kSortThreads) run[d] += tot[d];
    __syncthreads();
  }
}

// ---- negative-prefix over the sorted keys -------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* wsum, uint32_t* block_total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();   // wsum may still be read by the previous call
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  uint32_t off = 0, all = 0;
#pragma unroll
  for (int ww = 0; ww < kSortWarps; ++ww) {
    if (ww < w) off += wsum[ww];
    all += wsum[ww];
  }
  *block_total = all;
  return off + incl - v;
}

__global__ void __launch_bounds__(kSortThreads)
neg_count_kernel(const uint64_t* __restrict__ keys, long long n, long long chunk, uint32_t* __restrict__ block_neg) {
  __shared__ uint32_t wsum[kSortWarps];
  const long long begin = blockIdx.x * chunk;
  long long end = begin + chunk;
  if (end > n) end = n;
  uint32_t c = 0;
  for (long long i = begin + threadIdx.x; i < end; i += kSortThreads) c += (keys[i] & 1ull) == 0ull;
  uint32_t all;
  (void)block_excl_scan_256(c, wsum, &all);
  if (threadIdx.x == 0) block_neg[blockIdx.x] = all;
}

// negx[i] = negatives among sorted positions [0, i); every thread owns kItems consecutive positions of a tile
__global__ void __launch_bounds__(kSortThreads)
neg_prefix_kernel(const uint64_t* __restrict__ keys, long long n, long long chunk,
                  const uint32_t* __restrict__ block_neg_excl, uint32_t* __restrict__ negx) {
  __shared__ uint32_t wsum[kSortWarps];
  const long long begin = blockIdx.x * chunk;
  long long end = begin + chunk;
  if (end > n) end = n;
  uint32_t carry = block_neg_excl[blockIdx.x];
  for (long long base = begin; base < end; base += kTile) {
    const long long i0 = base + static_cast<long long>(threadIdx.x) * kItems;
    uint32_t isneg[kItems];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      isneg[k] = (i0 + k < end) ? static_cast<uint32_t>((keys[i0 + k] & 1ull) == 0ull) : 0u;
      c += isneg[k];
    }
    uint32_t all;
    uint32_t run = carry + block_excl_scan_256(c, wsum, &all);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      if (i0 + k < end) negx[i0 + k] = run;
      run += isneg[k];
    }
    carry += all;
  }
}

// ---- areas --------------------------------------------------------------------------------------------------
// partial[b] = {2U share (as uint64 bits), PR share (double bits)} of block b
__global__ void __launch_bounds__(256)
roc_area_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ negx, long long n,
                const uint32_t* __restrict__ total_neg, unsigned long long* __restrict__ partial_u,
                double* __restrict__ partial_pr) {
  const uint64_t N = *total_neg, P = static_cast<uint64_t>(n) - N;
  unsigned long long u2 = 0;
  double pr = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    const uint64_t ki = keys[i];
    const uint64_t s = ki >> 1;
    if (i + 1 < n && (keys[i + 1] >> 1) == s) continue;        // not the last position of its group
    long long h = i;                                           // first position of the group
    if (i > 0 && (keys[i - 1] >> 1) == s) {
      long long step = 2, lo, hi = i - 1;                      // keys[hi] is in the group
      for (;;) {
        lo = i - step;
        if (lo <= 0) { lo = 0; break; }
        if ((keys[lo] >> 1) != s) break;
        hi = lo;
        step <<= 1;
      }
      if ((keys[lo] >> 1) == s) {
        h = lo;                                                // reached position 0 inside the group
      } else {
        while (hi - lo > 1) {                                  // keys[lo] below the group, keys[hi] inside it
          const long long mid = (lo + hi) >> 1;
          if ((keys[mid] >> 1) == s) hi = mid; else lo = mid;
        }
        h = hi;
      }
    }
    const uint64_t neg_b = negx[h], pos_b = static_cast<uint64_t>(h) - neg_b;
    const uint64_t neg_e = static_cast<uint64_t>(negx[i]) + ((ki & 1ull) == 0ull), pos_e = static_cast<uint64_t>(i) + 1 - neg_e;
    u2 += (pos_e - pos_b) * (neg_b + neg_e);                   // pos_g * (2 neg_below + neg_g)
    if (P > 0) {
      const double tp0 = static_cast<double>(P - pos_b), fp0 = static_cast<double>(N - neg_b);
      const double r0 = tp0 / static_cast<double>(P), p0 = tp0 / (tp0 + fp0);
      double r1 = 0.0, p1 = 1.0;
      if (i + 1 < n) {
        const double tp1 = static_cast<double>(P - pos_e), fp1 = static_cast<double>(N - neg_e);
        r1 = tp1 / static_cast<double>(P);
        p1 = tp1 / (tp1 + fp1);
      }
      pr += (r0 - r1) * (p0 + p1) * 0.5;
    }
  }
  __shared__ unsigned long long su[8];
  __shared__ double sp[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    u2 += __shfl_xor_sync(0xffffffffu, u2, o);
    pr += __shfl_xor_sync(0xffffffffu, pr, o);
  }
  if ((threadIdx.x & 31) == 0) {
    su[threadIdx.x >> 5] = u2;
    sp[threadIdx.x >> 5] = pr;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0;
    double b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += su[w];
      b += sp[w];
    }
    partial_u[blockIdx.x] = a;
    partial_pr[blockIdx.x] = b;
  }
}

__global__ void roc_finalize_kernel(const unsigned long long* __restrict__ partial_u,
                                    const double* __restrict__ partial_pr, int nblocks, long long n,
                                    const uint32_t* __restrict__ total_neg, double* __restrict__ out) {
  if (threadIdx.x != 0) return;
  unsigned long long u2 = 0;
  double pr = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    u2 += partial_u[b];
    pr += partial_pr[b];
  }
  const double N = static_cast<double>(*total_neg), P = static_cast<double>(n) - N;
  out[0] = (P > 0 && N > 0) ? (static_cast<double>(u2) * 0.5) / (P * N) : nan("");
  out[1] = P;
  out[2] = N;
  out[3] = P > 0 ? pr : nan("");
}

struct Layout {
  long long keys_a, keys_b, counts, block_neg, total_neg, partial_u, partial_pr, bytes;
  int nblocks;
  long long chunk;
};

inline long long align256(long long v) { return (v + 255) & ~255ll; }

Layout layout_for(long long n) {
  Layout L;
  long long tiles = (n + kTile - 1) / kTile;
  if (tiles < 1) tiles = 1;
  L.nblocks = static_cast<int>(tiles < kMaxBlocks ? tiles : kMaxBlocks);
  long long tiles_per_block = (tiles + L.nblocks - 1) / L.nblocks;
  if (tiles_per_block < 8) tiles_per_block = 8;   // small inputs: fewer blocks, a shorter [512][blocks] table to scan
  L.chunk = tiles_per_block * kTile;
  L.nblocks = static_cast<int>((n + L.chunk - 1) / L.chunk);
  if (L.nblocks < 1) L.nblocks = 1;
  long long off = 0;
  L.keys_a = off; off = align256(off + 8 * n);
  L.keys_b = off; off = align256(off + 8 * n);
  L.counts = off; off = align256(off + 4ll * kBins * L.nblocks);
  L.block_neg = off; off = align256(off + 4ll * L.nblocks);
  L.total_neg = off; off = align256(off + 4);
  L.partial_u = off; off = align256(off + 8ll * kMaxBlocks);
  L.partial_pr = off; off = align256(off + 8ll * kMaxBlocks);
  L.bytes = off;
  return L;
}

}  // namespace
}  // namespace vfd

using namespace vfd;

VFD_API long long vfd_roc_auc_large_workspace(long long n) {
  if (n < 0) return -1;
  return layout_for(n > 0 ? n : 1).bytes;
}

VFD_API int vfd_roc_auc_large(const float* scores, const float* labels, long long n, double* out, void* workspace,
                              long long ws_bytes, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (out == nullptr || n < 0 || (n > 0 && (scores == nullptr || labels == nullptr)))
    return set_error(VFD_ERR_ARG, "roc_auc_large: bad arguments");
  if (n >= (1ll << 31)) return set_error(VFD_ERR_ARG, "roc_auc_large: at most 2^31 - 1 pairs per call");
  const Layout L = layout_for(n > 0 ? n : 1);
  if (workspace == nullptr || ws_bytes < L.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(VFD_ERR_ARG, "roc_auc_large: workspace too small or not 256-byte aligned "
                                  "(vfd_roc_auc_large_workspace)");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint64_t* ka = reinterpret_cast<uint64_t*>(ws + L.keys_a);
  uint64_t* kb = reinterpret_cast<uint64_t*>(ws + L.keys_b);
  uint32_t* counts = reinterpret_cast<uint32_t*>(ws + L.counts);
  uint32_t* block_neg = reinterpret_cast<uint32_t*>(ws + L.block_neg);
  uint32_t* total_neg = reinterpret_cast<uint32_t*>(ws + L.total_neg);
  unsigned long long* partial_u = reinterpret_cast<unsigned long long*>(ws + L.partial_u);
  double* partial_pr = reinterpret_cast<double*>(ws + L.partial_pr);
  const int nb = L.nblocks;
  int grid = static_cast<int>((n + 255) / 256);
  if (grid > kMaxBlocks) grid = kMaxBlocks;
  if (grid < 1) grid = 1;
  if (n > 0) {
    roc_keys_kernel<<<grid, 256, 0, st>>>(scores, labels, n, ka);
    for (int pass = 0; pass < kPasses; ++pass) {
      const int shift = pass * kRadixBits;
      radix_hist_kernel<<<nb, kSortThreads, 0, st>>>(ka, n, L.chunk, shift, counts);
      scan_u32_kernel<<<1, 1024, 0, st>>>(counts, kBins * nb, nullptr);
      radix_scatter_kernel<<<nb, kSortThreads, 0, st>>>(ka, kb, n, L.chunk, shift, counts);
      uint64_t* t = ka;
      ka = kb;
      kb = t;
    }
    // kPasses is even: the sorted keys are back in the first buffer, the second one becomes negx
    uint32_t* negx = reinterpret_cast<uint32_t*>(kb);
    neg_count_kernel<<<nb, kSortThreads, 0, st>>>(ka, n, L.chunk, block_neg);
    scan_u32_kernel<<<1, 1024, 0, st>>>(block_neg, nb, total_neg);
    neg_prefix_kernel<<<nb, kSortThreads, 0, st>>>(ka, n, L.chunk, block_neg, negx);
    roc_area_kernel<<<grid, 256, 0, st>>>(ka, negx, n, total_neg, partial_u, partial_pr);
    roc_finalize_kernel<<<1, 32, 0, st>>>(partial_u, partial_pr, grid, n, total_neg, out);
  } else {
    cudaError_t e = cudaMemsetAsync(total_neg, 0, 4, st);
    if (e != cudaSuccess) return set_cuda_error(e, "roc_auc_large: memset");
    roc_finalize_kernel<<<1, 32, 0, st>>>(partial_u, partial_pr, 0, 0, total_neg, out);
  }
  return check_launch("roc_auc_large");
}
