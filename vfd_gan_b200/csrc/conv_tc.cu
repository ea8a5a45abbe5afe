// Implicit-GEMM conv3d (stride 1, "same" zero padding, kernel taps in {1,3}^3) on tcgen05.
//
// Replaces the nn.Conv3d calls of the reference's R(2+1)D block
// (models/spatiotempconv.py:49-50,59-60,63-64), conv_last (models/mygannet.py:52,97) and the
// ConvLSTM gate conv (models/convlstm.py:36-40,46) -- forward, dgrad and wgrad.
//
// Data layout: activations are channels-last bf16 [N][D][H][W][C] (C a multiple of 8). One GEMM
// M-tile is a box of TW x TH x TD x TN = 128 voxels. For every filter tap the TMA engine loads
// the box shifted by the tap offset straight into a swizzled K-major smem tile (out-of-bounds
// voxels / channels are zero-filled by the TMA unit, which IS the conv zero padding), and one
// elected thread issues tcgen05.mma accumulating all taps x channel blocks into TMEM.
//
//   forward / dgrad : D[voxel][n]  = sum_{tap,c} X[voxel+tap][c] * Wp[n][tap][c]   (A,B K-major)
//   wgrad           : dW[tap][co][ci] = sum_voxel dY[voxel][co] * X[voxel+tap][ci] (A,B MN-major)
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"
#include "vfd_internal.h"

// Stage-isolation switches of tools/gpu_stage_probe.py (skip the TMA loads / the MMA issue / the epilogue ...).
// They exist only in the debug library (libvfd_b200_debug.so, -DVFD_DEBUG): in the product build the macros are the
// constant 0 and the compiler removes every branch they guard; vfd_set_debug is not exported.
#ifdef VFD_DEBUG
#define VFD_DBG(p, bit) ((p).dbg & (bit))
#define VFD_GDBG(bit) (g_dbg & (bit))
#else
#define VFD_DBG(p, bit) (0)
#define VFD_GDBG(bit) (0)
#endif

namespace vfd {

typedef double tc_stat_t;  // shared statistics accumulators: exact sums of fp32 partials, hence order-independent
constexpr int kTileM = 128;
constexpr int kFwdThreads = 192;  // warp0: TMA producer, warp1: MMA issuer, warps2-5: epilogue
constexpr int kMaxStages = 8;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

struct ConvGeom {
  int N, D, H, W;
  int TW, TH, TD, TN;
  int tilesW, tilesH, tilesD, tilesN;
  int ntaps;
  int8_t tap[27][4];  // (dd, dh, dw, unused): input offset of each tap relative to the output voxel
};

// Output side shared by the forward kernels.
struct Epilogue {
  int n_rows;        // rows in the packed weight matrix (valid bias entries)
  int out_cols;      // columns to store (multiple of 8)
  long long out_ld;  // elements between consecutive voxels in the output buffer
  int out_fp32;
  const float* bias;
  void* out;
  double* stats;     // optional [2][stats_ld]: per-channel sum / sum of squares of the stored bf16 values
  int stats_ld;
  // ConvLSTM step (models/convlstm.py:46-58) fused into the gate conv's epilogue: lstm_j > 0 means the packed weight
  // rows are ordered [n-tile][gate i,f,o,g][lstm_j hidden channels], so one 4*lstm_j-column accumulator tile holds all
  // four gates of hidden channels [nt*lstm_j, (nt+1)*lstm_j); bias is in the same (permuted) order
  int lstm_j, lstm_hid;
  const float* lstm_c_cur;   // [V][hid] fp32
  float* lstm_c_next;        // [V][hid] fp32
  float* lstm_act;           // [V][4*hid] fp32, gate-major (i | f | o | g) post-activation values for the backward pass
  void* lstm_h;              // bf16 [V][lstm_h_ld]: h' straight into the next step's [x, h] concat slice
  long long lstm_h_ld;
};

// column sums of a 32-row x 32-column register tile spread over a warp (row = lane): after the
// butterfly lane L holds the sum of column L. 31 shuffles.
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// One accumulator tile (this warp's 32 TMEM lanes x block_n columns) -> bias, convert, store, and
// optionally per-channel statistics into the CTA's shared accumulators. Warp-collective.
__device__ __forceinline__ void epilogue_tile(const Epilogue& e, uint32_t tacc, int block_n, int col0,
                                              bool valid, long long vox, int lane, tc_stat_t* s_sum,
                                              tc_stat_t* s_sq) {
  for (int c = 0; c < block_n; c += 32) {
    float v[32];
    if (c + 32 <= block_n) {
      tmem_ld32(tacc + c, v);
    } else {
      tmem_ld16(tacc + c, v);
#pragma unroll
      for (int i = 16; i < 32; ++i) v[i] = 0.f;
    }
    const int col = col0 + c;
    if (e.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col + i < e.n_rows) v[i] += __ldg(e.bias + col + i);
    }
    if (e.out_fp32) {
      if (valid) {
        float* o = reinterpret_cast<float*>(e.out) + vox * e.out_ld + col;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (c + i < block_n && col + i < e.out_cols)
            *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + vox * e.out_ld + col;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk;
        const __nv_bfloat162 t0 = __floats2bfloat162_rn(v[i], v[i + 1]);
        const __nv_bfloat162 t1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
        const __nv_bfloat162 t2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
        const __nv_bfloat162 t3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
        pk.x = *reinterpret_cast<const uint32_t*>(&t0);
        pk.y = *reinterpret_cast<const uint32_t*>(&t1);
        pk.z = *reinterpret_cast<const uint32_t*>(&t2);
        pk.w = *reinterpret_cast<const uint32_t*>(&t3);
        if (valid && c + i < block_n && col + i < e.out_cols) *reinterpret_cast<uint4*>(o + i) = pk;
        if (e.stats != nullptr) {  // statistics of exactly what is stored
          const float2 f0 = __bfloat1622float2(t0), f1 = __bfloat1622float2(t1);
          const float2 f2 = __bfloat1622float2(t2), f3 = __bfloat1622float2(t3);
          v[i] = f0.x; v[i + 1] = f0.y; v[i + 2] = f1.x; v[i + 3] = f1.y;
          v[i + 4] = f2.x; v[i + 5] = f2.y; v[i + 6] = f3.x; v[i + 7] = f3.y;
        }
      }
    }
    if (e.stats != nullptr) {
      float sq[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = valid ? v[i] : 0.f;
        sq[i] = v[i] * v[i];
      }
      const float cs = warp_colsum32(v, lane);
      const float cq = warp_colsum32(sq, lane);
      if (c + lane < block_n && col + lane < 1024) {
        // double accumulators: sums of fp32 partials are exact there (24-bit mantissas, a few thousand summands), so
        // the result does not depend on the order in which the epilogue warps arrive
        atomicAdd(&s_sum[col + lane], static_cast<tc_stat_t>(cs));
        atomicAdd(&s_sq[col + lane], static_cast<tc_stat_t>(cq));
      }
    }
  }
}

// After the last tile: the four epilogue warps publish the CTA's statistics.
__device__ __forceinline__ void epilogue_flush_stats(const Epilogue& e, int epi_thread, const tc_stat_t* s_sum,
                                                     const tc_stat_t* s_sq) {
  if (e.stats == nullptr) return;
  asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
  for (int j = epi_thread; j < e.stats_ld && j < 1024; j += 128) {
    if (s_sum[j] != 0 || s_sq[j] != 0) {
      atomicAdd(e.stats + j, static_cast<double>(s_sum[j]));
      atomicAdd(e.stats + e.stats_ld + j, static_cast<double>(s_sq[j]));
    }
  }
}

// ConvLSTM epilogue: this warp's 32 voxel rows x one n-tile = the four gates of lstm_j hidden channels.
//   i, f, o = sigmoid, g = tanh; c' = f * c + i * g; h' = o * tanh(c')      (models/convlstm.py:49-58)
// 16 channels at a time: four tcgen05.ld of 16 columns (one per gate), c read and c' written as fp32, h' written as
// bf16 into the next step's concat slice, activations saved gate-major for the cell backward kernel. The gates never
// go to HBM.
__device__ __forceinline__ void epilogue_tile_lstm(const Epilogue& e, uint32_t tacc, int nt, bool valid, long long vox) {
  const int J = e.lstm_j, hid = e.lstm_hid;
  for (int jb = 0; jb < J; jb += 16) {
    float gi[16], gf[16], go[16], gg[16];
    tmem_ld16(tacc + jb, gi);
    tmem_ld16(tacc + J + jb, gf);
    tmem_ld16(tacc + 2 * J + jb, go);
    tmem_ld16(tacc + 3 * J + jb, gg);
    if (!valid) continue;
    const int ch0 = nt * J + jb;
    if (e.bias != nullptr) {
      const float* b = e.bias + nt * 4 * J + jb;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        gi[i] += __ldg(b + i);
        gf[i] += __ldg(b + J + i);
        go[i] += __ldg(b + 2 * J + i);
        gg[i] += __ldg(b + 3 * J + i);
      }
    }
    const float* cc = e.lstm_c_cur + vox * hid + ch0;
    float c[16], h[16];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(cc + i);
      c[i] = t.x; c[i + 1] = t.y; c[i + 2] = t.z; c[i + 3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      gi[i] = 1.f / (1.f + expf(-gi[i]));
      gf[i] = 1.f / (1.f + expf(-gf[i]));
      go[i] = 1.f / (1.f + expf(-go[i]));
      gg[i] = tanhf(gg[i]);
      c[i] = gf[i] * c[i] + gi[i] * gg[i];
      h[i] = go[i] * tanhf(c[i]);
    }
    float* cn = e.lstm_c_next + vox * hid + ch0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(cn + i) = make_float4(c[i], c[i + 1], c[i + 2], c[i + 3]);
    __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(e.lstm_h) + vox * e.lstm_h_ld + ch0;
#pragma unroll
    for (int i = 0; i < 16; i += 8) {
      uint4 pk;
      const __nv_bfloat162 t0 = __floats2bfloat162_rn(h[i], h[i + 1]), t1 = __floats2bfloat162_rn(h[i + 2], h[i + 3]);
      const __nv_bfloat162 t2 = __floats2bfloat162_rn(h[i + 4], h[i + 5]), t3 = __floats2bfloat162_rn(h[i + 6], h[i + 7]);
      pk.x = *reinterpret_cast<const uint32_t*>(&t0);
      pk.y = *reinterpret_cast<const uint32_t*>(&t1);
      pk.z = *reinterpret_cast<const uint32_t*>(&t2);
      pk.w = *reinterpret_cast<const uint32_t*>(&t3);
      *reinterpret_cast<uint4*>(hp + i) = pk;
    }
    if (e.lstm_act != nullptr) {
      float* ap = e.lstm_act + vox * 4 * hid + ch0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        *reinterpret_cast<float4*>(ap + i) = make_float4(gi[i], gi[i + 1], gi[i + 2], gi[i + 3]);
        *reinterpret_cast<float4*>(ap + hid + i) = make_float4(gf[i], gf[i + 1], gf[i + 2], gf[i + 3]);
        *reinterpret_cast<float4*>(ap + 2 * hid + i) = make_float4(go[i], go[i + 1], go[i + 2], go[i + 3]);
        *reinterpret_cast<float4*>(ap + 3 * hid + i) = make_float4(gg[i], gg[i + 1], gg[i + 2], gg[i + 3]);
      }
    }
  }
}

struct FwdParams {
  ConvGeom g;
  int cblocks;    // channel blocks of KC per tap
  int n_tiles;    // tiles along the GEMM N (output channel) dimension
  int block_n;    // multiple of 16, <= 256
  int stages;
  Epilogue epi;
};

__device__ __forceinline__ void tile_origin(const ConvGeom& g, int mt, int& n0, int& d0, int& h0,
                                            int& w0) {
  int t = mt;
  w0 = (t % g.tilesW) * g.TW;
  t /= g.tilesW;
  h0 = (t % g.tilesH) * g.TH;
  t /= g.tilesH;
  d0 = (t % g.tilesD) * g.TD;
  t /= g.tilesD;
  n0 = t * g.TN;
}

template <int KC>
__global__ void __launch_bounds__(kFwdThreads, 1)
conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ tc_stat_t s_sum[1024], s_sq[1024];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  if (p.epi.stats != nullptr)
    for (int i = threadIdx.x; i < 1024; i += kFwdThreads) s_sum[i] = s_sq[i] = 0;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int kRowBytes = KC * 2;
  constexpr int kABytes = kTileM * kRowBytes;
  const int b_bytes = p.block_n * kRowBytes;
  const int stage_bytes = (kABytes + b_bytes + 1023) & ~1023;

  const int tiles_m = g.tilesW * g.tilesH * g.tilesD * g.tilesN;
  const int total_tiles = tiles_m * p.n_tiles;
  const int ksteps = g.ntaps * p.cblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        int n0, d0, h0, w0;
        tile_origin(g, mt, n0, d0, h0, w0);
        for (int tap = 0; tap < g.ntaps; ++tap) {
          const int dd = g.tap[tap][0], dh = g.tap[tap][1], dw = g.tap[tap][2];
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
            uint8_t* sb = sa + kABytes;
            mbar_expect_tx(&full_bar[s], kABytes + b_bytes);
            tma_load_5d(&tmA, &full_bar[s], sa, cb * KC, w0 + dw, h0 + dh, d0 + dd, n0);
            tma_load_2d(&tmB, &full_bar[s], sb, (tap * p.cblocks + cb) * KC, nt * p.block_n);
            if (++s == p.stages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform; one elected lane issues)
    {
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128(p.block_n, false, false);
      const uint64_t d0 = sdesc_kmajor(smem_u32(smem), kRowBytes);  // A and B tiles share the layout
      const uint32_t dlo0 = desc_lo(d0), dhi = desc_hi(d0);
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * kAccStride;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t alo = dlo0 + static_cast<uint32_t>(s) * stage16;
          const uint32_t blo = alo + (kABytes >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k)
              umma_bf16(tacc, desc_join(alo + 2 * k, dhi), desc_join(blo + 2 * k, dhi), idesc, (ks | k) != 0);
            umma_commit(&empty_bar[s]);
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (leader) umma_commit(&acc_full[as]);
        if (++as == 2) {
          as = 0;
          aph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------ epilogue (4 warps, one TMEM lane quarter each)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      int n0, d0, h0, w0;
      tile_origin(g, mt, n0, d0, h0, w0);
      int r = row;
      const int w = w0 + r % g.TW;
      r /= g.TW;
      const int h = h0 + r % g.TH;
      r /= g.TH;
      const int d = d0 + r % g.TD;
      const int n = n0 + r / g.TD;
      const bool valid = (w < g.W) && (h < g.H) && (d < g.D) && (n < g.N);
      const long long vox = ((static_cast<long long>(n) * g.D + d) * g.H + h) * g.W + w;
      const int col0 = nt * p.block_n;

      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
      if (p.epi.lstm_j > 0) epilogue_tile_lstm(p.epi, tacc, nt, valid, vox);
      else epilogue_tile(p.epi, tacc, p.block_n, col0, valid, vox, lane, s_sum, s_sq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == 2) {
        as = 0;
        aph ^= 1;
      }
    }
    epilogue_flush_stats(p.epi, (warp - 2) * 32 + lane, s_sum, s_sq);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ resident-weight forward
// Persistent forward/dgrad kernel for layers whose whole packed weight matrix fits in shared memory
// (every "small-N" layer of the GAN: first/last convs, uconv1, temporal and 1x1x1 convs ...):
//   * the weights are loaded ONCE per CTA and stay resident, so per tile only the input moves;
//   * a tile is G consecutive d-planes of one 8 x 16 voxel window (G x 128 GEMM rows, G in {1,2,4}). The
//     input arrives as ONE TMA box per channel block: (G + kd - 1) halo planes of (16+kh-1) x (8+kw-1)
//     voxels. Every (sub-tile, tap) pair is a row-shifted UMMA descriptor into that box (UMMA applies the
//     swizzle to absolute smem address bits, so a descriptor may start at any row of the pattern), which
//     amortises the mbarrier / tcgen05.commit hand-shakes over G sub-tiles and re-uses the temporal halo;
//   * the MMA warp runs warp-uniformly over a host-built table of descriptor offsets held in the kernel
//     parameters (constant bank -> uniform registers), one elected lane issuing tcgen05.mma back to back;
//   * eight epilogue warps (two per TMEM lane quarter, alternating sub-tiles) drain the accumulators:
//     TMEM -> registers -> (bias) -> swizzled smem staging -> TMA store, which writes whole 64/128-byte
//     rows and clips partial tiles. For a following BatchNorm they also accumulate the per-channel sum
//     / sum of squares of the stored bf16 values: lane j sums column j of the staged 32 x 32 chunk into
//     registers that live for the whole kernel, so the tensor is not read again for the statistics.
constexpr int kRes2Threads = 320;   // warp0: TMA producer, warp1: MMA issuer, warps2-9: epilogue
constexpr int kResMaxASlots = 8;
constexpr int kMaxChunks = 8;       // block_n <= 256

struct Res2Params {
  int N, D, H, W;
  int tilesW, tilesH, dgroups, G;
  int kd, kh, kw;
  int cblocks, block_n;
  int a_slots, a_slot_bytes, a_box_bytes, b_tile_bytes;
  int acc_stages, acc_stride;  // TMEM columns per sub-tile accumulator; one stage = G * acc_stride columns
  int nchunks;                 // 32-column output chunks
  int stage_bufs;              // staging buffers per epilogue warp (1, 2 or 4)
  int out_fp32;
  int n_rows;                  // valid bias entries
  int dbg;
  const float* bias;
  double* stats;
  int stats_ld;
  int stage_bytes;             // bytes of one epilogue staging buffer (32 rows x 64 or 128 B)
  // thin bf16 outputs (one 32-column chunk): the epilogue warps store their rows straight from registers (64 bytes per
  // voxel, whole 32-byte sectors) instead of staging them for a TMA store -- a 2 KB box of 64-byte rows costs the TMA
  // unit ~3 cycles per row, which made the store the slowest stage of the HBM-bound layers (profiles/r2_thin_epilogue.txt)
  int direct_out;
  void* out;
  long long out_ld;            // elements
  int out_cols;                // channels of the output tensor (multiple of 8)
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// KHW: spatial extent of the filter (kh == kw == KHW, 1 or 3); kd and G are run-time loops around the
// fully unrolled (kh, kw, K16) issue sequence whose descriptor offsets are immediates.
template <int KC, int KHW>
__global__ void __launch_bounds__(kRes2Threads, 1)
conv_fwd_res_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ Res2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[kResMaxASlots], a_empty[kResMaxASlots];
  __shared__ uint64_t b_full, acc_full[4], acc_empty[4];
  __shared__ uint32_t tmem_base_slot;


  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr int kRowBytes = KC * 2;
  const int ntaps = p.kd * p.kh * p.kw;
  const int nbt = ntaps * p.cblocks;  // resident weight tiles
  uint8_t* smem_stage = smem + static_cast<size_t>(nbt) * p.b_tile_bytes;
  uint8_t* smem_a = smem_stage + 8 * p.stage_bufs * p.stage_bytes;
  const int total_tiles = p.N * p.dgroups * p.tilesH * p.tilesW;
  const int stage_cols = p.G * p.acc_stride;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_slots; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(&b_full, 1);
    for (int s = 0; s < p.acc_stages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], p.G == 1 ? 4 : 8);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(&b_full, nbt * p.block_n * kRowBytes);
      for (int t = 0; t < nbt; ++t)
        tma_load_2d(&tmB, &b_full, smem + static_cast<size_t>(t) * p.b_tile_bytes, t * KC, 0);
      int sa = 0;
      uint32_t pha = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int w0 = (t % p.tilesW) * 8;
        t /= p.tilesW;
        const int h0 = (t % p.tilesH) * 16;
        t /= p.tilesH;
        const int d0 = (t % p.dgroups) * p.G;
        const int n = t / p.dgroups;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          mbar_wait(&a_empty[sa], pha ^ 1);
          if VFD_DBG(p, 1) {
            mbar_arrive(&a_full[sa]);
          } else {
            mbar_expect_tx(&a_full[sa], p.a_box_bytes);
            tma_load_5d(&tmA, &a_full[sa], smem_a + static_cast<size_t>(sa) * p.a_slot_bytes, cb * KC,
                        w0 - p.kw / 2, h0 - p.kh / 2, d0 - p.kd / 2, n);
          }
          if (++sa == p.a_slots) {
            sa = 0;
            pha ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-uniform; one elected lane issues)
    const bool leader = elect_one();
    const uint32_t idesc = idesc_bf16_m128(p.block_n, false, false);
    constexpr int PWc = 8 + KHW - 1, PHc = 16 + KHW - 1;
    constexpr uint32_t row16 = kRowBytes >> 4;
    constexpr uint32_t plane16 = PWc * PHc * row16;
    const uint64_t da0 = sdesc_kmajor_ex(smem_u32(smem_a), kRowBytes, PWc * kRowBytes, 0);
    const uint64_t db0 = sdesc_kmajor(smem_u32(smem), kRowBytes);
    const uint32_t alo0 = desc_lo(da0), ahi = desc_hi(da0), blo0 = desc_lo(db0), bhi = desc_hi(db0);
    const uint32_t a_slot16 = p.a_slot_bytes >> 4, b_tile16 = p.b_tile_bytes >> 4;
    const uint32_t tap16 = static_cast<uint32_t>(p.cblocks) * b_tile16;  // next tap's weight tile (same channel block)
    const bool issue = leader && !VFD_DBG(p, 2);
    int sa = 0, as = 0;
    uint32_t pha = 0, aph = 0;
    mbar_wait(&b_full, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[as], aph ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * stage_cols;
      for (int cb = 0; cb < p.cblocks; ++cb) {
        mbar_wait(&a_full[sa], pha);
        tc_fence_after();
        const uint32_t abase = alo0 + static_cast<uint32_t>(sa) * a_slot16;
        const uint32_t bbase = blo0 + static_cast<uint32_t>(cb) * b_tile16;
        for (int g = 0; g < p.G; ++g) {
          const uint32_t d = tacc + g * p.acc_stride;
          uint32_t bt = bbase;
          for (int a = 0; a < p.kd; ++a) {
            const uint32_t ap = abase + static_cast<uint32_t>(g + a) * plane16;
            const uint32_t first_acc = (cb | a) != 0;
            if (issue) {
#pragma unroll
              for (int b = 0; b < KHW; ++b)
#pragma unroll
                for (int c = 0; c < KHW; ++c)
#pragma unroll
                  for (int k = 0; k < KC / 16; ++k)
                    umma_bf16(d, desc_join(ap + (b * PWc + c) * row16 + 2 * k, ahi),
                              desc_join(bt + (b * KHW + c) * tap16 + 2 * k, bhi), idesc,
                              (b | c | k) == 0 ? first_acc : 1u);
            }
            bt += KHW * KHW * tap16;
          }
        }
        if (leader) umma_commit(&a_empty[sa]);
        if (++sa == p.a_slots) {
          sa = 0;
          pha ^= 1;
        }
      }
      if (leader) umma_commit(&acc_full[as]);
      if (++as == p.acc_stages) {
        as = 0;
        aph ^= 1;
      }
    }
  } else {
    // ------------------------------------------------ epilogue (8 warps: group e, TMEM lane quarter q)
    const int ew = warp - 2;
    const int e = ew >> 2;
    const int q = warp & 3;
    uint8_t* stg = smem_stage + static_cast<size_t>(ew) * p.stage_bufs * p.stage_bytes;
    const bool do_stats = p.stats != nullptr;
    // statistics: lane = (column pair cp, row parity rh); it sums columns 2cp, 2cp+1 over rows rh, rh+2, ...
    float ssum[kMaxChunks][2], ssq[kMaxChunks][2];
#pragma unroll
    for (int i = 0; i < kMaxChunks; ++i) ssum[i][0] = ssum[i][1] = ssq[i][0] = ssq[i][1] = 0.f;
    // single-chunk layers (N <= 32, the thin HBM-bound ones) take their statistics from warp-level mma.sync
    // instead: ones x tile gives the column sums, tile^T x tile (diagonal) the sums of squares; the fragments
    // come from the staged chunk through ldmatrix.trans, rows outside the tensor are masked to zero
    float msum[4][4], msq[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) msum[i][j] = msq[i][j] = 0.f;
    const bool mma_stats = p.stats != nullptr && p.nchunks == 1 && !p.out_fp32;
    const int cp = lane & 15, rh = lane >> 4;
    // byte offset of the pair inside a staged row for each swizzle phase (64-byte swizzle: 16-byte chunk
    // index ^ ((row >> 1) & 3)); rows 2i + rh have phase i & 3
    uint32_t coff[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) coff[x] = rh * 64 + ((((cp >> 2) ^ x) << 4) + ((cp & 3) << 2));
    int as = 0, buf = 0;
    uint32_t aph = 0;
    int iter = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      int t = tile;
      const int w0 = (t % p.tilesW) * 8;
      t /= p.tilesW;
      const int h0 = (t % p.tilesH) * 16;
      t /= p.tilesH;
      const int d0 = (t % p.dgroups) * p.G;
      const int n = t / p.dgroups;
      const bool mine = p.G > 1 || ((iter & 1) == e);
      if (mine) {
        mbar_wait(&acc_full[as], aph);
        tc_fence_after();
        const bool hw_ok = (w0 + (lane & 7) < p.W) && (h0 + 4 * q + (lane >> 3) < p.H);
        for (int g = (p.G > 1 ? e : 0); g < p.G; g += 2) {
          const int d = d0 + g;
          if (d >= p.D || VFD_DBG(p, 4)) continue;
          const unsigned vmask = __ballot_sync(0xffffffffu, hw_ok);
          const uint32_t tacc = tmem_base + as * stage_cols + g * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
          for (int ch = 0; ch < p.nchunks; ++ch) {
            float v[32];
            if VFD_DBG(p, 32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = 0.f;
            } else if (ch * 32 + 32 <= p.block_n) {
              tmem_ld32(tacc + ch * 32, v);
            } else {
              tmem_ld16(tacc + ch * 32, v);
#pragma unroll
              for (int i = 16; i < 32; ++i) v[i] = 0.f;
            }
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (ch * 32 + i < p.n_rows) v[i] += __ldg(p.bias + ch * 32 + i);
            }
            uint8_t* sb = stg + buf * p.stage_bytes;
            if (p.direct_out) {
              uint4 pk[4];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                pk[c].x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
                pk[c].y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                pk[c].z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
                pk[c].w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
              }
              if (hw_ok && !VFD_DBG(p, 8)) {
                const long long vox = ((static_cast<long long>(n) * p.D + d) * p.H + (h0 + 4 * q + (lane >> 3))) * p.W +
                                      (w0 + (lane & 7));
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * p.out_ld;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                  if (8 * c < p.out_cols) *reinterpret_cast<uint4*>(op + 8 * c) = pk[c];
              }
              if (mma_stats) {   // the statistics fragments come from the staged copy
                __syncwarp();
                uint8_t* rowp = sb + lane * 64;
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(rowp + ((c ^ ((lane >> 1) & 3)) << 4)) = pk[c];
                __syncwarp();
              }
            } else {
            // the staging buffer may still be the source of an earlier TMA store
            if (lane == 0 && !VFD_DBG(p, 8)) {
              if (p.stage_bufs == 4) bulk_wait_group_read<3>();
              else if (p.stage_bufs == 2) bulk_wait_group_read<1>();
              else bulk_wait_group_read<0>();
            }
            __syncwarp();
            if (p.out_fp32) {
              uint8_t* rowp = sb + lane * 128;
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(rowp + ((c ^ (lane & 7)) << 4)) =
                    make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            } else {
              uint8_t* rowp = sb + lane * 64;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                uint4 pk;
                pk.x = pack_bf16x2(v[8 * c], v[8 * c + 1]);
                pk.y = pack_bf16x2(v[8 * c + 2], v[8 * c + 3]);
                pk.z = pack_bf16x2(v[8 * c + 4], v[8 * c + 5]);
                pk.w = pack_bf16x2(v[8 * c + 6], v[8 * c + 7]);
                *reinterpret_cast<uint4*>(rowp + ((c ^ ((lane >> 1) & 3)) << 4)) = pk;
              }
            }
            if (!VFD_DBG(p, 16)) fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !VFD_DBG(p, 8)) {
              tma_store_5d(&tmC, sb, ch * 32, w0, h0 + 4 * q, d, n);
              bulk_commit_group();
            }
            }
            if (mma_stats) {
              const uint32_t sb_s = smem_u32(sb);
              const int mat = lane >> 3, mr = lane & 7, t2 = (lane & 3) * 2;
              constexpr uint32_t kOnes = 0x3F803F80u;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t vb = vmask >> (16 * ks);
                const uint32_t m_lo = (((vb >> t2) & 1u) ? 0xFFFFu : 0u) | (((vb >> (t2 + 1)) & 1u) ? 0xFFFF0000u : 0u);
                const uint32_t m_hi = (((vb >> (t2 + 8)) & 1u) ? 0xFFFFu : 0u) | (((vb >> (t2 + 9)) & 1u) ? 0xFFFF0000u : 0u);
                const int row = 16 * ks + (mat & 1) * 8 + mr;
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                  const int cidx = 2 * np + (mat >> 1);
                  // b0/b1: voxel rows 0-7 / 8-15 of channel block 2np; b2/b3: the same rows of block 2np+1.
                  // Read as an A fragment (channels on M, voxels on K) the same four registers are the
                  // transposed tile, so tile^T x tile is the Gram matrix whose diagonal is the exact
                  // (bf16 x bf16 products, fp32 accumulate) sum of squares.
                  uint32_t b0, b1, b2, b3;
                  ldmatrix_x4_trans(sb_s + row * 64 + ((cidx ^ ((row >> 1) & 3)) << 4), b0, b1, b2, b3);
                  b0 &= m_lo;
                  b2 &= m_lo;
                  b1 &= m_hi;
                  b3 &= m_hi;
                  mma_bf16_16816(msum[2 * np], kOnes, kOnes, kOnes, kOnes, b0, b1);
                  mma_bf16_16816(msum[2 * np + 1], kOnes, kOnes, kOnes, kOnes, b2, b3);
                  mma_bf16_16816(msq[2 * np], b0, b2, b1, b3, b0, b1);
                  mma_bf16_16816(msq[2 * np + 1], b0, b2, b1, b3, b2, b3);
                }
              }
            } else if (do_stats) {
              float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
              const uint8_t* colp = sb;
              if (vmask == 0xffffffffu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t u = *reinterpret_cast<const uint32_t*>(colp + i * 128 + coff[i & 3]);
                  const float f0 = __uint_as_float(u << 16), f1 = __uint_as_float(u & 0xFFFF0000u);
                  a0 += f0;
                  a1 += f1;
                  q0 = fmaf(f0, f0, q0);
                  q1 = fmaf(f1, f1, q1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const uint32_t u = *reinterpret_cast<const uint32_t*>(colp + i * 128 + coff[i & 3]);
                  const bool rv = (vmask >> (2 * i + rh)) & 1u;
                  const float f0 = rv ? __uint_as_float(u << 16) : 0.f, f1 = rv ? __uint_as_float(u & 0xFFFF0000u) : 0.f;
                  a0 += f0;
                  a1 += f1;
                  q0 = fmaf(f0, f0, q0);
                  q1 = fmaf(f1, f1, q1);
                }
              }
#pragma unroll
              for (int i = 0; i < kMaxChunks; ++i)
                if (i == ch) {
                  ssum[i][0] += a0;
                  ssum[i][1] += a1;
                  ssq[i][0] += q0;
                  ssq[i][1] += q1;
                }
            }
            buf = (buf + 1) & (p.stage_bufs - 1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
      }
      if (++as == p.acc_stages) {
        as = 0;
        aph ^= 1;
      }
    }
    if (lane == 0) bulk_wait_group<0>();
    if (do_stats) {
      // The CTA's statistics meet in fp64 (exact sums of the warps' fp32 partials: order-independent). The 4 KB live in
      // the epilogue staging area, which is free once every epilogue warp has drained its TMA stores -- static shared
      // memory for them cost 2 KB of the input-slot budget and with it up to 3.7 % on the STCNN step.
      double* s_sum = reinterpret_cast<double*>(smem_stage);
      double* s_sq = s_sum + 256;
      __syncwarp();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int i = ew * 32 + lane; i < 512; i += 256) s_sum[i] = 0.0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (mma_stats) {
        if (lane < 4) {   // every accumulator row holds the same column sums: lanes 0-3 own columns 8n + 2t, +1
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            atomicAdd(&s_sum[8 * n + 2 * lane], static_cast<double>(msum[n][0]));
            atomicAdd(&s_sum[8 * n + 2 * lane + 1], static_cast<double>(msum[n][1]));
          }
        }
        // Gram diagonals: accumulator element (row g [+8], column 2t + j) is on the diagonal when g == 2t + j
        const int gq = lane >> 2, dj = gq - 2 * (lane & 3);
        if (dj == 0 || dj == 1) {
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            atomicAdd(&s_sq[16 * np + gq], static_cast<double>(dj ? msq[2 * np][1] : msq[2 * np][0]));
            atomicAdd(&s_sq[16 * np + 8 + gq], static_cast<double>(dj ? msq[2 * np + 1][3] : msq[2 * np + 1][2]));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < kMaxChunks; ++i)
          if (i < p.nchunks) {
            atomicAdd(&s_sum[i * 32 + 2 * cp], static_cast<double>(ssum[i][0]));
            atomicAdd(&s_sum[i * 32 + 2 * cp + 1], static_cast<double>(ssum[i][1]));
            atomicAdd(&s_sq[i * 32 + 2 * cp], static_cast<double>(ssq[i][0]));
            atomicAdd(&s_sq[i * 32 + 2 * cp + 1], static_cast<double>(ssq[i][1]));
          }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight epilogue warps
      const int j = ew * 32 + lane;
      if (j < p.stats_ld && j < p.nchunks * 32 && (s_sum[j] != 0.0 || s_sq[j] != 0.0)) {
        atomicAdd(p.stats + j, s_sum[j]);
        atomicAdd(p.stats + p.stats_ld + j, s_sq[j]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ wgrad
constexpr int kWgThreads = 192;
constexpr int kWgBoxBytes = 128 * 128;  // one TMA box: 128 voxel rows x 64 channels bf16

struct WgradParams {
  ConvGeom g;
  int cout, cin;        // valid channel counts
  int co_tiles;         // tiles of 128 output channels
  int ci_tiles;         // tiles of block_n input channels
  int block_n;          // multiple of 16, <= 256
  int splits;           // voxel-chunk splits
  int stages;
  int tmem_cols;        // power of two >= block_n
  int co_pad;           // leading dimension of the accumulation buffer
  int ci_pad;
  float* acc;           // [ntaps][ci_pad][co_pad] fp32, accumulated with red.add
  long long det_stride; // deterministic mode: split s adds into acc + s * det_stride (0 = all splits into acc)
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));

  // work item: tap fastest so the CTAs that share a dY voxel range run together (L2 reuse)
  int wi = blockIdx.x;
  const int tap = wi % g.ntaps;
  wi /= g.ntaps;
  const int ct = wi % p.ci_tiles;
  wi /= p.ci_tiles;
  const int mt = wi % p.co_tiles;
  const int split = wi / p.co_tiles;

  const int chunks = g.tilesW * g.tilesH * g.tilesD * g.tilesN;
  const int per = (chunks + p.splits - 1) / p.splits;
  const int c_begin = split * per;
  const int c_end = min(chunks, c_begin + per);
  const int co0 = mt * 128;
  const int ci0 = ct * p.block_n;
  const int na = min(2, (p.cout - co0 + 63) / 64);                 // dY boxes with valid channels
  const int nb = (min(p.block_n, p.cin - ci0) + 63) / 64;          // X boxes with valid channels
  const int nb_slots = (p.block_n + 63) / 64;
  const int stage_bytes = (2 + nb_slots) * kWgBoxBytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const bool has_work = c_begin < c_end;

  if (warp == 0) {
    if (lane == 0 && has_work) {
      const int dd = g.tap[tap][0], dh = g.tap[tap][1], dw = g.tap[tap][2];
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        int n0, d0, h0, w0;
        tile_origin(g, c, n0, d0, h0, w0);
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + static_cast<size_t>(s) * stage_bytes;
        uint8_t* sb = sa + 2 * kWgBoxBytes;
        mbar_expect_tx(&full_bar[s], (na + nb) * kWgBoxBytes);
        for (int i = 0; i < na; ++i)
          tma_load_5d(&tmDY, &full_bar[s], sa + i * kWgBoxBytes, co0 + i * 64, w0, h0, d0, n0);
        for (int i = 0; i < nb; ++i)
          tma_load_5d(&tmX, &full_bar[s], sb + i * kWgBoxBytes, ci0 + i * 64, w0 + dw, h0 + dh,
                      d0 + dd, n0);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // not guarded by has_work: an empty range simply runs no iterations, and the guard would make
       // ptxas treat the loop as divergent (no uniform-datapath MMA issue)
      const bool leader = elect_one();
      const uint32_t idesc = idesc_bf16_m128(p.block_n, true, true);
      const uint64_t d0 = sdesc_mnmajor128(smem_u32(smem), kWgBoxBytes);
      const uint32_t dlo0 = desc_lo(d0), dhi = desc_hi(d0);
      const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t alo = dlo0 + static_cast<uint32_t>(s) * stage16;
        const uint32_t blo = alo + ((2 * kWgBoxBytes) >> 4);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 128 voxels per chunk = 8 x K16; 16 rows = 2048 B = 128 x 16 B
            umma_bf16(tmem_base, desc_join(alo + 128 * k, dhi), desc_join(blo + 128 * k, dhi), idesc,
                      (c > c_begin) || (k != 0));
          umma_commit(&empty_bar[s]);
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(&acc_full);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // rows >= 64 hold garbage when only one dY box was loaded; they are masked by co < cout
    for (int c = 0; c < p.block_n; c += 16) {
      float v[16];
      tmem_ld16(tacc + c, v);
      if (co < p.cout) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int ci = ci0 + c + i;
          if (ci < p.cin)
            atomicAdd(p.acc + static_cast<size_t>(split) * p.det_stride + (static_cast<size_t>(tap) * p.ci_pad + ci) * p.co_pad + co, v[i]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ wgrad v2
// One CTA = (voxel range, 128 output channels, ci_n input channels, one d-plane of the filter). Per
// 8x16-voxel chunk it loads the dY tile and ONE halo plane of X; each of the plane's kh*kw taps has its
// own TMEM accumulator (kh*kw*ci_n <= 512 columns) fed from row-shifted MN-major descriptors into the
// halo plane. dY and X are each read once per chunk instead of once per tap.
struct Wg2Params {
  int N, D, H, W;
  int tilesW, tilesH;
  int kd, kh, kw;
  int cout, cin;
  int co_tiles, ci_tiles, ci_n;
  int nb;            // 64-channel X boxes per stage
  int splits, stages, tmem_cols;
  int plane_bytes, plane_stride, stage_bytes;  // plane_stride: plane_bytes rounded up to 1024
  int co_pad, ci_pad;
  int dbg;
  float* acc;
  long long det_stride;   // deterministic mode: split s adds into acc + s * det_stride (0 = off)
};

template <int KHW>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad2_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ Wg2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  int wi = blockIdx.x;
  const int a = wi % p.kd;
  wi /= p.kd;
  const int ct = wi % p.ci_tiles;
  wi /= p.ci_tiles;
  const int mt = wi % p.co_tiles;
  const int split = wi / p.co_tiles;

  const int khw = p.kh * p.kw;
  const int PWc = 8 + p.kw - 1;
  const int chunks = p.N * p.D * p.tilesH * p.tilesW;
  const int per = (chunks + p.splits - 1) / p.splits;
  const int c_begin = split * per;
  const int c_end = min(chunks, c_begin + per);
  const int co0 = mt * 128;
  const int ci0 = ct * p.ci_n;
  const int na = min(2, (p.cout - co0 + 63) / 64);
  const int nbv = min(p.nb, (p.cin - ci0 + 63) / 64);   // X boxes that contain valid channels
  const bool has_work = c_begin < c_end;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0 && has_work) {
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        int t = c;
        const int w0 = (t % p.tilesW) * 8;
        t /= p.tilesW;
        const int h0 = (t % p.tilesH) * 16;
        t /= p.tilesH;
        const int d = t % p.D;
        const int n = t / p.D;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + static_cast<size_t>(s) * p.stage_bytes;
        uint8_t* sb = sa + 2 * kWgBoxBytes;
        if VFD_DBG(p, 1) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], na * kWgBoxBytes + nbv * p.plane_bytes);
          for (int i = 0; i < na; ++i)
            tma_load_5d(&tmDY, &full_bar[s], sa + i * kWgBoxBytes, co0 + i * 64, w0, h0, d, n);
          for (int i = 0; i < nbv; ++i)
            tma_load_5d(&tmX, &full_bar[s], sb + static_cast<size_t>(i) * p.plane_stride, ci0 + i * 64,
                        w0 - p.kw / 2, h0 - p.kh / 2, d + a - p.kd / 2, n);
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // not guarded by has_work: an empty range simply runs no iterations, and the guard would make
       // ptxas treat the loop as divergent (no uniform-datapath MMA issue)
      // warp-uniform; the (tap, K16-step) sequence is fully unrolled with immediate descriptor offsets and
      // one elected lane issues
      const bool leader = elect_one();
      const bool issue = leader && !VFD_DBG(p, 2);
      constexpr int PWk = 8 + KHW - 1;
      const uint32_t idesc = idesc_bf16_m128(p.ci_n, true, true);
      const uint64_t da0 = sdesc_mnmajor128_ex(smem_u32(smem), kWgBoxBytes, 1024);
      const uint64_t db0 = sdesc_mnmajor128_ex(smem_u32(smem) + 2 * kWgBoxBytes, p.plane_stride, PWk * 128);
      const uint32_t alo0 = desc_lo(da0), ahi = desc_hi(da0), blo0 = desc_lo(db0), bhi = desc_hi(db0);
      const uint32_t stage16 = p.stage_bytes >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t alo = alo0 + static_cast<uint32_t>(s) * stage16;
        const uint32_t blo = blo0 + static_cast<uint32_t>(s) * stage16;
        const uint32_t accumulate = c > c_begin;
        if (issue) {
#pragma unroll
          for (int b = 0; b < KHW; ++b)
#pragma unroll
            for (int cc = 0; cc < KHW; ++cc) {
              const uint32_t tacc = tmem_base + (b * KHW + cc) * p.ci_n;
#pragma unroll
              for (int k = 0; k < 8; ++k)  // dY advances 2 h-rows = 2048 B per K16, X advances 2*PWk halo rows
                umma_bf16(tacc, desc_join(alo + 128 * k, ahi),
                          desc_join(blo + (b * PWk + cc) * 8 + 16 * PWk * k, bhi), idesc, k == 0 ? accumulate : 1u);
            }
        }
        if (leader) umma_commit(&empty_bar[s]);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(&acc_full);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int tapidx = 0; tapidx < (VFD_DBG(p, 4) ? 0 : khw); ++tapidx) {
      const int tap = a * khw + tapidx;
      for (int c = 0; c < p.ci_n; c += 16) {
        float v[16];
        tmem_ld16(tq + tapidx * p.ci_n + c, v);
        if (co < p.cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c + i;
            if (ci < p.cin) atomicAdd(p.acc + static_cast<size_t>(split) * p.det_stride + (static_cast<size_t>(tap) * p.ci_pad + ci) * p.co_pad + co, v[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ wgrad v3 (swapped)
// Same data flow as wgrad v2 with the GEMM roles swapped: M = 128 INPUT channels (A = the X halo plane, row-
// shifted per tap), N = co_n OUTPUT channels (B = dY), K = voxels. An SS-mode MMA re-reads its 128-row A tile
// from shared memory for every K16 step (32 cycles) whatever N is, so the wide dimension belongs on N: with
// N = cout (up to 256) instead of a TMEM-limited ci_n = 512 / 9 taps the same FLOPs take ~1.6-2.3x fewer MMA
// cycles on the decoder's spatial convs. TMEM holds `tpc` taps x co_n columns; the 9 spatial taps are split
// over `tgroups` CTAs. Accumulates into acc laid out [tap][co_pad][ci_pad] (lanes = consecutive ci).
struct Wg3Params {
  int N, D, H, W, tilesW, tilesH;
  int kd;
  int cout, cin;
  int ci_tiles;        // tiles of 128 input channels (GEMM M)
  int co_tiles, co_n;  // tiles of co_n output channels (GEMM N), co_n % 16 == 0, <= 256
  int tpc, tgroups;    // spatial taps per CTA / number of tap groups
  int nbY;             // 64-channel dY boxes per stage
  int splits, stages, tmem_cols;
  int plane_bytes, plane_stride, stage_bytes;
  int co_pad, ci_pad;
  int dbg;
  float* acc;
  long long det_stride;   // deterministic mode: split s adds into acc + s * det_stride (0 = off)
};

template <int KHW>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad3_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ Wg3Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  int wi = blockIdx.x;
  const int tg = wi % p.tgroups;
  wi /= p.tgroups;
  const int a = wi % p.kd;
  wi /= p.kd;
  const int nt = wi % p.co_tiles;
  wi /= p.co_tiles;
  const int mt = wi % p.ci_tiles;
  const int split = wi / p.ci_tiles;

  constexpr int KT = KHW * KHW;
  constexpr int PWk = 8 + KHW - 1;
  const int t0 = tg * p.tpc;
  const int ntaps = min(p.tpc, KT - t0);
  const int chunks = p.N * p.D * p.tilesH * p.tilesW;
  const int per = (chunks + p.splits - 1) / p.splits;
  const int c_begin = split * per;
  const int c_end = min(chunks, c_begin + per);
  const int ci0 = mt * 128;
  const int co0 = nt * p.co_n;
  const int nx = min(2, (p.cin - ci0 + 63) / 64);                   // X boxes with valid channels
  const int ny = min(p.nbY, (p.cout - co0 + 63) / 64);              // dY boxes with valid channels
  const bool has_work = c_begin < c_end;
  uint8_t* smem_y_off = nullptr;
  (void)smem_y_off;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0 && has_work) {
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        int t = c;
        const int w0 = (t % p.tilesW) * 8;
        t /= p.tilesW;
        const int h0 = (t % p.tilesH) * 16;
        t /= p.tilesH;
        const int d = t % p.D;
        const int n = t / p.D;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sx = smem + static_cast<size_t>(s) * p.stage_bytes;
        uint8_t* sy = sx + 2 * p.plane_stride;
        if VFD_DBG(p, 1) {
          mbar_arrive(&full_bar[s]);
        } else {
          mbar_expect_tx(&full_bar[s], nx * p.plane_bytes + ny * kWgBoxBytes);
          for (int i = 0; i < nx; ++i)
            tma_load_5d(&tmX, &full_bar[s], sx + static_cast<size_t>(i) * p.plane_stride, ci0 + i * 64,
                        w0 - KHW / 2, h0 - KHW / 2, d + a - p.kd / 2, n);
          for (int i = 0; i < ny; ++i)
            tma_load_5d(&tmDY, &full_bar[s], sy + i * kWgBoxBytes, co0 + i * 64, w0, h0, d, n);
        }
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform MMA issue (see conv_wgrad2_kernel); not guarded by has_work
      const bool leader = elect_one();
      const bool issue = leader && !VFD_DBG(p, 2);
      const uint32_t idesc = idesc_bf16_m128(p.co_n, true, true);
      const uint64_t da0 = sdesc_mnmajor128_ex(smem_u32(smem), p.plane_stride, PWk * 128);
      const uint64_t db0 = sdesc_mnmajor128_ex(smem_u32(smem) + 2 * p.plane_stride, kWgBoxBytes, 1024);
      const uint32_t alo0 = desc_lo(da0), ahi = desc_hi(da0), blo0 = desc_lo(db0), bhi = desc_hi(db0);
      const uint32_t stage16 = p.stage_bytes >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t alo = alo0 + static_cast<uint32_t>(s) * stage16;
        const uint32_t blo = blo0 + static_cast<uint32_t>(s) * stage16;
        const uint32_t accumulate = c > c_begin;
        for (int t = 0; t < ntaps; ++t) {
          const int tap = t0 + t;
          const int b = tap / KHW, cc = tap - b * KHW;
          const uint32_t at = alo + static_cast<uint32_t>((b * PWk + cc) * 8);  // tap shift: rows of 128 B
          const uint32_t tacc = tmem_base + t * p.co_n;
          if (issue) {
#pragma unroll
            for (int k = 0; k < 8; ++k)  // X advances 2*PWk halo rows per K16, dY advances 2 h-rows = 2048 B
              umma_bf16(tacc, desc_join(at + 16 * PWk * k, ahi), desc_join(blo + 128 * k, bhi), idesc,
                        k == 0 ? accumulate : 1u);
          }
        }
        if (leader) umma_commit(&empty_bar[s]);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(&acc_full);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const int ci = ci0 + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int t = 0; t < (VFD_DBG(p, 4) ? 0 : ntaps); ++t) {
      const int tap = a * KT + t0 + t;
      for (int c = 0; c < p.co_n; c += 16) {
        float v[16];
        tmem_ld16(tq + t * p.co_n + c, v);
        if (ci < p.cin) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = co0 + c + i;
            if (co < p.cout) atomicAdd(p.acc + static_cast<size_t>(split) * p.det_stride + (static_cast<size_t>(tap) * p.co_pad + co) * p.ci_pad + ci, v[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ wgrad, temporal (3x1x1)
// Weight gradient of the temporal convs: dW[a][co][ci] = sum_v dY[v][co] * X[v + (a-1)*H*W][ci]. A CTA walks
// whole (n, h-tile, w-tile) columns along d and keeps a ring of pipeline stages; stage j of a column holds the
// X tile of plane j and the dY tile of plane j-1. Output plane d = j-1 then finds X_{d-1}, X_d, X_{d+1} in the
// stages j-2, j-1, j and dY_d in stage j, so every X and dY tile is loaded exactly once (the per-tap CTAs of
// conv_wgrad2_kernel read each of them three times through L2). Three TMEM accumulators (one per tap).
struct WgtParams {
  int N, D, H, W, tilesW, tilesH;
  int cout, cin;
  int co_tiles, ci_tiles, ci_n;
  int na, nb;            // 64-channel dY / X boxes per stage
  int splits, stages, tmem_cols, stage_bytes;
  int co_pad, ci_pad;
  int dbg;
  float* acc;            // [3][ci_pad][co_pad]
  long long det_stride;  // deterministic mode: split s adds into acc + s * det_stride (0 = off)
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_t_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ WgtParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ uint64_t acc_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  int wi = blockIdx.x;
  const int ct = wi % p.ci_tiles;
  wi /= p.ci_tiles;
  const int mt = wi % p.co_tiles;
  const int split = wi / p.co_tiles;

  const int columns = p.N * p.tilesH * p.tilesW;
  const int per = (columns + p.splits - 1) / p.splits;
  const int col_begin = split * per;
  const int col_end = min(columns, col_begin + per);
  const int co0 = mt * 128;
  const int ci0 = ct * p.ci_n;
  const int nay = min(p.na, (p.cout - co0 + 63) / 64);   // dY boxes with valid channels
  const int nbx = min(p.nb, (p.cin - ci0 + 63) / 64);    // X boxes with valid channels
  const bool has_work = col_begin < col_end;
  const int D = p.D;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, p.tmem_cols);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (lane == 0 && has_work) {
      int s = 0;
      uint32_t ph = 0;
      for (int col = col_begin; col < col_end; ++col) {
        int t = col;
        const int w0 = (t % p.tilesW) * 8;
        t /= p.tilesW;
        const int h0 = (t % p.tilesH) * 16;
        const int n = t / p.tilesH;
        for (int j = 0; j <= D; ++j) {       // stage j: X plane j (j < D) and dY plane j-1 (j >= 1)
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sy = smem + static_cast<size_t>(s) * p.stage_bytes;
          uint8_t* sx = sy + p.na * kWgBoxBytes;
          if VFD_DBG(p, 1) {
            mbar_arrive(&full_bar[s]);
          } else {
            mbar_expect_tx(&full_bar[s], ((j >= 1 ? nay : 0) + (j < D ? nbx : 0)) * kWgBoxBytes);
            if (j >= 1)
              for (int i = 0; i < nay; ++i)
                tma_load_5d(&tmDY, &full_bar[s], sy + i * kWgBoxBytes, co0 + i * 64, w0, h0, j - 1, n);
            if (j < D)
              for (int i = 0; i < nbx; ++i)
                tma_load_5d(&tmX, &full_bar[s], sx + i * kWgBoxBytes, ci0 + i * 64, w0, h0, j, n);
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform MMA issue; not guarded by has_work (see conv_wgrad2_kernel)
      const bool leader = elect_one();
      const bool issue = leader && !VFD_DBG(p, 2);
      const uint32_t idesc = idesc_bf16_m128(p.ci_n, true, true);
      const uint64_t da0 = sdesc_mnmajor128_ex(smem_u32(smem), kWgBoxBytes, 1024);
      const uint64_t db0 = sdesc_mnmajor128_ex(smem_u32(smem) + p.na * kWgBoxBytes, kWgBoxBytes, 1024);
      const uint32_t alo0 = desc_lo(da0), ahi = desc_hi(da0), blo0 = desc_lo(db0), bhi = desc_hi(db0);
      const uint32_t stage16 = p.stage_bytes >> 4;
      int s = 0;            // ring index of the stage being consumed (stage j)
      uint32_t ph = 0;
      uint32_t started0 = 0, started1 = 0, started2 = 0;
      for (int col = col_begin; col < col_end; ++col) {
        for (int j = 0; j <= D; ++j) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const int s1 = s == 0 ? p.stages - 1 : s - 1;       // stage j-1
          const int s2 = s1 == 0 ? p.stages - 1 : s1 - 1;     // stage j-2
          if (j >= 1) {                                        // output plane d = j-1
            const int d = j - 1;
            const uint32_t alo = alo0 + static_cast<uint32_t>(s) * stage16;   // dY_d lives in stage j
            const uint32_t b0 = blo0 + static_cast<uint32_t>(s2) * stage16;   // X_{d-1}
            const uint32_t b1 = blo0 + static_cast<uint32_t>(s1) * stage16;   // X_d
            const uint32_t b2 = blo0 + static_cast<uint32_t>(s) * stage16;    // X_{d+1}
            const bool t0 = d >= 1, t2 = d + 1 < D;
            if (issue) {
              if (t0) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  umma_bf16(tmem_base, desc_join(alo + 128 * k, ahi), desc_join(b0 + 128 * k, bhi), idesc,
                            k == 0 ? started0 : 1u);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem_base + p.ci_n, desc_join(alo + 128 * k, ahi), desc_join(b1 + 128 * k, bhi), idesc,
                          k == 0 ? started1 : 1u);
              if (t2) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                  umma_bf16(tmem_base + 2 * p.ci_n, desc_join(alo + 128 * k, ahi), desc_join(b2 + 128 * k, bhi),
                            idesc, k == 0 ? started2 : 1u);
              }
            }
            if (t0) started0 = 1;
            started1 = 1;
            if (t2) started2 = 1;
          }
          // releases: stage j-2 is no longer needed once plane j-1 has been issued; the last plane of a
          // column also frees the two younger stages
          if (leader) {
            if (j >= 2) umma_commit(&empty_bar[s2]);
            if (j == D) {
              if (D >= 1) umma_commit(&empty_bar[s1]);
              umma_commit(&empty_bar[s]);
            }
          }
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
      if (leader) umma_commit(&acc_full);
    }
  } else if (has_work) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    const uint32_t tq = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int a = 0; a < (VFD_DBG(p, 4) ? 0 : 3); ++a) {
      if (D == 1 && a != 1) continue;   // taps that never received a plane hold no data
      for (int c = 0; c < p.ci_n; c += 16) {
        float v[16];
        tmem_ld16(tq + a * p.ci_n + c, v);
        if (co < p.cout) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c + i;
            if (ci < p.cin) atomicAdd(p.acc + static_cast<size_t>(split) * p.det_stride + (static_cast<size_t>(a) * p.ci_pad + ci) * p.co_pad + co, v[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  // cuTensorMapEncodeTiled is a driver call: the calling thread (e.g. an autograd worker that has
  // not launched anything yet) must have the primary context bound, which a runtime no-op does.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(0);
    ctx_bound = true;
  }
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 5-D map over a channels-last activation [N][D][H][W][ld] exposing `channels` channels.
static int make_act_map(CUtensorMap* tm, const void* ptr, long long ld, int channels, int N, int D,
                        int H, int W, int boxC, int TW, int TH, int TD, int TN) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VFD_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8) || (channels % 8))
    return set_error(VFD_ERR_ARG, "activation pointer/ld/channels must be 16-byte aligned");
  cuuint64_t dims[5] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D,
                        (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * W, (cuuint64_t)ld * 2 * W * H,
                           (cuuint64_t)ld * 2 * W * H * D};
  cuuint32_t box[5] = {(cuuint32_t)boxC, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TD,
                       (cuuint32_t)TN};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(boxC * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[384];
    snprintf(msg, sizeof(msg),
             "cuTensorMapEncodeTiled(activation) failed (%d): ptr %p ld %lld ch %d dims N%d D%d H%d W%d box C%d W%d H%d D%d N%d",
             (int)r, ptr, ld, channels, N, D, H, W, boxC, TW, TH, TD, TN);
    return set_error(VFD_ERR_DRIVER, msg);
  }
  return 0;
}

static int make_weight_map(CUtensorMap* tm, const void* ptr, int rows, long long kcols, int boxK,
                           int boxRows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VFD_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (kcols % 8))
    return set_error(VFD_ERR_ARG, "packed weights must be 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kcols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(boxK * 2),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(VFD_ERR_DRIVER, "cuTensorMapEncodeTiled(weights) failed");
  return 0;
}

static int fill_geom(ConvGeom& g, int N, int D, int H, int W, int kd, int kh, int kw) {
  if ((kd != 1 && kd != 3) || (kh != 1 && kh != 3) || (kw != 1 && kw != 3))
    return set_error(VFD_ERR_ARG, "kernel extents must be 1 or 3");
  g.N = N;
  g.D = D;
  g.H = H;
  g.W = W;
  // pick the 128-voxel box (powers of two) that needs the fewest tiles; prefer wide-in-W boxes
  long long best = -1;
  for (int tw = 1; tw <= 128; tw *= 2)
    for (int th = 1; tw * th <= 128; th *= 2)
      for (int td = 1; tw * th * td <= 128; td *= 2) {
        const int tn = 128 / (tw * th * td);
        const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th) *
                                ((D + td - 1) / td) * ((N + tn - 1) / tn);
        // tie-break: larger tw, then th (longer contiguous runs)
        const long long score = tiles * 1000000 - tw * 1000 - th * 10 - td;
        if (best < 0 || score < best) {
          best = score;
          g.TW = tw;
          g.TH = th;
          g.TD = td;
          g.TN = tn;
        }
      }
  g.tilesW = (W + g.TW - 1) / g.TW;
  g.tilesH = (H + g.TH - 1) / g.TH;
  g.tilesD = (D + g.TD - 1) / g.TD;
  g.tilesN = (N + g.TN - 1) / g.TN;
  g.ntaps = kd * kh * kw;
  int t = 0;
  for (int a = 0; a < kd; ++a)
    for (int b = 0; b < kh; ++b)
      for (int c = 0; c < kw; ++c) {
        g.tap[t][0] = (int8_t)(a - kd / 2);
        g.tap[t][1] = (int8_t)(b - kh / 2);
        g.tap[t][2] = (int8_t)(c - kw / 2);
        g.tap[t][3] = 0;
        ++t;
      }
  return 0;
}

static int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// ring-buffer budgets of the streaming kernels (diagnostics: VFD_TC_BUDGET_KB for conv_fwd_tc, <= 208 because of its 16 KB
// of static statistics; VFD_WG_BUDGET_KB for the weight-gradient kernels, <= 222)
static int env_kb(const char* name, int def_kb, int lo, int hi) {
  if (const char* e = getenv(name)) {
    const int kb = atoi(e);
    if (kb >= lo && kb <= hi) return kb * 1024;
  }
  return def_kb * 1024;
}
static int tc_budget() {
  static int b = env_kb("VFD_TC_BUDGET_KB", 200, 64, 208);
  return b;
}
static int wg_budget() {
  static int b = env_kb("VFD_WG_BUDGET_KB", 200, 64, 222);
  return b;
}

template <int KC>
static int launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, FwdParams& p,
                      cudaStream_t stream) {
  const int stage_bytes = ((kTileM + p.block_n) * KC * 2 + 1023) & ~1023;
  int stages = tc_budget() / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(VFD_ERR_ARG, "conv tile does not fit in shared memory");
  p.stages = stages;
  // >113 KB of dynamic smem keeps the kernel at one CTA per SM (each CTA owns all 512 TMEM columns)
  size_t smem = (size_t)stages * stage_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_tc_kernel<KC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);   // + 16.2 KB static
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_fwd_tc)");
    attr_set = true;
  }
  const ConvGeom& g = p.g;
  const long long tiles = (long long)g.tilesW * g.tilesH * g.tilesD * g.tilesN * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv_fwd_tc_kernel<KC><<<grid, kFwdThreads, smem, stream>>>(tmA, tmB, p);
  return check_launch("conv_fwd_tc");
}

// dynamic shared memory the resident-weight kernel may plan with: + 1 KB alignment slack + 208 B static <= 227 KB.
// VFD_RES_BUDGET_KB (diagnostics) overrides it within [64, 225].
static int res_budget() {
  static int b = 0;
  if (!b) {
    b = 225 * 1024;   // 223 -> 225 KB: cfg2 34.79 -> 34.50 ms, cfg4 45.58 -> 44.88 ms (one more input slot on some layers)
    if (const char* e = getenv("VFD_RES_BUDGET_KB")) {
      const int kb = atoi(e);
      if (kb >= 64 && kb <= 225) b = kb * 1024;
    }
  }
  return b;
}

// 5-D map over the conv output for the epilogue's TMA stores: box = 32 channels x 8 (w) x 4 (h) voxels,
// i.e. the 32 accumulator rows one epilogue warp owns.
static int make_out_map(CUtensorMap* tm, const void* ptr, long long ld, int cols, int fp32, int N, int D, int H,
                        int W) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(VFD_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es = fp32 ? 4 : 2;
  cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)ld * es, (cuuint64_t)ld * es * W, (cuuint64_t)ld * es * W * H,
                           (cuuint64_t)ld * es * W * H * D};
  cuuint32_t box[5] = {32, 8, 4, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   fp32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[256];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled(output) failed (%d): ptr %p ld %lld cols %d fp32 %d", (int)r,
             ptr, ld, cols, fp32);
    return set_error(VFD_ERR_DRIVER, msg);
  }
  return 0;
}

static int g_dbg = 0;

// Resident-weight kernel: returns 1 when the geometry does not qualify (caller falls back), 0 when it
// launched or failed (status in *err).
template <int KC, int KHW>
static int try_launch_res(const void* x, long long x_ld, int cin, const void* w_packed, int w_rows, int cin_k,
                          const Epilogue& epi, int N, int D, int H, int W, int kd, int kh, int kw,
                          cudaStream_t stream, int* err) {
  *err = 0;
  Res2Params p;
  const int PWc = 8 + kw - 1, PHc = 16 + kh - 1;
  const int ntaps = kd * kh * kw;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.tilesW = (W + 7) / 8;
  p.tilesH = (H + 15) / 16;
  p.kd = kd; p.kh = kh; p.kw = kw;
  p.cblocks = cin_k / KC;
  p.block_n = w_rows;
  p.b_tile_bytes = (p.block_n * KC * 2 + 1023) & ~1023;
  const long long b_total = (long long)ntaps * p.cblocks * p.b_tile_bytes;
  if ((long long)ntaps * p.cblocks * p.block_n * KC * 2 >= (1 << 20)) return 1;  // mbarrier tx-count limit
  p.acc_stride = (p.block_n + 31) & ~31;
  const int mmas_per_sub = ntaps * p.cblocks * (KC / 16);
  int g0 = (kd == 3 || mmas_per_sub < 24) ? 4 : 1;
  if (kd == 1 && mmas_per_sub < 24 && p.acc_stride <= 32) g0 = 8;   // thin N: amortise the hand-shakes further
  if (const char* e = getenv("VFD_RES_G")) {   // diagnostics: force the sub-tile count
    const int f = atoi(e);
    if (f == 1 || f == 2 || f == 4 || f == 8) g0 = f;
  }
  while (g0 > D) g0 >>= 1;
  const int stage_bytes = epi.out_fp32 ? 4096 : 2048;  // 32 rows x 128 / 64 B
  // pick (G, staging buffers): first choice with >= 3 input slots, else the first with >= 2
  int bestG = 0, bestBufs = 0, bestSlots = 0;
  int max_bufs = 2;
  if (const char* e = getenv("VFD_RES_BUFS")) {   // diagnostics: staging buffers per epilogue warp
    const int f = atoi(e);
    if (f == 1 || f == 2 || f == 4) max_bufs = f;
  }
  for (int pass = 0; pass < 2 && !bestG; ++pass)
    for (int G = g0; G >= 1 && !bestG; G >>= 1)
      for (int bufs = max_bufs; bufs >= 1 && !bestG; bufs >>= 1) {
        if (G * p.acc_stride * 2 > 512) continue;
        const long long box = (long long)(G + kd - 1) * PWc * PHc * KC * 2;
        const long long slot = (box + 1023) & ~1023LL;
        const long long avail = (long long)res_budget() - b_total - 8LL * bufs * stage_bytes;
        if (avail <= 0 || box >= (1 << 20)) continue;
        const int slots = (int)(avail / slot);
        if (slots >= (pass == 0 ? 3 : 2)) {
          bestG = G; bestBufs = bufs; bestSlots = slots;
        }
      }
  if (!bestG) return 1;
  p.G = bestG;
  p.stage_bufs = bestBufs;
  p.a_slots = bestSlots > kResMaxASlots ? kResMaxASlots : bestSlots;
  p.a_box_bytes = (p.G + kd - 1) * PWc * PHc * KC * 2;
  p.a_slot_bytes = (p.a_box_bytes + 1023) & ~1023;
  p.dgroups = (D + p.G - 1) / p.G;
  p.acc_stages = 512 / (p.G * p.acc_stride);
  if (p.acc_stages > 4) p.acc_stages = 4;
  p.nchunks = (p.block_n + 31) / 32;
  p.out_fp32 = epi.out_fp32;
  p.n_rows = epi.n_rows;
  p.dbg = g_dbg;
  p.bias = epi.bias;
  p.stats = epi.stats;
  p.stats_ld = epi.stats_ld;
  p.stage_bytes = stage_bytes;
  // measured (profiles/r2_thin_epilogue.txt): rows of <= 48 bytes gain 15-25 % from the direct store, full 64-byte
  // rows lose 10-18 % (the TMA store is asynchronous, the register store holds the epilogue warp)
  static const int direct_env = getenv("VFD_RES_DIRECT") ? atoi(getenv("VFD_RES_DIRECT")) : -1;
  p.direct_out = p.nchunks == 1 && !epi.out_fp32 &&
                 (direct_env == 1 || (direct_env != 0 && epi.out_cols <= 24));
  p.out = epi.out;
  p.out_ld = epi.out_ld;
  p.out_cols = epi.out_cols;
  CUtensorMap tmA, tmB, tmC;
  if ((*err = make_act_map(&tmA, x, x_ld, cin, N, D, H, W, KC, PWc, PHc, p.G + kd - 1, 1))) return 0;
  if ((*err = make_weight_map(&tmB, w_packed, w_rows, (long long)ntaps * cin_k, KC, p.block_n))) return 0;
  if ((*err = make_out_map(&tmC, epi.out, epi.out_ld, epi.out_cols, epi.out_fp32, N, D, H, W))) return 0;
  size_t smem = (size_t)b_total + 8u * p.stage_bufs * stage_bytes + (size_t)p.a_slots * p.a_slot_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM: each CTA owns all 512 TMEM columns
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_res_kernel<KC, KHW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         226 * 1024);
    if (e != cudaSuccess) {
      *err = set_cuda_error(e, "cudaFuncSetAttribute(conv_fwd_res)");
      return 0;
    }
    attr_set = true;
  }
  const long long tiles = (long long)N * p.dgroups * p.tilesH * p.tilesW;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv_fwd_res_kernel<KC, KHW><<<grid, kRes2Threads, smem, stream>>>(tmA, tmB, tmC, p);
  *err = check_launch("conv_fwd_res");
  return 0;
}

// MMA-cycle model of one (tap, K16 step) of the weight-gradient GEMM: tiles x (A re-read 32 + N/4 cycles).
struct WgradPlan {
  int swap;                 // 1: M = input channels (conv_wgrad3_kernel), acc is [tap][co][ci]
  int ci_tiles, ci_n;       // v2: N tiling of the input channels
  int co_tiles, co_n;       // v3: N tiling of the output channels
  int tpc, tgroups;
};
static WgradPlan plan_wgrad(int cout, int cin, int kh) {
  WgradPlan w;
  const int khw = kh * kh;
  const int cin16 = (cin + 15) & ~15, cout16 = (cout + 15) & ~15;
  int max_n = (512 / khw) & ~15;
  if (max_n > 256) max_n = 256;
  w.ci_tiles = (cin16 + max_n - 1) / max_n;
  w.ci_n = (((cin16 + w.ci_tiles - 1) / w.ci_tiles) + 15) & ~15;
  const double cost2 = (double)((cout + 127) / 128) * w.ci_tiles * (32.0 + w.ci_n / 4.0);
  w.co_tiles = (cout16 + 191) / 192;   // <= 3 dY boxes per stage, so two pipeline stages fit in shared memory
  w.co_n = (((cout16 + w.co_tiles - 1) / w.co_tiles) + 15) & ~15;
  w.tpc = 512 / w.co_n;
  if (w.tpc > khw) w.tpc = khw;
  w.tgroups = (khw + w.tpc - 1) / w.tpc;
  const double cost3 = (double)((cin + 127) / 128) * w.co_tiles * (32.0 + w.co_n / 4.0);
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("VFD_CONV_WGRAD3");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  w.swap = enabled && kh == 3 && cost3 < 0.85 * cost2;
  return w;
}

// VFD_CONV_WGRADT=0 disables the temporal ring wgrad kernel (debug / A-B timing)
static bool wgrad_t_enabled() {
  static int mode = -1;
  if (mode == -1) {
    const char* e = getenv("VFD_CONV_WGRADT");
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode == 1;
}

// VFD_CONV_WGRAD2=0 disables the multi-tap halo wgrad kernel (debug / A-B timing)
static bool wgrad2_enabled() {
  static int mode = -1;
  if (mode == -1) {
    const char* e = getenv("VFD_CONV_WGRAD2");
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode == 1;
}

// VFD_CONV_RES=0 disables the resident-weight kernel (debug / A-B timing)
static bool res_enabled() {
  static int mode = -1;
  if (mode == -1) {
    const char* e = getenv("VFD_CONV_RES");
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode == 1;
}

}  // namespace vfd

namespace vfd {
int launch_tiny_pointwise(const void* x, long long x_ld, const void* w_packed, int cin_k, const float* bias, void* out,
                          long long out_ld, long long V, double* stats, int stats_ld, cudaStream_t stream);
}

using namespace vfd;

#ifdef VFD_DEBUG
VFD_API int vfd_set_debug(int flags) {
  g_dbg = flags;
  return 0;
}
#endif

VFD_API int vfd_conv3d_fwd(const void* x, long long x_ld, int cin, const void* w_packed,
                              int w_rows, int cin_k, const float* bias, void* out,
                              long long out_ld, int out_cols, int out_fp32, double* stats,
                              int stats_ld, int N, int D, int H, int W, int kd, int kh, int kw, int kc,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (stats != nullptr && (out_fp32 || stats_ld < out_cols || stats_ld > 1024))
    return set_error(VFD_ERR_ARG, "conv3d_fwd: fused statistics need a bf16 output and out_cols <= stats_ld <= 1024");
  if (kc != 16 && kc != 32 && kc != 64) return set_error(VFD_ERR_ARG, "kc must be 16, 32 or 64");
  if (cin_k % kc || w_rows % 16 || out_cols % 8 || out_cols > w_rows + 8)
    return set_error(VFD_ERR_ARG, "conv3d_fwd: bad channel padding");
  if ((out_fp32 ? (out_ld % 4) : (out_ld % 8)) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(VFD_ERR_ARG, "conv3d_fwd: output must be 16-byte aligned");
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;  // empty batch: nothing to do
  if (kd == 1 && kh == 1 && kw == 1 && cin <= 8 && out_cols == 8 && w_rows == 16 && !out_fp32 && !VFD_GDBG(64) &&
      (x_ld % 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    return launch_tiny_pointwise(x, x_ld, w_packed, cin_k, bias, out, out_ld, (long long)N * D * H * W, stats,
                                 stats_ld, stream);
  Epilogue epi;
  epi.n_rows = w_rows;
  epi.out_cols = out_cols;
  epi.out_ld = out_ld;
  epi.out_fp32 = out_fp32;
  epi.bias = bias;
  epi.out = out;
  epi.stats = stats;
  epi.stats_ld = stats_ld;
  epi.lstm_j = 0; epi.lstm_hid = 0; epi.lstm_c_cur = nullptr; epi.lstm_c_next = nullptr; epi.lstm_act = nullptr;
  epi.lstm_h = nullptr; epi.lstm_h_ld = 0;
  if ((kd != 1 && kd != 3) || (kh != 1 && kh != 3) || (kw != 1 && kw != 3))
    return set_error(VFD_ERR_ARG, "kernel extents must be 1 or 3");
  if (res_enabled() && w_rows <= 256 && (stats == nullptr || stats_ld <= 256)) {
    // tiles are 8 x 16 voxels of one (n, d) plane: require a reasonable fill
    const double fill = (double)W * H / ((double)((W + 7) / 8) * 8 * ((H + 15) / 16) * 16);
    if (fill >= 0.7 && kh == kw) {
      int err = 0, fb;
#define VFD_RES_ARGS x, x_ld, cin, w_packed, w_rows, cin_k, epi, N, D, H, W, kd, kh, kw, stream, &err
      if (kh == 3) {
        if (kc == 64) fb = try_launch_res<64, 3>(VFD_RES_ARGS);
        else if (kc == 32) fb = try_launch_res<32, 3>(VFD_RES_ARGS);
        else fb = try_launch_res<16, 3>(VFD_RES_ARGS);
      } else {
        if (kc == 64) fb = try_launch_res<64, 1>(VFD_RES_ARGS);
        else if (kc == 32) fb = try_launch_res<32, 1>(VFD_RES_ARGS);
        else fb = try_launch_res<16, 1>(VFD_RES_ARGS);
      }
#undef VFD_RES_ARGS
      if (!fb) return err;
    }
  }
  FwdParams p;
  if (int e = fill_geom(p.g, N, D, H, W, kd, kh, kw)) return e;
  p.cblocks = cin_k / kc;
  p.n_tiles = (w_rows + 255) / 256;
  p.block_n = (((w_rows + p.n_tiles - 1) / p.n_tiles) + 15) & ~15;
  p.epi = epi;
  CUtensorMap tmA, tmB;
  if (int e = make_act_map(&tmA, x, x_ld, cin, N, D, H, W, kc, p.g.TW, p.g.TH, p.g.TD, p.g.TN))
    return e;
  if (int e = make_weight_map(&tmB, w_packed, w_rows, (long long)p.g.ntaps * cin_k, kc, p.block_n))
    return e;
  if (kc == 64) return launch_fwd<64>(tmA, tmB, p, stream);
  if (kc == 32) return launch_fwd<32>(tmA, tmB, p, stream);
  return launch_fwd<16>(tmA, tmB, p, stream);
}

static bool wgrad_halo_ok(int H, int W, int kh, int kw) {
  const int tilesW = (W + 7) / 8, tilesH = (H + 15) / 16;
  const double fill = (double)W * H / ((double)tilesW * 8 * tilesH * 16);
  return fill >= 0.7 && kh == kw;
}

VFD_API int vfd_conv3d_wgrad_layout(int cout, int cin, int kd, int kh, int kw, int H, int W) {
  (void)kd;
  if ((kh != 1 && kh != 3) || kh != kw || !wgrad2_enabled() || !wgrad_halo_ok(H, W, kh, kw)) return 0;
  return plan_wgrad(cout, cin, kh).swap;
}

static int launch_wgrad3(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, int cin,
                         float* acc, int co_pad, int ci_pad, int N, int D, int H, int W, int kd, int kh,
                         const WgradPlan& w, cudaStream_t stream, long long det_stride, int* splits_out) {
  Wg3Params q;
  q.N = N; q.D = D; q.H = H; q.W = W;
  q.tilesW = (W + 7) / 8; q.tilesH = (H + 15) / 16;
  q.kd = kd; q.cout = cout; q.cin = cin;
  q.ci_tiles = (cin + 127) / 128;
  q.co_tiles = w.co_tiles; q.co_n = w.co_n; q.tpc = w.tpc; q.tgroups = w.tgroups;
  q.nbY = (w.co_n + 63) / 64;
  q.plane_bytes = (16 + kh - 1) * (8 + kh - 1) * 128;
  q.plane_stride = (q.plane_bytes + 1023) & ~1023;
  q.stage_bytes = 2 * q.plane_stride + q.nbY * kWgBoxBytes;
  int stages = wg_budget() / q.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(VFD_ERR_ARG, "conv3d_wgrad: swapped tile does not fit in shared memory");
  q.stages = stages;
  q.tmem_cols = 32;
  while (q.tmem_cols < q.tpc * q.co_n) q.tmem_cols *= 2;
  const long long chunks = (long long)N * D * q.tilesH * q.tilesW;
  const long long base = (long long)kd * q.ci_tiles * q.co_tiles * q.tgroups;
  long long splits = (2LL * num_sms()) / base;
  if (splits < 1) splits = (base <= num_sms()) ? num_sms() / base : 1;
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  q.splits = (int)splits;
  if (splits_out != nullptr) { *splits_out = q.splits; return 0; }
  q.co_pad = co_pad; q.ci_pad = ci_pad; q.acc = acc; q.dbg = g_dbg; q.det_stride = det_stride;
  const int dy_ch = (cout + 7) & ~7, x_ch = (cin + 7) & ~7;
  CUtensorMap tmDY, tmX;
  if (int e = make_act_map(&tmDY, dy, dy_ld, dy_ch, N, D, H, W, 64, 8, 16, 1, 1)) return e;
  if (int e = make_act_map(&tmX, x, x_ld, x_ch, N, D, H, W, 64, 8 + kh - 1, 16 + kh - 1, 1, 1)) return e;
  auto kfn = conv_wgrad3_kernel<3>;
  static bool attr3 = false;
  if (!attr3) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_wgrad3)");
    attr3 = true;
  }
  size_t smem = (size_t)stages * q.stage_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;
  kfn<<<(unsigned)(base * q.splits), kWgThreads, smem, stream>>>(tmDY, tmX, q);
  return check_launch("conv_wgrad3");
}

// One body for the three entry points: the plain launch (det_stride = 0, splits_out = nullptr), the split-count
// query of the deterministic variant (splits_out != nullptr: nothing is launched) and its launch (acc = the partial
// buffers, det_stride = elements per partial).
static int wgrad_dispatch(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, int cin,
                          float* acc, int co_pad, int ci_pad, int layout, int N, int D, int H, int W, int kd, int kh,
                          int kw, cudaStream_t stream, long long det_stride, int* splits_out) {
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) {
    if (splits_out != nullptr) *splits_out = 0;
    return 0;
  }
  if (co_pad < cout || ci_pad < cin) return set_error(VFD_ERR_ARG, "conv3d_wgrad: bad acc padding");
  if ((kd != 1 && kd != 3) || (kh != 1 && kh != 3) || (kw != 1 && kw != 3))
    return set_error(VFD_ERR_ARG, "kernel extents must be 1 or 3");
  if (layout == 1) {
    if (!vfd_conv3d_wgrad_layout(cout, cin, kd, kh, kw, H, W))
      return set_error(VFD_ERR_ARG, "conv3d_wgrad: layout 1 ([tap][co][ci]) only as reported by vfd_conv3d_wgrad_layout");
    return launch_wgrad3(dy, dy_ld, cout, x, x_ld, cin, acc, co_pad, ci_pad, N, D, H, W, kd, kh,
                         plan_wgrad(cout, cin, kh), stream, det_stride, splits_out);
  }
  if (layout != 0) return set_error(VFD_ERR_ARG, "conv3d_wgrad: layout must be 0 or 1");
  if (kd == 3 && kh == 1 && kw == 1 && wgrad_t_enabled() && wgrad_halo_ok(H, W, 1, 1) && D >= 2) {
    WgtParams q;
    q.N = N; q.D = D; q.H = H; q.W = W;
    q.tilesW = (W + 7) / 8; q.tilesH = (H + 15) / 16;
    q.cout = cout; q.cin = cin;
    const int cin16 = (cin + 15) & ~15;
    const int max_n = 160;                                   // 3 taps x ci_n <= 512 TMEM columns
    q.ci_tiles = (cin16 + max_n - 1) / max_n;
    q.ci_n = (((cin16 + q.ci_tiles - 1) / q.ci_tiles) + 15) & ~15;
    q.co_tiles = (cout + 127) / 128;
    q.na = cout > 64 ? 2 : 1;
    q.nb = (q.ci_n + 63) / 64;
    q.stage_bytes = (q.na + q.nb) * kWgBoxBytes;
    int stages = wg_budget() / q.stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages >= 4) {
      q.stages = stages;
      q.tmem_cols = 32;
      while (q.tmem_cols < 3 * q.ci_n) q.tmem_cols *= 2;
      const long long columns = (long long)N * q.tilesH * q.tilesW;
      const long long base = (long long)q.co_tiles * q.ci_tiles;
      long long splits = (2LL * num_sms()) / base;
      if (splits < 1) splits = (base <= num_sms()) ? num_sms() / base : 1;
      if (splits > columns) splits = columns;
      if (splits < 1) splits = 1;
      q.splits = (int)splits;
      if (splits_out != nullptr) { *splits_out = q.splits; return 0; }
      q.co_pad = co_pad; q.ci_pad = ci_pad; q.acc = acc; q.dbg = g_dbg; q.det_stride = det_stride;
      const int dy_ch = (cout + 7) & ~7, x_ch = (cin + 7) & ~7;
      CUtensorMap tmDY, tmX;
      if (int e = make_act_map(&tmDY, dy, dy_ld, dy_ch, N, D, H, W, 64, 8, 16, 1, 1)) return e;
      if (int e = make_act_map(&tmX, x, x_ld, x_ch, N, D, H, W, 64, 8, 16, 1, 1)) return e;
      static bool attrt = false;
      if (!attrt) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             224 * 1024);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_wgrad_t)");
        attrt = true;
      }
      size_t smem = (size_t)stages * q.stage_bytes + 1024;
      if (smem < 120 * 1024) smem = 120 * 1024;
      conv_wgrad_t_kernel<<<(unsigned)(base * q.splits), kWgThreads, smem, stream>>>(tmDY, tmX, q);
      return check_launch("conv_wgrad_t");
    }
  }
  {
    const int tilesW = (W + 7) / 8, tilesH = (H + 15) / 16;
    const double fill = (double)W * H / ((double)tilesW * 8 * tilesH * 16);
    if (wgrad2_enabled() && fill >= 0.7 && kh == kw) {
      Wg2Params q;
      q.N = N; q.D = D; q.H = H; q.W = W; q.tilesW = tilesW; q.tilesH = tilesH;
      q.kd = kd; q.kh = kh; q.kw = kw; q.cout = cout; q.cin = cin;
      const int khw = kh * kw;
      const int cin16 = (cin + 15) & ~15;
      int max_n = (512 / khw) & ~15;            // TMEM: khw accumulators of ci_n columns
      if (max_n > 256) max_n = 256;
      q.ci_tiles = (cin16 + max_n - 1) / max_n;
      q.ci_n = (((cin16 + q.ci_tiles - 1) / q.ci_tiles) + 15) & ~15;
      q.co_tiles = (cout + 127) / 128;
      q.nb = (q.ci_n + 63) / 64;
      q.plane_bytes = (16 + kh - 1) * (8 + kw - 1) * 128;
      q.plane_stride = (q.plane_bytes + 1023) & ~1023;
      q.stage_bytes = 2 * kWgBoxBytes + q.nb * q.plane_stride;
      int stages = wg_budget() / q.stage_bytes;
      if (stages > kMaxStages) stages = kMaxStages;
      q.tmem_cols = 32;
      while (q.tmem_cols < khw * q.ci_n) q.tmem_cols *= 2;
      if (stages >= 2 && q.tmem_cols <= 512) {
        q.stages = stages;
        const long long chunks = (long long)N * D * tilesH * tilesW;
        const long long base = (long long)kd * q.co_tiles * q.ci_tiles;
        // one CTA per SM at a time: size the grid to whole waves (2 when the (co, ci, kd) tiling allows it)
        long long splits = (2LL * num_sms()) / base;
        if (splits < 1) splits = (base <= num_sms()) ? num_sms() / base : 1;
        if (splits > chunks) splits = chunks;
        if (splits < 1) splits = 1;
        q.splits = (int)splits;
        if (splits_out != nullptr) { *splits_out = q.splits; return 0; }
        q.co_pad = co_pad; q.ci_pad = ci_pad; q.acc = acc; q.dbg = g_dbg; q.det_stride = det_stride;
        const int dy_ch = (cout + 7) & ~7, x_ch = (cin + 7) & ~7;
        CUtensorMap tmDY, tmX;
        if (int e = make_act_map(&tmDY, dy, dy_ld, dy_ch, N, D, H, W, 64, 8, 16, 1, 1)) return e;
        if (int e = make_act_map(&tmX, x, x_ld, x_ch, N, D, H, W, 64, 8 + kw - 1, 16 + kh - 1, 1, 1)) return e;
        auto kfn = kh == 3 ? conv_wgrad2_kernel<3> : conv_wgrad2_kernel<1>;
        static bool attr2[2] = {false, false};
        if (!attr2[kh == 3]) {
          cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
          if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_wgrad2)");
          attr2[kh == 3] = true;
        }
        size_t smem = (size_t)stages * q.stage_bytes + 1024;
        if (smem < 120 * 1024) smem = 120 * 1024;  // one CTA per SM (TMEM and wave accounting below assume it)
        kfn<<<(unsigned)(base * q.splits), kWgThreads, smem, stream>>>(tmDY, tmX, q);
        return check_launch("conv_wgrad2");
      }
    }
  }
  WgradParams p;
  if (int e = fill_geom(p.g, N, D, H, W, kd, kh, kw)) return e;
  p.cout = cout;
  p.cin = cin;
  p.co_tiles = (cout + 127) / 128;
  const int cin16 = (cin + 15) & ~15;
  p.ci_tiles = (cin16 + 255) / 256;
  p.block_n = (((cin16 + p.ci_tiles - 1) / p.ci_tiles) + 15) & ~15;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.block_n) p.tmem_cols *= 2;
  p.co_pad = co_pad;
  p.ci_pad = ci_pad;
  p.acc = acc;
  p.det_stride = det_stride;
  const int nb_slots = (p.block_n + 63) / 64;
  const int stage_bytes = (2 + nb_slots) * kWgBoxBytes;
  int stages = wg_budget() / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(VFD_ERR_ARG, "wgrad tile does not fit in shared memory");
  p.stages = stages;
  const long long chunks = (long long)p.g.tilesW * p.g.tilesH * p.g.tilesD * p.g.tilesN;
  const long long base = (long long)p.g.ntaps * p.co_tiles * p.ci_tiles;
  long long splits = (2LL * num_sms() + base - 1) / base;  // aim for about two waves of CTAs
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  if (splits_out != nullptr) { *splits_out = p.splits; return 0; }
  // cin/cout here are the valid counts; the TMA maps expose the channel-padded widths
  const int dy_ch = (cout + 7) & ~7, x_ch = (cin + 7) & ~7;
  CUtensorMap tmDY, tmX;
  if (int e = make_act_map(&tmDY, dy, dy_ld, dy_ch, N, D, H, W, 64, p.g.TW, p.g.TH, p.g.TD, p.g.TN))
    return e;
  if (int e = make_act_map(&tmX, x, x_ld, x_ch, N, D, H, W, 64, p.g.TW, p.g.TH, p.g.TD, p.g.TN))
    return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_wgrad_tc)");
    attr_set = true;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  const long long grid = base * p.splits;
  conv_wgrad_tc_kernel<<<(unsigned)grid, kWgThreads, smem, stream>>>(tmDY, tmX, p);
  return check_launch("conv_wgrad_tc");
}

VFD_API int vfd_conv3d_wgrad(const void* dy, long long dy_ld, int cout, const void* x,
                                long long x_ld, int cin, float* acc, int co_pad, int ci_pad, int layout, int N,
                                int D, int H, int W, int kd, int kh, int kw, void* stream_) {
  return wgrad_dispatch(dy, dy_ld, cout, x, x_ld, cin, acc, co_pad, ci_pad, layout, N, D, H, W, kd, kh, kw,
                        reinterpret_cast<cudaStream_t>(stream_), 0, nullptr);
}

// ---- deterministic variant: per-split partial accumulators + an ordered second pass -------------------------------
namespace vfd {
namespace {
// acc[i] += partial[0][i] + partial[1][i] + ... in split order (one thread per element, coalesced over i)
__global__ void __launch_bounds__(256)
wgrad_ordered_reduce_kernel(const float* __restrict__ partial, long long elems, int splits, float* __restrict__ acc) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < elems; i += gridDim.x * 256ll) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<size_t>(k) * elems + i];
    acc[i] += s;
  }
}
}  // namespace
}  // namespace vfd

VFD_API long long vfd_conv3d_wgrad_det_workspace(int cout, int cin, int co_pad, int ci_pad, int layout, int N, int D,
                                                 int H, int W, int kd, int kh, int kw) {
  int splits = 0;
  if (wgrad_dispatch(nullptr, 0, cout, nullptr, 0, cin, nullptr, co_pad, ci_pad, layout, N, D, H, W, kd, kh, kw,
                     nullptr, 0, &splits))
    return -1;
  return 4ll * splits * kd * kh * kw * co_pad * ci_pad;
}

VFD_API int vfd_conv3d_wgrad_det(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, int cin,
                                 float* acc, int co_pad, int ci_pad, int layout, int N, int D, int H, int W, int kd,
                                 int kh, int kw, void* workspace, long long ws_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int splits = 0;
  if (int e = wgrad_dispatch(nullptr, 0, cout, nullptr, 0, cin, nullptr, co_pad, ci_pad, layout, N, D, H, W, kd, kh,
                             kw, nullptr, 0, &splits))
    return e;
  if (splits == 0) return 0;
  const long long elems = (long long)kd * kh * kw * co_pad * ci_pad;
  if (workspace == nullptr || ws_bytes < 4ll * splits * elems || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return set_error(VFD_ERR_ARG, "conv3d_wgrad_det: workspace too small (vfd_conv3d_wgrad_det_workspace)");
  cudaError_t ce = cudaMemsetAsync(workspace, 0, 4ull * splits * elems, stream);
  if (ce != cudaSuccess) return set_cuda_error(ce, "conv3d_wgrad_det: memset");
  if (int e = wgrad_dispatch(dy, dy_ld, cout, x, x_ld, cin, static_cast<float*>(workspace), co_pad, ci_pad, layout, N,
                             D, H, W, kd, kh, kw, stream, elems, nullptr))
    return e;
  long long blocks = (elems + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_ordered_reduce_kernel<<<(int)blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), elems, splits, acc);
  return check_launch("wgrad_ordered_reduce");
}

// ---- ConvLSTM step: gate conv (Conv2d == kd 1) with the cell update in its epilogue --------------------------------
VFD_API int vfd_convlstm_step_fwd(const void* comb, long long comb_ld, int cin, const void* w_packed_perm, int cin_k,
                                  const float* bias_perm, const float* c_cur, int hid, float* c_next, void* h_out,
                                  long long h_ld, float* act, int N, int H, int W, int kh, int kw, int kc,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (N <= 0 || H <= 0 || W <= 0) return 0;
  if (comb == nullptr || w_packed_perm == nullptr || c_cur == nullptr || c_next == nullptr || h_out == nullptr)
    return set_error(VFD_ERR_ARG, "convlstm_step_fwd: null pointer");
  if (hid <= 0 || hid % 64) return set_error(VFD_ERR_ARG, "convlstm_step_fwd: hidden_dim must be a multiple of 64");
  if ((kh != 1 && kh != 3) || kh != kw) return set_error(VFD_ERR_ARG, "convlstm_step_fwd: kernel must be 1x1 or 3x3");
  if (kc != 16 && kc != 32 && kc != 64) return set_error(VFD_ERR_ARG, "kc must be 16, 32 or 64");
  if (cin_k % kc || comb_ld % 8 || h_ld % 8 || h_ld < hid || (reinterpret_cast<uintptr_t>(comb) & 15) ||
      (reinterpret_cast<uintptr_t>(h_out) & 15) || (reinterpret_cast<uintptr_t>(c_cur) & 15) ||
      (reinterpret_cast<uintptr_t>(c_next) & 15) || (reinterpret_cast<uintptr_t>(act) & 15))
    return set_error(VFD_ERR_ARG, "convlstm_step_fwd: tensors must be 16-byte aligned, channel strides multiples of 8");
  const int w_rows = 4 * hid;
  FwdParams p;
  if (int e = fill_geom(p.g, N, 1, H, W, 1, kh, kw)) return e;
  p.cblocks = cin_k / kc;
  p.block_n = 256;                    // four gates x 64 hidden channels per accumulator tile
  p.n_tiles = w_rows / 256;
  Epilogue epi;
  epi.n_rows = w_rows; epi.out_cols = 0; epi.out_ld = 0; epi.out_fp32 = 1; epi.bias = bias_perm; epi.out = nullptr;
  epi.stats = nullptr; epi.stats_ld = 0;
  epi.lstm_j = 64; epi.lstm_hid = hid; epi.lstm_c_cur = c_cur; epi.lstm_c_next = c_next; epi.lstm_act = act;
  epi.lstm_h = h_out; epi.lstm_h_ld = h_ld;
  p.epi = epi;
  CUtensorMap tmA, tmB;
  if (int e = make_act_map(&tmA, comb, comb_ld, cin, N, 1, H, W, kc, p.g.TW, p.g.TH, p.g.TD, p.g.TN)) return e;
  if (int e = make_weight_map(&tmB, w_packed_perm, w_rows, (long long)p.g.ntaps * cin_k, kc, p.block_n)) return e;
  if (kc == 64) return launch_fwd<64>(tmA, tmB, p, stream);
  if (kc == 32) return launch_fwd<32>(tmA, tmB, p, stream);
  return launch_fwd<16>(tmA, tmB, p, stream);
}
