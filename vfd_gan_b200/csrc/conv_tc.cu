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
#include "ptx.cuh"
#include "vfd_internal.h"

namespace vfd {

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

struct FwdParams {
  ConvGeom g;
  int cblocks;    // channel blocks of KC per tap
  int n_tiles;    // tiles along the GEMM N (output channel) dimension
  int block_n;    // multiple of 16, <= 256
  int stages;
  int n_rows;     // rows in the packed weight matrix (valid bias entries)
  int out_cols;   // columns to store (multiple of 8)
  long long out_ld;  // elements between consecutive voxels in the output buffer
  int out_fp32;
  const float* bias;
  void* out;
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

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const ConvGeom& g = p.g;

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
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_m128(p.block_n, false, false);
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
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            umma_bf16(tacc, sdesc_kmajor(sa + k * 32, kRowBytes), sdesc_kmajor(sb + k * 32, kRowBytes),
                      idesc, (ks | k) != 0);
          }
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&acc_full[as]);
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
      for (int c = 0; c < p.block_n; c += 16) {
        float v[16];
        tmem_ld16(tacc + c, v);  // warp-collective: every lane participates
        const int col = col0 + c;
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (col + i < p.n_rows) v[i] += __ldg(p.bias + col + i);
        }
        if (valid) {
          if (p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + vox * p.out_ld + col;
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (col + i < p.out_cols)
                *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + vox * p.out_ld + col;
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              if (col + i < p.out_cols) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[i], v[i + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(o + i) = pk;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == 2) {
        as = 0;
        aph ^= 1;
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
    if (lane == 0 && has_work) {
      const uint32_t idesc = idesc_bf16_m128(p.block_n, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t sb = sa + 2 * kWgBoxBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 128 voxels per chunk = 8 x K16; 16 rows = 2048 B
          umma_bf16(tmem_base, sdesc_mnmajor128(sa + k * 2048, kWgBoxBytes),
                    sdesc_mnmajor128(sb + k * 2048, kWgBoxBytes), idesc, (c > c_begin) || (k != 0));
        }
        umma_commit(&empty_bar[s]);
        if (++s == p.stages) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit(&acc_full);
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
            atomicAdd(p.acc + (static_cast<size_t>(tap) * p.ci_pad + ci) * p.co_pad + co, v[i]);
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
  if (r != CUDA_SUCCESS) return set_error(VFD_ERR_DRIVER, "cuTensorMapEncodeTiled(activation) failed");
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

constexpr int kSmemBudget = 200 * 1024;

template <int KC>
static int launch_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, FwdParams& p,
                      cudaStream_t stream) {
  const int stage_bytes = ((kTileM + p.block_n) * KC * 2 + 1023) & ~1023;
  int stages = kSmemBudget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(VFD_ERR_ARG, "conv tile does not fit in shared memory");
  p.stages = stages;
  // >113 KB of dynamic smem keeps the kernel at one CTA per SM (each CTA owns all 512 TMEM columns)
  size_t smem = (size_t)stages * stage_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_tc_kernel<KC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv_fwd_tc)");
    attr_set = true;
  }
  const ConvGeom& g = p.g;
  const long long tiles = (long long)g.tilesW * g.tilesH * g.tilesD * g.tilesN * p.n_tiles;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv_fwd_tc_kernel<KC><<<grid, kFwdThreads, smem, stream>>>(tmA, tmB, p);
  return check_launch("conv_fwd_tc");
}

}  // namespace vfd

using namespace vfd;

VFD_API int vfd_conv3d_fwd(const void* x, long long x_ld, int cin, const void* w_packed,
                              int w_rows, int cin_k, const float* bias, void* out,
                              long long out_ld, int out_cols, int out_fp32, int N, int D, int H,
                              int W, int kd, int kh, int kw, int kc, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (kc != 16 && kc != 32 && kc != 64) return set_error(VFD_ERR_ARG, "kc must be 16, 32 or 64");
  if (cin_k % kc || w_rows % 16 || out_cols % 8 || out_cols > w_rows + 8)
    return set_error(VFD_ERR_ARG, "conv3d_fwd: bad channel padding");
  if ((out_fp32 ? (out_ld % 4) : (out_ld % 8)) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(VFD_ERR_ARG, "conv3d_fwd: output must be 16-byte aligned");
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;  // empty batch: nothing to do
  FwdParams p;
  if (int e = fill_geom(p.g, N, D, H, W, kd, kh, kw)) return e;
  p.cblocks = cin_k / kc;
  p.n_tiles = (w_rows + 255) / 256;
  p.block_n = (((w_rows + p.n_tiles - 1) / p.n_tiles) + 15) & ~15;
  p.n_rows = w_rows;
  p.out_cols = out_cols;
  p.out_ld = out_ld;
  p.out_fp32 = out_fp32;
  p.bias = bias;
  p.out = out;
  CUtensorMap tmA, tmB;
  if (int e = make_act_map(&tmA, x, x_ld, cin, N, D, H, W, kc, p.g.TW, p.g.TH, p.g.TD, p.g.TN))
    return e;
  if (int e = make_weight_map(&tmB, w_packed, w_rows, (long long)p.g.ntaps * cin_k, kc, p.block_n))
    return e;
  if (kc == 64) return launch_fwd<64>(tmA, tmB, p, stream);
  if (kc == 32) return launch_fwd<32>(tmA, tmB, p, stream);
  return launch_fwd<16>(tmA, tmB, p, stream);
}

VFD_API int vfd_conv3d_wgrad(const void* dy, long long dy_ld, int cout, const void* x,
                                long long x_ld, int cin, float* acc, int co_pad, int ci_pad, int N,
                                int D, int H, int W, int kd, int kh, int kw, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (N <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  if (co_pad < cout || ci_pad < cin) return set_error(VFD_ERR_ARG, "conv3d_wgrad: bad acc padding");
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
  const int nb_slots = (p.block_n + 63) / 64;
  const int stage_bytes = (2 + nb_slots) * kWgBoxBytes;
  int stages = kSmemBudget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(VFD_ERR_ARG, "wgrad tile does not fit in shared memory");
  p.stages = stages;
  const long long chunks = (long long)p.g.tilesW * p.g.tilesH * p.g.tilesD * p.g.tilesN;
  const long long base = (long long)p.g.ntaps * p.co_tiles * p.ci_tiles;
  long long splits = (2LL * num_sms() + base - 1) / base;  // aim for about two waves of CTAs
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
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
