// tcgen05 / TMEM / TMA weight gradient of the codec's conv and transposed-conv layers (the codec update of
// `train.py --adv`, reference train.py:357-366: out_criterion["loss"].backward() fills every Conv2d /
// ConvTranspose2d .weight.grad through ATen's cuDNN wgrad).
//
// Both layer kinds reduce to one form over a LOW-resolution and a HIGH-resolution channels-last tensor
// (conv: LOW = output gradient, HIGH = layer input; transposed conv: LOW = layer input, HIGH = output gradient):
//     dW[(kh,kw)][n][k] = sum over images, LOW pixels (i,j):  N-side[px][n] * K-side[px'][k],
//     px' = HIGH pixel (s*i + kh - p, s*j + kw - p)
// i.e. a GEMM whose REDUCTION axis is the pixel axis.  In channels-last tensors the channel axis is the contiguous
// one, so both operands are MN-major: a TMA box [32 ch][pixels] lands as 128-byte rows (one pixel each,
// 128-byte swizzle with 32-byte atoms) which is exactly the canonical MN-major operand of tcgen05.mma kind::tf32 --
// 32 channels contiguous, 8 pixel rows per K = 8 instruction, 32-channel chunks LBO bytes apart.  No transposition
// anywhere.
//
// One CTA owns up to four taps that read the same stride-parity plane and kernel row of HIGH (its TMEM holds
// one [128 x Nk] fp32 accumulator per tap, 512 columns in all) and walks a contiguous range of 8x8-pixel LOW
// tiles.  Per tile it loads the LOW tile [64 px][C] and ONE halo patch [8 x (8 + span) px][C] of the plane; the
// operand of tap dx is that patch read through a descriptor whose start is shifted by dx rows (the tensor core
// swizzles on absolute address bits, profiles/r1_umma_descriptor_shift_probe.txt).  Out-of-bounds pixels are
// zero-filled by TMA = the layer's zero padding and the ragged tile edge at once.
// Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> per-split partial in HBM).
// A second kernel adds the pixel splits in fixed order (deterministic) into the packed layout
// [taps][n_ch][k_ch] that icadv_unpack_weight turns into the torch layout.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

namespace {

constexpr int kWgThreads = 192;
constexpr int kWgTile = 8;           // LOW tile = 8 x 8 pixels = 64 reduction rows = 8 K-steps
constexpr int kWgMaxTaps = 4;        // accumulators per CTA
constexpr int kWgMaxGroups = 16;
constexpr int kWgBarBytes = 1024;
constexpr int kWgSmemMax = 227 * 1024;

struct WgGroup {
  int plane;                 // index into WgParams::high
  int ntaps;
  int dy, dx0;               // patch origin relative to the tile origin (plane coordinates)
  int dxr[kWgMaxTaps];       // tap column inside the patch
  int wtap[kWgMaxTaps];      // kh * ksize + kw
};

struct WgParams {
  CUtensorMap low;           // box [32, 8, 8, 1]
  CUtensorMap high[4];       // parity planes of HIGH, box [32, pw, 8, 1]
  float* partial;            // [splits][taps_total][n_ch][k_ch]
  int n_tiles, tiles_x, tiles_y, tiles_per_split;
  int n_ch, k_ch, taps_total;
  int low_is_n;              // 1: LOW holds the n (M-side) channels (conv), 0: HIGH does (transposed conv)
  int m_blocks, k_blocks;    // 128-row blocks of n_ch, <= 256-column blocks of k_ch
  int pw;                    // patch width in pixels
  int stages;
  int n_groups;
  int variant;               // developer probe (ICADV_WG_VARIANT): descriptor variants, 0 = the shipped form
  WgGroup groups[kWgMaxGroups];
};

// MN-major operand of kind::tf32.  Measured on the B200 (scripts/wgrad_probe.py): with the MN-major bits set the tensor
// core accepts 32-bit operands only in the "128-byte swizzle with 32-byte atoms" layout (descriptor layout type 1;
// with the plain SWIZZLE_128B type the instruction yields zeros).  Rows are 128 B (one pixel, 32 channels); the XOR is
// on 32-byte granules with (row mod 4), taken from absolute address bits, repeating every 512 B -- TMA writes exactly
// this with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  A K = 8 instruction spans two 4-row groups SBO = 512 B apart;
// 32-channel chunks sit lbo_bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, int variant = 0) {
  uint64_t d = 0;
  const uint32_t sbo = (variant & 1) ? 1024 : 512;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>((variant & 2) ? 2 : 1) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 4;
  uint64_t* acc_full = full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 9);
  uint8_t* stage0 = smem + kWgBarBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = ((split * n_groups + group) * m_blocks + mb) * k_blocks + kb
  int bid = blockIdx.x;
  const int kb = bid % p.k_blocks; bid /= p.k_blocks;
  const int mb = bid % p.m_blocks; bid /= p.m_blocks;
  const int gi = bid % p.n_groups;
  const int split = bid / p.n_groups;
  const WgGroup& grp = p.groups[gi];

  const int k0 = kb * 256;
  const int nk = min(256, p.k_ch - k0);          // accumulator columns per tap (multiple of 32)
  const int n0 = mb * 128;
  // channel chunks of the two smem operands
  const int low_c0 = p.low_is_n ? n0 : k0, high_c0 = p.low_is_n ? k0 : n0;
  const int low_chunks = p.low_is_n ? 4 : nk / 32, high_chunks = p.low_is_n ? nk / 32 : 4;
  const uint32_t tile_chunk_bytes = kWgTile * kWgTile * 128;
  const uint32_t patch_chunk_bytes = kWgTile * p.pw * 128;
  const uint32_t tile_bytes = low_chunks * tile_chunk_bytes;
  const uint32_t stage_bytes = tile_bytes + high_chunks * patch_chunk_bytes;

  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.n_tiles, t_begin + p.tiles_per_split);
  const int my_tiles = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&p.low);
    tma_prefetch_desc(&p.high[grp.plane]);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one_sync()) {
      const int per_img = p.tiles_x * p.tiles_y;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % p.stages, use = it / p.stages;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        const int t = t_begin + it;
        const int img = t / per_img, r = t - img * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        uint8_t* st = stage0 + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], stage_bytes);
        for (int c = 0; c < low_chunks; ++c)
          tma_load_4d(st + c * tile_chunk_bytes, &p.low, &full[s], low_c0 + c * 32, tx * kWgTile, ty * kWgTile, img);
        for (int c = 0; c < high_chunks; ++c)
          tma_load_4d(st + tile_bytes + c * patch_chunk_bytes, &p.high[grp.plane], &full[s], high_c0 + c * 32,
                      tx * kWgTile + grp.dx0, ty * kWgTile + grp.dy, img);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one_sync()) {
      // both operands MN-major: bits 15 (A) and 16 (B)
      const uint32_t idesc = umma_idesc_tf32(128, nk) | ((p.variant & 4) ? 0u : ((1u << 15) | (1u << 16)));
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % p.stages, use = it / p.stages;
        mbar_wait(&full[s], use & 1);
        tc_fence_after_sync();
        const uint32_t tile_addr = smem_u32(stage0 + (size_t)s * stage_bytes);
        const uint32_t patch_addr = tile_addr + tile_bytes;
        const uint64_t tile_desc = umma_desc_mn_sw128(tile_addr, tile_chunk_bytes, p.variant);
        const uint64_t patch_desc = umma_desc_mn_sw128(patch_addr, patch_chunk_bytes, p.variant);
        for (int t = 0; t < grp.ntaps; ++t) {
          const uint32_t dcol = tmem_base + t * nk;
#pragma unroll
          for (int ks = 0; ks < kWgTile; ++ks) {
            const uint64_t td = tile_desc + (uint64_t)((ks * kWgTile * 128) >> 4);
            const uint64_t pd = patch_desc + (uint64_t)(((ks * p.pw + grp.dxr[t]) * 128) >> 4);
            tc_mma_tf32(dcol, p.low_is_n ? td : pd, p.low_is_n ? pd : td, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        tc_commit(&empty[s]);
      }
      tc_commit(acc_full);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: TMEM -> partial
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int n = n0 + q * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
    }
    for (int t = 0; t < grp.ntaps; ++t) {
      float* dst = p.partial + (((size_t)split * p.taps_total + grp.wtap[t]) * p.n_ch + n) * p.k_ch + k0;
      for (int c = 0; c < nk / 32; ++c) {
        float v[32];
        if (my_tiles > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + t * nk + c * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0.f;
        }
        if (n < p.n_ch) {
          float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e) d4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

// ---------------------------------------------------------------------------------------------------------------
// RGB end layers (g_a.0: 3 -> N conv, g_s.6: N -> 3 transposed conv; 5x5 stride 2).  The 3-channel tensor is HIGH; it
// is first copied into the padded RGB0 layout of the forward first-layer form ([n][H+4][W+8][4], pixel (h,w) at
// (h+2,w+2)), where the window of LOW pixel (i,j) and kernel row kh -- HIGH pixels (2i+kh-2, 2j-2 .. 2j+5), 4 floats
// each -- is 32 contiguous floats at row 2i+kh, float offset 8j: a 5-D tensor map with OVERLAPPING 128-byte windows
// (32-byte step along j) delivers [64 LOW px][32] boxes, i.e. an MN-major operand whose 32 "channels" are (kw, c).
// One CTA holds the five kernel rows (5 x 32 TMEM columns): D_kh[m][kw*4+c] = sum_px LOW[px][m] * window_kh[px][kw*4+c].
struct WgRgbParams {
  CUtensorMap low;           // box [32, 8, 8, 1]
  CUtensorMap win;           // 5-D: [32 floats | j | row pair | row parity | image], box [32, 8, 8, 1, 1]
  float* partial;            // [splits][5][c_low][32]
  int n_tiles, tiles_x, tiles_y, tiles_per_split;
  int c_low, m_blocks, stages, variant;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_rgb_kernel(const __grid_constant__ WgRgbParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 4;
  uint64_t* acc_full = full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 9);
  uint8_t* stage0 = smem + kWgBarBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x % p.m_blocks, split = blockIdx.x / p.m_blocks;
  constexpr uint32_t kBox = kWgTile * kWgTile * 128;          // 8 KB: one 64-row box
  constexpr uint32_t kStage = 4 * kBox + 5 * kBox;             // LOW tile (128 channels) + five windows
  const int t_begin = split * p.tiles_per_split;
  const int my_tiles = min(p.n_tiles, t_begin + p.tiles_per_split) - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&p.low);
    tma_prefetch_desc(&p.win);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      const int per_img = p.tiles_x * p.tiles_y;
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % p.stages, use = it / p.stages;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        const int t = t_begin + it;
        const int img = t / per_img, r = t - img * per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        uint8_t* st = stage0 + (size_t)s * kStage;
        mbar_arrive_expect_tx(&full[s], kStage);
        for (int c = 0; c < 4; ++c)
          tma_load_4d(st + c * kBox, &p.low, &full[s], mb * 128 + c * 32, tx * kWgTile, ty * kWgTile, img);
        for (int kh = 0; kh < 5; ++kh)
          tma_load_5d(st + (4 + kh) * kBox, &p.win, &full[s], 0, tx * kWgTile, ty * kWgTile + (kh >> 1), kh & 1, img);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, 32) | (1u << 15) | (1u << 16);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % p.stages, use = it / p.stages;
        mbar_wait(&full[s], use & 1);
        tc_fence_after_sync();
        const uint32_t tile_addr = smem_u32(stage0 + (size_t)s * kStage);
        const uint64_t tile_desc = umma_desc_mn_sw128(tile_addr, kBox, p.variant);
#pragma unroll
        for (int kh = 0; kh < 5; ++kh) {
          const uint64_t win_desc = umma_desc_mn_sw128(tile_addr + (4 + kh) * kBox, kBox, p.variant);
#pragma unroll
          for (int ks = 0; ks < kWgTile; ++ks)
            tc_mma_tf32(tmem_base + kh * 32, tile_desc + (uint64_t)((ks * 1024) >> 4), win_desc + (uint64_t)((ks * 1024) >> 4),
                        idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int m = mb * 128 + q * 32 + lane;
    if (my_tiles > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after_sync();
    }
    for (int kh = 0; kh < 5; ++kh) {
      float v[32];
      if (my_tiles > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kh * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0.f;
      }
      if (m < p.c_low) {
        float4* d4 = reinterpret_cast<float4*>(p.partial + (((size_t)split * 5 + kh) * p.c_low + m) * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e) d4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
      }
    }
    tc_fence_before_sync();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 256); }
}

// partial [splits][5 kh][c_low][kw*4+c] -> dwpack [kh*5+kw][n_ch][k_ch]; sconv: (n, k) = (m, c), else (c, m)
__global__ void __launch_bounds__(256) wgrad_tc_rgb_reduce_kernel(const float* __restrict__ partial,
                                                                  float* __restrict__ dw, int c_low, int splits,
                                                                  int sconv) {
  const int total = 25 * c_low * 3;
  const int per_split = 5 * c_low * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % 3, m = (i / 3) % c_low, tap = i / (3 * c_low);
    const int kh = tap / 5, kw = tap % 5;
    const float* src = partial + ((size_t)kh * c_low + m) * 32 + kw * 4 + c;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += src[(size_t)s * per_split];
    dw[sconv ? ((size_t)tap * c_low + m) * 3 + c : ((size_t)tap * 3 + c) * c_low + m] = a;
  }
}

// dwpack[i] = sum over splits (fixed order) of partial[split][i]
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float4* __restrict__ partial,
                                                              float4* __restrict__ dw, int64_t total4, int splits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = partial[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = partial[(int64_t)s * total4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn wg_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  });
  return fn;
}

// channels-last [N][H][W][C] sub-sampled with step s from pixel (a, b); box [32, box_w, 8, 1]
int wg_encode(CUtensorMap* m, const float* base, int C, int W, int H, int N, int s, int a, int b, int box_w,
              int variant) {
  EncodeTiledFn enc = wg_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
  const int Wp = (W - b + s - 1) / s, Hp = (H - a + s - 1) / s;
  if (Wp <= 0 || Hp <= 0) { set_error("wgrad: empty parity plane"); return ICADV_EINVAL; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)s * C * 4, (cuuint64_t)s * W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)kWgTile, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base + ((int64_t)a * W + b) * C), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   (variant & 8) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("wgrad: cuTensorMapEncodeTiled(C=%d W=%d H=%d N=%d s=%d) failed: %d", C, W, H, N, s, (int)r);
    return ICADV_ECUDA;
  }
  return ICADV_OK;
}

inline int wg_mod2(int v) { return ((v % 2) + 2) % 2; }

}  // namespace

static bool wg_rgb_form(const icadv_conv_desc* d) {
  if (d->ksize != 5 || d->stride != 2 || d->active != nullptr || d->in_pad4) return false;
  if (d->form == ICADV_FORM_SCONV) return d->k_ch == 3 && d->n_ch % 32 == 0 && d->in_h % 2 == 0 && d->in_w % 2 == 0;
  return d->n_ch == 3 && d->k_ch % 32 == 0;
}

int wgrad_tc_supported(const icadv_conv_desc* d) {
  if (wg_rgb_form(d)) return 1;
  return d->k_ch % 32 == 0 && d->n_ch % 32 == 0 && d->active == nullptr && !d->in_pad4 &&
         (d->ksize == 1 || d->ksize == 3 || d->ksize == 5) && (d->stride == 1 || d->stride == 2);
}

// stream-ordered scratch comes from the device's default memory pool; keep what it has handed out cached instead of
// returning it to the driver at every synchronisation (a fresh reservation costs milliseconds)
static void wg_keep_pool() {
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  });
}

static int wgrad_tc_rgb(const icadv_conv_desc* d, const Geometry& g, const float* gout, float* dwpack,
                        cudaStream_t stream) {
  const bool sconv = d->form == ICADV_FORM_SCONV;
  const float* low = sconv ? gout : d->in;
  const float* rgb = sconv ? d->in : gout;
  const int c_low = sconv ? d->n_ch : d->k_ch;
  const int low_h = sconv ? g.out_h : d->in_h, low_w = sconv ? g.out_w : d->in_w;
  const int H = 2 * low_h, W = 2 * low_w;     // spatial size of the RGB tensor
  WgRgbParams p;
  memset(&p, 0, sizeof(p));
  if (const char* e = getenv("ICADV_WG_VARIANT")) p.variant = atoi(e);
  p.c_low = c_low;
  p.m_blocks = (c_low + 127) / 128;
  p.tiles_x = (low_w + kWgTile - 1) / kWgTile;
  p.tiles_y = (low_h + kWgTile - 1) / kWgTile;
  p.n_tiles = d->n_img * p.tiles_x * p.tiles_y;
  p.stages = 3;
  int splits = 148 / p.m_blocks;
  if (splits > p.n_tiles) splits = p.n_tiles;
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;

  // scratch: padded RGB0 copy (TF32-rounded) + split partials
  const int64_t Wp = W + 8, Hp = H + 4;
  const size_t pad_bytes = (size_t)d->n_img * Hp * Wp * 16;
  const size_t part_floats = (size_t)splits * 5 * c_low * 32;
  uint8_t* scratch = nullptr;
  ICADV_CUDA_TRY(cudaMallocAsync(&scratch, pad_bytes + part_floats * sizeof(float), stream));
  float* pad = reinterpret_cast<float*>(scratch);
  p.partial = reinterpret_cast<float*>(scratch + pad_bytes);
  ICADV_CUDA_TRY(cudaMemsetAsync(pad, 0, pad_bytes, stream));
  int rc = icadv_pad_rgb4(rgb, pad, d->n_img, H, W, 1, nullptr, nullptr, reinterpret_cast<icadv_stream_t>(stream));
  if (rc) return rc;

  rc = wg_encode(&p.low, low, c_low, low_w, low_h, d->n_img, 1, 0, 0, kWgTile, p.variant);
  if (rc) return rc;
  {
    EncodeTiledFn enc = wg_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
    cuuint64_t dims[5] = {32, (cuuint64_t)(W / 2), (cuuint64_t)(Hp / 2), 2, (cuuint64_t)d->n_img};
    cuuint64_t strides[4] = {32, (cuuint64_t)(2 * Wp * 16), (cuuint64_t)(Wp * 16), (cuuint64_t)(Hp * Wp * 16)};
    cuuint32_t box[5] = {32, (cuuint32_t)kWgTile, (cuuint32_t)kWgTile, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&p.win, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, pad, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (p.variant & 8) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("wgrad: cuTensorMapEncodeTiled(rgb windows) failed: %d", (int)r); return ICADV_ECUDA; }
  }
  const int smem_bytes = kWgBarBytes + p.stages * 9 * kWgTile * kWgTile * 128;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_rgb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemMax);
  });
  ICADV_CUDA_TRY(attr_err);
  wgrad_tc_rgb_kernel<<<splits * p.m_blocks, kWgThreads, smem_bytes, stream>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  const int total = 25 * c_low * 3;
  wgrad_tc_rgb_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(p.partial, dwpack, c_low, splits, sconv ? 1 : 0);
  ICADV_CUDA_TRY(cudaGetLastError());
  ICADV_CUDA_TRY(cudaFreeAsync(scratch, stream));
  return ICADV_OK;
}

// dwpack [taps][n_ch][k_ch] := weight gradient of the contraction `d` describes (d->in = its input, gout = gradient of
// its output).  Operands are read as TF32 (the caller rounds them to nearest where they are produced).
int wgrad_tc(const icadv_conv_desc* d, const Geometry& g, const float* gout, float* dwpack, cudaStream_t stream) {
  int rc = icadv_check_device();
  if (rc) return rc;
  ICADV_REQUIRE(wgrad_tc_supported(d), "wgrad_tc: shape not eligible");
  wg_keep_pool();
  if (wg_rgb_form(d)) return wgrad_tc_rgb(d, g, gout, dwpack, stream);
  const int k = d->ksize, s = d->stride, pad = k / 2;
  const bool sconv = d->form == ICADV_FORM_SCONV;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.n_ch = d->n_ch; p.k_ch = d->k_ch; p.taps_total = k * k;
  p.low_is_n = sconv ? 1 : 0;
  p.m_blocks = (d->n_ch + 127) / 128;
  p.k_blocks = (d->k_ch + 255) / 256;
  const int nk_max = d->k_ch < 256 ? d->k_ch : 256;
  int tmax = 512 / nk_max;
  if (tmax > kWgMaxTaps) tmax = kWgMaxTaps;

  // LOW / HIGH tensors
  const float* low = sconv ? gout : d->in;
  const float* high = sconv ? d->in : gout;
  const int low_c = sconv ? d->n_ch : d->k_ch, high_c = sconv ? d->k_ch : d->n_ch;
  const int low_h = sconv ? g.out_h : d->in_h, low_w = sconv ? g.out_w : d->in_w;
  const int high_h = sconv ? d->in_h : g.out_h, high_w = sconv ? d->in_w : g.out_w;

  // groups: taps of one kernel row that read the same column parity, at most tmax per group
  int span = 0;
  p.n_groups = 0;
  for (int kh = 0; kh < k; ++kh) {
    const int offy = kh - pad, a = s == 2 ? wg_mod2(offy) : 0, dy = s == 2 ? (offy - a) / 2 : offy;
    for (int b = 0; b < s; ++b) {
      int cnt = 0;
      WgGroup* cur = nullptr;
      for (int kw = 0; kw < k; ++kw) {
        const int offx = kw - pad, bb = s == 2 ? wg_mod2(offx) : 0, dx = s == 2 ? (offx - bb) / 2 : offx;
        if (bb != b) continue;
        if (cur == nullptr || cnt == tmax) {
          ICADV_REQUIRE(p.n_groups < kWgMaxGroups, "wgrad_tc: too many tap groups");
          cur = &p.groups[p.n_groups++];
          cur->plane = a * 2 + b; cur->ntaps = 0; cur->dy = dy; cur->dx0 = dx;
          cnt = 0;
        }
        cur->dxr[cnt] = dx - cur->dx0;
        cur->wtap[cnt] = kh * k + kw;
        if (cur->dxr[cnt] > span) span = cur->dxr[cnt];
        cur->ntaps = ++cnt;
      }
    }
  }
  p.pw = kWgTile + span;
  if (const char* e = getenv("ICADV_WG_VARIANT")) p.variant = atoi(e);

  rc = wg_encode(&p.low, low, low_c, low_w, low_h, d->n_img, 1, 0, 0, kWgTile, p.variant);
  if (rc) return rc;
  for (int a = 0; a < s; ++a)
    for (int b = 0; b < s; ++b) {
      if ((high_h - a + s - 1) / s <= 0 || (high_w - b + s - 1) / s <= 0) continue;
      rc = wg_encode(&p.high[a * 2 + b], high, high_c, high_w, high_h, d->n_img, s, a, b, p.pw, p.variant);
      if (rc) return rc;
    }

  p.tiles_x = (low_w + kWgTile - 1) / kWgTile;
  p.tiles_y = (low_h + kWgTile - 1) / kWgTile;
  p.n_tiles = d->n_img * p.tiles_x * p.tiles_y;

  const int low_chunks_max = p.low_is_n ? 4 : nk_max / 32, high_chunks_max = p.low_is_n ? nk_max / 32 : 4;
  const int stage_bytes = low_chunks_max * kWgTile * kWgTile * 128 + high_chunks_max * kWgTile * p.pw * 128;
  p.stages = (kWgSmemMax - kWgBarBytes) / stage_bytes;
  if (p.stages > 4) p.stages = 4;
  ICADV_REQUIRE(p.stages >= 1, "wgrad_tc: stage of %d bytes does not fit shared memory", stage_bytes);
  const int smem_bytes = kWgBarBytes + p.stages * stage_bytes;

  // pixel splits: about one CTA per SM in all
  const int ctas_per_split = p.n_groups * p.m_blocks * p.k_blocks;
  int splits = 148 / ctas_per_split;
  if (splits < 1) splits = 1;
  if (splits > p.n_tiles) splits = p.n_tiles;
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;

  const size_t per_split = (size_t)p.taps_total * d->n_ch * d->k_ch;
  float* partial = dwpack;
  if (splits > 1) ICADV_CUDA_TRY(cudaMallocAsync(&partial, per_split * splits * sizeof(float), stream));
  p.partial = partial;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemMax);
  });
  ICADV_CUDA_TRY(attr_err);
  wgrad_tc_kernel<<<splits * ctas_per_split, kWgThreads, smem_bytes, stream>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  if (splits > 1) {
    const int64_t total4 = (int64_t)per_split / 4;
    int blocks = (int)((total4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_tc_reduce_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(partial),
                                                       reinterpret_cast<float4*>(dwpack), total4, splits);
    ICADV_CUDA_TRY(cudaGetLastError());
    ICADV_CUDA_TRY(cudaFreeAsync(partial, stream));
  }
  return ICADV_OK;
}

}  // namespace icadv
