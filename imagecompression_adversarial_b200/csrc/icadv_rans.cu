// Entropy coding on the GPU: range-ANS encode / decode of the quantised latents with per-element CDF indexes
// (SURVEY.md section 8f rank 4).  Replaces compressai's CPU coder (compressai/cpp_exts/rans/rans_interface.cpp:
// RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes on ryg_rans rans64.h, and
// compressai/cpp_exts/ops/ops.cpp pmf_to_quantized_cdf) behind EntropyBottleneck / GaussianConditional
// .compress() / .decompress() (reference call sites: attack_TIC.py:106-110, InvCompress/attack_inv.py:112-116,
// InvCompress/ours.py:100-175).
//
// The arithmetic per symbol is compressai's, bit for bit: 64-bit state, 32-bit renormalisation, 16-bit
// probabilities, out-of-range symbols escape to the last CDF slot followed by 4-bit bypass digits.  What is
// B200-native is the stream layout: an image's symbol sequence is dealt round-robin to `lanes` independent coders
// (one thread each, thousands of images x lanes in flight), each writing its own word stream; a second kernel packs
// the lanes of an image into one string  [lane word counts (lanes x u32) | lane 0 words | lane 1 words | ...].
// With lanes == 1 the header is omitted and the string is exactly the one compressai emits.
// Two sequence orders:
//   mode 0  flat (c, h, w) order (compressai's order for EntropyBottleneck / GaussianConditional.compress);
//           lane l takes flat indexes l, l + lanes, ...
//   mode 1  position-major: for every position of a caller-given list (raster, or the wavefront schedule of the
//           autoregressive models), the channels; lane l takes channels l, l + lanes, ... of every position.
//           With lanes == 1 and a raster list this is compressai's _compress_ar order.
// Tensors are channels-last [n_img][hw][C] int32 / fp32 like everything else in the library.
#include <stdint.h>

#include "icadv_common.cuh"

namespace icadv {

namespace {

constexpr uint64_t kRansL = 1ull << 31;
constexpr int kPrec = 16;
constexpr int kBypass = 4;
constexpr uint32_t kMaxBypass = 15;

struct RansTables {
  const int* cdf;        // [n_cdf][stride]
  const int* sizes;      // [n_cdf] entries per row (pmf length + 2)
  const int* offsets;    // [n_cdf]
  int stride;
};

struct SeqGeom {
  int hw, C, mode, lanes, n_pos;
  const int* order;      // mode 1: position list (hw indexes) or NULL = raster
};

// number of symbols lane l codes per image
__device__ __forceinline__ int lane_count(const SeqGeom& g, int l) {
  if (g.mode == 0) {
    const int64_t n = (int64_t)g.hw * g.C;
    return (int)((n - l + g.lanes - 1) / g.lanes);
  }
  return g.n_pos * ((g.C - l + g.lanes - 1) / g.lanes);
}
// element offset ([hw][C] layout) of the k-th symbol of lane l
__device__ __forceinline__ int64_t lane_elem(const SeqGeom& g, int l, int k) {
  if (g.mode == 0) {
    const int64_t i = (int64_t)l + (int64_t)k * g.lanes;
    const int c = (int)(i / g.hw), px = (int)(i % g.hw);
    return (int64_t)px * g.C + c;
  }
  const int cpl = (g.C - l + g.lanes - 1) / g.lanes;
  const int p = k / cpl, c = l + (k - p * cpl) * g.lanes;
  const int px = g.order != nullptr ? g.order[p] : p;
  return (int64_t)px * g.C + c;
}

__device__ __forceinline__ void enc_put(uint64_t& x, uint32_t*& ptr, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((kRansL >> kPrec) << 32) * freq;
  if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
  x = ((x / freq) << kPrec) + (x % freq) + start;
}
__device__ __forceinline__ void enc_put_bits(uint64_t& x, uint32_t*& ptr, uint32_t val) {
  const uint32_t freq = 1u << (16 - kBypass);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
  x = (x << kBypass) | val;
}

// one thread = one (image, lane): codes its symbols last to first, writing words backwards from the end of its
// scratch region; lane_words[img][lane] = words written (the stream is the TAIL of the region)
__global__ void __launch_bounds__(128) rans_encode_kernel(const int* __restrict__ symbols, const int* __restrict__ indexes,
                                                          int n_img, SeqGeom g, RansTables t,
                                                          uint32_t* __restrict__ scratch, int lane_cap,
                                                          int* __restrict__ lane_words) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_img * g.lanes) return;
  const int img = tid / g.lanes, l = tid - img * g.lanes;
  const int64_t base = (int64_t)img * g.hw * g.C;
  uint32_t* const end = scratch + ((int64_t)tid + 1) * lane_cap;
  uint32_t* ptr = end;
  uint64_t x = kRansL;
  for (int k = lane_count(g, l) - 1; k >= 0; --k) {
    const int64_t e = base + lane_elem(g, l, k);
    const int ci = indexes[e];
    const int* cdf = t.cdf + (int64_t)ci * t.stride;
    const int max_value = t.sizes[ci] - 2;
    int value = symbols[e] - t.offsets[ci];
    uint32_t raw = 0;
    bool esc = false;
    if (value < 0) { raw = (uint32_t)(-2 * value - 1); value = max_value; esc = true; }
    else if (value >= max_value) { raw = (uint32_t)(2 * (value - max_value)); value = max_value; esc = true; }
    if (esc) {
      // pushed after the symbol: digit count (base-15 unary chunks), then the digits low to high; coded in reverse
      int n_bypass = 0;
      while (n_bypass < 8 && (raw >> (n_bypass * kBypass)) != 0) ++n_bypass;
      for (int j = n_bypass - 1; j >= 0; --j) enc_put_bits(x, ptr, (raw >> (j * kBypass)) & kMaxBypass);
      enc_put_bits(x, ptr, (uint32_t)(n_bypass % (int)kMaxBypass));
      for (int q = n_bypass / (int)kMaxBypass; q > 0; --q) enc_put_bits(x, ptr, kMaxBypass);
    }
    const uint32_t start = (uint32_t)cdf[value];
    enc_put(x, ptr, start, (uint32_t)cdf[value + 1] - start);
  }
  ptr -= 2;
  ptr[0] = (uint32_t)x;
  ptr[1] = (uint32_t)(x >> 32);
  lane_words[tid] = (int)(end - ptr);
}

// one block per image: header + concatenation of the lane streams; total_words[img] = words of the string
__global__ void __launch_bounds__(256) rans_pack_kernel(const uint32_t* __restrict__ scratch, int lane_cap,
                                                        const int* __restrict__ lane_words, int lanes,
                                                        uint32_t* __restrict__ packed, int packed_stride,
                                                        int* __restrict__ total_words) {
  extern __shared__ int pre[];          // exclusive prefix of the lane word counts
  const int img = blockIdx.x;
  const int* lw = lane_words + (int64_t)img * lanes;
  if (threadIdx.x == 0) {
    int acc = lanes > 1 ? lanes : 0;    // header words
    for (int l = 0; l < lanes; ++l) { pre[l] = acc; acc += lw[l]; }
    pre[lanes] = acc;
    total_words[img] = acc;
  }
  __syncthreads();
  uint32_t* out = packed + (int64_t)img * packed_stride;
  if (lanes > 1)
    for (int l = threadIdx.x; l < lanes; l += blockDim.x) out[l] = (uint32_t)lw[l];
  for (int l = 0; l < lanes; ++l) {
    const int n = lw[l];
    const uint32_t* src = scratch + ((int64_t)img * lanes + l + 1) * lane_cap - n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[pre[l] + i] = src[i];
  }
}

struct DecState { uint64_t x; const uint32_t* ptr; };

__device__ __forceinline__ uint32_t dec_bits(DecState& s) {
  const uint32_t val = (uint32_t)s.x & kMaxBypass;
  s.x >>= kBypass;
  if (s.x < kRansL) { s.x = (s.x << 32) | *s.ptr++; }
  return val;
}

__device__ __forceinline__ int dec_symbol(DecState& s, const RansTables& t, int ci) {
  const int* cdf = t.cdf + (int64_t)ci * t.stride;
  const int size = t.sizes[ci];
  const int max_value = size - 2;
  const uint32_t cum = (uint32_t)s.x & 0xFFFFu;
  // largest v with cdf[v] <= cum (cdf is non-decreasing, cdf[0] = 0, cdf[size-1] = 65536)
  int lo = 0, hi = size - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((uint32_t)cdf[mid] <= cum) lo = mid; else hi = mid;
  }
  const uint32_t start = (uint32_t)cdf[lo], freq = (uint32_t)cdf[lo + 1] - start;
  s.x = (uint64_t)freq * (s.x >> kPrec) + cum - start;
  if (s.x < kRansL) { s.x = (s.x << 32) | *s.ptr++; }
  int value = lo;
  if (value == max_value) {
    uint32_t val = dec_bits(s);
    int n_bypass = (int)val;
    while (val == kMaxBypass) { val = dec_bits(s); n_bypass += (int)val; }
    uint32_t raw = 0;
    for (int j = 0; j < n_bypass; ++j) raw |= dec_bits(s) << (j * kBypass);
    value = (int)(raw >> 1);
    if (raw & 1u) value = -value - 1; else value += max_value;
  }
  return value + t.offsets[ci];
}

__device__ __forceinline__ const uint32_t* lane_stream(const uint32_t* str, int lanes, int l) {
  if (lanes == 1) return str;
  int off = lanes;
  for (int i = 0; i < l; ++i) off += (int)str[i];
  return str + off;
}

// one thread = one (image, lane): decodes its whole sequence; out = symbol (+ mean), symbols_out optional
__global__ void __launch_bounds__(128) rans_decode_kernel(const uint32_t* __restrict__ packed, int packed_stride,
                                                          const int* __restrict__ indexes,
                                                          const float* __restrict__ means, float* __restrict__ out,
                                                          int n_img, SeqGeom g, RansTables t) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_img * g.lanes) return;
  const int img = tid / g.lanes, l = tid - img * g.lanes;
  const int64_t base = (int64_t)img * g.hw * g.C;
  DecState s;
  s.ptr = lane_stream(packed + (int64_t)img * packed_stride, g.lanes, l);
  s.x = (uint64_t)s.ptr[0] | ((uint64_t)s.ptr[1] << 32);
  s.ptr += 2;
  const int n = lane_count(g, l);
  for (int k = 0; k < n; ++k) {
    const int64_t e = base + lane_elem(g, l, k);
    const int v = dec_symbol(s, t, indexes[e]);
    out[e] = (float)v + (means != nullptr ? means[e] : 0.f);
  }
}

// incremental decoding for the autoregressive models: the per-lane coder state lives in device memory between steps
__global__ void __launch_bounds__(128) rans_decode_init_kernel(const uint32_t* __restrict__ packed, int packed_stride,
                                                               int n_img, int lanes, uint64_t* __restrict__ state,
                                                               int* __restrict__ cursor) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_img * lanes) return;
  const int img = tid / lanes, l = tid - img * lanes;
  const uint32_t* str = packed + (int64_t)img * packed_stride;
  const uint32_t* p = lane_stream(str, lanes, l);
  state[tid] = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
  cursor[tid] = (int)(p - str) + 2;
}

// one step: every lane decodes its channels of the listed positions (in list order)
__global__ void __launch_bounds__(128) rans_decode_step_kernel(const uint32_t* __restrict__ packed, int packed_stride,
                                                               uint64_t* __restrict__ state, int* __restrict__ cursor,
                                                               const int* __restrict__ indexes,
                                                               const float* __restrict__ means, float* __restrict__ out,
                                                               int n_img, int hw, int C, int lanes,
                                                               const int* __restrict__ positions, int n_positions,
                                                               RansTables t) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_img * lanes) return;
  const int img = tid / lanes, l = tid - img * lanes;
  const uint32_t* str = packed + (int64_t)img * packed_stride;
  DecState s;
  s.x = state[tid];
  s.ptr = str + cursor[tid];
  for (int p = 0; p < n_positions; ++p) {
    const int64_t row = ((int64_t)img * hw + positions[p]) * C;
    for (int c = l; c < C; c += lanes) {
      const int v = dec_symbol(s, t, indexes[row + c]);
      out[row + c] = (float)v + (means != nullptr ? means[row + c] : 0.f);
    }
  }
  state[tid] = s.x;
  cursor[tid] = (int)(s.ptr - str);
}

// compressai GaussianConditional.build_indexes: index = number of table entries (all but the last) below the
// lower-bounded scale
__global__ void build_indexes_kernel(const float* __restrict__ scales, const float* __restrict__ table, int levels,
                                     float bound, int* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = fmaxf(scales[i], bound);
    int idx = levels - 1;
    for (int k = 0; k < levels - 1; ++k) idx -= (s <= __ldg(table + k)) ? 1 : 0;
    out[i] = idx;
  }
}

int check_geom(int n_img, int hw, int C, int mode, int n_pos, int lanes) {
  ICADV_REQUIRE(n_img > 0 && hw > 0 && C > 0, "rans: bad sizes");
  ICADV_REQUIRE(mode == 0 || mode == 1, "rans: mode must be 0 (flat c,h,w) or 1 (position-major)");
  ICADV_REQUIRE(lanes >= 1 && lanes <= 1024, "rans: lanes must be in [1, 1024]");
  ICADV_REQUIRE(mode == 0 || (n_pos > 0 && n_pos <= hw), "rans: bad position count");
  ICADV_REQUIRE(mode == 0 || lanes <= C, "rans: position-major order deals CHANNELS to lanes (lanes <= C)");
  return ICADV_OK;
}

}  // namespace

}  // namespace icadv

using namespace icadv;

extern "C" {

/* host: compressai/cpp_exts/ops/ops.cpp pmf_to_quantized_cdf; cdf has n + 1 entries */
int icadv_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int* cdf) {
  ICADV_REQUIRE(pmf && cdf && n > 0 && precision > 0 && precision <= 16, "pmf_to_quantized_cdf: bad arguments");
  cdf[0] = 0;
  uint64_t total = 0;
  for (int i = 0; i < n; ++i) {
    ICADV_REQUIRE(pmf[i] >= 0.f && pmf[i] == pmf[i] && pmf[i] < 1e30f, "pmf_to_quantized_cdf: invalid probability");
    cdf[i + 1] = (int)roundf(pmf[i] * (float)(1 << precision));
    total += (uint64_t)cdf[i + 1];
  }
  ICADV_REQUIRE(total > 0, "pmf_to_quantized_cdf: zero total");
  for (int i = 0; i <= n; ++i) cdf[i] = (int)((((uint64_t)1 << precision) * (uint64_t)cdf[i]) / total);
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
  cdf[n] = 1 << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < n; ++j) {
      const uint32_t freq = (uint32_t)(cdf[j + 1] - cdf[j]);
      if (freq > 1 && freq < best_freq) { best_freq = freq; best = j; }
    }
    ICADV_REQUIRE(best != -1, "pmf_to_quantized_cdf: no symbol can spare a count");
    if (best < i) for (int j = best + 1; j <= i; ++j) cdf[j]--;
    else for (int j = i + 1; j <= best; ++j) cdf[j]++;
  }
  return ICADV_OK;
}

int icadv_rans_lane_capacity(int n_img, int hw, int C, int mode, int n_pos, int lanes) {
  (void)n_img;
  const int64_t per = mode == 0 ? ((int64_t)hw * C + lanes - 1) / lanes : (int64_t)n_pos * ((C + lanes - 1) / lanes);
  return (int)(2 * per + 4);   // <= 52 bits per symbol (16 + 4 + 8 x 4) + the 64-bit flush
}

int icadv_rans_encode(const int* symbols, const int* indexes, int n_img, int hw, int C, const int* cdf,
                      const int* cdf_sizes, const int* offsets, int cdf_stride, int mode, const int* order, int n_pos,
                      int lanes, uint32_t* scratch, int* lane_words, uint32_t* packed, int packed_stride,
                      int* total_words, icadv_stream_t stream) {
  ICADV_REQUIRE(symbols && indexes && cdf && cdf_sizes && offsets && scratch && lane_words && packed && total_words,
                "rans_encode: null pointer");
  int rc = check_geom(n_img, hw, C, mode, n_pos, lanes);
  if (rc) return rc;
  const int cap = icadv_rans_lane_capacity(n_img, hw, C, mode, n_pos, lanes);
  ICADV_REQUIRE((int64_t)packed_stride >= (int64_t)lanes * cap + lanes, "rans_encode: packed_stride too small");
  SeqGeom g{hw, C, mode, lanes, n_pos, order};
  RansTables t{cdf, cdf_sizes, offsets, cdf_stride};
  const int threads = n_img * lanes;
  rans_encode_kernel<<<(threads + 127) / 128, 128, 0, as_stream(stream)>>>(symbols, indexes, n_img, g, t, scratch, cap,
                                                                           lane_words);
  ICADV_CUDA_TRY(cudaGetLastError());
  rans_pack_kernel<<<n_img, 256, (lanes + 1) * sizeof(int), as_stream(stream)>>>(scratch, cap, lane_words, lanes, packed,
                                                                                 packed_stride, total_words);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_rans_decode(const uint32_t* packed, int packed_stride, const int* indexes, const float* means, float* out,
                      int n_img, int hw, int C, const int* cdf, const int* cdf_sizes, const int* offsets,
                      int cdf_stride, int mode, const int* order, int n_pos, int lanes, icadv_stream_t stream) {
  ICADV_REQUIRE(packed && indexes && out && cdf && cdf_sizes && offsets, "rans_decode: null pointer");
  int rc = check_geom(n_img, hw, C, mode, n_pos, lanes);
  if (rc) return rc;
  SeqGeom g{hw, C, mode, lanes, n_pos, order};
  RansTables t{cdf, cdf_sizes, offsets, cdf_stride};
  const int threads = n_img * lanes;
  rans_decode_kernel<<<(threads + 127) / 128, 128, 0, as_stream(stream)>>>(packed, packed_stride, indexes, means, out,
                                                                           n_img, g, t);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_rans_decode_init(const uint32_t* packed, int packed_stride, int n_img, int lanes, uint64_t* state,
                           int* cursor, icadv_stream_t stream) {
  ICADV_REQUIRE(packed && state && cursor && n_img > 0 && lanes >= 1 && lanes <= 1024, "rans_decode_init: bad arguments");
  const int threads = n_img * lanes;
  rans_decode_init_kernel<<<(threads + 127) / 128, 128, 0, as_stream(stream)>>>(packed, packed_stride, n_img, lanes,
                                                                                state, cursor);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_rans_decode_step(const uint32_t* packed, int packed_stride, uint64_t* state, int* cursor, const int* indexes,
                           const float* means, float* out, int n_img, int hw, int C, const int* cdf,
                           const int* cdf_sizes, const int* offsets, int cdf_stride, const int* positions,
                           int n_positions, int lanes, icadv_stream_t stream) {
  ICADV_REQUIRE(packed && state && cursor && indexes && out && cdf && cdf_sizes && offsets && positions,
                "rans_decode_step: null pointer");
  ICADV_REQUIRE(n_img > 0 && hw > 0 && C > 0 && n_positions > 0 && lanes >= 1 && lanes <= C, "rans_decode_step: bad sizes");
  RansTables t{cdf, cdf_sizes, offsets, cdf_stride};
  const int threads = n_img * lanes;
  rans_decode_step_kernel<<<(threads + 127) / 128, 128, 0, as_stream(stream)>>>(
      packed, packed_stride, state, cursor, indexes, means, out, n_img, hw, C, lanes, positions, n_positions, t);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_build_indexes(const float* scales, const float* table, int levels, float bound, int* out, long long n,
                        icadv_stream_t stream) {
  ICADV_REQUIRE(scales && table && out && levels > 0 && n > 0, "build_indexes: bad arguments");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  build_indexes_kernel<<<blocks, 256, 0, as_stream(stream)>>>(scales, table, levels, bound, out, n);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
