// Error plumbing, contraction geometry, weight packing, layout conversion, GDN reparametrisation.
#include <mutex>
#include <string>

#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int mod2(int v) { return ((v % 2) + 2) % 2; }

int make_geometry(const icadv_conv_desc* d, Geometry* g) {
  ICADV_REQUIRE(d != nullptr, "null descriptor");
  ICADV_REQUIRE(d->ksize == 1 || d->ksize == 3 || d->ksize == 5, "ksize %d unsupported", d->ksize);
  ICADV_REQUIRE(d->stride == 1 || d->stride == 2, "stride %d unsupported", d->stride);
  ICADV_REQUIRE(d->in_h > 0 && d->in_w > 0 && d->n_img > 0 && d->k_ch > 0 && d->n_ch > 0, "bad sizes");
  const int k = d->ksize, s = d->stride, p = k / 2;
  g->form = d->form; g->ksize = k; g->stride = s; g->pad = p;
  g->in_h = d->in_h; g->in_w = d->in_w;
  for (int l = 0; l < 4; ++l) { g->n_taps[l] = 0; g->out_a[l] = g->out_b[l] = 0; }
  if (d->form == ICADV_FORM_SCONV) {
    g->out_h = (d->in_h + 2 * p - k) / s + 1;
    g->out_w = (d->in_w + 2 * p - k) / s + 1;
    g->tile_h = g->out_h; g->tile_w = g->out_w;
    g->n_launch = 1;
    int n = 0;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        Tap t;
        if (s == 1) {
          t.plane = 0; t.dy = (int16_t)(kh - p); t.dx = (int16_t)(kw - p);
        } else {
          int a = mod2(kh - p), b = mod2(kw - p);
          t.plane = (int16_t)(a * 2 + b);
          t.dy = (int16_t)((kh - p - a) / 2);
          t.dx = (int16_t)((kw - p - b) / 2);
        }
        t.wtap = (int16_t)(kh * k + kw);
        g->taps[0][n++] = t;
      }
    g->n_taps[0] = n;
  } else if (d->form == ICADV_FORM_TCONV) {
    g->out_h = d->in_h * s; g->out_w = d->in_w * s;
    g->tile_h = d->in_h; g->tile_w = d->in_w;
    g->n_launch = s * s;
    for (int a = 0; a < s; ++a)
      for (int b = 0; b < s; ++b) {
        int l = a * s + b, n = 0;
        g->out_a[l] = a; g->out_b[l] = b;
        for (int kh = 0; kh < k; ++kh) {
          if (mod2(a + p - kh) != 0 && s == 2) continue;
          for (int kw = 0; kw < k; ++kw) {
            if (mod2(b + p - kw) != 0 && s == 2) continue;
            Tap t;
            t.plane = 0;
            t.dy = (int16_t)((a + p - kh) / s);
            t.dx = (int16_t)((b + p - kw) / s);
            t.wtap = (int16_t)(kh * k + kw);
            g->taps[l][n++] = t;
          }
        }
        g->n_taps[l] = n;
      }
  } else {
    ICADV_REQUIRE(false, "unknown form %d", d->form);
  }
  return ICADV_OK;
}

// ------------------------------------------------------------------ small kernels
// kind 0: Conv2d fwd      w[co][ci][t] -> P[t][n=co][k=ci]
// kind 1: Conv2d dgrad    w[co][ci][t] -> P[t][n=ci][k=co]
// kind 2: ConvT  fwd      w[ci][co][t] -> P[t][n=co][k=ci]
// kind 3: ConvT  dgrad    w[ci][co][t] -> P[t][n=ci][k=co]
// (c_out, c_in) are the LAYER's out/in channels; torch dims: Conv2d [c_out,c_in], ConvT [c_in,c_out].
__device__ __forceinline__ int64_t torch_w_index(int kind, int n, int k, int t, int c_out, int c_in, int taps) {
  int co, ci;
  if (kind == 0 || kind == 2) { co = n; ci = k; } else { ci = n; co = k; }
  if (kind <= 1) return ((int64_t)co * c_in + ci) * taps + t;
  return ((int64_t)ci * c_out + co) * taps + t;
}

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ P, int kind, int c_out, int c_in,
                                   int taps, int round) {
  const int N = (kind == 0 || kind == 2) ? c_out : c_in;
  const int K = (kind == 0 || kind == 2) ? c_in : c_out;
  int64_t total = (int64_t)taps * N * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    int n = (int)((i / K) % N);
    int t = (int)(i / ((int64_t)K * N));
    const float v = w[torch_w_index(kind, n, k, t, c_out, c_in, taps)];
    P[i] = round ? round_tf32(v) : v;
  }
}

__global__ void unpack_weight_kernel(const float* __restrict__ P, float* __restrict__ w, int kind, int c_out, int c_in,
                                     int taps, int accumulate) {
  const int N = (kind == 0 || kind == 2) ? c_out : c_in;
  const int K = (kind == 0 || kind == 2) ? c_in : c_out;
  int64_t total = (int64_t)taps * N * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = (int)(i % K);
    int n = (int)((i / K) % N);
    int t = (int)(i / ((int64_t)K * N));
    int64_t j = torch_w_index(kind, n, k, t, c_out, c_in, taps);
    w[j] = accumulate ? w[j] + P[i] : P[i];
  }
}

__global__ void pack_weight_rgb_kernel(const float* __restrict__ w, float* __restrict__ P, int n_ch, int round) {
  const int total = 5 * n_ch * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % 32, n = (i / 32) % n_ch, kh = i / (32 * n_ch);
    const int kw = k >> 2, c = k & 3;
    float v = 0.f;
    if (kw < 5 && c < 3) v = w[((n * 3 + c) * 5 + kh) * 5 + kw];
    P[i] = round ? round_tf32(v) : v;
  }
}

__global__ void pad_rgb4_kernel(const float* __restrict__ src, float4* __restrict__ dst, int h, int w, int round,
                                const int* __restrict__ active, const int* __restrict__ n_active) {
  const int slot = blockIdx.y;
  if (n_active != nullptr && slot >= *n_active) return;
  const int img = active != nullptr ? active[slot] : slot;
  const int wp = w + 8, hp = h + 4;
  const float* s = src + (int64_t)img * h * w * 3;
  float4* d = dst + (int64_t)img * hp * wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < h * w; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    float a = s[3 * i], b = s[3 * i + 1], c = s[3 * i + 2];
    if (round) { a = round_tf32(a); b = round_tf32(b); c = round_tf32(c); }
    d[(int64_t)(y + 2) * wp + x + 2] = make_float4(a, b, c, 0.f);
  }
}

// NCHW <-> NHWC through a 32x32 shared tile: both sides coalesced.
__global__ void transpose_cp_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // src: [batch][rows][cols] -> dst: [batch][cols][rows]
  __shared__ float tile[32][33];
  const int64_t base = (int64_t)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int rr = blockIdx.y * 32 + r;
    if (rr < rows && c < cols) tile[r][threadIdx.x] = src[base + (int64_t)rr * cols + c];
  }
  __syncthreads();
  int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int cc = threadIdx.y; cc < 32; cc += 8) {
    int c2 = blockIdx.x * 32 + cc;
    if (c2 < cols && r2 < rows) dst[base + (int64_t)c2 * rows + r2] = tile[threadIdx.x][cc];
  }
}

// Few channels (RGB images: C <= 4): the 32x32 tile above leaves 29 of 32 lanes idle on the channel axis (measured
// 0.44 TB/s on the 151 MB images of the ms-ssim attack loop, scripts/msssim_profile.py).  Here a block moves 256 pixels:
// the interleaved side is one contiguous run of 256*C floats (coalesced), the planar side C runs of 256 floats; the
// shared-memory side reads at stride C, conflict-free for C = 1, 3 (odd) and 2-way for C = 2, 4.
// FUSE: the [0, 1] clamp of the reconstruction (attack_rd.py:353-356, Up_bound(Low_bound(x, 0), 1)) rides on the copy --
// forward on the way to planes, and its backward (utils/ops.py:28-56: a gradient passes a bound it sits on only if it
// points back inside) on the way back, gated by the unclamped interleaved x.  The ms-ssim attack loop then has one launch
// on each side of the value-and-gradient pyramid instead of three and a copy.
__device__ __forceinline__ float clamp01_grad(float g, float x) {
  const float lo = fmaxf(x, 0.f);
  g = ((lo <= 1.f) || (g > 0.f)) ? g : 0.f;     // Up_bound(., 1) backward, evaluated at Low_bound(x, 0)
  g = ((x >= 0.f) || (g < 0.f)) ? g : 0.f;      // Low_bound(., 0) backward
  return g;
}

template <int C, bool TO_PLANAR, bool FUSE>
__global__ void __launch_bounds__(256) interleave_small_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                               const float* __restrict__ gate_x, int hw) {
  __shared__ float tile[256 * C];
  const int64_t img = (int64_t)blockIdx.y * hw * C;
  const int p0 = blockIdx.x * 256, np = min(256, hw - p0), t = threadIdx.x;
  if (TO_PLANAR) {
    const float* in = src + img + (int64_t)p0 * C;
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (t + k * 256 < np * C) {
        const float v = in[t + k * 256];
        tile[t + k * 256] = FUSE ? fminf(fmaxf(v, 0.f), 1.f) : v;
      }
    __syncthreads();
    if (t < np) {
#pragma unroll
      for (int c = 0; c < C; ++c) dst[img + (int64_t)c * hw + p0 + t] = tile[t * C + c];
    }
  } else {
    if (t < np) {
#pragma unroll
      for (int c = 0; c < C; ++c) tile[t * C + c] = src[img + (int64_t)c * hw + p0 + t];
    }
    __syncthreads();
    float* out = dst + img + (int64_t)p0 * C;
    const float* gx = FUSE ? gate_x + img + (int64_t)p0 * C : nullptr;
#pragma unroll
    for (int k = 0; k < C; ++k)
      if (t + k * 256 < np * C) {
        const float v = tile[t + k * 256];
        out[t + k * 256] = FUSE ? clamp01_grad(v, gx[t + k * 256]) : v;
      }
  }
}

template <bool TO_PLANAR, bool FUSE>
static int launch_interleave_small(const float* src, float* dst, const float* gate_x, int n, int c, int hw, cudaStream_t s) {
  dim3 grid((hw + 255) / 256, n);
  ICADV_REQUIRE(n <= 65535, "transpose grid too large");
  switch (c) {
    case 1: interleave_small_kernel<1, TO_PLANAR, FUSE><<<grid, 256, 0, s>>>(src, dst, gate_x, hw); break;
    case 2: interleave_small_kernel<2, TO_PLANAR, FUSE><<<grid, 256, 0, s>>>(src, dst, gate_x, hw); break;
    case 3: interleave_small_kernel<3, TO_PLANAR, FUSE><<<grid, 256, 0, s>>>(src, dst, gate_x, hw); break;
    default: interleave_small_kernel<4, TO_PLANAR, FUSE><<<grid, 256, 0, s>>>(src, dst, gate_x, hw); break;
  }
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

__global__ void gdn_reparam_kernel(const float* __restrict__ raw, float* __restrict__ eff, int rows, int cols,
                                   float bound, float pedestal, int transpose, int round) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int r = i / cols, c = i % cols;
  float v = fmaxf(raw[i], bound);
  v = v * v - pedestal;
  if (round) v = round_tf32(v);
  if (transpose) eff[c * rows + r] = v; else eff[i] = v;
}

// nn.PixelShuffle(r) on channels-last tensors (compressai subpel_conv3x3): out[n, h*r+i, w*r+j, c] = in[n, h, w, c*r*r + i*r + j]
// inverse != 0 runs the permutation backwards (its gradient).
// Both sides coalesced: a block takes kShufT consecutive LOW-resolution pixels of one row (one contiguous run of
// kShufT * c_in floats) through shared memory; each of the r HIGH-resolution rows they map to is again one contiguous run
// (r * kShufT pixels x c_out floats).  The previous form (one thread per LOW element, 4-byte accesses c_out floats apart on
// the HIGH side) measured 1.0 TB/s on the cheng2020 upsampling blocks (profiles/r2_bench_config4.json).
// A block takes T consecutive LOW pixels (T chosen by the host so that a tile is ~24 KB whatever the channel count) --
// element indexes are stepped without divisions (the first form spent ~40 integer instructions per element on o / c_out and
// xl / r: 2.4 TB/s on the 768-channel blocks, 0.3 TB/s on the 12-channel one).
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_seg,
                                                            int h, int w, int c_out, int r, int inverse, int T) {
  extern __shared__ float tile[];   // [T][c_in]
  const int c_in = c_out * r * r;
  const int segs_x = (w + T - 1) / T;
  const int dq = 256 / c_out, dr = 256 % c_out;          // element step of a thread: (xl, c) += (dq, dr) with carry
  const int xl0 = threadIdx.x / c_out, c0 = threadIdx.x % c_out;
  for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
    const int sx = seg % segs_x, y = (seg / segs_x) % h;
    const int64_t n = seg / ((int64_t)segs_x * h);
    const int x0 = sx * T, nx = min(T, w - x0);
    const int64_t lo = ((n * h + y) * (int64_t)w + x0) * c_in;              // first LOW element of the segment
    const int run = nx * r * c_out;                                          // floats of one HIGH row of the segment
    const int n_lo = nx * c_in;
    const bool v4 = (c_in & 3) == 0;                                         // then lo and n_lo are multiples of 4
    if (!inverse) {
      if (v4) {
        const float4* s4 = reinterpret_cast<const float4*>(src + lo);
        float4* t4 = reinterpret_cast<float4*>(tile);
        for (int i = threadIdx.x; i < n_lo / 4; i += 256) t4[i] = s4[i];
      } else {
        for (int i = threadIdx.x; i < n_lo; i += 256) tile[i] = src[lo + i];
      }
      __syncthreads();
    }
    for (int di = 0; di < r; ++di) {
      const int64_t hi = ((n * h * r + (int64_t)y * r + di) * ((int64_t)w * r) + (int64_t)x0 * r) * c_out;
      int xl = xl0, c = c0;
      for (int o = threadIdx.x; o < run; o += 256) {
        const int x = r == 2 ? (xl >> 1) : xl / r, dj = r == 2 ? (xl & 1) : xl - (xl / r) * r;
        const int t = x * c_in + c * r * r + di * r + dj;
        if (!inverse) dst[hi + o] = tile[t]; else tile[t] = src[hi + o];
        xl += dq; c += dr;
        if (c >= c_out) { c -= c_out; ++xl; }
      }
    }
    if (inverse) {
      __syncthreads();
      if (v4) {
        float4* d4 = reinterpret_cast<float4*>(dst + lo);
        const float4* t4 = reinterpret_cast<const float4*>(tile);
        for (int i = threadIdx.x; i < n_lo / 4; i += 256) d4[i] = t4[i];
      } else {
        for (int i = threadIdx.x; i < n_lo; i += 256) dst[lo + i] = tile[i];
      }
    }
    __syncthreads();
  }
}

// copy `count` channels of every pixel: dst[px, dst_off + c] = src[px, src_off + c]   (torch.cat / chunk on dim 1)
__global__ void copy_channels_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n_px, int c_src,
                                     int c_dst, int src_off, int dst_off, int count) {
  const int64_t total = n_px * count;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % count);
    const int64_t px = i / count;
    dst[px * c_dst + dst_off + c] = src[px * c_src + src_off + c];
  }
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_pixel_shuffle(const float* src, float* dst, int n, int h, int w, int c_out, int r, int inverse,
                        icadv_stream_t stream) {
  ICADV_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c_out > 0 && r >= 1, "bad pixel_shuffle args");
  const int c_in = c_out * r * r;
  ICADV_REQUIRE(c_in <= 12288, "pixel_shuffle: c_out * r * r too large");
  int T = 6144 / c_in;                       // ~24 KB of LOW pixels per block iteration
  if (T < 2) T = 2;
  if (T > w) T = w;
  const size_t smem = (size_t)T * c_in * sizeof(float);
  const int64_t n_seg = (int64_t)n * h * ((w + T - 1) / T);
  ICADV_REQUIRE(n_seg < (1ll << 31), "pixel_shuffle: tensor too large");
  if (smem > 48 * 1024) {
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
      err = cudaFuncSetAttribute(pixel_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(pixel_shuffle) failed: %s", cudaGetErrorString(err)); return ICADV_ECUDA; }
  }
  int64_t blocks = n_seg < 148 * 8 ? n_seg : 148 * 8;
  pixel_shuffle_kernel<<<(int)blocks, 256, smem, as_stream(stream)>>>(src, dst, (int)n_seg, h, w, c_out, r, inverse, T);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_copy_channels(const float* src, float* dst, int64_t n_px, int c_src, int c_dst, int src_off, int dst_off,
                        int count, icadv_stream_t stream) {
  ICADV_REQUIRE(src && dst && n_px > 0 && count > 0 && src_off >= 0 && dst_off >= 0 && src_off + count <= c_src &&
                    dst_off + count <= c_dst, "bad copy_channels args");
  int64_t blocks = (n_px * count + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  copy_channels_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, dst, n_px, c_src, c_dst, src_off, dst_off, count);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

const char* icadv_last_error(void) { return g_err; }
int icadv_version(void) { return 100; }

int icadv_check_device(void) {
  // cudaGetDeviceProperties costs milliseconds and this is called from every plan creation: cache the verdict per
  // device ordinal (0 = unknown, 1 = sm_100, 2 = other); benign race: every thread computes the same value
  static int verdict[64] = {0};
  int dev = 0;
  ICADV_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && verdict[dev] == 1) return ICADV_OK;
  int major = 0, minor = 0;
  ICADV_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  ICADV_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    set_error("device %d is sm_%d%d; this library is sm_100a only (no fallback)", dev, major, minor);
    return ICADV_EARCH;
  }
  if (dev >= 0 && dev < 64) verdict[dev] = 1;
  return ICADV_OK;
}

int icadv_conv_out_hw(const icadv_conv_desc* d, int* out_h, int* out_w) {
  Geometry g;
  int rc = make_geometry(d, &g);
  if (rc) return rc;
  if (out_h) *out_h = g.out_h;
  if (out_w) *out_w = g.out_w;
  return ICADV_OK;
}

int icadv_pack_weight(const float* w, float* wpack, int kind, int c_out, int c_in, int ksize, int round_tf32,
                      icadv_stream_t stream) {
  ICADV_REQUIRE(w && wpack && kind >= 0 && kind <= 3, "bad pack_weight args");
  int taps = ksize * ksize;
  int64_t total = (int64_t)taps * c_out * c_in;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, wpack, kind, c_out, c_in, taps, round_tf32);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_pack_weight_rgb(const float* w, float* wpack, int n_ch, int round_tf32, icadv_stream_t stream) {
  ICADV_REQUIRE(w && wpack && n_ch > 0, "bad pack_weight_rgb args");
  pack_weight_rgb_kernel<<<(5 * n_ch * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(w, wpack, n_ch, round_tf32);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_pad_rgb4(const float* src, float* dst, int n_img, int h, int w, int round_tf32, const int* active,
                   const int* n_active, icadv_stream_t stream) {
  ICADV_REQUIRE(src && dst && n_img > 0 && h > 0 && w > 0 && n_img <= 65535, "bad pad_rgb4 args");
  int bx = (h * w + 255) / 256;
  if (bx > 592) bx = 592;
  pad_rgb4_kernel<<<dim3(bx, n_img), 256, 0, as_stream(stream)>>>(src, reinterpret_cast<float4*>(dst), h, w, round_tf32,
                                                                  active, n_active);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_unpack_weight(const float* dwpack, float* dw, int kind, int c_out, int c_in, int ksize, int accumulate,
                        icadv_stream_t stream) {
  ICADV_REQUIRE(dw && dwpack && kind >= 0 && kind <= 3, "bad unpack_weight args");
  int taps = ksize * ksize;
  int64_t total = (int64_t)taps * c_out * c_in;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  unpack_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dwpack, dw, kind, c_out, c_in, taps, accumulate);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

static int launch_transpose(const float* src, float* dst, int batch, int rows, int cols, cudaStream_t s) {
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch), block(32, 8);
  ICADV_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "transpose grid too large");
  transpose_cp_kernel<<<grid, block, 0, s>>>(src, dst, rows, cols);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, icadv_stream_t stream) {
  ICADV_REQUIRE(src && dst, "null pointer");
  if (c <= 4) return launch_interleave_small<false, false>(src, dst, nullptr, n, c, h * w, as_stream(stream));
  return launch_transpose(src, dst, n, c, h * w, as_stream(stream));  // [n][c][hw] -> [n][hw][c]
}

int icadv_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, icadv_stream_t stream) {
  ICADV_REQUIRE(src && dst, "null pointer");
  if (c <= 4) return launch_interleave_small<true, false>(src, dst, nullptr, n, c, h * w, as_stream(stream));
  return launch_transpose(src, dst, n, h * w, c, as_stream(stream));  // [n][hw][c] -> [n][c][hw]
}

int icadv_clamp01_nhwc_to_nchw(const float* x_nhwc, float* y_nchw, int n, int c, int h, int w, icadv_stream_t stream) {
  ICADV_REQUIRE(x_nhwc && y_nchw, "null pointer");
  ICADV_REQUIRE(c >= 1 && c <= 4, "clamp01_nhwc_to_nchw takes images (1..4 channels)");
  return launch_interleave_small<true, true>(x_nhwc, y_nchw, nullptr, n, c, h * w, as_stream(stream));
}

int icadv_clamp01_backward_nchw_to_nhwc(const float* g_nchw, const float* x_nhwc, float* gx_nhwc, int n, int c, int h,
                                        int w, icadv_stream_t stream) {
  ICADV_REQUIRE(g_nchw && x_nhwc && gx_nhwc, "null pointer");
  ICADV_REQUIRE(c >= 1 && c <= 4, "clamp01_backward_nchw_to_nhwc takes images (1..4 channels)");
  return launch_interleave_small<false, true>(g_nchw, gx_nhwc, x_nhwc, n, c, h * w, as_stream(stream));
}

int icadv_gdn_reparam(const float* raw, float* eff, int rows, int cols, float bound, float pedestal, int transpose,
                      int round_tf32, icadv_stream_t stream) {
  ICADV_REQUIRE(raw && eff && rows > 0 && cols > 0, "bad gdn_reparam args");
  int total = rows * cols;
  gdn_reparam_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(raw, eff, rows, cols, bound, pedestal,
                                                                          transpose, round_tf32);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
