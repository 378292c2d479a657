// CUDA-core contraction path for the shapes the tcgen05 path does not take: the 3-channel end layers
// (g_a.0: 3 -> N, g_s.6: N -> 3 and their input-gradients), odd channel counts, and the weight gradient.
// Same geometry (icadv_common.cuh) and packed-weight layout as the tensor path; ICADV_EPI_LINEAR only.
// Reference ops replaced: nn.Conv2d / nn.ConvTranspose2d forward + autograd (anchors/utils.py:112-130).
#include "icadv_common.cuh"

namespace icadv {

struct SimtParams {
  const float* in; const float* w; const float* bias; float* out;
  int n_img, in_h, in_w, out_h, out_w, tile_h, tile_w;
  int K, N, num_taps, stride_in, stride_out, out_a, out_b, act;
  int sconv;  // 1: input pixel = stride*(i,j)+(dy,dx) with dy,dx = kh-p ; 0: input pixel = (i+dy, j+dx)
  const int* active; const int* n_active;
  int8_t dy[kMaxTaps], dx[kMaxTaps], wt[kMaxTaps];
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ICADV_ACT_RELU) return fmaxf(v, 0.f);
  if (act == ICADV_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == ICADV_ACT_ABS) return fabsf(v);
  return v;
}

// Tiled fp32 implicit GEMM: block = 64 tile-space pixels x 64 output channels, reduction over
// r = tap*K + k in chunks of 16.  256 threads, 4x4 outputs per thread.
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int slot = blockIdx.z;
  if (p.n_active != nullptr && slot >= *p.n_active) return;
  const int img = p.active != nullptr ? p.active[slot] : slot;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int npx = p.tile_h * p.tile_w;
  const int R = p.num_taps * p.K;

  // loader roles: each thread loads 4 consecutive r for one pixel (A) and one channel (B)
  const int lp = tid >> 2, lr = (tid & 3) * 4;
  const int m = m0 + lp;
  const bool m_ok = m < npx;
  const int pi = m_ok ? m / p.tile_w : 0, pj = m_ok ? m % p.tile_w : 0;
  const int bn = n0 + lp;
  const bool bn_ok = bn < p.N;
  const float* in_img = p.in + (int64_t)img * p.in_h * p.in_w * p.K;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int r0 = 0; r0 < R; r0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = r0 + lr + e;
      float av = 0.f, bv = 0.f;
      if (r < R) {
        const int t = r / p.K, k = r - t * p.K;
        if (m_ok) {
          const int ih = p.sconv ? p.stride_in * pi + p.dy[t] : pi + p.dy[t];
          const int iw = p.sconv ? p.stride_in * pj + p.dx[t] : pj + p.dx[t];
          if (ih >= 0 && ih < p.in_h && iw >= 0 && iw < p.in_w)
            av = __ldg(in_img + ((int64_t)ih * p.in_w + iw) * p.K + k);
        }
        if (bn_ok) bv = __ldg(p.w + ((int64_t)p.wt[t] * p.N + bn) * p.K + k);
      }
      As[lr + e][lp] = av;
      Bs[lr + e][lp] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  float* out_img = p.out + (int64_t)img * p.out_h * p.out_w * p.N;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int mm = m0 + ty * 4 + x;
    if (mm >= npx) continue;
    const int i = mm / p.tile_w, j = mm % p.tile_w;
    const int oh = p.stride_out * i + p.out_a, ow = p.stride_out * j + p.out_b;
    if (oh >= p.out_h || ow >= p.out_w) continue;
    float* o = out_img + ((int64_t)oh * p.out_w + ow) * p.N;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int n = n0 + tx * 4 + y;
      if (n < p.N) o[n] = act_apply(acc[x][y] + (p.bias ? __ldg(p.bias + n) : 0.f), p.act);
    }
  }
}

// Narrow-output variant (N <= 4, e.g. the RGB end layers with K = 128/192): one warp per tile-space
// pixel, lanes split K, fixed-order butterfly reduction.  Weights for the launch's taps sit in smem.
__global__ void __launch_bounds__(256) conv_simt_narrow_kernel(const SimtParams p, int px_per_block) {
  extern __shared__ float wsm[];  // [num_taps][N][K]
  const int slot = blockIdx.z;
  if (p.n_active != nullptr && slot >= *p.n_active) return;
  const int img = p.active != nullptr ? p.active[slot] : slot;
  const int K = p.K, N = p.N;
  for (int i = threadIdx.x; i < p.num_taps * N * K; i += blockDim.x) {
    const int t = i / (N * K), rem = i - t * N * K;
    wsm[i] = __ldg(p.w + (int64_t)p.wt[t] * N * K + rem);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int npx = p.tile_h * p.tile_w;
  const float* in_img = p.in + (int64_t)img * p.in_h * p.in_w * K;
  float* out_img = p.out + (int64_t)img * p.out_h * p.out_w * N;
  const int m_end = min(npx, (int)(blockIdx.x + 1) * px_per_block);
  for (int m = blockIdx.x * px_per_block + warp; m < m_end; m += nwarp) {
    const int i = m / p.tile_w, j = m % p.tile_w;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < p.num_taps; ++t) {
      const int ih = p.sconv ? p.stride_in * i + p.dy[t] : i + p.dy[t];
      const int iw = p.sconv ? p.stride_in * j + p.dx[t] : j + p.dx[t];
      if (ih < 0 || ih >= p.in_h || iw < 0 || iw >= p.in_w) continue;  // warp-uniform
      const float* src = in_img + ((int64_t)ih * p.in_w + iw) * K;
      const float* wt = wsm + t * N * K;
      for (int k = lane * 4; k < K; k += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + k));
        for (int n = 0; n < N; ++n) {
          const float4 w4 = *reinterpret_cast<const float4*>(wt + n * K + k);
          acc[n] = fmaf(a.x, w4.x, fmaf(a.y, w4.y, fmaf(a.z, w4.z, fmaf(a.w, w4.w, acc[n]))));
        }
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    const int oh = p.stride_out * i + p.out_a, ow = p.stride_out * j + p.out_b;
    if (lane < N && oh < p.out_h && ow < p.out_w) {
      float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      out_img[((int64_t)oh * p.out_w + ow) * N + lane] = act_apply(v + (p.bias ? __ldg(p.bias + lane) : 0.f), p.act);
    }
  }
}

// Weight gradient: dW[wt][n][k] = sum over images, tile pixels of gout[out px][n] * in[tap px][k].
// One block per (tap, 32x32 (n,k) sub-tile); the pixel loop is split over blockIdx.z and reduced by a
// second pass in fixed order (deterministic).  Only used by the --adv training step.
__global__ void __launch_bounds__(256) conv_wgrad_partial_kernel(const SimtParams p, const float* __restrict__ gout,
                                                                 float* __restrict__ partial, int splits) {
  __shared__ float Gs[32][33];
  __shared__ float Is[32][33];
  const int t = blockIdx.x;
  const int tiles_k = (p.K + 31) / 32;
  const int n0 = (blockIdx.y / tiles_k) * 32, k0 = (blockIdx.y % tiles_k) * 32;
  const int split = blockIdx.z;
  const int npx = p.tile_h * p.tile_w;
  const int64_t total = (int64_t)p.n_img * npx;
  const int64_t per = (total + splits - 1) / splits;
  const int64_t beg = split * per, end = min(total, beg + per);
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t b0 = beg; b0 < end; b0 += 32) {
    for (int rr = tyy; rr < 32; rr += 8) {
      const int64_t mm = b0 + rr;
      float gv = 0.f, iv = 0.f;
      if (mm < end) {
        const int img = (int)(mm / npx), m = (int)(mm % npx);
        const int i = m / p.tile_w, j = m % p.tile_w;
        const int oh = p.stride_out * i + p.out_a, ow = p.stride_out * j + p.out_b;
        const int ih = p.sconv ? p.stride_in * i + p.dy[t] : i + p.dy[t];
        const int iw = p.sconv ? p.stride_in * j + p.dx[t] : j + p.dx[t];
        if (oh < p.out_h && ow < p.out_w && n0 + tx < p.N)
          gv = __ldg(gout + (((int64_t)img * p.out_h + oh) * p.out_w + ow) * p.N + n0 + tx);
        if (ih >= 0 && ih < p.in_h && iw >= 0 && iw < p.in_w && k0 + tx < p.K)
          iv = __ldg(p.in + (((int64_t)img * p.in_h + ih) * p.in_w + iw) * p.K + k0 + tx);
      }
      Gs[rr][tx] = gv;
      Is[rr][tx] = iv;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float iv = Is[r][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(Gs[r][tyy * 4 + q], iv, acc[q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int n = n0 + tyy * 4 + q, k = k0 + tx;
    if (n < p.N && k < p.K)
      partial[((int64_t)split * p.num_taps + t) * p.N * p.K + (int64_t)n * p.K + k] = acc[q];
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, const SimtParams p,
                                    int splits, int accumulate) {
  const int64_t per_tap = (int64_t)p.N * p.K;
  const int64_t total = (int64_t)p.num_taps * per_tap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += partial[(int64_t)sp * total + i];
    const int t = (int)(i / per_tap);
    const int64_t dst = (int64_t)p.wt[t] * per_tap + (i - (int64_t)t * per_tap);
    dw[dst] = accumulate ? dw[dst] + s : s;
  }
}

// dbias[n] = sum over all pixels of gout[px][n].  Two stages, both fixed-order: block (x = 32 channels, y = pixel split)
// writes partial[y][n]; the second kernel adds the splits.
constexpr int kBiasSplits = 128;
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ gout, float* __restrict__ partial,
                                                        int64_t npx, int N) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, tyy = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int64_t per = (npx + gridDim.y - 1) / gridDim.y;
  const int64_t m0 = (int64_t)blockIdx.y * per, m1 = m0 + per < npx ? m0 + per : npx;
  float s = 0.f;
  if (n < N)
    for (int64_t m = m0 + tyy; m < m1; m += 8) s += __ldg(gout + m * N + n);
  red[tyy][tx] = s;
  __syncthreads();
  if (tyy == 0 && n < N) {
    float a = 0.f;
    for (int r = 0; r < 8; ++r) a += red[r][tx];
    partial[(int64_t)blockIdx.y * N + n] = a;
  }
}
__global__ void bias_grad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dbias, int N, int splits) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float a = 0.f;
  for (int y = 0; y < splits; ++y) a += partial[(int64_t)y * N + n];
  dbias[n] = a;
}

static int fill_params(const icadv_conv_desc* d, const Geometry& g, int l, SimtParams* p) {
  p->in = d->in; p->w = d->wpack; p->bias = d->bias; p->out = d->out;
  p->n_img = d->n_img; p->in_h = d->in_h; p->in_w = d->in_w; p->out_h = g.out_h; p->out_w = g.out_w;
  p->tile_h = g.tile_h; p->tile_w = g.tile_w; p->K = d->k_ch; p->N = d->n_ch;
  p->num_taps = g.n_taps[l]; p->act = d->act; p->active = d->active; p->n_active = d->n_active;
  p->sconv = d->form == ICADV_FORM_SCONV ? 1 : 0;
  if (p->sconv) { p->stride_in = d->stride; p->stride_out = 1; p->out_a = p->out_b = 0; }
  else { p->stride_in = 1; p->stride_out = d->stride; p->out_a = g.out_a[l]; p->out_b = g.out_b[l]; }
  for (int t = 0; t < p->num_taps; ++t) {
    const Tap& tp = g.taps[l][t];
    if (p->sconv) {  // undo the parity-plane encoding: offsets relative to stride*(i,j)
      const int kh = tp.wtap / g.ksize, kw = tp.wtap % g.ksize;
      p->dy[t] = (int8_t)(kh - g.pad); p->dx[t] = (int8_t)(kw - g.pad);
    } else {
      p->dy[t] = (int8_t)tp.dy; p->dx[t] = (int8_t)tp.dx;
    }
    p->wt[t] = (int8_t)tp.wtap;
  }
  return ICADV_OK;
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_conv_simt(const icadv_conv_desc* d, icadv_stream_t stream) {
  ICADV_REQUIRE(d && d->in && d->out && d->wpack, "null pointer");
  ICADV_REQUIRE(d->epi == ICADV_EPI_LINEAR && !d->acc_from_in, "conv_simt: linear epilogue only");
  Geometry g;
  int rc = make_geometry(d, &g);
  if (rc) return rc;
  ICADV_REQUIRE(d->n_img <= 65535, "n_img too large");
  for (int l = 0; l < g.n_launch; ++l) {
    SimtParams p;
    fill_params(d, g, l, &p);
    if (p.num_taps == 0) continue;
    const int npx = g.tile_h * g.tile_w;
    const size_t wbytes = (size_t)p.num_taps * p.N * p.K * sizeof(float);
    if (p.N <= 4 && p.K % 4 == 0 && wbytes <= 96 * 1024) {
      static bool attr_done = false;
      if (!attr_done) {
        ICADV_CUDA_TRY(cudaFuncSetAttribute(conv_simt_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            96 * 1024));
        attr_done = true;
      }
      const int px_per_block = 256;
      dim3 grid((npx + px_per_block - 1) / px_per_block, 1, d->n_img);
      conv_simt_narrow_kernel<<<grid, 256, wbytes, as_stream(stream)>>>(p, px_per_block);
    } else {
      dim3 grid((npx + BM - 1) / BM, (p.N + BN - 1) / BN, d->n_img);
      conv_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
    }
    ICADV_CUDA_TRY(cudaGetLastError());
  }
  return ICADV_OK;
}

// `d` describes the FORWARD contraction (in, geometry); gout has the shape of d->out.
// dwpack: [taps][n_ch][k_ch] (overwritten), dbias: [n_ch] or NULL.  Workspace is allocated
// stream-ordered (cudaMallocAsync) -- this entry point is not on the attack path.
// path: ICADV_WGRAD_AUTO = tcgen05 kernel where the shape is eligible (icadv_wgrad_tc.cu; operands read as TF32),
// ICADV_WGRAD_SIMT = fp32 CUDA-core kernel (parity mode; shapes the tensor path does not take), ICADV_WGRAD_TC = tensor
// path or an error.
int icadv_conv_wgrad_ex(const icadv_conv_desc* d, const float* gout, float* dwpack, float* dbias, int path,
                        icadv_stream_t stream) {
  ICADV_REQUIRE(d && d->in && gout && dwpack, "null pointer");
  ICADV_REQUIRE(d->active == nullptr, "wgrad does not take an image indirection");
  Geometry g;
  int rc = make_geometry(d, &g);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const int taps_total = d->ksize * d->ksize;
  const bool tc = path == ICADV_WGRAD_TC || (path == ICADV_WGRAD_AUTO && wgrad_tc_supported(d));
  if (tc) {
    rc = wgrad_tc(d, g, gout, dwpack, s);
    if (rc) return rc;
  } else {
  ICADV_CUDA_TRY(cudaMemsetAsync(dwpack, 0, (size_t)taps_total * d->n_ch * d->k_ch * sizeof(float), s));
  for (int l = 0; l < g.n_launch; ++l) {
    SimtParams p;
    fill_params(d, g, l, &p);
    if (p.num_taps == 0) continue;
    const int64_t total = (int64_t)d->n_img * g.tile_h * g.tile_w;
    int splits = (int)((total + 4095) / 4096);
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    float* partial = nullptr;
    const size_t bytes = (size_t)splits * p.num_taps * p.N * p.K * sizeof(float);
    ICADV_CUDA_TRY(cudaMallocAsync(&partial, bytes, s));
    dim3 grid(p.num_taps, ((p.N + 31) / 32) * ((p.K + 31) / 32), splits);
    conv_wgrad_partial_kernel<<<grid, 256, 0, s>>>(p, gout, partial, splits);
    ICADV_CUDA_TRY(cudaGetLastError());
    const int64_t tot = (int64_t)p.num_taps * p.N * p.K;
    int blocks = (int)((tot + 255) / 256);
    if (blocks > 2048) blocks = 2048;
    wgrad_reduce_kernel<<<blocks, 256, 0, s>>>(partial, dwpack, p, splits, 0);
    ICADV_CUDA_TRY(cudaGetLastError());
    ICADV_CUDA_TRY(cudaFreeAsync(partial, s));
  }
  }
  if (dbias != nullptr) {
    const int64_t npx = (int64_t)d->n_img * g.out_h * g.out_w;
    float* partial = nullptr;
    ICADV_CUDA_TRY(cudaMallocAsync(&partial, (size_t)kBiasSplits * d->n_ch * sizeof(float), s));
    bias_grad_kernel<<<dim3((d->n_ch + 31) / 32, kBiasSplits), 256, 0, s>>>(gout, partial, npx, d->n_ch);
    ICADV_CUDA_TRY(cudaGetLastError());
    bias_grad_reduce_kernel<<<(d->n_ch + 127) / 128, 128, 0, s>>>(partial, dbias, d->n_ch, kBiasSplits);
    ICADV_CUDA_TRY(cudaGetLastError());
    ICADV_CUDA_TRY(cudaFreeAsync(partial, s));
  }
  return ICADV_OK;
}

int icadv_conv_wgrad(const icadv_conv_desc* d, const float* gout, float* dwpack, float* dbias,
                     icadv_stream_t stream) {
  return icadv_conv_wgrad_ex(d, gout, dwpack, dbias, ICADV_WGRAD_AUTO, stream);
}

int icadv_conv_wgrad_tc_supported(const icadv_conv_desc* d) { return d != nullptr && wgrad_tc_supported(d) ? 1 : 0; }

}  // extern "C"
