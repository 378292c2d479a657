// Kernels that only the codec UPDATE of adversarial training needs (train.py:335-366, `--adv`): the backward of
// the entropy models and of the GDN parameters, the rate / distortion terms of RateDistortionLoss (train.py:37-96)
// with their gradients, and the post-all-reduce step clip_grad_norm_(1.0) + Adam (train.py:360-361, coder.py:50-86)
// fused over ONE flat parameter buffer.  The attack loop never calls any of these.
// All reductions are two-stage with a fixed order (deterministic, no float atomics).
#include "icadv_common.cuh"

namespace icadv {

constexpr int kRedBlocksT = ICADV_RED_BLOCKS;

__device__ __forceinline__ float block_sum_256t(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
  }
  __syncthreads();
  return s;  // valid in thread 0
}

__device__ __forceinline__ float sigmoid_t(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------ EntropyBottleneck backward
// Forward (icadv_entropy.cu): v = x + noise; lo = logits(v - .5), up = logits(v + .5); sg = -sign(lo + up);
// lik = max(|sigmoid(sg up) - sigmoid(sg lo)|, bound).  Table rows as written by eb_prepare_kernel.
struct EbChain {
  float t[4][3], h[4][3], u;
};

__device__ __forceinline__ float eb_chain_fwd(float u, const float* __restrict__ T, int C, int c, EbChain& s) {
  s.u = u;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s.t[0][k] = __ldg(T + (int64_t)k * C + c) * u + __ldg(T + (int64_t)(3 + k) * C + c);
    s.h[0][k] = s.t[0][k] + __ldg(T + (int64_t)(6 + k) * C + c) * tanhf(s.t[0][k]);
  }
#pragma unroll
  for (int i = 1; i <= 3; ++i) {
    const int base = 9 + 15 * (i - 1);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float t = __ldg(T + (int64_t)(base + 9 + r) * C + c);
#pragma unroll
      for (int q = 0; q < 3; ++q) t += __ldg(T + (int64_t)(base + 3 * r + q) * C + c) * s.h[i - 1][q];
      s.t[i][r] = t;
      s.h[i][r] = t + __ldg(T + (int64_t)(base + 12 + r) * C + c) * tanhf(t);
    }
  }
  float out = __ldg(T + (int64_t)57 * C + c);
#pragma unroll
  for (int q = 0; q < 3; ++q) out += __ldg(T + (int64_t)(54 + q) * C + c) * s.h[3][q];
  return out;
}

// accumulates d(table) into acc[58] and returns d u
__device__ __forceinline__ float eb_chain_bwd(float g_out, const float* __restrict__ T, int C, int c, const EbChain& s,
                                              float* acc) {
  float dh[3];
  acc[57] += g_out;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    acc[54 + q] += g_out * s.h[3][q];
    dh[q] = g_out * __ldg(T + (int64_t)(54 + q) * C + c);
  }
#pragma unroll
  for (int i = 3; i >= 1; --i) {
    const int base = 9 + 15 * (i - 1);
    float dprev[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float th = tanhf(s.t[i][r]);
      const float tf = __ldg(T + (int64_t)(base + 12 + r) * C + c);
      acc[base + 12 + r] += dh[r] * th;
      const float dt = dh[r] * (1.f + tf * (1.f - th * th));
      acc[base + 9 + r] += dt;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        acc[base + 3 * r + q] += dt * s.h[i - 1][q];
        dprev[q] += dt * __ldg(T + (int64_t)(base + 3 * r + q) * C + c);
      }
    }
    dh[0] = dprev[0]; dh[1] = dprev[1]; dh[2] = dprev[2];
  }
  float du = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float th = tanhf(s.t[0][k]);
    const float tf = __ldg(T + (int64_t)(6 + k) * C + c);
    acc[6 + k] += dh[k] * th;
    const float dt = dh[k] * (1.f + tf * (1.f - th * th));
    acc[3 + k] += dt;
    acc[k] += dt * s.u;
    du += dt * __ldg(T + (int64_t)k * C + c);
  }
  return du;
}

// thread = channel (coalesced channels-last rows); block b owns rows [b*rows_per_block, ...); partial [b][58][C]
__global__ void eb_backward_kernel(const float* __restrict__ x_hat, const float* __restrict__ g_lik,
                                   const float* __restrict__ table, float* __restrict__ g_x,
                                   float* __restrict__ partial, int64_t rows, int rows_per_block, int C, float lik_bound) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float acc[58];
#pragma unroll
  for (int j = 0; j < 58; ++j) acc[j] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int64_t r = r0; r < r1; ++r) {
    const int64_t i = r * C + c;
    const float v = x_hat[i];
    float G = g_lik[i];
    EbChain lo_s, up_s;
    const float lo = eb_chain_fwd(v - 0.5f, table, C, c, lo_s), up = eb_chain_fwd(v + 0.5f, table, C, c, up_s);
    const float sum = lo + up;
    const float sg = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    const float su = sigmoid_t(sg * up), sl = sigmoid_t(sg * lo);
    const float d = su - sl;
    if (!(fabsf(d) >= lik_bound || G < 0.f)) G = 0.f;            // LowerBound backward rule
    const float gd = G * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    const float g_up = gd * sg * su * (1.f - su), g_lo = -gd * sg * sl * (1.f - sl);
    float dv = eb_chain_bwd(g_up, table, C, c, up_s, acc);
    dv += eb_chain_bwd(g_lo, table, C, c, lo_s, acc);
    g_x[i] = dv;                                                 // x_hat = x + noise
  }
#pragma unroll
  for (int j = 0; j < 58; ++j) partial[((int64_t)blockIdx.x * 58 + j) * C + c] = acc[j];
}

// raw parameter pointers as in eb_prepare
struct EbRawT {
  const float* matrix[5];
  const float* factor[4];
};

// fixed-order sum over blocks, then the chain rule through softplus / tanh of the raw parameters
__global__ void eb_backward_finalize_kernel(const float* __restrict__ partial, int n_blocks, EbRawT raw,
                                            float* __restrict__ graw, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int j = 0; j < 58; ++j) {
    float s = 0.f;
    for (int b = 0; b < n_blocks; ++b) s += partial[((int64_t)b * 58 + j) * C + c];
    // which raw parameter does row j belong to?
    float scale = 1.f;
    if (j < 3) scale = sigmoid_t(raw.matrix[0][c * 3 + j]);
    else if (j >= 6 && j < 9) { const float th = tanhf(raw.factor[0][c * 3 + (j - 6)]); scale = 1.f - th * th; }
    else if (j >= 9 && j < 54) {
      const int i = 1 + (j - 9) / 15, k = (j - 9) % 15;
      if (k < 9) scale = sigmoid_t(raw.matrix[i][c * 9 + k]);
      else if (k >= 12) { const float th = tanhf(raw.factor[i][c * 3 + (k - 12)]); scale = 1.f - th * th; }
    } else if (j >= 54 && j < 57) scale = sigmoid_t(raw.matrix[4][c * 3 + (j - 54)]);
    graw[(int64_t)j * C + c] = s * scale;
  }
}

// ------------------------------------------------------------------------------------------ GaussianConditional backward
__global__ void gc_backward_kernel(const float* __restrict__ y_hat, const float* __restrict__ scales,
                                   const float* __restrict__ means, const float* __restrict__ g_lik,
                                   float* __restrict__ g_y, float* __restrict__ g_scales, float* __restrict__ g_means,
                                   int64_t n, float scale_bound, float lik_bound) {
  const float kInvSqrt2 = 0.70710678118654752440f, kInvSqrt2Pi = 0.39894228040143267794f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mu = means != nullptr ? means[i] : 0.f;
    const float sraw = scales[i];
    const float s = fmaxf(sraw, scale_bound);
    const float dq = y_hat[i] - mu;
    const float v = fabsf(dq);
    const float a_up = (0.5f - v) / s, a_lo = (-0.5f - v) / s;
    const float lik = 0.5f * erfcf(-kInvSqrt2 * a_up) - 0.5f * erfcf(-kInvSqrt2 * a_lo);
    float G = g_lik[i];
    if (!(lik >= lik_bound || G < 0.f)) G = 0.f;                 // LowerBound backward rule on the likelihood
    const float p_up = kInvSqrt2Pi * expf(-0.5f * a_up * a_up), p_lo = kInvSqrt2Pi * expf(-0.5f * a_lo * a_lo);
    const float dl_dv = (p_lo - p_up) / s;
    const float dl_ds = -(p_up * a_up - p_lo * a_lo) / s;
    const float sgn = dq > 0.f ? 1.f : (dq < 0.f ? -1.f : 0.f);
    const float gy = G * dl_dv * sgn;
    g_y[i] = gy;
    if (g_means != nullptr) g_means[i] = -gy;
    float gs = G * dl_ds;
    if (!(sraw >= scale_bound || gs < 0.f)) gs = 0.f;            // LowerBound backward rule on the scale
    g_scales[i] = gs;
  }
}

// ------------------------------------------------------------------------------------------ RD-loss terms
// sum log(max(lik, floor)) over everything (train.py:62-64); partial sums per block, then one fixed-order pass
__global__ void __launch_bounds__(256) log_sum_kernel(const float* __restrict__ lik, float* __restrict__ ws, int64_t n,
                                                      float floor_) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += logf(fmaxf(lik[i], floor_));
  const float s = block_sum_256t(acc, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}
__global__ void sum_blocks_kernel(const float* __restrict__ ws, float* __restrict__ out, int n_blocks) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += (double)ws[b];
    out[0] = (float)s;
  }
}
// g_lik = scale / lik where lik >= floor (torch.clamp passes the gradient inside the range), else 0
__global__ void log_sum_backward_kernel(const float* __restrict__ lik, float* __restrict__ g, int64_t n, float floor_,
                                        const float* __restrict__ scale_dev, float scale_host) {
  const float sc = scale_host * (scale_dev != nullptr ? scale_dev[0] : 1.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float l = lik[i];
    g[i] = l >= floor_ ? sc / l : 0.f;
  }
}
// out = scale * (a - b)
__global__ void scaled_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                   int64_t n, const float* __restrict__ scale_dev, float scale_host) {
  const float sc = scale_host * (scale_dev != nullptr ? scale_dev[0] : 1.f);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = sc * (a[i] - b[i]);
}

// ------------------------------------------------------------------------------------------ GDN parameter gradients
// y = x * n^(-+1/2), n_i = beta_i + sum_j gamma_ij x_j^2.  With t_i = g_i y_i sc_i^2 (GDN) or g_i y_i / sc_i^2 (IGDN):
// d beta_i = -+1/2 sum_px t_i,   d gamma_ij = -+1/2 sum_px t_i x_j^2,   x_j = y_j / sc_j.
// grid (C/32, C/32, splits); block 32 x 8: thread (tx, ty) owns gamma rows i = 32*bx + ty*4 .. +3, column j = 32*by + tx.
__global__ void __launch_bounds__(256) gdn_param_grad_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                             const float* __restrict__ sc, float* __restrict__ partial,
                                                             int64_t n_px, int C, int inverse) {
  __shared__ float Ts[32][33], Xs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int64_t per = (n_px + gridDim.z - 1) / gridDim.z;
  const int64_t p0 = (int64_t)blockIdx.z * per, p1 = p0 + per < n_px ? p0 + per : n_px;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float accb = 0.f;   // beta partial: threads with ty == 0 and blockIdx.y == 0, channel i0 + tx
  for (int64_t p = p0; p < p1; p += 32) {
    // load 32 pixels x 32 channels of T (rows i0..) and X^2 (rows j0..): thread (tx = channel, ty*4.. = pixel)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t px = p + ty * 4 + k;
      float tv = 0.f, xv = 0.f;
      if (px < p1) {
        const int64_t a = px * C + i0 + tx, b = px * C + j0 + tx;
        const float s1 = sc[a], s2 = sc[b];
        const float s1sq = s1 * s1;
        tv = inverse ? (s1sq > 0.f ? g[a] * y[a] / s1sq : 0.f) : g[a] * y[a] * s1sq;
        const float xj = s2 > 0.f ? y[b] / s2 : 0.f;
        xv = xj * xj;
      }
      Ts[ty * 4 + k][tx] = tv;
      Xs[ty * 4 + k][tx] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const float xq = Xs[q][tx];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += Ts[q][ty * 4 + k] * xq;
      if (ty == 0) accb += Ts[q][tx];
    }
    __syncthreads();
  }
  const float sign = inverse ? 0.5f : -0.5f;
  float* pg = partial + (int64_t)blockIdx.z * ((int64_t)C * C + C);
#pragma unroll
  for (int k = 0; k < 4; ++k) pg[(int64_t)(i0 + ty * 4 + k) * C + j0 + tx] = sign * acc[k];
  if (ty == 0 && blockIdx.y == 0) pg[(int64_t)C * C + i0 + tx] = sign * accb;
}

// Tensor-path form of the same gradients: d gamma_ij = sum_px T_i X2_j is a contraction whose reduction axis is the pixel
// axis -- exactly the 1x1 case of the tcgen05 weight gradient (icadv_wgrad_tc.cu: "input" X2 = x^2, "output gradient"
// T = -+1/2 t), and d beta_i = sum_px T_i is its bias gradient.  This kernel writes the two operands (rounded to TF32 where
// they are produced, like every operand of the tensor path); the CUDA-core kernel above stays for the parity mode and for
// channel counts the tensor path does not take.  Config-5 update step: 6 x 0.30 ms -> see DESIGN.md.
__global__ void __launch_bounds__(256) gdn_param_operands_kernel(const float4* __restrict__ g, const float4* __restrict__ y,
                                                                 const float4* __restrict__ sc, float4* __restrict__ T,
                                                                 float4* __restrict__ X2, int64_t n4, int inverse) {
  const float half = inverse ? 0.5f : -0.5f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 gv = __ldg(g + i), yv = __ldg(y + i), sv = __ldg(sc + i);
    const float ga[4] = {gv.x, gv.y, gv.z, gv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w}, sa[4] = {sv.x, sv.y, sv.z, sv.w};
    float t[4], x2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s2 = sa[k] * sa[k];
      const float tv = inverse ? (s2 > 0.f ? ga[k] * ya[k] / s2 : 0.f) : ga[k] * ya[k] * s2;   // as gdn_param_grad_kernel
      const float xj = sa[k] > 0.f ? ya[k] / sa[k] : 0.f;
      uint32_t a, b;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(half * tv));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(xj * xj));
      t[k] = __uint_as_float(a); x2[k] = __uint_as_float(b);
    }
    T[i] = make_float4(t[0], t[1], t[2], t[3]);
    X2[i] = make_float4(x2[0], x2[1], x2[2], x2[3]);
  }
}

// fixed-order sum over splits + chain rule through the non-negative reparametrisation eff = max(raw, bound)^2 - pedestal
// with the LowerBound backward rule (pass if raw >= bound or the incoming gradient is negative).
// partial: per split `stride` floats; this launch finalises the `count` entries starting at `offset` of every split.
__global__ void gdn_param_grad_finalize_kernel(const float* __restrict__ partial, int splits, int64_t stride,
                                               int64_t offset, int64_t count, const float* __restrict__ raw,
                                               float* __restrict__ g_raw, float bound) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * stride + offset + i];
  const float r = raw[i];
  const float glb = s * 2.f * fmaxf(r, bound);
  g_raw[i] = (r >= bound || glb < 0.f) ? glb : 0.f;
}

// ------------------------------------------------------------------------------------------ clip_grad_norm_ + Adam
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, float* __restrict__ ws, int64_t n) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) acc += g[i] * g[i];
  const float s = block_sum_256t(acc, red);
  if (threadIdx.x == 0) ws[blockIdx.x] = s;
}

// torch.nn.utils.clip_grad_norm_(params, max_norm): coef = min(1, max_norm / (norm + 1e-6)); then torch.optim.Adam
// (no weight decay, no amsgrad).  sumsq (device scalar, may be NULL = no clipping) = sum of squares of ALL gradients.
__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t n, const float* __restrict__ sumsq, float max_norm,
                                 float grad_scale, float omb1, float b2, float omb2, float eps, float step_size,
                                 float bc2_sqrt) {
  float coef = grad_scale;   // grad_scale: 1/world after a SUM all-reduce (1 after an AVG all-reduce)
  if (sumsq != nullptr) {
    const float c = max_norm / (sqrtf(sumsq[0]) * grad_scale + 1e-6f);
    coef = (c < 1.f ? c : 1.f) * grad_scale;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * omb1;
    const float vi = v[i] * b2 + omb2 * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

static inline int ew_blocks_t(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_eb_backward_workspace_floats(int64_t rows, int C) {
  const int rows_per_block = 64;
  const int64_t nb = (rows + rows_per_block - 1) / rows_per_block;
  return (int)(nb * 58 * C);
}

int icadv_eb_backward(const float* x_hat, const float* g_lik, const float* table, const float* const* matrices,
                      const float* const* factors, float* g_x, float* g_raw, float* ws, int64_t rows, int C,
                      float lik_bound, icadv_stream_t stream) {
  ICADV_REQUIRE(x_hat && g_lik && table && matrices && factors && g_x && g_raw && ws, "null pointer");
  ICADV_REQUIRE(C >= 1 && C <= 1024 && rows >= 1, "eb_backward: C must be in [1,1024]");
  const int rows_per_block = 64;
  const int nb = (int)((rows + rows_per_block - 1) / rows_per_block);
  const int threads = (C + 31) / 32 * 32;
  eb_backward_kernel<<<nb, threads, 0, as_stream(stream)>>>(x_hat, g_lik, table, g_x, ws, rows, rows_per_block, C,
                                                           lik_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  EbRawT raw;
  for (int i = 0; i < 5; ++i) raw.matrix[i] = matrices[i];
  for (int i = 0; i < 4; ++i) raw.factor[i] = factors[i];
  eb_backward_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(ws, nb, raw, g_raw, C);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gc_backward(const float* y_hat, const float* scales, const float* means, const float* g_lik, float* g_y,
                      float* g_scales, float* g_means, int64_t n, float scale_bound, float lik_bound,
                      icadv_stream_t stream) {
  ICADV_REQUIRE(y_hat && scales && g_lik && g_y && g_scales, "null pointer");
  ICADV_REQUIRE((means == nullptr) == (g_means == nullptr), "gc_backward: means and g_means go together");
  gc_backward_kernel<<<ew_blocks_t(n), 256, 0, as_stream(stream)>>>(y_hat, scales, means, g_lik, g_y, g_scales, g_means,
                                                                  n, scale_bound, lik_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_log_sum(const float* lik, float* ws, float* out, int64_t n, float floor_, icadv_stream_t stream) {
  ICADV_REQUIRE(lik && ws && out && n >= 1, "bad log_sum args");
  log_sum_kernel<<<kRedBlocksT, 256, 0, as_stream(stream)>>>(lik, ws, n, floor_);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_blocks_kernel<<<1, 32, 0, as_stream(stream)>>>(ws, out, kRedBlocksT);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_log_sum_backward(const float* lik, float* g_lik, int64_t n, float floor_, const float* scale_dev,
                           float scale_host, icadv_stream_t stream) {
  ICADV_REQUIRE(lik && g_lik && n >= 1, "bad log_sum_backward args");
  log_sum_backward_kernel<<<ew_blocks_t(n), 256, 0, as_stream(stream)>>>(lik, g_lik, n, floor_, scale_dev, scale_host);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_scaled_diff(const float* a, const float* b, float* out, int64_t n, const float* scale_dev, float scale_host,
                      icadv_stream_t stream) {
  ICADV_REQUIRE(a && b && out && n >= 1, "bad scaled_diff args");
  scaled_diff_kernel<<<ew_blocks_t(n), 256, 0, as_stream(stream)>>>(a, b, out, n, scale_dev, scale_host);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_param_grad_workspace_floats(int C) { return 32 * (C * C + C); }

int icadv_gdn_param_grad(const float* g, const float* y, const float* sc, const float* beta_raw, const float* gamma_raw,
                         float* g_beta, float* g_gamma, float* ws, int64_t n_px, int C, int inverse, float beta_bound,
                         float gamma_bound, icadv_stream_t stream) {
  ICADV_REQUIRE(g && y && sc && beta_raw && gamma_raw && g_beta && g_gamma && ws, "null pointer");
  ICADV_REQUIRE(C % 32 == 0 && C >= 32 && n_px >= 1, "gdn_param_grad: C must be a multiple of 32");
  const int splits = 32;
  dim3 grid(C / 32, C / 32, splits);
  gdn_param_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, y, sc, ws, n_px, C, inverse);
  ICADV_CUDA_TRY(cudaGetLastError());
  const int64_t stride = (int64_t)C * C + C, cnt = (int64_t)C * C;   // per split: [C*C gamma | C beta]
  gdn_param_grad_finalize_kernel<<<(int)((cnt + 255) / 256), 256, 0, as_stream(stream)>>>(ws, splits, stride, 0, cnt,
                                                                                          gamma_raw, g_gamma, gamma_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  gdn_param_grad_finalize_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(ws, splits, stride, cnt, C, beta_raw,
                                                                                 g_beta, beta_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_param_operands(const float* g, const float* y, const float* sc, float* T, float* X2, int64_t n,
                             int inverse, icadv_stream_t stream) {
  ICADV_REQUIRE(g && y && sc && T && X2 && n >= 4 && n % 4 == 0, "bad gdn_param_operands args");
  gdn_param_operands_kernel<<<ew_blocks_t(n / 4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(sc),
      reinterpret_cast<float4*>(T), reinterpret_cast<float4*>(X2), n / 4, inverse);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_param_grad_finalize(const float* d_gamma_eff, const float* d_beta_eff, const float* beta_raw,
                                  const float* gamma_raw, float* g_beta, float* g_gamma, int C, float beta_bound,
                                  float gamma_bound, icadv_stream_t stream) {
  ICADV_REQUIRE(d_gamma_eff && d_beta_eff && beta_raw && gamma_raw && g_beta && g_gamma && C >= 1, "null pointer");
  const int64_t cnt = (int64_t)C * C;
  gdn_param_grad_finalize_kernel<<<(int)((cnt + 255) / 256), 256, 0, as_stream(stream)>>>(d_gamma_eff, 1, 0, 0, cnt, gamma_raw,
                                                                                          g_gamma, gamma_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  gdn_param_grad_finalize_kernel<<<(C + 255) / 256, 256, 0, as_stream(stream)>>>(d_beta_eff, 1, 0, 0, C, beta_raw, g_beta,
                                                                                 beta_bound);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_sumsq(const float* g, float* ws, float* out, int64_t n, icadv_stream_t stream) {
  ICADV_REQUIRE(g && ws && out && n >= 1, "bad sumsq args");
  sumsq_kernel<<<kRedBlocksT, 256, 0, as_stream(stream)>>>(g, ws, n);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_blocks_kernel<<<1, 32, 0, as_stream(stream)>>>(ws, out, kRedBlocksT);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_adam_clip_step(float* params, const float* grads, float* m, float* v, int64_t n, const float* sumsq,
                         float max_norm, float grad_scale, double lr, double beta1, double beta2, double eps, int step,
                         icadv_stream_t stream) {
  ICADV_REQUIRE(params && grads && m && v && n >= 1 && step >= 1, "bad adam_clip_step args");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  adam_clip_kernel<<<ew_blocks_t(n), 256, 0, as_stream(stream)>>>(params, grads, m, v, n, sumsq, max_norm,
                                                                grad_scale, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                                (float)eps, (float)(lr / bc1), (float)sqrt(bc2));
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
