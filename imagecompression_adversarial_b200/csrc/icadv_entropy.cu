// Entropy-model kernels (bandwidth bound: one pass over the latent, coalesced channels-last access,
// per-image rate reduced in fixed order):
//   * EntropyBottleneck factorised density (compressai) -- logits chain 1-3-3-3-3-1 held in
//     registers, softplus/tanh of the parameters hoisted into a prepared [58][C] table;
//   * GaussianConditional discretised Gaussian with erfc (compressai; A.4; reference pin
//     visual_distribution.py:85-100);
//   * the round / additive-noise quantisers (utils/ops.py:8-25; compressai quantize, anchors/model.py:102);
//   * rate:  bits[n] = -sum log2(max(lik, floor))   (attack_rd.py:419, self_ensemble.py:222, train.py:60-64).
#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

constexpr int kRedBlocks = ICADV_RED_BLOCKS;
constexpr int kEbTable = 58;  // prepared floats per channel

__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float block_sum_256e(float v, float* red) {
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
  return s;
}

// raw parameter pointers (torch layouts): matrix_i [C][f_{i+1}][f_i], bias_i [C][f_{i+1}][1], factor_i [C][f_{i+1}][1]
struct EbRaw {
  const float* matrix[5];
  const float* bias[5];
  const float* factor[4];
};

// table layout (row j, column c): rows 0-2 softplus(M0)[3], 3-5 b0, 6-8 tanh(f0);
// for i=1..3: base 9+15(i-1): 9 softplus(M_i) row-major [out][in], 3 b_i, 3 tanh(f_i); rows 54-56 softplus(M4)[3], 57 b4
__global__ void eb_prepare_kernel(EbRaw raw, float* __restrict__ table, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  auto put = [&](int row, float v) { table[(int64_t)row * C + c] = v; };
  for (int k = 0; k < 3; ++k) {
    put(k, softplusf(raw.matrix[0][c * 3 + k]));
    put(3 + k, raw.bias[0][c * 3 + k]);
    put(6 + k, tanhf(raw.factor[0][c * 3 + k]));
  }
  for (int i = 1; i <= 3; ++i) {
    const int base = 9 + 15 * (i - 1);
    for (int k = 0; k < 9; ++k) put(base + k, softplusf(raw.matrix[i][c * 9 + k]));
    for (int k = 0; k < 3; ++k) {
      put(base + 9 + k, raw.bias[i][c * 3 + k]);
      put(base + 12 + k, tanhf(raw.factor[i][c * 3 + k]));
    }
  }
  for (int k = 0; k < 3; ++k) put(54 + k, softplusf(raw.matrix[4][c * 3 + k]));
  put(57, raw.bias[4][c]);
}

__device__ __forceinline__ float eb_logits(float v, const float* __restrict__ T, int C, int c) {
  float h[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float t = __ldg(T + (int64_t)k * C + c) * v + __ldg(T + (int64_t)(3 + k) * C + c);
    h[k] = t + __ldg(T + (int64_t)(6 + k) * C + c) * tanhf(t);
  }
#pragma unroll
  for (int i = 1; i <= 3; ++i) {
    const int base = 9 + 15 * (i - 1);
    float o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float t = __ldg(T + (int64_t)(base + 3 * r) * C + c) * h[0] + __ldg(T + (int64_t)(base + 3 * r + 1) * C + c) * h[1] +
                __ldg(T + (int64_t)(base + 3 * r + 2) * C + c) * h[2] + __ldg(T + (int64_t)(base + 9 + r) * C + c);
      o[r] = t + __ldg(T + (int64_t)(base + 12 + r) * C + c) * tanhf(t);
    }
    h[0] = o[0]; h[1] = o[1]; h[2] = o[2];
  }
  return __ldg(T + (int64_t)54 * C + c) * h[0] + __ldg(T + (int64_t)55 * C + c) * h[1] +
         __ldg(T + (int64_t)56 * C + c) * h[2] + __ldg(T + (int64_t)57 * C + c);
}

// x: [n_img][px][C] channels-last.  mode 0: x_hat = round(x - med) + med ; mode 1: x_hat = x + noise
__global__ void __launch_bounds__(256) eb_forward_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                         const float* __restrict__ table,
                                                         const float* __restrict__ medians, float* __restrict__ x_hat,
                                                         float* __restrict__ lik, float* __restrict__ ws,
                                                         int64_t per_img, int C, int mode, float lik_bound,
                                                         float bits_floor) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const float xv = x[base + i];
    float v;
    if (mode == 0) {
      const float med = __ldg(medians + c);
      v = rintf(xv - med) + med;  // torch.round = round-half-to-even
    } else {
      v = xv + noise[base + i];
    }
    const float lo = eb_logits(v - 0.5f, table, C, c), up = eb_logits(v + 0.5f, table, C, c);
    const float sum = lo + up;
    const float sg = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    float l = fabsf(sigmoidf(sg * up) - sigmoidf(sg * lo));
    l = fmaxf(l, lik_bound);
    x_hat[base + i] = v;
    lik[base + i] = l;
    acc -= log2f(fmaxf(l, bits_floor));
  }
  const float s = block_sum_256e(acc, red);
  if (threadIdx.x == 0) ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) gc_forward_kernel(const float* __restrict__ y, const float* __restrict__ scales,
                                                         const float* __restrict__ means,
                                                         const float* __restrict__ noise, float* __restrict__ y_hat,
                                                         float* __restrict__ lik, float* __restrict__ ws,
                                                         int64_t per_img, int mode, float scale_bound, float lik_bound,
                                                         float bits_floor) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img;
  float acc = 0.f;
  const float kInvSqrt2 = 0.70710678118654752440f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img; i += (int64_t)gridDim.x * 256) {
    const float yv = y[base + i];
    const float mu = means != nullptr ? means[base + i] : 0.f;
    float q;
    if (mode == 0) q = rintf(yv - mu) + mu; else q = yv + noise[base + i];
    const float s = fmaxf(scales[base + i], scale_bound);
    const float v = fabsf(q - mu);
    // Phi(t) = 0.5 erfc(-t / sqrt(2))
    const float up = 0.5f * erfcf(-kInvSqrt2 * ((0.5f - v) / s));
    const float lo = 0.5f * erfcf(-kInvSqrt2 * ((-0.5f - v) / s));
    float l = fmaxf(up - lo, lik_bound);
    y_hat[base + i] = q;
    lik[base + i] = l;
    acc -= log2f(fmaxf(l, bits_floor));
  }
  const float s = block_sum_256e(acc, red);
  if (threadIdx.x == 0) ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
}

__global__ void sum_ws_kernel(const float* __restrict__ ws, float* __restrict__ out, int n_img) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_img) return;
  float s = 0.f;
  for (int b = 0; b < kRedBlocks; ++b) s += ws[(int64_t)n * kRedBlocks + b];
  out[n] = s;
}

// elementwise helpers at the operator surface: 0 abs, 1 relu, 2 leaky(0.01), 3 round, 4 add (y = x + b), 5 round to TF32,
// 6 clamp to [0, 1] (torch.clamp of attack_rd.py:411, self_ensemble.py:182,207), 7 copy; + 256: the result is rounded to TF32
// on store (its only readers are tensor-path contractions: saves their separate rounding pass)
template <int OP, bool RND>
__device__ __forceinline__ void unary_loop(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y,
                                           int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    float r;
    if (OP == 0) r = fabsf(v);
    else if (OP == 1) r = fmaxf(v, 0.f);
    else if (OP == 2) r = v > 0.f ? v : 0.01f * v;
    else if (OP == 3) r = rintf(v);
    else if (OP == 5) r = round_tf32(v);
    else if (OP == 6) r = fminf(fmaxf(v, 0.f), 1.f);
    else if (OP == 7) r = v;
    else r = v + b[i];
    y[i] = RND ? round_tf32(r) : r;
  }
}
// the operation is resolved OUTSIDE the element loop (one specialised loop per op: a switch inside the loop cost 1.6x on
// these bandwidth-bound passes)
__global__ void unary_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y, int64_t n,
                             int op) {
  const bool rnd = (op & 256) != 0;
  switch (op & 255) {
    case 0: rnd ? unary_loop<0, true>(x, b, y, n) : unary_loop<0, false>(x, b, y, n); break;
    case 1: rnd ? unary_loop<1, true>(x, b, y, n) : unary_loop<1, false>(x, b, y, n); break;
    case 2: rnd ? unary_loop<2, true>(x, b, y, n) : unary_loop<2, false>(x, b, y, n); break;
    case 3: unary_loop<3, false>(x, b, y, n); break;
    case 5: unary_loop<5, false>(x, b, y, n); break;
    case 6: rnd ? unary_loop<6, true>(x, b, y, n) : unary_loop<6, false>(x, b, y, n); break;
    case 7: rnd ? unary_loop<7, true>(x, b, y, n) : unary_loop<7, false>(x, b, y, n); break;
    default: rnd ? unary_loop<4, true>(x, b, y, n) : unary_loop<4, false>(x, b, y, n); break;
  }
}

template <int OP, bool RND>
__device__ __forceinline__ void act_backward_loop(const float* __restrict__ x, const float* __restrict__ g,
                                                  float* __restrict__ gx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i], gv = g[i];
    float r;
    if (OP == 0) r = v > 0.f ? gv : (v < 0.f ? -gv : 0.f);
    else if (OP == 1) r = v > 0.f ? gv : 0.f;
    else r = v > 0.f ? gv : 0.01f * gv;
    gx[i] = RND ? round_tf32(r) : r;
  }
}
// gradient of abs / relu / leaky given the forward INPUT (or, for relu/leaky, equivalently the output); + 256: gradient
// rounded to TF32 on store (it feeds a tensor-path input gradient)
__global__ void act_backward_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx,
                                    int64_t n, int op) {
  const bool rnd = (op & 256) != 0;
  switch (op & 255) {
    case 0: rnd ? act_backward_loop<0, true>(x, g, gx, n) : act_backward_loop<0, false>(x, g, gx, n); break;
    case 1: rnd ? act_backward_loop<1, true>(x, g, gx, n) : act_backward_loop<1, false>(x, g, gx, n); break;
    default: rnd ? act_backward_loop<2, true>(x, g, gx, n) : act_backward_loop<2, false>(x, g, gx, n); break;
  }
}

// compressai.layers.AttentionBlock gate (cheng2020_attn): y = a * sigmoid(b) + x and its gradients
__global__ void attention_gate_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                      const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] * (1.f / (1.f + expf(-b[i]))) + x[i];
}
__global__ void attention_gate_backward_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                               const float* __restrict__ g, float* __restrict__ ga,
                                               float* __restrict__ gb, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = 1.f / (1.f + expf(-b[i])), gv = g[i];
    ga[i] = gv * s;
    gb[i] = gv * a[i] * s * (1.f - s);
  }
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_attention_gate(const float* a, const float* b, const float* x, float* y, int64_t n, icadv_stream_t stream) {
  ICADV_REQUIRE(a && b && x && y && n > 0, "bad attention_gate args");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attention_gate_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(a, b, x, y, n);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_attention_gate_backward(const float* a, const float* b, const float* g, float* ga, float* gb, int64_t n,
                                  icadv_stream_t stream) {
  ICADV_REQUIRE(a && b && g && ga && gb && n > 0, "bad attention_gate_backward args");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  attention_gate_backward_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(a, b, g, ga, gb, n);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_eb_prepare(const float* const* matrices, const float* const* biases, const float* const* factors,
                     float* table, int C, icadv_stream_t stream) {
  ICADV_REQUIRE(matrices && biases && factors && table && C > 0, "bad eb_prepare args");
  EbRaw raw;
  for (int i = 0; i < 5; ++i) { raw.matrix[i] = matrices[i]; raw.bias[i] = biases[i]; }
  for (int i = 0; i < 4; ++i) raw.factor[i] = factors[i];
  eb_prepare_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(raw, table, C);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_eb_forward(const float* x, const float* noise, const float* table, const float* medians, float* x_hat,
                     float* lik, float* ws, float* bits, int n_img, int64_t per_img, int C, int mode,
                     float lik_bound, float bits_floor, icadv_stream_t stream) {
  ICADV_REQUIRE(x && table && medians && x_hat && lik && ws && bits, "null pointer");
  ICADV_REQUIRE(mode == 0 || noise != nullptr, "train mode needs a noise tensor");
  ICADV_REQUIRE(per_img % C == 0, "per_img must be a multiple of C (channels-last)");
  dim3 grid(kRedBlocks, n_img);
  eb_forward_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, noise, table, medians, x_hat, lik, ws, per_img, C, mode,
                                                         lik_bound, bits_floor);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_ws_kernel<<<(n_img + 127) / 128, 128, 0, as_stream(stream)>>>(ws, bits, n_img);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gc_forward(const float* y, const float* scales, const float* means, const float* noise, float* y_hat,
                     float* lik, float* ws, float* bits, int n_img, int64_t per_img, int mode, float scale_bound,
                     float lik_bound, float bits_floor, icadv_stream_t stream) {
  ICADV_REQUIRE(y && scales && y_hat && lik && ws && bits, "null pointer");
  ICADV_REQUIRE(mode == 0 || noise != nullptr, "train mode needs a noise tensor");
  dim3 grid(kRedBlocks, n_img);
  gc_forward_kernel<<<grid, 256, 0, as_stream(stream)>>>(y, scales, means, noise, y_hat, lik, ws, per_img, mode,
                                                         scale_bound, lik_bound, bits_floor);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_ws_kernel<<<(n_img + 127) / 128, 128, 0, as_stream(stream)>>>(ws, bits, n_img);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_unary(const float* x, const float* b, float* y, int64_t n, int op, icadv_stream_t stream) {
  ICADV_REQUIRE(x && y && op >= 0 && (op & 255) <= 7 && (op & ~511) == 0 && ((op & 255) != 4 || b), "bad unary args");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  unary_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, b, y, n, op);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_act_backward(const float* x, const float* g, float* gx, int64_t n, int op, icadv_stream_t stream) {
  ICADV_REQUIRE(x && g && gx && op >= 0 && (op & 255) <= 2 && (op & ~511) == 0, "bad act_backward args");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  act_backward_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, g, gx, n, op);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
