// Bandwidth-bound kernels of the perturbation step: eps-clamp / [0,1]-clamp projection with the custom
// Low_bound/Up_bound backward rule, per-image budget test and branch compaction (device side: no
// host sync), Adam or sign update, output clamp + distortion + gradient seed.
// Reference: attack_rd.py:333-334,353-364,507,517,546-554; utils/ops.py:28-56; torch.optim.Adam;
//            attack_ifgsm.py:348-362,409-418.
// All reductions are two-stage with a fixed summation order (deterministic, no float atomics).
#include "icadv_common.cuh"

namespace icadv {

constexpr int kRedBlocks = ICADV_RED_BLOCKS;  // partial sums per image

__device__ __forceinline__ float block_sum_256(float v, float* red) {
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
  return s;  // valid in thread 0
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---- forward: noise -> im_in, partial sums of (im_s - im_in)^2; the LAST block to finish finalises: per-image loss_i,
// branch, compaction, LR schedule, Adam bias corrections (and the condition of the loop's graph IF node).  Until round 2
// session 3 the finalisation was its own one-block launch (and the condition a third): three dependent launches ahead of
// every iteration, which is what a budget-branch iteration of a small batch consists of.  The per-image sums run over
// the blocks' partials in fixed order, so the result does not depend on which block comes last.
struct FinalizeArgs {
  int n_img, force_branch, sched_period, ge_test;
  double inv_per_img, lr0, lr_gamma, beta1, beta2;
  float budget;
};

__device__ __forceinline__ void perturb_finalize(const float* ws, const icadv_perturb_state& st, const FinalizeArgs& fa,
                                                 int* s_branch, float* s_loss) {
  for (int n = threadIdx.x; n < fa.n_img; n += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < kRedBlocks; ++b) s += __ldcg(ws + (int64_t)n * kRedBlocks + b);   // written by other blocks
    const float loss_i = (float)((double)s * fa.inv_per_img);
    st.sum_d2[n] = s;
    st.loss_i[n] = loss_i;
    s_loss[n] = loss_i;
  }
  __syncthreads();
  // budget test on the BATCH mean (ge_test bit 1): the reference's attack_our takes torch.mean over the whole batch
  // (attack_rd.py:333-334), which is what train.py:342 runs on a training batch -- one shared branch per iteration
  const bool batch_budget = (fa.ge_test & 2) != 0;
  if (batch_budget) {
    if (threadIdx.x == 0) {
      double acc = 0.0;
      for (int k = 0; k < fa.n_img; ++k) acc += (double)s_loss[k];
      s_loss[1024] = (float)(acc / (double)fa.n_img);
    }
    __syncthreads();
  }
  for (int n = threadIdx.x; n < fa.n_img; n += blockDim.x) {
    const float loss_i = batch_budget ? s_loss[1024] : s_loss[n];
    // attack_rd.py:334 -- A when over budget (">"); the ROI variant switches on ">=" (attack_data.py:219)
    int br = ((fa.ge_test & 1) ? (loss_i >= fa.budget) : (loss_i > fa.budget)) ? 0 : 1;
    if (fa.force_branch >= 0) br = fa.force_branch;
    st.branch[n] = br;
    s_branch[n] = br;
    const int i = st.step[n];  // 0-based iteration index
    // MultiStepLR([1,2,3], gamma) stepped when i % period == 0 (attack_rd.py:503,553-554)
    int nsched = (i == 0) ? 0 : 1 + (i - 1) / fa.sched_period;
    if (nsched > 3) nsched = 3;
    double lr = fa.lr0;
    for (int k = 0; k < nsched; ++k) lr *= fa.lr_gamma;
    const int t = i + 1;
    const double bc1 = 1.0 - pow(fa.beta1, (double)t), bc2 = 1.0 - pow(fa.beta2, (double)t);
    st.lr[n] = (float)lr;
    st.step_size[n] = (float)(lr / bc1);
    st.bc2_sqrt[n] = (float)sqrt(bc2);
    st.step[n] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int k = 0; k < fa.n_img; ++k)
      if (s_branch[k]) st.active[cnt++] = k;
    *st.n_active = cnt;
    for (int k = cnt; k < fa.n_img; ++k) st.active[k] = 0;
    if (st.cond_handle != 0ull) cudaGraphSetConditional((cudaGraphConditionalHandle)st.cond_handle, cnt > 0 ? 1u : 0u);
    *st.counter = 0u;           // ready for the next launch (launches on one state are stream-ordered)
  }
}

__global__ void __launch_bounds__(256) perturb_forward_kernel(const float4* __restrict__ im_s,
                                                              const float4* __restrict__ noise,
                                                              float4* __restrict__ im_in, float* __restrict__ ws,
                                                              int64_t per_img4, float eps,
                                                              const float4* __restrict__ w_in, icadv_perturb_state st,
                                                              FinalizeArgs fa) {
  __shared__ float red[8];
  __shared__ int s_branch[1024];
  __shared__ float s_loss[1025];
  __shared__ int s_last;
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img4;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 s = __ldg(im_s + base + i), z = __ldg(noise + base + i);
    // ROI attack (attack_cv.py:149-163): per-pixel weight mask_tar + lamb_bkg_in * mask_bkg, shared by all images
    const float4 w = w_in != nullptr ? __ldg(w_in + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 o;
    o.x = clampf(s.x + clampf(z.x, -eps, eps), 0.f, 1.f);
    o.y = clampf(s.y + clampf(z.y, -eps, eps), 0.f, 1.f);
    o.z = clampf(s.z + clampf(z.z, -eps, eps), 0.f, 1.f);
    o.w = clampf(s.w + clampf(z.w, -eps, eps), 0.f, 1.f);
    im_in[base + i] = o;
    const float dx = s.x - o.x, dy = s.y - o.y, dz = s.z - o.z, dw = s.w - o.w;
    acc += w.x * dx * dx + w.y * dy * dy + w.z * dz * dz + w.w * dw * dw;
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) {
    ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
    __threadfence();                                           // the partial is visible before the arrival is counted
    const unsigned int arrived = atomicAdd(st.counter, 1u);
    s_last = (arrived == gridDim.x * gridDim.y - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  perturb_finalize(ws, st, fa, s_branch, s_loss);
}

// ---- backward of both clamp pairs + Adam, fused
__device__ __forceinline__ float clamp_chain_grad(float g, float s, float z, float eps) {
  // im_in = Up(Low(t, 0), 1), t = s + nc ; nc = Up(Low(z, -eps), eps)      (utils/ops.py:36-41, 51-56)
  const float nc = clampf(z, -eps, eps);
  const float t = s + nc;
  const float u = fmaxf(t, 0.f);
  g = ((u <= 1.f) || (g > 0.f)) ? g : 0.f;      // Up_bound(., 1) backward
  g = ((t >= 0.f) || (g < 0.f)) ? g : 0.f;      // Low_bound(., 0) backward
  const float l = fmaxf(z, -eps);
  g = ((l <= eps) || (g > 0.f)) ? g : 0.f;      // Up_bound(., eps) backward
  g = ((z >= -eps) || (g < 0.f)) ? g : 0.f;     // Low_bound(., -eps) backward
  return g;
}

struct AdamCoef { float omb1, b2, omb2, eps; };  // (1-beta1), beta2, (1-beta2) rounded from double like torch does

__device__ __forceinline__ float adam_step(float& z, float& m, float& v, float g, const AdamCoef c, float step_size,
                                           float bc2_sqrt) {
  const float adam_eps = c.eps;
  m = m + (g - m) * c.omb1;                     // exp_avg.lerp_(grad, 1 - beta1)
  v = v * c.b2 + c.omb2 * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + adam_eps;
  z = z - step_size * (m / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
  return z;
}

__global__ void __launch_bounds__(256) perturb_update_adam_kernel(const float4* __restrict__ im_s, float4* noise,
                                                                  const float4* __restrict__ g_in,
                                                                  const float4* __restrict__ g_a_ext, float4* m4,
                                                                  float4* v4, icadv_perturb_state st,
                                                                  int64_t per_img4, float eps, AdamCoef coef,
                                                                  float gradA_scale, float gradB_scale,
                                                                  const float4* __restrict__ w_in) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img4;
  const int br = st.branch[n];
  const float step_size = st.step_size[n], bc2_sqrt = st.bc2_sqrt[n];
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 s = __ldg(im_s + base + i);
    float4 z = noise[base + i], m = m4[base + i], v = v4[base + i];
    float4 g;
    if (br) {
      const float4 gi = __ldg(g_in + base + i);
      g = make_float4(gi.x * gradB_scale, gi.y * gradB_scale, gi.z * gradB_scale, gi.w * gradB_scale);
    } else if (g_a_ext != nullptr) {   // budget branch with an externally computed gradient (1 - ms_ssim)
      const float4 gi = __ldg(g_a_ext + base + i);
      g = make_float4(gi.x * gradA_scale, gi.y * gradA_scale, gi.z * gradA_scale, gi.w * gradA_scale);
    } else {
      // loss = mean((im_s - im_in)^2): d/d im_in = 2 (im_in - im_s) / P
      g.x = 2.f * (clampf(s.x + clampf(z.x, -eps, eps), 0.f, 1.f) - s.x) * gradA_scale;
      g.y = 2.f * (clampf(s.y + clampf(z.y, -eps, eps), 0.f, 1.f) - s.y) * gradA_scale;
      g.z = 2.f * (clampf(s.z + clampf(z.z, -eps, eps), 0.f, 1.f) - s.z) * gradA_scale;
      g.w = 2.f * (clampf(s.w + clampf(z.w, -eps, eps), 0.f, 1.f) - s.w) * gradA_scale;
      if (w_in != nullptr) {   // weighted budget term of the ROI attack
        const float4 w = __ldg(w_in + i);
        g.x *= w.x; g.y *= w.y; g.z *= w.z; g.w *= w.w;
      }
    }
    g.x = clamp_chain_grad(g.x, s.x, z.x, eps);
    g.y = clamp_chain_grad(g.y, s.y, z.y, eps);
    g.z = clamp_chain_grad(g.z, s.z, z.z, eps);
    g.w = clamp_chain_grad(g.w, s.w, z.w, eps);
    adam_step(z.x, m.x, v.x, g.x, coef, step_size, bc2_sqrt);
    adam_step(z.y, m.y, v.y, g.y, coef, step_size, bc2_sqrt);
    adam_step(z.z, m.z, v.z, g.z, coef, step_size, bc2_sqrt);
    adam_step(z.w, m.w, v.w, g.w, coef, step_size, bc2_sqrt);
    noise[base + i] = z; m4[base + i] = m; v4[base + i] = v;
  }
}

// ---- I-FGSM / PGD step
__global__ void ifgsm_update_kernel(const float* __restrict__ im_s, float* im_adv, const float* __restrict__ g,
                                    int64_t n, float alpha, float eps) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gv = g[i];
    const float sg = (gv > 0.f) ? 1.f : (gv < 0.f ? -1.f : 0.f);
    float x = im_adv[i] + alpha * sg;
    const float s = im_s[i];
    x = x > s + eps ? s + eps : x;     // attack_ifgsm.py:417
    x = x < s - eps ? s - eps : x;     // attack_ifgsm.py:418
    im_adv[i] = x;
  }
}

// ---- MI-FGSM (attack_ifgsm.py:348-362): g_m = mu g_m + grad / ||grad||_1 ; x = clamp(x + alpha sign(g_m), 0, 1); project
__global__ void __launch_bounds__(256) sum_abs_kernel(const float4* __restrict__ a, float* __restrict__ ws,
                                                      int64_t per_img4) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img4;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 p = __ldg(a + base + i);
    acc += fabsf(p.x) + fabsf(p.y) + fabsf(p.z) + fabsf(p.w);
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
}

__global__ void mifgsm_update_kernel(const float* __restrict__ im_s, float* im_adv, const float* __restrict__ g,
                                     float* gm, const float* __restrict__ l1, int64_t per_img, float alpha, float eps,
                                     float mu) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img;
  const float inv = 1.f / l1[n];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_img; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = mu * gm[base + i] + g[base + i] * inv;
    gm[base + i] = m;
    const float sg = (m > 0.f) ? 1.f : (m < 0.f ? -1.f : 0.f);
    float x = clampf(im_adv[base + i] + alpha * sg, 0.f, 1.f);
    const float s = im_s[base + i];
    x = x > s + eps ? s + eps : x;
    x = x < s - eps ? s - eps : x;
    im_adv[base + i] = x;
  }
}

// ---- C&W loss (attack_cw.py:111-140): loss = loss_i + c (1 - MSE_o), with c = 0 for an image whose reconstruction
//      error already exceeds 1.1 x its target level.  Gradient wrt im_in = 2 (im_in - im_s) / P + c_n g_net, where g_net
//      is the network-branch gradient of (1 - MSE_o) the stack backward produced.
__global__ void __launch_bounds__(256) cw_combine_kernel(const float4* __restrict__ g_net, const float4* __restrict__ im_in,
                                                         const float4* __restrict__ im_s, float4* __restrict__ g_out,
                                                         const float* __restrict__ c, const float* __restrict__ level,
                                                         const float* __restrict__ sum_d2, int64_t per_img4,
                                                         float inv_per_img) {
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img4;
  const float mse_o = sum_d2[n] * inv_per_img;
  const float cn = (mse_o > level[n] * 1.1f) ? 0.f : c[n];          // attack_cw.py:138-139
  const float a = 2.f * inv_per_img;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 g = __ldg(g_net + base + i), x = __ldg(im_in + base + i), s = __ldg(im_s + base + i);
    g_out[base + i] = make_float4(a * (x.x - s.x) + cn * g.x, a * (x.y - s.y) + cn * g.y, a * (x.z - s.z) + cn * g.z,
                                  a * (x.w - s.w) + cn * g.w);
  }
}

// ---- output clamp + distortion + gradient seed (attack_rd.py:353-364)
__global__ void __launch_bounds__(256) output_loss_kernel(const float4* __restrict__ x, const float4* __restrict__ ref,
                                                          float4* __restrict__ g_x, float* __restrict__ ws,
                                                          int64_t per_img4, int do_clamp, float grad_scale,
                                                          const int* __restrict__ active,
                                                          const int* __restrict__ n_active,
                                                          const float4* __restrict__ w_out) {
  __shared__ float red[8];
  const int slot = blockIdx.y;
  if (n_active != nullptr && slot >= *n_active) return;
  const int n = active != nullptr ? active[slot] : slot;
  const int64_t base = (int64_t)n * per_img4;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 xv = __ldg(x + base + i), rv = __ldg(ref + base + i);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, rs[4] = {rv.x, rv.y, rv.z, rv.w};
    const float4 wv = w_out != nullptr ? __ldg(w_out + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float wk[4] = {wv.x, wv.y, wv.z, wv.w};
    float gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float o = do_clamp ? clampf(xs[k], 0.f, 1.f) : xs[k];
      const float d = rs[k] - o;
      acc += wk[k] * d * d;
      float g = 2.f * wk[k] * d * grad_scale;   // d(1 - mean(w d^2)) / d o  (grad_scale < 0: d(+mean(w d^2)) / d o)
      if (do_clamp) {
        const float l = fmaxf(xs[k], 0.f);
        g = ((l <= 1.f) || (g > 0.f)) ? g : 0.f;
        g = ((xs[k] >= 0.f) || (g < 0.f)) ? g : 0.f;
      }
      gs[k] = g;
    }
    if (g_x != nullptr) g_x[base + i] = make_float4(gs[0], gs[1], gs[2], gs[3]);
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
}

__global__ void sum_finalize_kernel(const float* __restrict__ ws, float* __restrict__ out, int n_img,
                                    const int* __restrict__ active, const int* __restrict__ n_active) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_img) return;
  if (n_active != nullptr && slot >= *n_active) return;
  const int n = active != nullptr ? active[slot] : slot;
  float s = 0.f;
  for (int b = 0; b < kRedBlocks; ++b) s += ws[(int64_t)n * kRedBlocks + b];
  out[n] = s;
}

__global__ void __launch_bounds__(256) sum_sqdiff_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                                         float* __restrict__ ws, int64_t per_img4) {
  __shared__ float red[8];
  const int n = blockIdx.y;
  const int64_t base = (int64_t)n * per_img4;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < per_img4; i += (int64_t)gridDim.x * 256) {
    const float4 p = __ldg(a + base + i), q = __ldg(b + base + i);
    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z, dw = p.w - q.w;
    acc += dx * dx + dy * dy + dz * dz + dw * dw;
  }
  const float s = block_sum_256(acc, red);
  if (threadIdx.x == 0) ws[(int64_t)n * kRedBlocks + blockIdx.x] = s;
}

__global__ void bound_forward_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float bound,
                                     int upper) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = upper ? fminf(x[i], bound) : fmaxf(x[i], bound);
}
__global__ void bound_backward_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                      float* __restrict__ gx, int64_t n, float bound, int upper) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = gy[i], xv = x[i];
    const bool keep = upper ? ((xv <= bound) || (g > 0.f)) : ((xv >= bound) || (g < 0.f));
    gx[i] = keep ? g : 0.f;
  }
}

static inline int ew_blocks(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_perturb_forward(const float* im_s, const float* noise, float* im_in, float* ws,
                          const icadv_perturb_state* st, int n_img, int64_t per_img, float eps, float noise_budget,
                          int force_branch, double lr0, double lr_gamma, int sched_period, double beta1,
                          double beta2, icadv_stream_t stream) {
  return icadv_perturb_forward_roi(im_s, noise, im_in, ws, st, n_img, per_img, eps, noise_budget, force_branch, lr0,
                                   lr_gamma, sched_period, beta1, beta2, nullptr, 0, stream);
}

int icadv_perturb_forward_roi(const float* im_s, const float* noise, float* im_in, float* ws,
                              const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                              float noise_budget, int force_branch, double lr0, double lr_gamma, int sched_period,
                              double beta1, double beta2, const float* w_in, int ge_test, icadv_stream_t stream) {
  ICADV_REQUIRE(im_s && noise && im_in && ws && st, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0, "per_img must be a multiple of 4");
  ICADV_REQUIRE(n_img >= 1 && n_img <= 1024, "n_img must be in [1,1024]");
  ICADV_REQUIRE(sched_period >= 1, "sched_period (steps//3) must be >= 1 (ZeroDivisionError in the reference)");
  ICADV_REQUIRE(st->counter != nullptr, "perturb state without an arrival counter");
  dim3 grid(kRedBlocks, n_img);
  FinalizeArgs fa;
  fa.n_img = n_img; fa.force_branch = force_branch; fa.sched_period = sched_period; fa.ge_test = ge_test;
  fa.inv_per_img = 1.0 / (double)per_img; fa.lr0 = lr0; fa.lr_gamma = lr_gamma; fa.beta1 = beta1; fa.beta2 = beta2;
  fa.budget = noise_budget;
  perturb_forward_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(im_s), reinterpret_cast<const float4*>(noise),
      reinterpret_cast<float4*>(im_in), ws, per_img / 4, eps, reinterpret_cast<const float4*>(w_in), *st, fa);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_perturb_update_adam(const float* im_s, float* noise, const float* g_in, const float* g_a_ext, float* m,
                              float* v,
                              const icadv_perturb_state* st, int n_img, int64_t per_img, float eps, double beta1,
                              double beta2, double adam_eps, float gradA_scale, float gradB_scale,
                              icadv_stream_t stream) {
  return icadv_perturb_update_adam_roi(im_s, noise, g_in, g_a_ext, m, v, st, n_img, per_img, eps, beta1, beta2, adam_eps,
                                       gradA_scale, gradB_scale, nullptr, stream);
}

int icadv_perturb_update_adam_roi(const float* im_s, float* noise, const float* g_in, const float* g_a_ext, float* m,
                                  float* v, const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                                  double beta1, double beta2, double adam_eps, float gradA_scale, float gradB_scale,
                                  const float* w_in, icadv_stream_t stream) {
  ICADV_REQUIRE(im_s && noise && m && v && st, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0, "per_img must be a multiple of 4");
  dim3 grid(kRedBlocks, n_img);
  AdamCoef coef;
  coef.omb1 = (float)(1.0 - beta1); coef.b2 = (float)beta2; coef.omb2 = (float)(1.0 - beta2); coef.eps = (float)adam_eps;
  perturb_update_adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(im_s), reinterpret_cast<float4*>(noise),
      reinterpret_cast<const float4*>(g_in), reinterpret_cast<const float4*>(g_a_ext), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), *st,
      per_img / 4, eps, coef, gradA_scale, gradB_scale, reinterpret_cast<const float4*>(w_in));
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_ifgsm_update(const float* im_s, float* im_adv, const float* g, int64_t n, float alpha, float eps,
                       icadv_stream_t stream) {
  ICADV_REQUIRE(im_s && im_adv && g, "null pointer");
  ifgsm_update_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(im_s, im_adv, g, n, alpha, eps);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_mifgsm_update(const float* im_s, float* im_adv, const float* g, float* g_mom, float* ws, float* l1,
                        int n_img, int64_t per_img, float alpha, float eps, float mu, icadv_stream_t stream) {
  ICADV_REQUIRE(im_s && im_adv && g && g_mom && ws && l1, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0 && n_img >= 1 && n_img <= 65535, "bad sizes");
  dim3 grid(kRedBlocks, n_img);
  sum_abs_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(g), ws, per_img / 4);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_finalize_kernel<<<(n_img + 127) / 128, 128, 0, as_stream(stream)>>>(ws, l1, n_img, nullptr, nullptr);
  ICADV_CUDA_TRY(cudaGetLastError());
  mifgsm_update_kernel<<<grid, 256, 0, as_stream(stream)>>>(im_s, im_adv, g, g_mom, l1, per_img, alpha, eps, mu);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_cw_combine(const float* g_net, const float* im_in, const float* im_s, float* g_out, const float* c,
                     const float* level, const float* sum_d2, int n_img, int64_t per_img, icadv_stream_t stream) {
  ICADV_REQUIRE(g_net && im_in && im_s && g_out && c && level && sum_d2, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0 && n_img >= 1 && n_img <= 65535, "bad sizes");
  dim3 grid(kRedBlocks, n_img);
  cw_combine_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(g_net),
                                                        reinterpret_cast<const float4*>(im_in),
                                                        reinterpret_cast<const float4*>(im_s),
                                                        reinterpret_cast<float4*>(g_out), c, level, sum_d2, per_img / 4,
                                                        1.f / (float)per_img);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_output_loss(const float* x, const float* ref, float* g_x, float* ws, float* sum_d2, int n_img,
                      int64_t per_img, int do_clamp, float grad_scale, const int* active, const int* n_active,
                      icadv_stream_t stream) {
  return icadv_output_loss_roi(x, ref, g_x, ws, sum_d2, n_img, per_img, do_clamp, grad_scale, active, n_active, nullptr,
                               stream);
}

int icadv_output_loss_roi(const float* x, const float* ref, float* g_x, float* ws, float* sum_d2, int n_img,
                          int64_t per_img, int do_clamp, float grad_scale, const int* active, const int* n_active,
                          const float* w_out, icadv_stream_t stream) {
  ICADV_REQUIRE(x && ref && ws && sum_d2, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0, "per_img must be a multiple of 4");
  dim3 grid(kRedBlocks, n_img);
  output_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x),
                                                          reinterpret_cast<const float4*>(ref),
                                                          reinterpret_cast<float4*>(g_x), ws, per_img / 4, do_clamp,
                                                          grad_scale, active, n_active,
                                                          reinterpret_cast<const float4*>(w_out));
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_finalize_kernel<<<(n_img + 127) / 128, 128, 0, as_stream(stream)>>>(ws, sum_d2, n_img, active, n_active);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_sum_sqdiff(const float* a, const float* b, float* ws, float* out, int n_img, int64_t per_img,
                     icadv_stream_t stream) {
  ICADV_REQUIRE(a && b && ws && out, "null pointer");
  ICADV_REQUIRE(per_img % 4 == 0, "per_img must be a multiple of 4");
  dim3 grid(kRedBlocks, n_img);
  sum_sqdiff_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                         reinterpret_cast<const float4*>(b), ws, per_img / 4);
  ICADV_CUDA_TRY(cudaGetLastError());
  sum_finalize_kernel<<<(n_img + 127) / 128, 128, 0, as_stream(stream)>>>(ws, out, n_img, nullptr, nullptr);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_bound_forward(const float* x, float* y, int64_t n, float bound, int upper, icadv_stream_t stream) {
  ICADV_REQUIRE(x && y, "null pointer");
  bound_forward_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, y, n, bound, upper);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_bound_backward(const float* x, const float* gy, float* gx, int64_t n, float bound, int upper,
                         icadv_stream_t stream) {
  ICADV_REQUIRE(x && gy && gx, "null pointer");
  bound_backward_kernel<<<ew_blocks(n), 256, 0, as_stream(stream)>>>(x, gy, gx, n, bound, upper);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
