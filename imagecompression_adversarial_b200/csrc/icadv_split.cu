// 3xTF32 parity mode of the tensor path, and the in-library Philox noise of the training quantiser.
//
// Parity mode ("3xtf32"): every contraction of the codec stacks still runs on conv_tc_kernel (tcgen05 kind::tf32,
// fp32 accumulation in TMEM), but each fp32 operand is split into two TF32 terms
//     x = hi + lo,  hi = rn_tf32(x),  lo = rn_tf32(x - hi)          (|lo| <= 2^-11 |x|)
// and the product is formed from three tensor-core terms  hi*Whi + lo*Whi + hi*Wlo  (the dropped lo*Wlo term is
// 2^-22 relative).  The split is a DATA LAYOUT: activations are expanded along the channel (K) axis to
// [hi | lo | hi | 0-pad], weights to [Whi | Whi | Wlo | 0-pad], so a contraction with K input channels becomes the
// same kernel launched with K' = roundup(3K, 32) channels -- three MMAs per original K-step accumulate in one TMEM
// accumulator.  GDN / IGDN are unfused in this mode (square / backward operand -> split -> 1x1 contraction with
// gamma -> elementwise apply) so that the normalisation GEMM gets the same treatment.
// Reference semantics: nn.Conv2d / nn.ConvTranspose2d / compressai GDN in fp32 (anchors/utils.py:112-130,
// utils/ops.py:58-97); this mode exists so that per-step losses of attack_rd.py:332-379,506-560 match the fp32
// reference to the 1e-3 the north star states.
#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

__device__ __forceinline__ float split_pick(float v, int seg, int layout) {
  const float hi = round_tf32(v);
  if (seg >= 3) return 0.f;
  const bool want_lo = layout == 0 ? seg == 1 : seg == 2;
  return want_lo ? round_tf32(v - hi) : hi;
}

// out[G][px][Ks] holding split channel kk = g * Ks + k of [hi | lo | hi | 0] (layout 0, activation) or
// [hi | hi | lo | 0] (layout 1, weight); op 1: v = x^2
__global__ void split3_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n_px, int C, int Ks, int G,
                              int op, int layout) {
  const int64_t total = n_px * Ks * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kl = (int)(i % Ks);
    const int64_t px = (i / Ks) % n_px;
    const int k = (int)(i / ((int64_t)Ks * n_px)) * Ks + kl;
    const int seg = k / C, c = k - seg * C;
    float v = seg < 3 ? x[px * C + c] : 0.f;
    if (op == 1) v = v * v;
    out[i] = split_pick(v, seg, layout);
  }
}

// operand of the normalisation GEMM of the GDN / IGDN backward, split: t = g y sc^2 (GDN) or g y / sc^2 (IGDN)
__global__ void gdn_bwd_operand_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                       const float* __restrict__ sc, float* __restrict__ out, int64_t n_px, int C, int Ks,
                                       int G, int inverse) {
  const int64_t total = n_px * Ks * G;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kl = (int)(i % Ks);
    const int64_t px = (i / Ks) % n_px;
    const int k = (int)(i / ((int64_t)Ks * n_px)) * Ks + kl;
    const int seg = k / C, c = k - seg * C;
    float t = 0.f;
    if (seg < 3) {
      const int64_t j = px * C + c;
      const float s = sc[j], s2 = s * s;
      t = inverse ? (s2 > 0.f ? g[j] * y[j] / s2 : 0.f) : g[j] * y[j] * s2;
    }
    out[i] = split_pick(t, seg, 0);
  }
}

// y = x * sc, sc = (norm)^(-1/2) (GDN) or (norm)^(+1/2) (IGDN); norm = beta + gamma x^2 from the 1x1 contraction
__global__ void gdn_apply_kernel(const float* __restrict__ x, const float* __restrict__ nrm, float* __restrict__ y,
                                 float* __restrict__ sc, int64_t n, int inverse) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float r = sqrtf(nrm[i]);
    const float s = inverse ? r : 1.0f / r;
    sc[i] = s;
    y[i] = x[i] * s;
  }
}

// out = g sc -+ (y / sc) w   (w = gamma^T t from the 1x1 contraction)
__global__ void gdn_bwd_combine_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                       const float* __restrict__ sc, const float* __restrict__ w, float* __restrict__ out,
                                       int64_t n, int inverse) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = sc[i];
    const float xs = s > 0.f ? y[i] / s : 0.f;
    out[i] = inverse ? g[i] * s + xs * w[i] : g[i] * s - xs * w[i];
  }
}

// out = sum over the G partial outputs of the K-slices, added in order in round-to-nearest fp32
__global__ void sum_slices_kernel(const float* __restrict__ parts, float* __restrict__ out, int64_t n, int G) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = parts[i];
    for (int g = 1; g < G; ++g) acc += parts[(int64_t)g * n + i];
    out[i] = acc;
  }
}

// ---- Philox4x32-10 (Salmon et al. 2011), one 128-bit block per four outputs
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void uniform_noise_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset, float lo,
                                     float width) {
  const int64_t blocks4 = (n + 3) / 4;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < blocks4; b += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + (uint64_t)b;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = 4 * b + j;
      // 24 random bits -> [0, 1) exactly representable, so lo + width * u stays inside [lo, lo + width)
      if (i < n) out[i] = lo + width * ((float)(c[j] >> 8) * (1.0f / 16777216.0f));
    }
  }
}

static inline int grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_split3(const float* x, float* out, int64_t n_px, int C, int Ks, int G, int op, int layout,
                 icadv_stream_t stream) {
  ICADV_REQUIRE(x && out && n_px > 0 && C > 0 && G >= 1 && Ks % 32 == 0 && (int64_t)Ks * G >= 3 * (int64_t)C,
                "split3: need Ks %% 32 == 0 and Ks * G >= 3C");
  ICADV_REQUIRE((op == 0 || op == 1) && (layout == 0 || layout == 1), "split3: bad op / layout");
  split3_kernel<<<grid_for(n_px * Ks * G), 256, 0, as_stream(stream)>>>(x, out, n_px, C, Ks, G, op, layout);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_bwd_operand_split3(const float* g, const float* y, const float* sc, float* out, int64_t n_px, int C, int Ks,
                                 int G, int inverse, icadv_stream_t stream) {
  ICADV_REQUIRE(g && y && sc && out && n_px > 0 && C > 0 && G >= 1 && Ks % 32 == 0 && (int64_t)Ks * G >= 3 * (int64_t)C,
                "gdn_bwd_operand: bad args");
  gdn_bwd_operand_kernel<<<grid_for(n_px * Ks * G), 256, 0, as_stream(stream)>>>(g, y, sc, out, n_px, C, Ks, G, inverse);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_sum_slices(const float* parts, float* out, int64_t n, int G, icadv_stream_t stream) {
  ICADV_REQUIRE(parts && out && n > 0 && G >= 1, "sum_slices: bad args");
  sum_slices_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(parts, out, n, G);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_apply(const float* x, const float* nrm, float* y, float* sc, int64_t n, int inverse,
                    icadv_stream_t stream) {
  ICADV_REQUIRE(x && nrm && y && sc && n > 0, "gdn_apply: bad args");
  gdn_apply_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(x, nrm, y, sc, n, inverse);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_gdn_bwd_combine(const float* g, const float* y, const float* sc, const float* w, float* out, int64_t n,
                          int inverse, icadv_stream_t stream) {
  ICADV_REQUIRE(g && y && sc && w && out && n > 0, "gdn_bwd_combine: bad args");
  gdn_bwd_combine_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(g, y, sc, w, out, n, inverse);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_uniform_noise(float* out, int64_t n, uint64_t seed, uint64_t offset, float lo, float hi,
                        icadv_stream_t stream) {
  ICADV_REQUIRE(out && n > 0 && hi > lo, "uniform_noise: bad args");
  uniform_noise_kernel<<<grid_for((n + 3) / 4), 256, 0, as_stream(stream)>>>(out, n, seed, offset, lo, hi - lo);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
