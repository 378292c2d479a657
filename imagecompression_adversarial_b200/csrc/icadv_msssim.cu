// MS-SSIM level kernel (bandwidth bound): one pass over an image pair produces the per-plane sums of the
// SSIM and contrast-structure maps.  The five Gaussian-filtered maps (X, Y, X^2, Y^2, XY) never leave the
// SM: a 32x32 output tile + halo is staged in shared memory, filtered separably, and reduced in fixed order.
//   variant 1 (valid, separable 11-tap sigma 1.5): pytorch_msssim.ms_ssim  (attack_rd.py:336,362;
//              self_ensemble.py:225,228; train.py:44,88)
//   variant 2 (zero "same" padding, 2-D = outer-product window): utils/torch_msssim.py:26-52
// plus the 2x2 average pooling between levels (padding (H%2, W%2) for variant 1; none for variant 2).
#include "icadv_common.cuh"

namespace icadv {

constexpr int kT = 32;          // output tile edge
constexpr int kMaxWin = 11;
constexpr int kIn = kT + kMaxWin - 1;  // 42

struct SsimParams {
  const float* X; const float* Y; float* ws;
  int h, w, out_h, out_w, win, pad;  // pad = 0 (valid) or win/2 (same, zero fill)
  int tiles_x, tiles_y;
  float c1, c2;
  float taps[kMaxWin];
};

__global__ void __launch_bounds__(256) ssim_level_kernel(const SsimParams p) {
  __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
  __shared__ float hm[5][kIn][kT + 1];
  __shared__ float red[2][8];
  const int plane = blockIdx.z;
  const int ty0 = blockIdx.y * kT, tx0 = blockIdx.x * kT;
  const float* X = p.X + (int64_t)plane * p.h * p.w;
  const float* Y = p.Y + (int64_t)plane * p.h * p.w;
  const int span = kT + p.win - 1;
  for (int i = threadIdx.x; i < span * span; i += 256) {
    const int r = i / span, c = i % span;
    const int gy = ty0 + r - p.pad, gx = tx0 + c - p.pad;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) { a = __ldg(X + (int64_t)gy * p.w + gx); b = __ldg(Y + (int64_t)gy * p.w + gx); }
    sx[r][c] = a; sy[r][c] = b;
  }
  __syncthreads();
  // horizontal pass
  for (int i = threadIdx.x; i < span * kT; i += 256) {
    const int r = i / kT, c = i % kT;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float g = p.taps[k], a = sx[r][c + k], b = sy[r][c + k];
      m1 = fmaf(g, a, m1); m2 = fmaf(g, b, m2);
      s11 = fmaf(g, a * a, s11); s22 = fmaf(g, b * b, s22); s12 = fmaf(g, a * b, s12);
    }
    hm[0][r][c] = m1; hm[1][r][c] = m2; hm[2][r][c] = s11; hm[3][r][c] = s22; hm[4][r][c] = s12;
  }
  __syncthreads();
  // vertical pass + maps
  float acc_s = 0.f, acc_c = 0.f;
  for (int i = threadIdx.x; i < kT * kT; i += 256) {
    const int r = i / kT, c = i % kT;
    if (ty0 + r >= p.out_h || tx0 + c >= p.out_w) continue;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float g = p.taps[k];
      m1 = fmaf(g, hm[0][r + k][c], m1); m2 = fmaf(g, hm[1][r + k][c], m2);
      s11 = fmaf(g, hm[2][r + k][c], s11); s22 = fmaf(g, hm[3][r + k][c], s22); s12 = fmaf(g, hm[4][r + k][c], s12);
    }
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
    const float cs = (2.f * v12 + p.c2) / (v1 + v2 + p.c2);
    const float ss = ((2.f * m12 + p.c1) / (m11 + m22 + p.c1)) * cs;
    acc_s += ss; acc_c += cs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = acc_c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int k = 0; k < 8; ++k) { s += red[0][k]; c += red[1][k]; }
    const int nb = p.tiles_x * p.tiles_y, b = blockIdx.y * p.tiles_x + blockIdx.x;
    p.ws[((int64_t)plane * nb + b) * 2 + 0] = s;
    p.ws[((int64_t)plane * nb + b) * 2 + 1] = c;
  }
}


// ---- 11-tap fast path (every level of a >= 176-pixel image): register-blocked separable filter.  The generic kernel
// above issues one shared-memory load per multiply-add (127 LDS per output pixel: measured 4.5 % DRAM, shared-memory
// bound, profiles/r2_ncu_aux_kernels.txt); here a thread produces FOUR consecutive outputs of a filter line from 14
// loaded values (3.1x fewer loads), conflict-free in both passes (row strides 43 / 33 floats).
constexpr int kRun = 4;

__global__ void __launch_bounds__(256) ssim_level_kernel11(const SsimParams p) {
  constexpr int W = kMaxWin, span = kT + W - 1;   // 11, 42
  __shared__ float sx[span][span + 1], sy[span][span + 1];
  __shared__ float hm[5][span][kT + 1];
  __shared__ float red[2][8];
  const int plane = blockIdx.z;
  const int ty0 = blockIdx.y * kT, tx0 = blockIdx.x * kT;
  const float* X = p.X + (int64_t)plane * p.h * p.w;
  const float* Y = p.Y + (int64_t)plane * p.h * p.w;
  for (int i = threadIdx.x; i < span * span; i += 256) {
    const int r = i / span, c = i % span;
    const int gy = ty0 + r - p.pad, gx = tx0 + c - p.pad;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) { a = __ldg(X + (int64_t)gy * p.w + gx); b = __ldg(Y + (int64_t)gy * p.w + gx); }
    sx[r][c] = a; sy[r][c] = b;
  }
  float taps[W];
#pragma unroll
  for (int k = 0; k < W; ++k) taps[k] = p.taps[k];
  __syncthreads();
  // horizontal pass: item = (row r, run of 4 columns)
  for (int i = threadIdx.x; i < span * (kT / kRun); i += 256) {
    const int r = i / (kT / kRun), c0 = (i % (kT / kRun)) * kRun;
    float a[kRun + W - 1], b[kRun + W - 1];
#pragma unroll
    for (int k = 0; k < kRun + W - 1; ++k) { a[k] = sx[r][c0 + k]; b[k] = sy[r][c0 + k]; }
#pragma unroll
    for (int o = 0; o < kRun; ++o) {
      float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float g = taps[k], u = a[o + k], v = b[o + k];
        m1 = fmaf(g, u, m1); m2 = fmaf(g, v, m2);
        s11 = fmaf(g, u * u, s11); s22 = fmaf(g, v * v, s22); s12 = fmaf(g, u * v, s12);
      }
      hm[0][r][c0 + o] = m1; hm[1][r][c0 + o] = m2; hm[2][r][c0 + o] = s11; hm[3][r][c0 + o] = s22; hm[4][r][c0 + o] = s12;
    }
  }
  __syncthreads();
  // vertical pass + maps: item = (column c, run of 4 rows); 32 x 8 items = one per thread
  float acc_s = 0.f, acc_c = 0.f;
  {
    const int c = threadIdx.x % kT, r0 = (threadIdx.x / kT) * kRun;
    float o5[5][kRun];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      float v[kRun + W - 1];
#pragma unroll
      for (int k = 0; k < kRun + W - 1; ++k) v[k] = hm[m][r0 + k][c];
#pragma unroll
      for (int o = 0; o < kRun; ++o) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) t = fmaf(taps[k], v[o + k], t);
        o5[m][o] = t;
      }
    }
#pragma unroll
    for (int o = 0; o < kRun; ++o) {
      if (ty0 + r0 + o >= p.out_h || tx0 + c >= p.out_w) continue;
      const float m1 = o5[0][o], m2 = o5[1][o];
      const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
      const float v1 = o5[2][o] - m11, v2 = o5[3][o] - m22, v12 = o5[4][o] - m12;
      const float cs = (2.f * v12 + p.c2) / (v1 + v2 + p.c2);
      const float ss = ((2.f * m12 + p.c1) / (m11 + m22 + p.c1)) * cs;
      acc_s += ss; acc_c += cs;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = acc_c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int k = 0; k < 8; ++k) { s += red[0][k]; c += red[1][k]; }
    const int nb = p.tiles_x * p.tiles_y, b = blockIdx.y * p.tiles_x + blockIdx.x;
    p.ws[((int64_t)plane * nb + b) * 2 + 0] = s;
    p.ws[((int64_t)plane * nb + b) * 2 + 1] = c;
  }
}

__global__ void ssim_finalize_kernel(const float* __restrict__ ws, float* __restrict__ ssim_sum,
                                     float* __restrict__ cs_sum, int planes, int nb) {
  const int pl = blockIdx.x * blockDim.x + threadIdx.x;
  if (pl >= planes) return;
  float s = 0.f, c = 0.f;
  for (int b = 0; b < nb; ++b) { s += ws[((int64_t)pl * nb + b) * 2]; c += ws[((int64_t)pl * nb + b) * 2 + 1]; }
  ssim_sum[pl] = s; cs_sum[pl] = c;
}

// even width divisible by 4, no padding: a thread averages two float4 row pieces into one float2 (two outputs)
__global__ void __launch_bounds__(256) avgpool2_vec_kernel(const float* __restrict__ x, float* __restrict__ y, int h, int w,
                                                           int oh, int ow2) {
  const int plane = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= oh * ow2) return;
  const int oy = i / ow2, o2 = i - oy * ow2;
  const float4* r0 = reinterpret_cast<const float4*>(x + ((int64_t)plane * h + 2 * oy) * w);
  const float4 a = __ldg(r0 + o2), b = __ldg(r0 + (w >> 2) + o2);
  // same association as the scalar kernel: ((x00 + x01) + x10) + x11, then * 0.25
  reinterpret_cast<float2*>(y + ((int64_t)plane * oh + oy) * (2 * ow2))[o2] =
      make_float2(0.25f * (((a.x + a.y) + b.x) + b.y), 0.25f * (((a.z + a.w) + b.z) + b.w));
}

// F.avg_pool2d(x, kernel_size=2, stride=2, padding=(ph, pw)), count_include_pad=True
__global__ void avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int h, int w, int oh,
                                int ow, int ph, int pw) {
  const int64_t total = (int64_t)planes * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
    const int64_t pl = i / ((int64_t)ow * oh);
    const float* src = x + pl * h * w;
    float s = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int yy = 2 * oy - ph + dy, xx = 2 * ox - pw + dx;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) s += __ldg(src + (int64_t)yy * w + xx);
      }
    y[i] = 0.25f * s;
  }
}


// ---------------------------------------------------------------------------------------------------------
// Backward of one MS-SSIM level with respect to X: variant 1 (valid separable window) and variant 2 (zero "same"
// padding, utils/torch_msssim.py:26-52 -- its 2-D window is the outer product of the 1-D taps, so it is separable too).
//   value_l[plane] = mean_px map(X, Y),  map = cs (levels 0..3) or ssim (last level)
//   dX = coef_cs[plane] * d(sum cs)/dX + coef_ss[plane] * d(sum ssim)/dX + 0.25 * dXnext[pool parent]
// One block produces a 32x32 tile of dX.  The filtered statistics are RECOMPUTED from a 52x52 input tile (no
// saved maps): forward separable filter on 5 maps -> A12, A11, B on a 42x42 tile -> transposed separable filter.
//   cs   = (2 s12 + C2) / (s11 + s22 + C2),   ssim = lum * cs,   lum = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1)
//   dX = Y * Gt[A12] + 2 X * Gt[A11] + Gt[B],   A12 = a 2/D,  A11 = -a cs/D,  B = -mu2 A12 - 2 mu1 A11 (+ lum term)
// ---------------------------------------------------------------------------------------------------------
constexpr int kBT = 32;                 // dX tile edge
constexpr int kBA = kBT + kMaxWin - 1;  // 42: tile of the statistics maps
constexpr int kBX = kBA + kMaxWin - 1;  // 52: input tile

struct SsimBwdParams {
  const float* X; const float* Y; const float* coef_cs; const float* coef_ss; const float* dXnext; float* dX;
  int h, w, oh, ow, win, nh, nw, ph, pw;
  int so;   // statistics-map position of tile row 0 relative to the dX tile: win-1 (valid window) or win/2 (zero "same" padding)
  float c1, c2;
  float taps[kMaxWin];
};

__global__ void __launch_bounds__(256) ssim_level_bwd_kernel(const SsimBwdParams p) {
  extern __shared__ float sm[];
  float (*sx)[kBX + 1] = reinterpret_cast<float (*)[kBX + 1]>(sm);                       // [52][53]
  float (*sy)[kBX + 1] = reinterpret_cast<float (*)[kBX + 1]>(sm + kBX * (kBX + 1));
  float* hbase = sm + 2 * kBX * (kBX + 1);                                               // h5[5][52][43]
  auto h5 = [&](int m, int r, int c) -> float& { return hbase[(m * kBX + r) * (kBA + 1) + c]; };
  float* abase = hbase + 5 * kBX * (kBA + 1);                                            // a3[3][42][43]
  auto a3 = [&](int m, int r, int c) -> float& { return abase[(m * kBA + r) * (kBA + 1) + c]; };
  float* tbase = abase + 3 * kBA * (kBA + 1);                                            // ha[3][42][33]
  auto ha = [&](int m, int r, int c) -> float& { return tbase[(m * kBA + r) * (kBT + 1) + c]; };

  const int plane = blockIdx.z;
  const int ty0 = blockIdx.y * kBT, tx0 = blockIdx.x * kBT;
  const int R = p.win - 1;
  const float* X = p.X + (int64_t)plane * p.h * p.w;
  const float* Y = p.Y + (int64_t)plane * p.h * p.w;
  const float ccs = p.coef_cs[plane], css = p.coef_ss[plane];
  const int xs = kBT + 2 * R, as = kBT + R;   // used extents (52 / 42 for win = 11)
  for (int i = threadIdx.x; i < xs * xs; i += 256) {
    const int r = i / xs, c = i % xs;
    const int gy = ty0 - R + r, gx = tx0 - R + c;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) { a = __ldg(X + (int64_t)gy * p.w + gx); b = __ldg(Y + (int64_t)gy * p.w + gx); }
    sx[r][c] = a; sy[r][c] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < xs * as; i += 256) {   // horizontal forward filter
    const int r = i / as, c = i % as;
    float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float g = p.taps[k], a = sx[r][c + k], b = sy[r][c + k];
      m1 = fmaf(g, a, m1); m2 = fmaf(g, b, m2);
      s11 = fmaf(g, a * a, s11); s22 = fmaf(g, b * b, s22); s12 = fmaf(g, a * b, s12);
    }
    h5(0, r, c) = m1; h5(1, r, c) = m2; h5(2, r, c) = s11; h5(3, r, c) = s22; h5(4, r, c) = s12;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < as * as; i += 256) {   // vertical forward filter + per-position coefficients
    const int r = i / as, c = i % as;
    // position in the statistics map.  Valid window: map position o reads inputs o .. o+R; "same" padding (R = 2 pad):
    // o-pad .. o+pad -- in both cases the 52x52 input tile starts R before the dX tile, only this origin differs
    const int oy = ty0 - p.so + r, ox = tx0 - p.so + c;
    float A12 = 0.f, A11 = 0.f, B = 0.f;
    if (oy >= 0 && oy < p.oh && ox >= 0 && ox < p.ow) {
      float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
      for (int k = 0; k < p.win; ++k) {
        const float g = p.taps[k];
        m1 = fmaf(g, h5(0, r + k, c), m1); m2 = fmaf(g, h5(1, r + k, c), m2);
        s11 = fmaf(g, h5(2, r + k, c), s11); s22 = fmaf(g, h5(3, r + k, c), s22); s12 = fmaf(g, h5(4, r + k, c), s12);
      }
      const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
      const float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
      const float D = v1 + v2 + p.c2, cs = (2.f * v12 + p.c2) / D;
      const float Dl = m11 + m22 + p.c1, lum = (2.f * m12 + p.c1) / Dl;
      const float a_cs = ccs + css * lum;               // upstream weight on the cs factor
      A12 = a_cs * 2.f / D;
      A11 = -a_cs * cs / D;
      B = -m2 * A12 - 2.f * m1 * A11 + css * cs * 2.f * (m2 - lum * m1) / Dl;
    }
    a3(0, r, c) = A12; a3(1, r, c) = A11; a3(2, r, c) = B;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < as * kBT; i += 256) {  // horizontal transposed filter
    const int r = i / kBT, c = i % kBT;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float g = p.taps[k];
      t0 = fmaf(g, a3(0, r, c + R - k), t0); t1 = fmaf(g, a3(1, r, c + R - k), t1); t2 = fmaf(g, a3(2, r, c + R - k), t2);
    }
    ha(0, r, c) = t0; ha(1, r, c) = t1; ha(2, r, c) = t2;
  }
  __syncthreads();
  float* out = p.dX + (int64_t)plane * p.h * p.w;
  for (int i = threadIdx.x; i < kBT * kBT; i += 256) { // vertical transposed filter + combine
    const int r = i / kBT, c = i % kBT;
    const int gy = ty0 + r, gx = tx0 + c;
    if (gy >= p.h || gx >= p.w) continue;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < p.win; ++k) {
      const float g = p.taps[k];
      t0 = fmaf(g, ha(0, r + R - k, c), t0); t1 = fmaf(g, ha(1, r + R - k, c), t1); t2 = fmaf(g, ha(2, r + R - k, c), t2);
    }
    float d = sy[r + R][c + R] * t0 + 2.f * sx[r + R][c + R] * t1 + t2;
    if (p.dXnext != nullptr) {
      const int py = (gy + p.ph) >> 1, px = (gx + p.pw) >> 1;
      if (py < p.nh && px < p.nw) d += 0.25f * __ldg(p.dXnext + ((int64_t)plane * p.nh + py) * p.nw + px);
    }
    out[(int64_t)gy * p.w + gx] = d;
  }
}


// ---- 11-tap fast path of the backward level: the same five phases, every filter pass register-blocked (a thread
// produces a run of consecutive outputs of a filter line from one set of loaded values: 16 loads for 6 outputs instead of
// 66, 14 for 4 instead of 44).
__global__ void __launch_bounds__(256) ssim_level_bwd_kernel11(const SsimBwdParams p) {
  constexpr int W = kMaxWin, R = W - 1;           // 11, 10
  constexpr int XS = kBX, AS = kBA;               // 52 input rows / cols, 42 statistics rows / cols
  extern __shared__ float sm[];
  float (*sx)[kBX + 1] = reinterpret_cast<float (*)[kBX + 1]>(sm);
  float (*sy)[kBX + 1] = reinterpret_cast<float (*)[kBX + 1]>(sm + kBX * (kBX + 1));
  float* hbase = sm + 2 * kBX * (kBX + 1);
  auto h5 = [&](int m, int r, int c) -> float& { return hbase[(m * kBX + r) * (kBA + 1) + c]; };
  float* abase = hbase + 5 * kBX * (kBA + 1);
  auto a3 = [&](int m, int r, int c) -> float& { return abase[(m * kBA + r) * (kBA + 1) + c]; };
  float* tbase = abase + 3 * kBA * (kBA + 1);
  auto ha = [&](int m, int r, int c) -> float& { return tbase[(m * kBA + r) * (kBT + 1) + c]; };

  const int plane = blockIdx.z;
  const int ty0 = blockIdx.y * kBT, tx0 = blockIdx.x * kBT;
  const float* X = p.X + (int64_t)plane * p.h * p.w;
  const float* Y = p.Y + (int64_t)plane * p.h * p.w;
  const float ccs = p.coef_cs[plane], css = p.coef_ss[plane];
  for (int i = threadIdx.x; i < XS * XS; i += 256) {
    const int r = i / XS, c = i % XS;
    const int gy = ty0 - R + r, gx = tx0 - R + c;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < p.h && gx >= 0 && gx < p.w) { a = __ldg(X + (int64_t)gy * p.w + gx); b = __ldg(Y + (int64_t)gy * p.w + gx); }
    sx[r][c] = a; sy[r][c] = b;
  }
  float taps[W];
#pragma unroll
  for (int k = 0; k < W; ++k) taps[k] = p.taps[k];
  __syncthreads();
  // (1) horizontal forward filter: item = (input row, run of 6 statistics columns)
  constexpr int RUN6 = 6;
  for (int i = threadIdx.x; i < XS * (AS / RUN6); i += 256) {
    const int r = i / (AS / RUN6), c0 = (i % (AS / RUN6)) * RUN6;
    float a[RUN6 + W - 1], b[RUN6 + W - 1];
#pragma unroll
    for (int k = 0; k < RUN6 + W - 1; ++k) { a[k] = sx[r][c0 + k]; b[k] = sy[r][c0 + k]; }
#pragma unroll
    for (int o = 0; o < RUN6; ++o) {
      float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float g = taps[k], u = a[o + k], v = b[o + k];
        m1 = fmaf(g, u, m1); m2 = fmaf(g, v, m2);
        s11 = fmaf(g, u * u, s11); s22 = fmaf(g, v * v, s22); s12 = fmaf(g, u * v, s12);
      }
      h5(0, r, c0 + o) = m1; h5(1, r, c0 + o) = m2; h5(2, r, c0 + o) = s11; h5(3, r, c0 + o) = s22; h5(4, r, c0 + o) = s12;
    }
  }
  __syncthreads();
  // (2) vertical forward filter + per-position coefficients: item = (statistics column, run of 6 rows)
  for (int i = threadIdx.x; i < AS * (AS / RUN6); i += 256) {
    const int c = i % AS, r0 = (i / AS) * RUN6;
    float st[5][RUN6];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      float v[RUN6 + W - 1];
#pragma unroll
      for (int k = 0; k < RUN6 + W - 1; ++k) v[k] = h5(m, r0 + k, c);
#pragma unroll
      for (int o = 0; o < RUN6; ++o) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) t = fmaf(taps[k], v[o + k], t);
        st[m][o] = t;
      }
    }
#pragma unroll
    for (int o = 0; o < RUN6; ++o) {
      const int r = r0 + o;
      const int oy = ty0 - p.so + r, ox = tx0 - p.so + c;
      float A12 = 0.f, A11 = 0.f, B = 0.f;
      if (oy >= 0 && oy < p.oh && ox >= 0 && ox < p.ow) {
        const float m1 = st[0][o], m2 = st[1][o];
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float v1 = st[2][o] - m11, v2 = st[3][o] - m22, v12 = st[4][o] - m12;
        const float D = v1 + v2 + p.c2, cs = (2.f * v12 + p.c2) / D;
        const float Dl = m11 + m22 + p.c1, lum = (2.f * m12 + p.c1) / Dl;
        const float a_cs = ccs + css * lum;
        A12 = a_cs * 2.f / D;
        A11 = -a_cs * cs / D;
        B = -m2 * A12 - 2.f * m1 * A11 + css * cs * 2.f * (m2 - lum * m1) / Dl;
      }
      a3(0, r, c) = A12; a3(1, r, c) = A11; a3(2, r, c) = B;
    }
  }
  __syncthreads();
  // (3) horizontal transposed filter: t[c] = sum_k g[k] a[c + R - k]: item = (statistics row, run of 4 output columns)
  for (int i = threadIdx.x; i < AS * (kBT / kRun); i += 256) {
    const int r = i / (kBT / kRun), c0 = (i % (kBT / kRun)) * kRun;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float v[kRun + W - 1];
#pragma unroll
      for (int k = 0; k < kRun + W - 1; ++k) v[k] = a3(m, r, c0 + k);
#pragma unroll
      for (int o = 0; o < kRun; ++o) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) t = fmaf(taps[k], v[o + R - k], t);
        ha(m, r, c0 + o) = t;
      }
    }
  }
  __syncthreads();
  // (4) vertical transposed filter + combine: item = (output column, run of 4 rows); 32 x 8 items = one per thread
  {
    const int c = threadIdx.x % kBT, r0 = (threadIdx.x / kBT) * kRun;
    float t3[3][kRun];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      float v[kRun + W - 1];
#pragma unroll
      for (int k = 0; k < kRun + W - 1; ++k) v[k] = ha(m, r0 + k, c);
#pragma unroll
      for (int o = 0; o < kRun; ++o) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) t = fmaf(taps[k], v[o + R - k], t);
        t3[m][o] = t;
      }
    }
    float* out = p.dX + (int64_t)plane * p.h * p.w;
#pragma unroll
    for (int o = 0; o < kRun; ++o) {
      const int r = r0 + o;
      const int gy = ty0 + r, gx = tx0 + c;
      if (gy >= p.h || gx >= p.w) continue;
      float d = sy[r + R][c + R] * t3[0][o] + 2.f * sx[r + R][c + R] * t3[1][o] + t3[2][o];
      if (p.dXnext != nullptr) {
        const int py = (gy + p.ph) >> 1, px = (gx + p.pw) >> 1;
        if (py < p.nh && px < p.nw) d += 0.25f * __ldg(p.dXnext + ((int64_t)plane * p.nh + py) * p.nw + px);
      }
      out[(int64_t)gy * p.w + gx] = d;
    }
  }
}



// ---------------------------------------------------------------------------------------------------------
// Fused value + unit gradient of one level (11-tap window), row-marching form.  The tile kernels above recompute the
// statistics of a 32x32 tile from a 52x52 input tile (288 multiply-adds per pixel, one shared-memory load per 1-4 of
// them); here a block owns a strip of kVgT columns and walks down the rows, every thread keeping the last eleven rows
// of ITS column in registers:
//     input row n -> horizontal filter (shared-memory row) -> ring of 11 x 4 row-filtered statistics
//                 -> vertical filter = statistics of row n-10 -> per-position coefficients A12, A11, B -> ring of 11 x 3
//                 -> vertical transposed filter (registers) -> shared-memory row -> horizontal transposed filter
//                 -> U[n-10] = Y * T[A12] + 2 X * T[A11] + T[B]
// so the vertical passes read registers only and no statistic is computed twice (about 190 multiply-adds and 55 shared
// loads per pixel, forward value included).  Four statistics suffice: X^2 and Y^2 only occur summed.
//   U = d(sum over the level's map)/dX with unit upstream weight: the map is cs (levels 0..3: ccs = 1, css = 0) or
//   ssim = lum * cs (last level: ccs = 0, css = 1).  The per-plane weights depend on ALL levels' values, so they are
//   applied afterwards by ssim_combine_kernel together with the 2x2-average-pool chain between levels.
// Coordinates are those of the zero-padded frame (pad = 0: valid window, variant 1; pad = 5: zero "same" padding,
// variant 2): frame (hp, wp) = (h + 2 pad, w + 2 pad), statistics positions (oh, ow) = (hp - 10, wp - 10), position o
// reads frame rows / columns o .. o+10, the gradient of frame pixel i collects positions i-10 .. i.
// The row loop is unrolled eleven times so that the ring slot of a row is a compile-time register name.  Measured
// alternatives (B200, 32 x 3 x 512 x 768, level 0): this form 0.98 ms (ncu: issue slots 64 % busy; first stall reason
// "no instruction" -- the 74 KB loop body misses the 32 KB instruction cache); only the two vertical filters instantiated
// per slot behind a block-uniform switch (25 KB of code) 1.04 ms -- the 128-register cap then makes ptxas rematerialise
// the plane addresses every row; the same with 3 blocks per SM (138 registers) 1.44 ms: the kernel lives on its 16 warps
// per SM.  The two-pass tile kernels above take 0.56 + 1.73 ms for the same work.
// A block computes statistics columns c0 .. c0+kVgT-1 and the kVgT-10 output columns that need only those; a row
// segment [i0, i1) of outputs costs 20 extra rows of warm-up.  Per-block partial sums (fixed order) go to ws.
// ---------------------------------------------------------------------------------------------------------
constexpr int kVgT = 128;      // threads = statistics columns per strip
constexpr int kVgR = kMaxWin - 1;

struct SsimVgParams {
  const float* X; const float* Y; float* U; float* ws;
  int h, w, pad, hp, wp, oh, ow, seg_rows;
  float c1, c2, ccs, css;
  float taps[kMaxWin];
};

// Arithmetic is packed two-wide (sm_100 FFMA2, __ffma2_rn: one instruction = two fp32 fused multiply-adds, the tap a
// broadcast uniform-register operand): the statistics travel as the pairs (mu_x, mu_y) and (x^2 + y^2, xy), the gradient
// coefficients as (A12, A11) + B, and the shared rows hold (x, y, x^2 + y^2, xy) as float4 -- the products are made once per pixel, not once per tap --
// and (V12, V11) as float2.  The kernel is issue bound
// (ncu: FMA pipe 34 % busy at 64 % of the issue slots), so halving the multiply-add and shared-load instruction counts
// is what counts; every component is still one IEEE fma, bit-identical to the scalar form.
// 1 / x for the two SSIM denominators (both >= C1 or C2 > 0 up to rounding of the variance): MUFU.RCP + one Newton step,
// no slow-path branch -- an IEEE division would split the row body into several basic blocks.  <= 1 ulp from 1.f / x.
__device__ __forceinline__ float vg_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.f), r);
}

__global__ void __launch_bounds__(kVgT, 4) ssim_level_vg_kernel(const SsimVgParams p) {
  constexpr int W = kMaxWin, R = kVgR, T = kVgT;
  __shared__ float4 s_in[2][T + R];            // (x, y, x^2 + y^2, xy) of the frame row: products made once per pixel
  __shared__ float2 s_va[2][R + T];            // vertically filtered (A12, A11); R pad entries in front
  __shared__ float s_vb[2][R + T];             // vertically filtered B
  __shared__ float s_red[2][T / 32];
  const int t = threadIdx.x;
  const int plane = blockIdx.z;
  const int c0 = -R + (int)blockIdx.x * (T - R);
  const int i0 = (int)blockIdx.y * p.seg_rows, i1 = min(i0 + p.seg_rows, p.hp);
  const float* __restrict__ X = p.X + (int64_t)plane * p.h * p.w;
  const float* __restrict__ Y = p.Y + (int64_t)plane * p.h * p.w;
  float* __restrict__ U = p.U + (int64_t)plane * p.h * p.w;
  const int ox = c0 + t;                       // statistics column = output column (frame coordinates)
  const bool ox_valid = ox >= 0 && ox < p.ow;
  const bool own_col = t >= R && ox_valid;     // counted in the sums by this strip
  const int gx0 = ox - p.pad;                  // image column of frame column ox
  const bool gx0_in = gx0 >= 0 && gx0 < p.w;
  const int gx1 = gx0 + T;                     // the R extra columns right of the strip (threads 0..R-1)
  const bool gx1_in = t < R && gx1 >= 0 && gx1 < p.w;
  const bool out_col = t >= R && gx0_in;

  float2 hm[W], hs[W], aa[W];                  // rings: (mu_x, mu_y), (x^2 + y^2, xy) row-filtered; (A12, A11)
  float ab[W];                                 //        B
#pragma unroll
  for (int k = 0; k < W; ++k) {
    hm[k] = hs[k] = aa[k] = make_float2(0.f, 0.f);
    ab[k] = 0.f;
  }
  if (t < R) {
    s_va[0][t] = s_va[1][t] = make_float2(0.f, 0.f);
    s_vb[0][t] = s_vb[1][t] = 0.f;
  }
  float acc_s = 0.f, acc_c = 0.f;
  float2 in0 = make_float2(0.f, 0.f), in1 = make_float2(0.f, 0.f);
  // running offsets inside the plane (h * w < 2^31): the input row being fetched and the output row being produced
  int gy_in = i0 - R - p.pad, off_in = gy_in * p.w + gx0;
  if (gy_in >= 0 && gy_in < p.h) {
    if (gx0_in) { in0.x = __ldg(X + off_in); in0.y = __ldg(Y + off_in); }
    if (gx1_in) { in1.x = __ldg(X + off_in + T); in1.y = __ldg(Y + off_in + T); }
  }
  // One barrier per row: the row-filtered V of iteration n is consumed (horizontal transposed filter, output row n-11)
  // by iteration n+1, after that iteration's barrier; both shared rows are double-buffered.
  const int n_last = i1 + R;                   // consumer-only iteration that drains the last V row
  for (int base = i0 - R; base <= n_last; base += W) {
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const int n = base + s;                  // frame row entering the pipeline
      if (n > n_last) break;
      const int b = n & 1;
      const bool produce = n < n_last;         // block-uniform
      s_in[b][t] = make_float4(in0.x, in0.y, fmaf(in0.x, in0.x, in0.y * in0.y), in0.x * in0.y);
      if (t < R) s_in[b][T + t] = make_float4(in1.x, in1.y, fmaf(in1.x, in1.x, in1.y * in1.y), in1.x * in1.y);
      // next input row and the operands of the output row drained in this iteration: issued ahead of the arithmetic
      ++gy_in; off_in += p.w;
      in0 = in1 = make_float2(0.f, 0.f);
      if (n + 1 < n_last && gy_in >= 0 && gy_in < p.h) {
        if (gx0_in) { in0.x = __ldg(X + off_in); in0.y = __ldg(Y + off_in); }
        if (gx1_in) { in1.x = __ldg(X + off_in + T); in1.y = __ldg(Y + off_in + T); }
      }
      const int io = n - R - 1;                // output row drained by this iteration
      const int gyo = io - p.pad;
      const bool drain = io >= i0;             // block-uniform
      const bool do_out = drain && out_col && gyo >= 0 && gyo < p.h;
      const int off_o = gyo * p.w + gx0;
      float xo = 0.f, yo = 0.f;
      if (do_out) { xo = __ldg(X + off_o); yo = __ldg(Y + off_o); }
      __syncthreads();
      // One straight-line body per row: the four filter passes are independent multiply-add chains (the vertical ones
      // need this row's horizontal result only at their LAST tap) and the scheduler may interleave them; the
      // block-uniform conditions of the pipeline's warm-up rows only predicate the side effects.  Rows computed on a
      // partly filled ring hold finite garbage that no owned sum and no drained output ever reads (see the header).
      float2 h01 = make_float2(0.f, 0.f), h23 = make_float2(0.f, 0.f);
      float2 ta = make_float2(0.f, 0.f);
      float tb = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const float g = p.taps[k];
        const float2 gg = make_float2(g, g);
        const float4 e = s_in[b][t + k];
        h01 = __ffma2_rn(gg, make_float2(e.x, e.y), h01);
        h23 = __ffma2_rn(gg, make_float2(e.z, e.w), h23);
        // V row io was written by the previous iteration into buffer b^1 (R pad entries in front: threads < R read them)
        ta = __ffma2_rn(gg, s_va[b ^ 1][R + t - k], ta); tb = fmaf(g, s_vb[b ^ 1][R + t - k], tb);
      }
      hm[s] = h01; hs[s] = h23;
      if (do_out) U[off_o] = yo * ta.x + 2.f * xo * ta.y + tb;
      const int i = n - R;                     // statistics row finished by this iteration
      float2 m = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < W; ++k) {            // k = 0 is the oldest ring row (frame row n-10)
        const int r = (s + 1 + k) % W;
        const float g = p.taps[k];
        const float2 gg = make_float2(g, g);
        m = __ffma2_rn(gg, hm[r], m); q = __ffma2_rn(gg, hs[r], q);
      }
      float A12, A11, B;
      {
        const float m1 = m.x, m2 = m.y, sq = q.x, s12 = q.y;
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float rD = vg_rcp((sq - m11 - m22) + p.c2);
        const float cs = (2.f * (s12 - m12) + p.c2) * rD;
        const float rDl = vg_rcp(m11 + m22 + p.c1);
        const float lum = (2.f * m12 + p.c1) * rDl;
        const bool valid = produce && ox_valid && i >= 0 && i < p.oh;
        if (valid && own_col && i >= i0) { acc_c += cs; acc_s += lum * cs; }
        const float a_cs = p.ccs + p.css * lum;
        A12 = a_cs * 2.f * rD;
        A11 = -a_cs * cs * rD;
        B = -m2 * A12 - 2.f * m1 * A11 + p.css * cs * 2.f * (m2 - lum * m1) * rDl;
        A12 = valid ? A12 : 0.f; A11 = valid ? A11 : 0.f; B = valid ? B : 0.f;
      }
      aa[s] = make_float2(A12, A11); ab[s] = B;
      float2 va = make_float2(0.f, 0.f);
      float vb = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) {            // V[i] = sum_k g[k] A[i-k]; A[i-k] sits k slots back
        const int r = (s - k + W) % W;
        const float g = p.taps[k];
        va = __ffma2_rn(make_float2(g, g), aa[r], va); vb = fmaf(g, ab[r], vb);
      }
      s_va[b][R + t] = va; s_vb[b][R + t] = vb;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_s += __shfl_xor_sync(0xffffffffu, acc_s, o);
    acc_c += __shfl_xor_sync(0xffffffffu, acc_c, o);
  }
  if ((t & 31) == 0) { s_red[0][t >> 5] = acc_s; s_red[1][t >> 5] = acc_c; }
  __syncthreads();
  if (t == 0) {
    float s = 0.f, c = 0.f;
    for (int k = 0; k < T / 32; ++k) { s += s_red[0][k]; c += s_red[1][k]; }
    const int nb = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
    p.ws[((int64_t)plane * nb + blk) * 2 + 0] = s;
    p.ws[((int64_t)plane * nb + blk) * 2 + 1] = c;
  }
}

// D_l = coef[plane] * U_l + 0.25 * D_{l+1}[pool parent], in place, coarse to fine: the chain rule through
// value = prod_l relu(level value)^w_l and through F.avg_pool2d(2, padding (ph, pw)) between levels.
// w % 4 == 0 and an unpadded pool (pw = 0): four pixels per thread, one float4 of U and one float2 of parents
__global__ void __launch_bounds__(256) ssim_combine4_kernel(float* __restrict__ U, const float* __restrict__ coef,
                                                            const float* __restrict__ Dnext, int h, int w4, int nh, int nw,
                                                            int ph) {
  const int plane = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= h * w4) return;
  const int gy = i / w4, g4 = i - gy * w4;
  const float c = coef[plane];
  float4* row = reinterpret_cast<float4*>(U + ((int64_t)plane * h + gy) * (w4 * 4));
  float4 u = row[g4];
  float2 par = make_float2(0.f, 0.f);
  const int py = (gy + ph) >> 1;
  if (Dnext != nullptr && py < nh)
    par = __ldg(reinterpret_cast<const float2*>(Dnext + ((int64_t)plane * nh + py) * nw) + g4);
  u.x = fmaf(0.25f, par.x, c * u.x); u.y = fmaf(0.25f, par.x, c * u.y);
  u.z = fmaf(0.25f, par.y, c * u.z); u.w = fmaf(0.25f, par.y, c * u.w);
  row[g4] = u;
}

__global__ void __launch_bounds__(256) ssim_combine_kernel(float* __restrict__ U, const float* __restrict__ coef,
                                                           const float* __restrict__ Dnext, int h, int w, int nh, int nw,
                                                           int ph, int pw) {
  const int plane = blockIdx.z, gy = blockIdx.y;
  const float c = coef[plane];
  float* row = U + ((int64_t)plane * h + gy) * w;
  const int py = (gy + ph) >> 1;
  const float* nrow = (Dnext != nullptr && py < nh) ? Dnext + ((int64_t)plane * nh + py) * nw : nullptr;
  for (int gx = blockIdx.x * 256 + threadIdx.x; gx < w; gx += gridDim.x * 256) {
    float d = c * row[gx];
    if (nrow != nullptr) {
      const int px = (gx + pw) >> 1;
      if (px < nw) d = fmaf(0.25f, __ldg(nrow + px), d);
    }
    row[gx] = d;
  }
}


// All the scalar work between the level launches and the combine launches in ONE launch (it was five finalize launches and
// ~30 ATen kernels on [levels, B, C] tensors per composition -- at one image per call, as the reference CLI runs, those
// launches were a third of an ms-ssim iteration): per plane and level the block partials are summed in fixed order,
//   v_l = relu(mean of the cs map)  (ssim map at the last level),   P = prod_l v_l^w_l,   value[b] = mean_c P,
//   coef_l[plane] = upstream[b] / C * w_l * P / v_l / npx_l   (0 where v_l <= 0: the relu's gradient)
// which is autograd of pytorch_msssim.ms_ssim's `torch.prod(relu(mcs_and_ssim) ** weights)` and its channel mean.
constexpr int kMaxLevels = 8;
struct MsCoefParams {
  const float* ws; const float* upstream; float* value; float* coef;
  int levels, batch, channels;
  int ws_off[kMaxLevels], nb[kMaxLevels];
  float inv_npx[kMaxLevels], weight[kMaxLevels];
};

// One warp per image: the 32 lanes share every partial list (lane k takes partials k, k + 32, ...; xor-butterfly sum, the
// same order on every run), so the longest list -- 112 block partials per plane at one image -- is four loads deep.  (One
// thread per image walking its lists serially measured 24 us at 16 images: longer than the kernels it replaced.)
__global__ void __launch_bounds__(32) msssim_coef_kernel(const MsCoefParams p) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int planes = p.batch * p.channels;
  const float up = p.upstream[b] / (float)p.channels;
  float vsum = 0.f;
  for (int c = 0; c < p.channels; ++c) {
    const int pl = b * p.channels + c;
    float v[kMaxLevels];
    float P = 1.f;
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      if (l >= p.levels) break;
      const float* w = p.ws + p.ws_off[l] + (int64_t)pl * p.nb[l] * 2 + (l == p.levels - 1 ? 0 : 1);
      float s = 0.f;
      for (int k = lane; k < p.nb[l]; k += 32) s += __ldg(w + 2 * k);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      v[l] = fmaxf(s * p.inv_npx[l], 0.f);
      P *= powf(v[l], p.weight[l]);
    }
    vsum += P;
#pragma unroll
    for (int l = 0; l < kMaxLevels; ++l) {
      if (l >= p.levels) break;
      if (lane == 0) p.coef[l * planes + pl] = v[l] > 0.f ? up * p.weight[l] * P / v[l] * p.inv_npx[l] : 0.f;
    }
  }
  if (lane == 0) p.value[b] = vsum / (float)p.channels;
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_ssim_workspace_floats(int planes, int h, int w, int win, int same_pad) {
  const int oh = same_pad ? h : h - win + 1, ow = same_pad ? w : w - win + 1;
  if (oh <= 0 || ow <= 0) return 0;
  return planes * ((oh + kT - 1) / kT) * ((ow + kT - 1) / kT) * 2;
}

int icadv_ssim_level(const float* X, const float* Y, float* ws, float* ssim_sum, float* cs_sum, int planes, int h,
                     int w, const float* win_taps_host, int win, int same_pad, float c1, float c2,
                     icadv_stream_t stream) {
  ICADV_REQUIRE(X && Y && ws && ssim_sum && cs_sum && win_taps_host, "null pointer");
  ICADV_REQUIRE(win >= 1 && win <= kMaxWin && (win % 2 == 1 || !same_pad), "window must be odd and <= 11");
  SsimParams p;
  p.X = X; p.Y = Y; p.ws = ws; p.h = h; p.w = w; p.win = win; p.pad = same_pad ? win / 2 : 0;
  p.out_h = same_pad ? h : h - win + 1; p.out_w = same_pad ? w : w - win + 1;
  ICADV_REQUIRE(p.out_h > 0 && p.out_w > 0, "image smaller than the window");
  p.tiles_x = (p.out_w + kT - 1) / kT; p.tiles_y = (p.out_h + kT - 1) / kT;
  ICADV_REQUIRE(planes <= 65535 && p.tiles_y <= 65535, "grid too large");
  p.c1 = c1; p.c2 = c2;
  for (int k = 0; k < kMaxWin; ++k) p.taps[k] = k < win ? win_taps_host[k] : 0.f;
  dim3 grid(p.tiles_x, p.tiles_y, planes);
  if (win == kMaxWin) ssim_level_kernel11<<<grid, 256, 0, as_stream(stream)>>>(p);
  else ssim_level_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  ssim_finalize_kernel<<<(planes + 127) / 128, 128, 0, as_stream(stream)>>>(ws, ssim_sum, cs_sum, planes,
                                                                            p.tiles_x * p.tiles_y);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_avgpool2(const float* x, float* y, int planes, int h, int w, int pad_h, int pad_w, icadv_stream_t stream) {
  ICADV_REQUIRE(x && y && planes > 0 && h > 0 && w > 0, "bad avgpool2 args");
  const int oh = (h + 2 * pad_h - 2) / 2 + 1, ow = (w + 2 * pad_w - 2) / 2 + 1;
  if (pad_h == 0 && pad_w == 0 && w % 4 == 0 && h % 2 == 0 && planes <= 65535 &&
      reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 8 == 0) {
    dim3 grid((oh * (ow / 2) + 255) / 256, planes);
    avgpool2_vec_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, h, w, oh, ow / 2);
    ICADV_CUDA_TRY(cudaGetLastError());
    return ICADV_OK;
  }
  const int64_t total = (int64_t)planes * oh * ow;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  avgpool2_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, y, planes, h, w, oh, ow, pad_h, pad_w);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_ssim_level_backward(const float* X, const float* Y, const float* coef_cs, const float* coef_ss,
                              const float* dXnext, float* dX, int planes, int h, int w, int next_h, int next_w,
                              int pad_h, int pad_w, const float* win_taps_host, int win, int same_pad, float c1,
                              float c2, icadv_stream_t stream) {
  ICADV_REQUIRE(X && Y && coef_cs && coef_ss && dX && win_taps_host, "null pointer");
  ICADV_REQUIRE(win >= 1 && win <= kMaxWin && h >= win && w >= win, "bad window / image size");
  ICADV_REQUIRE(!same_pad || win % 2 == 1, "same padding needs an odd window");
  SsimBwdParams p;
  p.X = X; p.Y = Y; p.coef_cs = coef_cs; p.coef_ss = coef_ss; p.dXnext = dXnext; p.dX = dX;
  p.h = h; p.w = w; p.win = win;
  p.oh = same_pad ? h : h - win + 1; p.ow = same_pad ? w : w - win + 1;
  p.so = same_pad ? win / 2 : win - 1;
  p.nh = next_h; p.nw = next_w; p.ph = pad_h; p.pw = pad_w; p.c1 = c1; p.c2 = c2;
  for (int k = 0; k < kMaxWin; ++k) p.taps[k] = k < win ? win_taps_host[k] : 0.f;
  const size_t smem = sizeof(float) * (2 * kBX * (kBX + 1) + 5 * kBX * (kBA + 1) + 3 * kBA * (kBA + 1) + 3 * kBA * (kBT + 1));
  static bool attr_done = false;
  if (!attr_done) {
    ICADV_CUDA_TRY(cudaFuncSetAttribute(ssim_level_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ICADV_CUDA_TRY(cudaFuncSetAttribute(ssim_level_bwd_kernel11, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid((w + kBT - 1) / kBT, (h + kBT - 1) / kBT, planes);
  ICADV_REQUIRE(planes <= 65535 && grid.y <= 65535, "grid too large");
  if (win == kMaxWin) ssim_level_bwd_kernel11<<<grid, 256, smem, as_stream(stream)>>>(p);
  else ssim_level_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}


static int vg_segments(int planes, int hp, int wp, int* seg_rows) {
  const int strips = (wp + (kVgT - kVgR) - 1) / (kVgT - kVgR);
  // enough blocks for ~2 waves of 4 blocks per SM, rows per segment >= 32 (20 warm-up rows per segment)
  int nseg = (148 * 4 * 2 + planes * strips - 1) / (planes * strips);
  const int max_seg = (hp + 31) / 32;
  if (nseg > max_seg) nseg = max_seg;
  if (nseg < 1) nseg = 1;
  *seg_rows = (hp + nseg - 1) / nseg;
  return (hp + *seg_rows - 1) / *seg_rows;
}

int icadv_ssim_vg_workspace_floats(int planes, int h, int w, int same_pad) {
  const int pad = same_pad ? kMaxWin / 2 : 0, hp = h + 2 * pad, wp = w + 2 * pad;
  if (hp < kMaxWin || wp < kMaxWin || planes <= 0) return 0;
  int seg_rows;
  const int nseg = vg_segments(planes, hp, wp, &seg_rows);
  return planes * ((wp + (kVgT - kVgR) - 1) / (kVgT - kVgR)) * nseg * 2;
}

int icadv_ssim_level_value_grad(const float* X, const float* Y, float* U, float* ws, float* ssim_sum, float* cs_sum,
                                int planes, int h, int w, const float* win_taps_host, int win, int same_pad, float c1,
                                float c2, int last_level, icadv_stream_t stream) {
  ICADV_REQUIRE(X && Y && U && ws && win_taps_host, "null pointer");
  ICADV_REQUIRE((ssim_sum == nullptr) == (cs_sum == nullptr), "ssim_sum and cs_sum go together");
  ICADV_REQUIRE(win == kMaxWin, "the fused value + gradient level kernel takes the 11-tap window only");
  SsimVgParams p;
  p.X = X; p.Y = Y; p.U = U; p.ws = ws; p.h = h; p.w = w; p.pad = same_pad ? win / 2 : 0;
  p.hp = h + 2 * p.pad; p.wp = w + 2 * p.pad; p.oh = p.hp - kVgR; p.ow = p.wp - kVgR;
  ICADV_REQUIRE(planes > 0 && p.oh > 0 && p.ow > 0, "image smaller than the window");
  p.c1 = c1; p.c2 = c2; p.ccs = last_level ? 0.f : 1.f; p.css = last_level ? 1.f : 0.f;
  for (int k = 0; k < kMaxWin; ++k) p.taps[k] = win_taps_host[k];
  const int strips = (p.wp + (kVgT - kVgR) - 1) / (kVgT - kVgR);
  const int nseg = vg_segments(planes, p.hp, p.wp, &p.seg_rows);
  ICADV_REQUIRE(planes <= 65535 && nseg <= 65535, "grid too large");
  dim3 grid(strips, nseg, planes);
  ssim_level_vg_kernel<<<grid, kVgT, 0, as_stream(stream)>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  if (ssim_sum != nullptr) {   // NULL: the block partials stay in ws for icadv_msssim_coefficients
    ssim_finalize_kernel<<<(planes + 127) / 128, 128, 0, as_stream(stream)>>>(ws, ssim_sum, cs_sum, planes, strips * nseg);
    ICADV_CUDA_TRY(cudaGetLastError());
  }
  return ICADV_OK;
}

int icadv_msssim_coefficients(const float* ws_all, const int* ws_offset_host, const int* nblocks_host,
                              const float* npx_host, const float* weights_host, int levels, const float* upstream,
                              float* value, float* coef, int batch, int channels, icadv_stream_t stream) {
  ICADV_REQUIRE(ws_all && ws_offset_host && nblocks_host && npx_host && weights_host && upstream && value && coef,
                "null pointer");
  ICADV_REQUIRE(levels >= 1 && levels <= kMaxLevels && batch >= 1 && channels >= 1, "bad msssim_coefficients args");
  MsCoefParams p;
  p.ws = ws_all; p.upstream = upstream; p.value = value; p.coef = coef;
  p.levels = levels; p.batch = batch; p.channels = channels;
  for (int l = 0; l < levels; ++l) {
    ICADV_REQUIRE(nblocks_host[l] >= 1 && npx_host[l] > 0.f, "bad level geometry");
    p.ws_off[l] = ws_offset_host[l]; p.nb[l] = nblocks_host[l];
    p.inv_npx[l] = 1.f / npx_host[l]; p.weight[l] = weights_host[l];
  }
  msssim_coef_kernel<<<batch, 32, 0, as_stream(stream)>>>(p);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

int icadv_ssim_combine(float* U, const float* coef, const float* Dnext, int planes, int h, int w, int next_h, int next_w,
                       int pad_h, int pad_w, icadv_stream_t stream) {
  ICADV_REQUIRE(U && coef && planes > 0 && h > 0 && w > 0, "bad ssim_combine args");
  ICADV_REQUIRE(planes <= 65535 && h <= 65535, "grid too large");
  // the vector form needs 16-byte rows of U and 8-byte rows of parents that cover every pixel pair (nw = w / 2)
  const bool vec = w % 4 == 0 && (Dnext == nullptr || (pad_w == 0 && next_w == w / 2)) &&
                   reinterpret_cast<uintptr_t>(U) % 16 == 0 && reinterpret_cast<uintptr_t>(Dnext) % 8 == 0;
  if (vec) {
    dim3 grid4((h * (w / 4) + 255) / 256, planes);
    ssim_combine4_kernel<<<grid4, 256, 0, as_stream(stream)>>>(U, coef, Dnext, h, w / 4, next_h, next_w, pad_h);
    ICADV_CUDA_TRY(cudaGetLastError());
    return ICADV_OK;
  }
  dim3 grid((w + 1023) / 1024, h, planes);
  ssim_combine_kernel<<<grid, 256, 0, as_stream(stream)>>>(U, coef, Dnext, h, w, next_h, next_w, pad_h, pad_w);
  ICADV_CUDA_TRY(cudaGetLastError());
  return ICADV_OK;
}

}  // extern "C"
