// tcgen05 / TMEM / TMA implicit-GEMM for the codec's conv and transposed-conv stacks, with GDN / IGDN
// fused into the epilogue in both directions.  Replaces nn.Conv2d / nn.ConvTranspose2d + compressai GDN
// under net.g_a / net.g_s (reference: anchors/utils.py:112-130, utils/ops.py:58-97, attack_rd.py:344-349,547).
//
// One CTA = one 8x16 tile of "tile-space" pixels (M = 128 rows) x all n_ch output channels (N <= 256).
//   D[128, N] = sum over taps t, 32-channel chunks kc:  A_t,kc[128, 32] * W_t,kc[N, 32]^T       (kind::tf32)
// A_t,kc is one TMA box [32 ch, 16 w, 8 h, 1 img] of the channels-last input, shifted by the tap offset
// (out-of-bounds = zero padding); for stride-2 convs the input is addressed through four parity-plane
// tensor maps so every tap is a unit-stride box.  W_t,kc is a TMA box [32, N] of the packed weights.
// Both land in SWIZZLE_128B K-major layout and feed tcgen05.mma directly; the accumulator lives in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
// GDN epilogues append n_ch/32 extra K-blocks to the same smem ring: the A operand of those blocks
// (x^2 forward, g*y*sc^2 backward) is written by the epilogue warps, the B operand (gamma / gamma^T)
// arrives by TMA, and the product accumulates into a second TMEM region.
#include <mutex>

#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

constexpr int kTH = 8, kTW = 16, kTileM = 128;
constexpr int kABytes = kTileM * 128;   // one [128 x 32 fp32] operand / staging tile
constexpr int kEpiBufs = 4;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;      // 227 KB
constexpr int kBarBytes = 512;

struct TcParams {
  CUtensorMap a_map[4];
  CUtensorMap w_map, g_map, out_map, sc_map, yprev_map, scprev_map;
  Tap taps[kMaxTaps];
  int num_taps, k_chunks, n_ch, n_chunks;
  int tiles_x, tiles_y;
  int epi, act, acc_from_in, round_out;
  int num_stages, stage_bytes, tmem_cols;
  const float* bias;
  const float* beta;
  const int* active;
  const int* n_active;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ICADV_ACT_RELU) return fmaxf(v, 0.f);
  if (act == ICADV_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == ICADV_ACT_ABS) return fabsf(v);
  return v;
}

__device__ __forceinline__ void read_row32(const uint8_t* buf, int row, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 t = *reinterpret_cast<const float4*>(buf + sw128_off(row, j));
    v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ void write_row32(uint8_t* buf, int row, const float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(buf + sw128_off(row, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

__global__ void __launch_bounds__(192, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.y;
  if (p.n_active != nullptr && slot >= *p.n_active) return;  // whole CTA leaves together
  const int img = p.active != nullptr ? p.active[slot] : slot;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x % p.tiles_x;
  const int i0 = ty * kTH, j0 = tx * kTW;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_buf = smem + p.num_stages * p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_buf + kEpiBufs * kABytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;   // [2]
  uint64_t* a2_ready = acc_full + 2;         // [8]
  uint64_t* ld_full = a2_ready + 8;          // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ld_full + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    for (int c = 0; c < 8; ++c) mbar_init(&a2_ready[c], 128);
    mbar_init(ld_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, p.tmem_cols); tmem_relinquish(); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.out_map);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  const bool gdn = p.epi != ICADV_EPI_LINEAR;
  const bool bwd = p.epi == ICADV_EPI_GDN_BWD || p.epi == ICADV_EPI_IGDN_BWD;
  const int main_kb = p.acc_from_in ? 0 : p.num_taps * p.k_chunks;
  const int gdn_kb = gdn ? p.n_chunks : 0;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_ch) * 128u;
  const int S = p.num_stages;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int kb = 0;
      for (int t = 0; t < (p.acc_from_in ? 0 : p.num_taps); ++t) {
        const Tap tap = p.taps[t];
        for (int kc = 0; kc < p.k_chunks; ++kc, ++kb) {
          const int s = kb % S;
          mbar_wait(&empty[s], ((kb / S) & 1) ^ 1);
          uint8_t* st = smem + s * p.stage_bytes;
          mbar_arrive_expect_tx(&full[s], kABytes + b_bytes);
          tma_load_4d(st, &p.a_map[tap.plane], &full[s], kc * 32, j0 + tap.dx, i0 + tap.dy, img);
          tma_load_2d(st + kABytes, &p.w_map, &full[s], kc * 32, tap.wtap * p.n_ch);
        }
      }
      for (int c = 0; c < gdn_kb; ++c, ++kb) {
        const int s = kb % S;
        mbar_wait(&empty[s], ((kb / S) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], b_bytes);
        tma_load_2d(smem + s * p.stage_bytes + kABytes, &p.g_map, &full[s], c * 32, 0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      int kb = 0;
      for (; kb < main_kb; ++kb) {
        const int s = kb % S;
        mbar_wait(&full[s], (kb / S) & 1);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * p.stage_bytes);
        const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      if (main_kb > 0) tc_commit(&acc_full[0]);
      for (int c = 0; c < gdn_kb; ++c, ++kb) {
        const int s = kb % S;
        mbar_wait(&full[s], (kb / S) & 1);
        mbar_wait(&a2_ready[c], 0);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(smem + s * p.stage_bytes);
        const uint64_t ad = umma_desc_sw128(a_addr), bd = umma_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_tf32(tmem + p.n_ch, ad + 2 * k, bd + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      if (gdn_kb > 0) tc_commit(&acc_full[1]);
    }
    __syncwarp();
  } else {
    // ===================== epilogue (4 warps = 128 accumulator rows) =====================
    const int q = warp & 3;              // TMEM lane quadrant this warp may touch
    const int row = q * 32 + lane;       // tile pixel: (row / kTW, row % kTW)
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const bool leader = (row == 0);
    uint8_t* bufO = epi_buf;                  // out staging
    uint8_t* bufS = epi_buf + kABytes;        // FWD: scale staging;  BWD: y_prev chunk
    uint8_t* bufC = epi_buf + 2 * kABytes;    // BWD: sc_prev chunk
    uint8_t* bufX = epi_buf + 3 * kABytes;    // acc_from_in: x / g chunk
    const int n_loads = (p.acc_from_in ? 1 : 0) + (bwd ? 2 : 0);
    uint32_t ld_phase = 0;
    bool store_pending = false;

    auto fetch_chunk = [&](int c) {  // TMA-load the global operands of chunk c into staging
      if (n_loads == 0) return;
      named_bar_sync(1, 128);  // everyone finished reading the previous chunk's staging
      if (leader) {
        mbar_arrive_expect_tx(ld_full, n_loads * kABytes);
        if (p.acc_from_in) tma_load_4d(bufX, &p.a_map[0], ld_full, c * 32, j0, i0, img);
        if (bwd) {
          tma_load_4d(bufS, &p.yprev_map, ld_full, c * 32, j0, i0, img);
          tma_load_4d(bufC, &p.scprev_map, ld_full, c * 32, j0, i0, img);
        }
      }
      __syncwarp();
      mbar_wait(ld_full, ld_phase);
      ld_phase ^= 1;
    };
    auto load_acc1 = [&](int c, float* v) {
      if (p.acc_from_in) {
        read_row32(bufX, row, v);
      } else {
        tmem_ld32(t_lane + c * 32, v);
        tmem_ld_wait();
      }
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + c * 32 + j);
      }
    };
    auto store_chunk = [&](int c, const float* o, const float* o2) {  // o2: second output (scale) or null
      if (store_pending) {
        if (leader) tma_store_wait_read0();
        __syncwarp();
        named_bar_sync(1, 128);
      }
      if (p.round_out) {
        float r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = round_tf32(o[j]);
        write_row32(bufO, row, r);
      } else {
        write_row32(bufO, row, o);
      }
      if (o2 != nullptr) write_row32(bufS, row, o2);
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        tma_store_4d(&p.out_map, bufO, c * 32, j0, i0, img);
        if (o2 != nullptr) tma_store_4d(&p.sc_map, bufS, c * 32, j0, i0, img);
        tma_store_commit();
      }
      __syncwarp();
      store_pending = true;
    };

    if (main_kb > 0) {
      mbar_wait(&acc_full[0], 0);
      tc_fence_after_sync();
    }

    if (!gdn) {
      for (int c = 0; c < p.n_chunks; ++c) {
        float v[32];
        fetch_chunk(c);
        load_acc1(c, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
        store_chunk(c, v, nullptr);
      }
    } else {
      // ---- pass 1: build the A operand of the normalisation GEMM ----
      for (int c = 0; c < p.n_chunks; ++c) {
        float v[32], a2[32];
        fetch_chunk(c);
        load_acc1(c, v);
        if (!bwd) {
#pragma unroll
          for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * v[j]);
        } else {
          float yv[32], sv[32];
          read_row32(bufS, row, yv);
          read_row32(bufC, row, sv);
          if (p.epi == ICADV_EPI_GDN_BWD) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * yv[j] * sv[j] * sv[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              // out-of-image rows are zero-filled (sc = 0): keep them finite
              float s2 = sv[j] * sv[j];
              a2[j] = s2 > 0.f ? round_tf32(v[j] * yv[j] / s2) : 0.f;
            }
          }
        }
        const int kb = main_kb + c, s = kb % S;
        mbar_wait(&empty[s], ((kb / S) & 1) ^ 1);   // MMAs that last read this slot are done
        write_row32(smem + s * p.stage_bytes, row, a2);
        fence_proxy_async_smem();
        mbar_arrive(&a2_ready[c]);
      }
      // ---- pass 2: normalise ----
      mbar_wait(&acc_full[1], 0);
      tc_fence_after_sync();
      for (int c = 0; c < p.n_chunks; ++c) {
        float v[32], w[32];
        fetch_chunk(c);
        load_acc1(c, v);
        tmem_ld32(t_lane + p.n_ch + c * 32, w);
        tmem_ld_wait();
        if (!bwd) {
          float sc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float n = __ldg(p.beta + c * 32 + j) + w[j];
            sc[j] = (p.epi == ICADV_EPI_GDN_FWD) ? rsqrtf(n) : sqrtf(n);
            v[j] *= sc[j];
          }
          store_chunk(c, v, sc);
        } else {
          float yv[32], sv[32];
          read_row32(bufS, row, yv);
          read_row32(bufC, row, sv);
          const float sign = (p.epi == ICADV_EPI_GDN_BWD) ? -1.f : 1.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float xs = sv[j] > 0.f ? yv[j] / sv[j] : 0.f;   // x = y / sc
            v[j] = v[j] * sv[j] + sign * xs * w[j];
          }
          store_chunk(c, v, nullptr);
        }
      }
    }
    if (leader) tma_store_wait0();
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 4-D channels-last view: dims (C, W, H, N) with arbitrary pixel strides (in floats); box [32, TW, TH, 1]
static int encode_nhwc(CUtensorMap* m, const float* base, int C, int W, int H, int N, int64_t sw, int64_t sh,
                       int64_t sn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)kTW, (cuuint32_t)kTH, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d C=%d W=%d H=%d N=%d) failed: %d", C, W, H, N, (int)r);
    return ICADV_ECUDA;
  }
  return ICADV_OK;
}

// 2-D K-major matrix [rows][cols]; box [32, box_rows]
static int encode_mat(CUtensorMap* m, const float* base, int cols, int rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d cols=%d rows=%d) failed: %d", cols, rows, (int)r);
    return ICADV_ECUDA;
  }
  return ICADV_OK;
}

// parity plane (a,b) of a dense [N,H,W,C] tensor with spatial step s: pixels (s*i+a, s*j+b)
static int encode_plane(CUtensorMap* m, const float* base, int C, int W, int H, int N, int s, int a, int b) {
  const int Wp = (W - b + s - 1) / s, Hp = (H - a + s - 1) / s;
  if (Wp <= 0 || Hp <= 0) { set_error("empty parity plane"); return ICADV_EINVAL; }
  return encode_nhwc(m, base + ((int64_t)a * W + b) * C, C, Wp, Hp, N, (int64_t)s * C, (int64_t)s * W * C,
                     (int64_t)H * W * C);
}

}  // namespace icadv

using namespace icadv;

struct icadv_conv_plan {
  int n_launch;
  TcParams params[4];
  dim3 grid[4];
  int smem_bytes[4];
};

static int tc_supported(const icadv_conv_desc* d, bool report) {
#define TC_REQ(cond, msg)                     \
  do {                                        \
    if (!(cond)) {                            \
      if (report) set_error("conv_tc: " msg); \
      return 0;                               \
    }                                         \
  } while (0)
  TC_REQ(d != nullptr, "null descriptor");
  TC_REQ(d->k_ch % 32 == 0 && d->k_ch >= 32, "k_ch must be a multiple of 32");
  TC_REQ(d->n_ch % 32 == 0 && d->n_ch >= 32 && d->n_ch <= 256, "n_ch must be a multiple of 32 in [32,256]");
  TC_REQ(d->epi >= ICADV_EPI_LINEAR && d->epi <= ICADV_EPI_IGDN_BWD, "bad epilogue");
  if (d->epi != ICADV_EPI_LINEAR) TC_REQ(2 * d->n_ch <= 512, "GDN epilogue needs 2*n_ch <= 512 TMEM columns");
  if (d->acc_from_in) TC_REQ(d->k_ch == d->n_ch && d->ksize == 1 && d->stride == 1, "acc_from_in needs a 1x1 identity geometry");
  TC_REQ(d->ksize * d->ksize <= kMaxTaps, "too many taps");
#undef TC_REQ
  return 1;
}

extern "C" {

int icadv_conv_tc_supported(const icadv_conv_desc* d) { return tc_supported(d, false); }

int icadv_conv_plan_create(const icadv_conv_desc* d, icadv_conv_plan** out_plan) {
  ICADV_REQUIRE(out_plan != nullptr, "null plan pointer");
  *out_plan = nullptr;
  if (!tc_supported(d, true)) return ICADV_EINVAL;
  ICADV_REQUIRE(d->in && d->out && (d->acc_from_in || d->wpack), "null tensor pointer");
  const bool gdn = d->epi != ICADV_EPI_LINEAR;
  const bool bwd = d->epi == ICADV_EPI_GDN_BWD || d->epi == ICADV_EPI_IGDN_BWD;
  if (gdn) ICADV_REQUIRE(d->gmat != nullptr, "GDN epilogue needs gmat");
  if (gdn && !bwd) ICADV_REQUIRE(d->beta && d->out_scale, "GDN forward needs beta and out_scale");
  if (bwd) ICADV_REQUIRE(d->y_prev && d->sc_prev, "GDN backward needs y_prev and sc_prev");
  Geometry g;
  int rc = make_geometry(d, &g);
  if (rc) return rc;
  int dev_rc = icadv_check_device();
  if (dev_rc) return dev_rc;

  icadv_conv_plan* plan = new (std::nothrow) icadv_conv_plan();
  if (!plan) { set_error("out of host memory"); return ICADV_ENOMEM; }
  plan->n_launch = g.n_launch;
  const int K = d->k_ch, N = d->n_ch, taps_total = d->ksize * d->ksize;
  const int s = d->stride;
  for (int l = 0; l < g.n_launch; ++l) {
    TcParams& p = plan->params[l];
    memset(&p, 0, sizeof(p));
    // ---- input maps
    if (d->form == ICADV_FORM_SCONV && s == 2) {
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          if (d->in_h <= a || d->in_w <= b) { p.a_map[a * 2 + b] = p.a_map[0]; continue; }
          rc = encode_plane(&p.a_map[a * 2 + b], d->in, K, d->in_w, d->in_h, d->n_img, 2, a, b);
          if (rc) { delete plan; return rc; }
        }
    } else {
      rc = encode_nhwc(&p.a_map[0], d->in, K, d->in_w, d->in_h, d->n_img, K, (int64_t)d->in_w * K,
                       (int64_t)d->in_h * d->in_w * K);
      if (rc) { delete plan; return rc; }
      p.a_map[1] = p.a_map[2] = p.a_map[3] = p.a_map[0];
    }
    // ---- output-side maps (dense for SCONV, parity plane for TCONV)
    auto out_side = [&](CUtensorMap* m, const float* base) -> int {
      if (d->form == ICADV_FORM_TCONV && s == 2)
        return encode_plane(m, base, N, g.out_w, g.out_h, d->n_img, 2, g.out_a[l], g.out_b[l]);
      return encode_nhwc(m, base, N, g.out_w, g.out_h, d->n_img, N, (int64_t)g.out_w * N,
                         (int64_t)g.out_h * g.out_w * N);
    };
    rc = out_side(&p.out_map, d->out);
    if (!rc && gdn && !bwd) rc = out_side(&p.sc_map, d->out_scale);
    if (!rc && bwd) rc = out_side(&p.yprev_map, d->y_prev);
    if (!rc && bwd) rc = out_side(&p.scprev_map, d->sc_prev);
    if (!rc && !d->acc_from_in) rc = encode_mat(&p.w_map, d->wpack, K, taps_total * N, N);
    if (!rc && gdn) rc = encode_mat(&p.g_map, d->gmat, N, N, N);
    if (rc) { delete plan; return rc; }
    if (!(gdn && !bwd)) p.sc_map = p.out_map;
    if (!bwd) { p.yprev_map = p.out_map; p.scprev_map = p.out_map; }
    if (d->acc_from_in) p.w_map = p.out_map;
    if (!gdn) p.g_map = p.out_map;

    p.num_taps = g.n_taps[l];
    for (int t = 0; t < p.num_taps; ++t) p.taps[t] = g.taps[l][t];
    p.k_chunks = K / 32; p.n_ch = N; p.n_chunks = N / 32;
    p.tiles_x = (g.tile_w + kTW - 1) / kTW; p.tiles_y = (g.tile_h + kTH - 1) / kTH;
    p.epi = d->epi; p.act = d->act; p.acc_from_in = d->acc_from_in; p.round_out = d->round_out_tf32;
    p.stage_bytes = kABytes + N * 128;
    int avail = kSmemLimit - 1024 - kEpiBufs * kABytes - kBarBytes;
    p.num_stages = avail / p.stage_bytes;
    if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
    if (p.num_stages < 2) { delete plan; set_error("conv_tc: not enough shared memory for n_ch=%d", N); return ICADV_EINVAL; }
    int cols = gdn ? 2 * N : N, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    p.bias = d->bias; p.beta = d->beta; p.active = d->active; p.n_active = d->n_active;
    plan->grid[l] = dim3(p.tiles_x * p.tiles_y, d->n_img, 1);
    plan->smem_bytes[l] = 1024 + p.num_stages * p.stage_bytes + kEpiBufs * kABytes + kBarBytes;
    if (d->n_img > 65535) { delete plan; set_error("conv_tc: n_img too large"); return ICADV_EINVAL; }
  }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  });
  if (attr_err != cudaSuccess) {
    delete plan;
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    return ICADV_ECUDA;
  }
  *out_plan = plan;
  return ICADV_OK;
}

int icadv_conv_plan_launch(const icadv_conv_plan* plan, icadv_stream_t stream) {
  ICADV_REQUIRE(plan != nullptr, "null plan");
  for (int l = 0; l < plan->n_launch; ++l) {
    conv_tc_kernel<<<plan->grid[l], 192, plan->smem_bytes[l], as_stream(stream)>>>(plan->params[l]);
    ICADV_CUDA_TRY(cudaGetLastError());
  }
  return ICADV_OK;
}

int icadv_conv_plan_destroy(icadv_conv_plan* plan) {
  delete plan;
  return ICADV_OK;
}

int icadv_conv_tc(const icadv_conv_desc* d, icadv_stream_t stream) {
  icadv_conv_plan* plan = nullptr;
  int rc = icadv_conv_plan_create(d, &plan);
  if (rc) return rc;
  rc = icadv_conv_plan_launch(plan, stream);
  icadv_conv_plan_destroy(plan);
  return rc;
}

}  // extern "C"
