// tcgen05 / TMEM / TMA implicit-GEMM for the codec's conv and transposed-conv stacks, with GDN / IGDN
// fused into the epilogue in both directions.  Replaces nn.Conv2d / nn.ConvTranspose2d + compressai GDN
// under net.g_a / net.g_s (reference: anchors/utils.py:112-130, utils/ops.py:58-97, attack_rd.py:344-349,547).
//
// One CTA = one 16x8 tile (16 rows, 8 columns) of "tile-space" pixels (M = 128 rows) x all n_ch output
// channels (N <= 256).
//   D[128, N] = sum over taps t, 32-channel chunks kc:  A_t,kc[128, 32] * W_t,kc[N, 32]^T       (kind::tf32)
// The input is loaded ONCE per (parity plane, 32-channel chunk) as a halo patch: a TMA box
// [32 ch, 8 + halo w, 16 + halo h, 1 img] of the channels-last input (out-of-bounds = zero padding; stride-2
// convs address the input through four parity-plane tensor maps so every tap is a unit shift).  The patch
// lands as consecutive 128-byte pixel rows (SWIZZLE_128B); because a tile row is exactly one 8-row swizzle
// group, the A operand of tap (dy, dx) is the SAME patch read through a shared-memory descriptor whose start
// address is shifted by (dy * patch_w + dx) rows and whose group stride (SBO) is patch_w * 128 bytes -- the
// tensor core applies the swizzle on absolute address bits, so unaligned starts and strides are exact
// (profiles/r1_umma_descriptor_shift_probe.txt).  That cuts the L2 -> SM operand traffic of a 5x5/2 layer from
// 25 boxes to 4 patches per chunk; the weights W_t,kc stream through their own ring as TMA boxes [32, N].
// The accumulator lives in TMEM.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
// GDN epilogues append n_ch/32 extra K-blocks to the same smem ring: the A operand of those blocks
// (x^2 forward, g*y*sc^2 backward) is written by the epilogue warps, the B operand (gamma / gamma^T)
// arrives by TMA, and the product accumulates into a second TMEM region.
#include <stdlib.h>

#include <mutex>

#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

constexpr int kTH = 16, kTW = 8, kTileM = 128;   // kTW must be 8: one tile row = one 8-row swizzle group
constexpr int kABytes = kTileM * 128;   // one [128 x 32 fp32] operand / staging tile
constexpr int kMaxStages = 8;           // weight ring
constexpr int kMaxPatch = 8;            // patch ring
constexpr int kMaxGroups = 8;           // patches per 32-channel chunk (parity planes; kernel rows of the RGB form)
constexpr int kSmemLimit = 232448;      // 227 KB per CTA
constexpr int kSmemTwoCta = 115712;     // (228 KB per SM) / 2 - 1 KB system reservation per CTA
constexpr int kBarBlock = 448;          // up to 49 mbarriers + the TMEM base; followed by per-channel bias / beta copies (2 * n_ch floats)
constexpr int kEpiCol2im = 5;           // internal epilogue code: narrow-output transposed conv via col2im
constexpr int kZStride = 77;            // floats per row of the col2im staging tile (odd: conflict-free)
constexpr int kZStride4 = 100;          // persistent col2im, 3 output channels: rows of 25 taps x float4 (stride = 4 mod 32 words:
                                        // quarter-warp 16-byte accesses of consecutive rows hit distinct bank groups)

// One halo patch of the input: everything the taps [tap_begin, tap_end) read for one 32-channel chunk.
struct Group {
  int16_t map;        // index into a_map (parity plane)
  int16_t plane5;     // rank-5 (RGB first-layer) form: coordinate 3 of the box
  int16_t dy0, dx0;   // patch origin relative to the tile origin
  int16_t tap_begin, tap_end;
  int32_t sbo_bytes;  // stride between tile rows inside the patch = patch_w * 128
  int32_t bytes;      // patch_w * patch_h * 128
};

struct TcParams {
  CUtensorMap a_map[4];
  CUtensorMap w_map, g_map, out_map, sc_map, yprev_map, scprev_map;
  Group groups[kMaxGroups];
  int32_t tap_aoff[kMaxTaps + 1];   // byte offset of the tap's first row inside its patch (one slack entry: read ahead)
  int16_t tap_wtap[kMaxTaps];   // row block of the packed weight
  int num_groups, num_taps, k_chunks, n_ch, n_chunks;   // n_ch = MMA N (one N-tile); n_chunks = n_ch / 32
  int num_patch, patch_bytes;               // patch ring: slots and bytes per slot (multiple of 1024)
  int n_total;                              // all output channels (> n_ch when blockIdx.z tiles N; linear epilogue only)
  int tiles_x, tiles_y, tile_step_y, tile_step_x, tile_off;
  int epi, act, acc_from_in, round_out, a_rank5;
  int num_stages, tmem_cols, ld_bufs;       // num_stages: weight ring (n_ch * 128 bytes per stage)
  int ld_alias;                             // BWD: the saved-chunk staging pair is the LAST 32 KB of the weight ring (idle
                                            // once the main loop is done); gamma then cycles through stages [0, S-2)
  int t_h, t_w, o_h, o_w, o_s, o_a, o_b;    // tile-space extent; output geometry: pixel (o_s*i + o_a, o_s*j + o_b)
  const float* yprev; const float* scprev; const float* xin;   // epilogue operands read with plain loads
  int c2i_in_h, c2i_in_w, c2i_nch;         // col2im: input extent, real output channels
  float* c2i_out;                           // col2im: dense [n_img, 2*in_h, 2*in_w, c2i_nch]
  const float* bias;
  const float* beta;
  const int* active;
  const int* n_active;
  int bwd_lookahead;   // backward epilogues: L2-prefetch the saved tensors of the CTA this many launch slots ahead (0 = off)
  int dbg_flags;    // developer switches (env ICADV_TC_DBG): 1 = no prefetch of saved y/scale, 2 = single-buffered stores,
                    // 16 = weight boxes not loaded (timing experiment: what the weight stream costs; results are garbage)
  long long* dbg;   // optional per-CTA phase timestamps (16 slots per CTA), developer profiling only
};

#define TC_STAMP(slot)                                                          \
  do {                                                                          \
    if (p.dbg != nullptr) p.dbg[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 32 + (slot)] = clock64(); \
  } while (0)

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

constexpr int kThreads = 192;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue

// 128-byte row chunk (32 channels) of one pixel straight from global memory; zeros when the pixel is outside
__device__ __forceinline__ void ldg_row32(const float* __restrict__ ptr, bool valid, float* v) {
  if (valid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(ptr) + j);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
}

// EPI: ICADV_EPI_* or kEpiCol2im.  FROM_IN: accumulator comes from global memory instead of a contraction.
// Up to two CTAs share an SM (3 x 32 KB stages, 256 TMEM columns each): one CTA's prologue / epilogue / store
// drain overlaps the other's main loop.
template <int EPI, bool FROM_IN>
__global__ void __launch_bounds__(kThreads, 2) conv_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr bool gdn = EPI >= ICADV_EPI_GDN_FWD && EPI <= ICADV_EPI_IGDN_BWD;
  constexpr bool bwd = EPI == ICADV_EPI_GDN_BWD || EPI == ICADV_EPI_IGDN_BWD;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = blockIdx.y;
  if (p.n_active != nullptr && slot >= *p.n_active) return;  // whole CTA leaves together
  const int img = p.active != nullptr ? p.active[slot] : slot;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x % p.tiles_x;
  const int i0 = ty * p.tile_step_y + p.tile_off, j0 = tx * p.tile_step_x + p.tile_off;
  const int n_off = blockIdx.z * p.n_ch;    // first output channel of this CTA's N-tile

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int P = p.num_patch, S = p.num_stages;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_ch) * 128u;
  uint8_t* wring = smem + P * p.patch_bytes;               // weight ring behind the patch ring
  uint8_t* ring_end = wring + S * b_bytes;
  uint8_t* ld_buf = p.ld_alias ? ring_end - 2 * kABytes : ring_end;   // BWD: y_prev / sc_prev chunk staging (2 x 16 KB)
  uint64_t* full = reinterpret_cast<uint64_t*>(ring_end + p.ld_bufs * kABytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* pfull = empty + kMaxStages;      // [kMaxPatch]
  uint64_t* pempty = pfull + kMaxPatch;      // [kMaxPatch]
  uint64_t* acc_full = pempty + kMaxPatch;   // [2]
  uint64_t* a2_ready = acc_full + 2;         // [8]
  uint64_t* ld_full = a2_ready + 8;          // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ld_full + 1);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + kBarBlock);   // [n_ch]
  float* sbeta = sbias + p.n_ch;                                                            // [n_ch]

  if (threadIdx.x == 0) {
    TC_STAMP(0);
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < P; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    for (int c = 0; c < 8; ++c) mbar_init(&a2_ready[c], 128);
    mbar_init(ld_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, p.tmem_cols); tmem_relinquish(); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.out_map);
    if (gdn) tma_prefetch_desc(&p.g_map);
    if (EPI == ICADV_EPI_GDN_FWD || EPI == ICADV_EPI_IGDN_FWD) tma_prefetch_desc(&p.sc_map);
  }
  if (threadIdx.x >= 64) {   // parameters the epilogue reads per channel
    for (int i = threadIdx.x - 64; i < p.n_ch; i += kThreads - 64) {
      const bool has_bias = p.bias != nullptr && (EPI != kEpiCol2im || i < p.c2i_nch);   // col2im: n_ch is the padded 96
      sbias[i] = has_bias ? __ldg(p.bias + n_off + i) : 0.f;
      if (EPI == ICADV_EPI_GDN_FWD || EPI == ICADV_EPI_IGDN_FWD) sbeta[i] = __ldg(p.beta + i);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) TC_STAMP(1);

  const int main_pc = FROM_IN ? 0 : p.k_chunks * p.num_groups;   // patches of the main loop
  const int main_kb = FROM_IN ? 0 : p.k_chunks * p.num_taps;     // weight boxes (= K-blocks) of the main loop
  const int gdn_kb = gdn ? p.n_chunks : 0;

  // Both single-thread roles below are latency chains (one lane, no ILP): ring positions are carried as
  // (slot, phase) counters -- no division -- and everything that does not depend on a barrier is computed
  // before waiting on it, so the path barrier -> issue stays a handful of instructions.
  if (warp == 0) {
    // ===================== TMA producer =====================
    {   // all 32 lanes run the loop (uniform state); one elected lane issues
      if constexpr (bwd) {
        // the saved y / scale of this tile come from HBM (written by the forward pass long ago): pull them into L2
        // now, so the epilogue's chunk loads are L2 hits instead of eight serial DRAM round trips
        if ((p.dbg_flags & 4) && lane == 0)   // measured: no gain (the chunk round trip, not DRAM, is the latency) -- kept as a switch
          for (int c = 0; c < p.n_chunks; ++c) {
            tma_prefetch_4d(&p.yprev_map, c * 32, j0, i0, img);
            tma_prefetch_4d(&p.scprev_map, c * 32, j0, i0, img);
          }
        // Lookahead: the saved y / scale of the tile that will run `bwd_lookahead` CTAs after this one (about two waves
        // later) are pulled into L2 now, so that tile's pass-1 chunk loads are L2 hits (1.35 us per chunk) instead of DRAM
        // round trips (2.3 us per chunk; profiles/r1_tile_timeline_bwd_v10.txt)
        if (p.bwd_lookahead > 0 && lane == 0) {
          const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + p.bwd_lookahead;
          const int slot2 = (int)(lin / gridDim.x), tile2 = (int)(lin % gridDim.x);
          const int n_slots = p.n_active != nullptr ? *p.n_active : (int)gridDim.y;
          if (slot2 < n_slots) {
            const int img2 = p.active != nullptr ? p.active[slot2] : slot2;
            const int i2 = (tile2 / p.tiles_x) * p.tile_step_y + p.tile_off, j2 = (tile2 % p.tiles_x) * p.tile_step_x + p.tile_off;
            for (int c = 0; c < p.n_chunks; ++c) {
              tma_prefetch_4d(&p.yprev_map, c * 32, j2, i2, img2);
              tma_prefetch_4d(&p.scprev_map, c * 32, j2, i2, img2);
            }
          }
        }
      }
      int s = 0, ps = 0;
      uint32_t s_par = 1, p_par = 1;          // parity to wait for on the "empty" barriers (first pass: free)
      long long t_wait_e = 0;
      const long long t_prod0 = clock64();
      uint8_t* wdst = wring;
      uint8_t* pdst = smem;
      const int n_row0 = n_off;
      for (int kc = 0; kc < (FROM_IN ? 0 : p.k_chunks); ++kc) {
        const int c0 = kc * 32;
        for (int g = 0; g < p.num_groups; ++g) {
          const Group gr = p.groups[g];
          const int cx = j0 + gr.dx0, cy = i0 + gr.dy0;
          mbar_wait(&pempty[ps], p_par);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&pfull[ps], gr.bytes);
            if (p.a_rank5)
              tma_load_5d(pdst, &p.a_map[0], &pfull[ps], 0, cx, cy, gr.plane5, img);
            else
              tma_load_4d(pdst, &p.a_map[gr.map], &pfull[ps], c0, cx, cy, img);
          }
          if (++ps == P) { ps = 0; p_par ^= 1; pdst = smem; } else { pdst += p.patch_bytes; }
          int wrow = p.tap_wtap[gr.tap_begin] * p.n_total + n_row0;
          for (int t = gr.tap_begin; t < gr.tap_end; ++t) {
            const int wrow_next = (t + 1 < gr.tap_end ? p.tap_wtap[t + 1] : 0) * p.n_total + n_row0;
            if (p.dbg != nullptr) { const long long t0 = clock64(); mbar_wait(&empty[s], s_par); t_wait_e += clock64() - t0; }
            else mbar_wait(&empty[s], s_par);
            if (elect_one_sync()) {
              if (p.dbg_flags & 16) {   // developer experiment: the stage is handed over without loading the weights
                mbar_arrive(&full[s]);
              } else {
                mbar_arrive_expect_tx(&full[s], b_bytes);
                tma_load_2d(wdst, &p.w_map, &full[s], c0, wrow);
              }
            }
            wrow = wrow_next;
            if (++s == S) { s = 0; s_par ^= 1; wdst = wring; } else { wdst += b_bytes; }
          }
        }
      }
      if (p.dbg != nullptr && lane == 0) {
        long long* q = p.dbg + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 32;
        q[13] = q[0] + t_wait_e; q[14] = q[0] + (clock64() - t_prod0);   // producer: blocked on empty / whole main loop
      }
      if (p.ld_alias) {
        // gamma chunks cycle through stages [0, S-2) only; slot j has been used ceil((main_kb - j) / S) times so far
        const int Sg = S - 2;
        for (int c = 0; c < gdn_kb; ++c) {
          const int j = c % Sg;
          const int n = (main_kb > j ? (main_kb - j + S - 1) / S : 0) + c / Sg;   // use index of slot j
          mbar_wait(&empty[j], ((n & 1) ^ 1));
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full[j], b_bytes);
            tma_load_2d(wring + j * b_bytes, &p.g_map, &full[j], c * 32, 0);
          }
        }
      } else
      for (int c = 0; c < gdn_kb; ++c) {
        mbar_wait(&empty[s], s_par);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[s], b_bytes);
          tma_load_2d(wdst, &p.g_map, &full[s], c * 32, 0);
        }
        if (++s == S) { s = 0; s_par ^= 1; wdst = wring; } else { wdst += b_bytes; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {   // ONE thread runs the whole role: no per-K-block reconvergence (see conv_tcp_kernel)
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      // shared-memory descriptors: hi word = SBO>>4 | version 1 (bit 14) | SWIZZLE_128B (bit 30); lo word = addr>>4 | LBO 1
      const uint32_t hi_dense = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t w_lo0 = ((smem_u32(wring) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t p_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_step = b_bytes >> 4, p_step = static_cast<uint32_t>(p.patch_bytes) >> 4;
      int s = 0, ps = 0;
      uint32_t s_par = 0, p_par = 0, w_lo = w_lo0, p_lo = p_lo0, acc = 0;
      long long t_wait_w = 0, t_wait_p = 0;   // developer profiling: cycles blocked on the weight / patch rings (slow paths)
      bool w_ok = false;                      // early probe of full[s]
      TC_STAMP(2);
      for (int kc = 0; kc < (FROM_IN ? 0 : p.k_chunks); ++kc) {
        for (int g = 0; g < p.num_groups; ++g) {
          const uint32_t a_hi = (static_cast<uint32_t>(p.groups[g].sbo_bytes) >> 4) | (1u << 14) | (2u << 29);
          const int t_end = p.groups[g].tap_end;
          int t = p.groups[g].tap_begin;
          uint32_t aoff = static_cast<uint32_t>(p.tap_aoff[t]) >> 4;
          if (!mbar_test_wait(&pfull[ps], p_par)) {
            const long long t0 = clock64();
            mbar_wait(&pfull[ps], p_par);
            t_wait_p += clock64() - t0;
          }
          for (; t < t_end; ++t) {
            // tap (dy, dx): same patch, start shifted by whole pixel rows; tile rows are patch_w * 128 bytes apart
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (p_lo + aoff);
            const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | w_lo;
            if (!w_ok) {
              const long long t0 = clock64();
              mbar_wait(&full[s], s_par);
              t_wait_w += clock64() - t0;
            }
            uint64_t* wdone = &empty[s];
            if (++s == S) { s = 0; s_par ^= 1; w_lo = w_lo0; } else { w_lo += w_step; }
            tc_fence_after_sync();
            tc_mma_tf32(tmem, ad, bd, idesc, acc);
            tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
            tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u);
            tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
            tc_commit(wdone);
            acc = 1u;
            // off the critical path (the MMAs above are draining): next tap's offset, next stage's barrier
            aoff = static_cast<uint32_t>(p.tap_aoff[t + 1]) >> 4;   // tap_aoff has one slack entry
            w_ok = mbar_test_wait(&full[s], s_par);
          }
          tc_commit(&pempty[ps]);
          if (++ps == P) { ps = 0; p_par ^= 1; p_lo = p_lo0; } else { p_lo += p_step; }
        }
      }
      if (main_kb > 0) tc_commit(&acc_full[0]);
      TC_STAMP(3);
      if (p.dbg != nullptr) {
        long long* q = p.dbg + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 32;
        q[11] = q[0] + t_wait_w; q[12] = q[0] + t_wait_p;   // stored relative to slot 0 like the stamps
      }
      for (int c = 0; c < gdn_kb; ++c) {
        if (p.ld_alias) {   // gamma ring = stages [0, S-2): recompute the slot and its use parity (see the producer)
          const int Sg = S - 2;
          s = c % Sg;
          const int n = (main_kb > s ? (main_kb - s + S - 1) / S : 0) + c / Sg;
          s_par = static_cast<uint32_t>(n & 1);
          w_lo = w_lo0 + s * w_step;
        }
        const uint64_t ad = (static_cast<uint64_t>(hi_dense) << 32) | p_lo;
        const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | w_lo;
        mbar_wait(&full[s], s_par);
        mbar_wait(&a2_ready[c], 0);
        tc_fence_after_sync();
        // accumulator from global memory (FROM_IN): TMEM holds the normalisation product only, at column 0
        const uint32_t dn = tmem + (FROM_IN ? 0u : static_cast<uint32_t>(p.n_ch));
        tc_mma_tf32(dn, ad, bd, idesc, c > 0 ? 1u : 0u);
        tc_mma_tf32(dn, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(dn, ad + 4, bd + 4, idesc, 1u);
        tc_mma_tf32(dn, ad + 6, bd + 6, idesc, 1u);
        tc_commit(&empty[s]);
        tc_commit(&pempty[ps]);
        if (!p.ld_alias) { if (++s == S) { s = 0; s_par ^= 1; w_lo = w_lo0; } else { w_lo += w_step; } }
        if (++ps == P) { ps = 0; p_par ^= 1; p_lo = p_lo0; } else { p_lo += p_step; }
      }
      if (gdn_kb > 0) tc_commit(&acc_full[1]);
      TC_STAMP(4);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 4 warps = 128 accumulator rows =====================
    const int q = warp & 3;              // TMEM lane quadrant this warp may touch
    const int row = q * 32 + lane;       // tile pixel: (row / kTW, row % kTW)
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16);
    const bool leader = (row == 0);
    const int nC = p.n_chunks;

    if constexpr (EPI == kEpiCol2im) {
      // ---- Z[128 px][taps*nch] -> smem (the stage ring is idle by now), then gather the transposed-conv
      //      outputs of the tile interior ----
      float* Zs = reinterpret_cast<float*>(smem);
      const int zcols = 25 * p.c2i_nch;
      mbar_wait(&acc_full[0], 0);
      tc_fence_after_sync();
      if (leader) TC_STAMP(5);
      for (int c = 0; c < nC; ++c) {
        float v[32];
        tmem_ld32(t_lane + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c * 32 + j < zcols) Zs[row * kZStride + c * 32 + j] = v[j];
      }
      named_bar_sync(1, 128);
      if (leader) TC_STAMP(6);
      const int OH = 2 * p.c2i_in_h, OW = 2 * p.c2i_in_w, nch = p.c2i_nch;
      constexpr int kIy = kTH - 2, kIx = kTW - 2;   // tile interior (14 x 6 input pixels)
      for (int o = row; o < kIy * kIx * 4; o += 128) {
        const int cell = o >> 2, a = (o >> 1) & 1, b = o & 1;
        const int ti = 1 + cell / kIx, tj = 1 + cell % kIx;
        const int gi = i0 + ti, gj = j0 + tj;
        if (gi >= p.c2i_in_h || gj >= p.c2i_in_w) continue;
        // taps kh = a + 2u, kw = b + 2v (u, v < 3; kh, kw < 5); source pixel (ti + 1 - u, tj + 1 - v)
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 3; ++u) {
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            const int kh = a + 2 * u, kw = b + 2 * v;
            const bool ok = kh < 5 && kw < 5;
            const float* z = Zs + ((ti + 1 - u) * kTW + (tj + 1 - v)) * kZStride + (ok ? (kh * 5 + kw) * nch : 0);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (c < nch) acc[c] += ok ? z[c] : 0.f;
          }
        }
        float* dst = p.c2i_out + (((int64_t)img * OH + 2 * gi + a) * OW + 2 * gj + b) * nch;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch) dst[c] = acc[c] + sbias[c];
      }
      if (leader) TC_STAMP(8);
    } else {
      // global operands of the epilogue (saved y / scale of the matching forward; stand-alone input) are read
      // straight into registers: this thread's pixel, 128 contiguous bytes per 32-channel chunk
      const int gi = i0 + row / kTW, gj = j0 + row % kTW;
      const bool px_ok = gi < p.t_h && gj < p.t_w;
      const int64_t pix = (((int64_t)img * p.o_h + p.o_s * gi + p.o_a) * p.o_w + p.o_s * gj + p.o_b) * p.n_ch;
      // store staging aliases the patch + weight rings (idle once the last MMA has completed): 16 KB regions
      auto stg = [&](int i) -> uint8_t* { return smem + i * kABytes; };
      int stores = 0;
      uint32_t ld_cnt = 0;
      uint8_t* ldY = ld_buf;
      uint8_t* ldS = ld_buf + kABytes;
      // BWD: saved y / scale chunks arrive by TMA (single staging pair; the next chunk is requested as soon as every
      // thread has copied the current one into registers)
      auto fetch = [&](int c) {
        if (leader) {
          mbar_arrive_expect_tx(ld_full, 2 * kABytes);
          tma_load_4d(ldY, &p.yprev_map, ld_full, c * 32, j0, i0, img);
          tma_load_4d(ldS, &p.scprev_map, ld_full, c * 32, j0, i0, img);
        }
        __syncwarp();
      };
      int cur_chunk = 0;
      auto take = [&](float* yv, float* sv, int next) {   // next: chunk to request afterwards, or -1
        if (p.dbg_flags & 1) { named_bar_sync(1, 128); fetch(cur_chunk); }
        mbar_wait(ld_full, ld_cnt & 1);
        ++ld_cnt;
        read_row32(ldY, row, yv);
        read_row32(ldS, row, sv);
        // The staging pair is about to be overwritten through the async proxy: every shared-memory load above must
        // have RETURNED (not merely issued) before this thread arrives.  One word of each 16-byte load feeds the
        // barrier instruction, so the scoreboard holds the arrival until the data is in registers.
        uint32_t dep = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) dep |= __float_as_uint(yv[4 * j]) | __float_as_uint(sv[4 * j]);
        fence_proxy_async_smem();   // generic-proxy reads above vs the async-proxy (TMA) overwrite that follows
        asm volatile("{\n.reg .b32 t;\nand.b32 t, %2, 0;\nadd.u32 t, t, %0;\nbar.sync t, %1;\n}" ::"r"(1u), "r"(128u), "r"(dep)
                     : "memory");
        if (p.dbg_flags & 1) { cur_chunk = next; } else if (next >= 0) fetch(next);
      };
      if constexpr (bwd) { if (!(p.dbg_flags & 1) && !p.ld_alias) fetch(0); }   // overlaps the main loop

      auto load_acc1 = [&](int c, float* v) {
        if constexpr (FROM_IN) {
          ldg_row32(p.xin + pix + c * 32, px_ok, v);
        } else {
          tmem_ld32(t_lane + c * 32, v);
          tmem_ld_wait();
        }
        const float4* b4 = reinterpret_cast<const float4*>(sbias + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];
          v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
      };
      auto store_chunk = [&](int c, const float* o, const float* o2) {  // o2: second output (scale) or null
        const int sb = stores & 1;   // double-buffered staging: wait only for the store issued two chunks ago
        uint8_t* bufO = stg(2 * sb);
        uint8_t* bufS = stg(2 * sb + 1);
        if (stores >= 2 || ((p.dbg_flags & 2) && stores >= 1)) {
          if (leader) { if (p.dbg_flags & 2) tma_store_wait_read0(); else tma_store_wait_read1(); }
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
          tma_store_4d(&p.out_map, bufO, n_off + c * 32, j0, i0, img);
          if (o2 != nullptr) tma_store_4d(&p.sc_map, bufS, c * 32, j0, i0, img);
          tma_store_commit();
        }
        __syncwarp();
        ++stores;
      };

      if (main_kb > 0) {
        mbar_wait(&acc_full[0], 0);
        tc_fence_after_sync();
      }
      if constexpr (bwd) { if (!(p.dbg_flags & 1) && p.ld_alias) fetch(0); }   // the weight stages it lands in are idle now
      if (leader) TC_STAMP(5);

      if constexpr (!gdn) {
        for (int c = 0; c < nC; ++c) {
          float v[32];
          load_acc1(c, v);
          if (p.act == ICADV_ACT_RELU) {          // uniform branch outside the element loop (no per-element jump table)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.act == ICADV_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
          } else if (p.act == ICADV_ACT_ABS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fabsf(v[j]);
          }
          store_chunk(c, v, nullptr);
        }
      } else {
        // ---- pass 1: build the A operand of the normalisation GEMM (patch ring, continuing after the main loop) ----
        int e_ps = main_pc % P;
        uint32_t e_par = ((main_pc / P) & 1) ^ 1;
        for (int c = 0; c < nC; ++c) {
          float v[32], a2[32];
          load_acc1(c, v);
          if constexpr (!bwd) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * v[j]);
          } else {
            float yv[32], sv[32];
            take(yv, sv, c + 1 < nC ? c + 1 : 0);   // after the last chunk: chunk 0 again, for pass 2
            if constexpr (EPI == ICADV_EPI_GDN_BWD) {
#pragma unroll
              for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * yv[j] * sv[j] * sv[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                // pixels outside the image read as zeros (sc = 0): keep them finite
                const float s2 = sv[j] * sv[j];
                a2[j] = s2 > 0.f ? round_tf32(fast_div(v[j] * yv[j], s2)) : 0.f;
              }
            }
          }
          mbar_wait(&pempty[e_ps], e_par);   // MMAs that last read this patch slot are done
          write_row32(smem + e_ps * p.patch_bytes, row, a2);
          if (++e_ps == P) { e_ps = 0; e_par ^= 1; }
          fence_proxy_async_smem();
          mbar_arrive(&a2_ready[c]);
        }
        if (leader) TC_STAMP(6);
        // ---- pass 2: normalise ----
        mbar_wait(&acc_full[1], 0);
        tc_fence_after_sync();
        if (leader) TC_STAMP(7);
        for (int c = 0; c < nC; ++c) {
          float v[32], w[32];
          load_acc1(c, v);
          tmem_ld32(t_lane + (FROM_IN ? 0 : p.n_ch) + c * 32, w);
          tmem_ld_wait();
          if constexpr (!bwd) {
            float sc[32];
            const float4* be4 = reinterpret_cast<const float4*>(sbeta + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 be = be4[j];
              const float nn[4] = {be.x + w[4 * j], be.y + w[4 * j + 1], be.z + w[4 * j + 2], be.w + w[4 * j + 3]};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float r = rsqrtf(nn[t]);
                sc[4 * j + t] = (EPI == ICADV_EPI_GDN_FWD) ? r : nn[t] * r;   // sqrt(n) = n * rsqrt(n)
                v[4 * j + t] *= sc[4 * j + t];
              }
            }
            store_chunk(c, v, sc);
          } else {
            float yv[32], sv[32];
            take(yv, sv, c + 1 < nC ? c + 1 : -1);
            constexpr float sign = (EPI == ICADV_EPI_GDN_BWD) ? -1.f : 1.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float xs = sv[j] > 0.f ? fast_div(yv[j], sv[j]) : 0.f;   // x = y / sc
              v[j] = v[j] * sv[j] + sign * xs * w[j];
            }
            store_chunk(c, v, nullptr);
          }
        }
      }
      if (leader) TC_STAMP(8);
      if (leader) tma_store_wait0();
      if (leader) TC_STAMP(9);
      __syncwarp();
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
  if (threadIdx.x == 0) TC_STAMP(10);
}

// =====================================================================================================
// Persistent variant: ONE CTA per SM walks a static list of work items (tile x parity class x image).
//   * the accumulator is double-buffered in TMEM (2 x 256 columns: [acc N | norm N]), so the MMA issuer starts the
//     next item's main loop while the previous item's epilogue is still reading;
//   * two epilogue warpgroups (warps 2-5 / 6-9), each with its own 2 x 16 KB staging, share every item (half of the
//     32-channel chunks each; col2im: alternate items);
//   * the four output-parity classes of a stride-2 transposed conv are items of ONE launch (rotated so that every CTA
//     sees all classes: they differ in length 9/6/6/4 taps);
//   * gamma (the B operand of the normalisation GEMM) is loaded once per CTA and stays resident;
//   * normalisation K-blocks are issued by the MMA thread between main-loop K-blocks of the NEXT item, as soon as the
//     epilogue has written their A operand (polled with mbarrier.test_wait, which never suspends);
//   * backward epilogues read the saved y / scale rows with plain 128-byte-per-thread loads (L2-prefetched when the
//     item starts) instead of a TMA staging pair: that shared memory now holds the deeper rings.
// Reference semantics are those of conv_tc_kernel above (anchors/utils.py:112-130, utils/ops.py:58-97).
// =====================================================================================================
constexpr int kPThreads = 384;   // warp 0 weight producer, 1 main MMA issuer, 2-9 two epilogue groups, 10 normalisation MMA issuer, 11 patch producer
constexpr int kPNormWarp = 10, kPPatchWarp = 11;

struct TcpClass {
  int16_t g_begin, g_end;   // groups of this class
  int16_t o_a, o_b;         // output parity (TCONV), else 0
};

struct TcpParams {
  CUtensorMap a_map[4];
  CUtensorMap w_map, g_map;
  CUtensorMap out_map[4], sc_map[4];   // per class; backward epilogues: sc_map = saved scale, yprev_map = saved y
  CUtensorMap yprev_map[4];
  Group groups[kMaxGroups];
  int32_t tap_aoff[kMaxTaps + 2];   // slack entries: the issuer reads up to tap t + 2 ahead
  int16_t tap_wtap[kMaxTaps + 1];
  TcpClass cls[4];
  int n_class, k_chunks, n_ch, n_chunks, n_total;
  int num_patch, patch_bytes, num_stages;
  int grp_bytes;                              // per epilogue group: 2 x 16 KB slots, or the col2im scatter tile
  int tiles_x, tiles_y, n_img;
  int tile_step_y, tile_step_x, tile_off;    // col2im: tiles overlap by a 1-pixel halo
  int act, round_out, a_rank5;
  int t_h, t_w, o_h, o_w, o_s;
  int c2i_in_h, c2i_in_w, c2i_nch;
  float* c2i_out;
  const float* yprev; const float* scprev;
  const float* bias; const float* beta;
  const int* active; const int* n_active;
  long long* dbg;   // developer profiling: 16 cycle counters per CTA (see scripts/persistent_timeline.py)
  int ys_slots;     // streaming backward kernel: slots of the saved-tensor ring (each [y chunk | scale chunk] = 32 KB)
  float* out;       // streaming backward kernel: dense output base (plain 128-byte stores per pixel and chunk)
  int w_resident;   // col2im: the k_chunks weight boxes of the 1x1 GEMM are loaded once and stay in shared memory
  int dbg_skip_w;   // developer experiment (ICADV_TC_DBG bit 16): weight boxes are not loaded (timing only, results garbage)
};

struct TcpItem { int img, i0, j0, cls; };

__device__ __forceinline__ TcpItem tcp_decode(const TcpParams& p, int item) {
  // class rotates with the tile index so that a CTA's items (stride gridDim.x) cycle through all classes, while the
  // classes of one tile are adjacent items (their input patch is shared through L2)
  const int tile_all = item / p.n_class;
  TcpItem it;
  it.cls = (item + tile_all) % p.n_class;
  const int tiles = p.tiles_x * p.tiles_y;
  const int slot = tile_all / tiles, t = tile_all - slot * tiles;
  it.img = p.active != nullptr ? p.active[slot] : slot;
  it.i0 = (t / p.tiles_x) * p.tile_step_y + p.tile_off;
  it.j0 = (t % p.tiles_x) * p.tile_step_x + p.tile_off;
  return it;
}

template <int EPI>
__global__ void __launch_bounds__(kPThreads, 1) conv_tcp_kernel(const __grid_constant__ TcpParams p) {
  constexpr bool gdn = EPI >= ICADV_EPI_GDN_FWD && EPI <= ICADV_EPI_IGDN_BWD;
  constexpr bool bwd = EPI == ICADV_EPI_GDN_BWD || EPI == ICADV_EPI_IGDN_BWD;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int P = p.num_patch, S = p.num_stages, nC = p.n_chunks;
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_ch) * 128u;
  uint8_t* wring = smem + P * p.patch_bytes;
  uint8_t* gmat = wring + S * b_bytes;                              // gdn: nC boxes [32 x n_ch], resident
  uint8_t* grp0 = gmat + (gdn ? nC * b_bytes : 0u);                 // 2 groups x grp_bytes (2 slots x 16 KB; col2im: Z tile)
  uint64_t* wfull = reinterpret_cast<uint64_t*>(grp0 + 2 * p.grp_bytes);
  uint64_t* wempty = wfull + kMaxStages;
  uint64_t* pfull = wempty + kMaxStages;
  uint64_t* pempty = pfull + kMaxPatch;
  uint64_t* gfull = pempty + kMaxPatch;       // [1]
  uint64_t* acc_full = gfull + 1;             // [2]
  uint64_t* norm_full = acc_full + 2;         // [2]
  uint64_t* tmem_free = norm_full + 2;        // [2]
  uint64_t* a2_ready = tmem_free + 2;         // [2 groups][2 slots]
  uint64_t* a2_free = a2_ready + 4;           // [2][2]
  uint64_t* ld_full = a2_free + 4;            // [2] backward: saved y / scale chunk of each epilogue group
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ld_full + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(wfull) + kBarBlock);
  float* sbeta = sbias + p.n_ch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < P; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    mbar_init(gfull, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1); mbar_init(&norm_full[b], 1);
      mbar_init(&tmem_free[b], bwd ? 2 : 1);   // backward epilogues: both epilogue groups read every item
    }
    for (int k = 0; k < 4; ++k) { mbar_init(&a2_ready[k], 128); mbar_init(&a2_free[k], 1); }
    mbar_init(&ld_full[0], 1); mbar_init(&ld_full[1], 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.out_map[0]);
    if (gdn) tma_prefetch_desc(&p.g_map);
  }
  for (int i = threadIdx.x; i < p.n_ch; i += kPThreads) {
    sbias[i] = (p.bias != nullptr && (EPI != kEpiCol2im || i < p.c2i_nch)) ? __ldg(p.bias + i) : 0.f;
    if (EPI == ICADV_EPI_GDN_FWD || EPI == ICADV_EPI_IGDN_FWD) sbeta[i] = __ldg(p.beta + i);
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  const int n_slots = p.n_active != nullptr ? *p.n_active : p.n_img;
  const int total = n_slots * p.tiles_x * p.tiles_y * p.n_class;

  // Patch stream of this CTA: patches q = 0, 1, 2, ... in (item, channel chunk, group) order go to ring slot q % P.
  // `mask` = 0: the calling thread produces all of them; 1: those with (q & 1) == first (two producing threads).
  auto produce_patches = [&](int first, int mask) {
    int ps = 0, q = 0;
    uint32_t p_par = 1;
    uint8_t* pdst = smem;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      const TcpItem it = tcp_decode(p, item);
      const TcpClass cl = p.cls[it.cls];
      for (int kc = 0; kc < p.k_chunks; ++kc) {
        const int c0 = kc * 32;
        for (int g = cl.g_begin; g < cl.g_end; ++g, ++q) {
          if ((q & mask) == first) {
            const int cx = it.j0 + p.groups[g].dx0, cy = it.i0 + p.groups[g].dy0;
            mbar_wait(&pempty[ps], p_par);
            mbar_arrive_expect_tx(&pfull[ps], p.groups[g].bytes);
            if (p.a_rank5)
              tma_load_5d(pdst, &p.a_map[0], &pfull[ps], 0, cx, cy, p.groups[g].plane5, it.img);
            else
              tma_load_4d(pdst, &p.a_map[p.groups[g].map], &pfull[ps], c0, cx, cy, it.img);
          }
          if (++ps == P) { ps = 0; p_par ^= 1; pdst = smem; } else { pdst += p.patch_bytes; }
        }
      }
    }
  };

  if (warp == 0) {
    // ===================== TMA producer: weights (+ the resident gamma) =====================
    // Weights and patches have separate producing threads: a patch is requested as soon as its slot is free (one
    // whole patch period ahead of its use) instead of queueing behind the weight boxes of the previous patch.
    if (elect_one_sync()) {
      if constexpr (EPI == kEpiCol2im) {
        // the 1x1 GEMM's weights (k_chunks boxes, the same for every item) are loaded once; this thread then takes every
        // other patch of the input stream (one thread issues one box per ~343 cycles: four boxes per item from one thread
        // would sit right at the HBM time of an item)
        const int wrow0 = p.tap_wtap[p.groups[p.cls[0].g_begin].tap_begin] * p.n_total;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_arrive_expect_tx(&wfull[kc], b_bytes);
          tma_load_2d(wring + kc * b_bytes, &p.w_map, &wfull[kc], kc * 32, wrow0);
        }
        produce_patches(1, 1);
      } else {
        if (gdn) {
          mbar_arrive_expect_tx(gfull, nC * b_bytes);
          for (int c = 0; c < nC; ++c) tma_load_2d(gmat + c * b_bytes, &p.g_map, gfull, c * 32, 0);
        }
        int s = 0;
        uint32_t s_par = 1;
        uint8_t* wdst = wring;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
          const TcpItem it = tcp_decode(p, item);
          const TcpClass cl = p.cls[it.cls];
          const int t_begin = p.groups[cl.g_begin].tap_begin, t_end = p.groups[cl.g_end - 1].tap_end;   // a class's taps are contiguous
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            const int c0 = kc * 32;
            int wrow = p.tap_wtap[t_begin] * p.n_total;
            for (int t = t_begin; t < t_end; ++t) {
              mbar_wait(&wempty[s], s_par);
              mbar_arrive_expect_tx(&wfull[s], b_bytes);
              tma_load_2d(wdst, &p.w_map, &wfull[s], c0, wrow);
              wrow = p.tap_wtap[t + 1] * p.n_total;   // one slack entry
              if (++s == S) { s = 0; s_par ^= 1; wdst = wring; } else { wdst += b_bytes; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kPPatchWarp) {
    // ===================== TMA producer: input halo patches =====================
    if (elect_one_sync()) produce_patches(0, EPI == kEpiCol2im ? 1 : 0);
    __syncwarp();
  } else if (warp == 1) {
    // ===================== main-loop MMA issuer =====================
    // ONE thread runs the whole role (elected once, no per-K-block reconvergence): at N = 128 the tensor core retires
    // a K-block (4 MMAs) every 256 cycles and a lone warp issues roughly one dependent instruction per 4-6 cycles, so
    // the issue path has to stay at a few dozen instructions per K-block (a bare issue loop reaches the tensor floor,
    // scripts/probes/mma_rate_probe.cu; the previous form -- converged warp, elect + reconvergence per K-block,
    // normalisation hand-off polled in the same loop -- measured ~530 cycles per K-block, ncu source view:
    // 19 % of it on the reconvergence barrier waiting for the four UTCHMMA to drain).  The normalisation K-blocks have
    // their own issuing warp (below): they only depend on the epilogue, never on this loop.
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      const uint32_t hi_dense = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t w_lo0 = ((smem_u32(wring) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t p_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_step = b_bytes >> 4, p_step = static_cast<uint32_t>(p.patch_bytes) >> 4;
      int s = 0, ps = 0;
      uint32_t s_par = 0, p_par = 0, w_lo = w_lo0, p_lo = p_lo0;
      uint32_t free_bits = 3;                  // tmem_free parity per buffer (first use: free)
      int b = 0;
      const bool prof = p.dbg != nullptr;
      long long c_free = 0, n_items = 0, c_ring = 0, c_ring_p = 0;
      const long long c_start = prof ? clock64() : 0;
      bool w_ok = false;                       // result of the early probe of wfull[s]
      if constexpr (EPI == kEpiCol2im) {
        // Resident weights, one class, one patch per K-block: the loop needs no item geometry at all.  Per K-block: the
        // patch wait (normally complete: the ring is deep), four MMAs, one commit.
        const int g0 = p.cls[0].g_begin;
        const uint32_t a_hi = (static_cast<uint32_t>(p.groups[g0].sbo_bytes) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t aoff = static_cast<uint32_t>(p.tap_aoff[p.groups[g0].tap_begin]) >> 4;
        for (int kc = 0; kc < p.k_chunks; ++kc) mbar_wait(&wfull[kc], 0);
        for (int item = blockIdx.x; item < total; item += gridDim.x, b ^= 1) {
          ++n_items;
          if (!mbar_test_wait(&tmem_free[b], (free_bits >> b) & 1u)) {
            const long long t0 = clock64();
            mbar_wait(&tmem_free[b], (free_bits >> b) & 1u);
            c_free += clock64() - t0;
          }
          free_bits ^= 1u << b;
          const uint32_t d = tmem + b * 256;
          w_lo = w_lo0;
          for (int kc = 0; kc < p.k_chunks; ++kc) {
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (p_lo + aoff);
            const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | w_lo;
            if (!mbar_test_wait(&pfull[ps], p_par)) {
              const long long t0 = clock64();
              mbar_wait(&pfull[ps], p_par);
              c_ring_p += clock64() - t0;
            }
            tc_fence_after_sync();
            tc_mma_tf32(d, ad, bd, idesc, kc > 0 ? 1u : 0u);
            tc_mma_tf32(d, ad + 2, bd + 2, idesc, 1u);
            tc_mma_tf32(d, ad + 4, bd + 4, idesc, 1u);
            tc_mma_tf32(d, ad + 6, bd + 6, idesc, 1u);
            tc_commit(&pempty[ps]);
            w_lo += w_step;
            if (++ps == P) { ps = 0; p_par ^= 1; p_lo = p_lo0; } else { p_lo += p_step; }
          }
          tc_commit(&acc_full[b]);
        }
      } else
      for (int item = blockIdx.x; item < total; item += gridDim.x, b ^= 1) {
        const TcpItem it = tcp_decode(p, item);
        const TcpClass cl = p.cls[it.cls];
        ++n_items;
        // the epilogue of the item that last used TMEM buffer b must have finished reading it
        if (!mbar_test_wait(&tmem_free[b], (free_bits >> b) & 1u)) {
          const long long t0 = clock64();
          mbar_wait(&tmem_free[b], (free_bits >> b) & 1u);
          c_free += clock64() - t0;
        }
        free_bits ^= 1u << b;
        tc_fence_after_sync();
        const uint32_t d = tmem + b * 256;
        uint32_t acc = 0;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int g = cl.g_begin; g < cl.g_end; ++g) {
            const uint32_t a_hi = (static_cast<uint32_t>(p.groups[g].sbo_bytes) >> 4) | (1u << 14) | (2u << 29);
            const int t_end = p.groups[g].tap_end;
            int t = p.groups[g].tap_begin;
            uint32_t aoff = static_cast<uint32_t>(p.tap_aoff[t]) >> 4;
            if (!mbar_test_wait(&pfull[ps], p_par)) {   // slow path (timed for the developer counters)
              const long long t0 = clock64();
              mbar_wait(&pfull[ps], p_par);
              c_ring_p += clock64() - t0;
            }
            for (; t < t_end; ++t) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (p_lo + aoff);
              const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | w_lo;
              if (!w_ok) {
                const long long t0 = clock64();
                mbar_wait(&wfull[s], s_par);
                c_ring += clock64() - t0;
              }
              uint64_t* wdone = &wempty[s];
              if (++s == S) { s = 0; s_par ^= 1; w_lo = w_lo0; } else { w_lo += w_step; }
              tc_fence_after_sync();
              tc_mma_tf32(d, ad, bd, idesc, acc);
              tc_mma_tf32(d, ad + 2, bd + 2, idesc, 1u);
              tc_mma_tf32(d, ad + 4, bd + 4, idesc, 1u);
              tc_mma_tf32(d, ad + 6, bd + 6, idesc, 1u);
              tc_commit(wdone);
              acc = 1u;
              // off the critical path (the four MMAs above are draining): next tap's offset, next stage's barrier
              aoff = static_cast<uint32_t>(p.tap_aoff[t + 1]) >> 4;   // tap_aoff has slack entries
              w_ok = mbar_test_wait(&wfull[s], s_par);
            }
            tc_commit(&pempty[ps]);
            if (++ps == P) { ps = 0; p_par ^= 1; p_lo = p_lo0; } else { p_lo += p_step; }
          }
        }
        tc_commit(&acc_full[b]);
      }
      if (prof) {
        long long* q = p.dbg + (int64_t)blockIdx.x * 16;
        q[0] = clock64() - c_start; q[1] = c_ring + c_ring_p; q[2] = c_free; q[4] = c_ring_p; q[7] = n_items; q[8] = q[0];
      }
    }
    __syncwarp();
  } else if (warp == kPNormWarp) {
    // ===================== normalisation MMA issuer (GDN / IGDN epilogues) =====================
    // Walks the same item sequence; per item nC K-blocks [128 x 32] x gamma chunk -> the item's norm region.  The A
    // operand is written by the item's epilogue group (a2_ready), which in turn waited for the main accumulator.
    if (gdn && elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      const uint32_t hi_dense = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t g_lo0 = ((smem_u32(gmat) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t a2_lo0 = ((smem_u32(grp0) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_step = b_bytes >> 4;
      uint32_t rdy_bits = 0;                   // a2_ready parity, bit (buffer * 2 + slot)
      int b = 0;
      mbar_wait(gfull, 0);
      for (int item = blockIdx.x; item < total; item += gridDim.x, b ^= 1) {
        const uint32_t dn = tmem + b * 256 + p.n_ch;
        // forward: the item belongs to epilogue group b, chunk c in slot c & 1.  backward: chunk c is produced by group g
        // (first / second half of the channels) in that group's slot 0, and the two groups' chunks are taken alternately
        // (a fixed order: the accumulation order stays deterministic)
        const int half = bwd ? (nC + 1) >> 1 : nC;
        int issued = 0;
        for (int j = 0; j < half; ++j) {
          for (int gg = 0; gg < (bwd ? 2 : 1); ++gg) {
            const int c = gg ? half + j : j;
            if (c >= nC) continue;
            const int g = bwd ? gg : b;
            const int slot = bwd ? 0 : (j & 1);   // backward: the operand replaces the staged y chunk in place (slot 0)
            const int k = g * 2 + slot;
            const uint64_t ad = (static_cast<uint64_t>(hi_dense) << 32) |
                                (a2_lo0 + g * (static_cast<uint32_t>(p.grp_bytes) >> 4) + slot * (kABytes >> 4));
            const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | (g_lo0 + c * w_step);
            mbar_wait(&a2_ready[k], (rdy_bits >> k) & 1u);
            rdy_bits ^= 1u << k;
            tc_fence_after_sync();
            tc_mma_tf32(dn, ad, bd, idesc, issued > 0 ? 1u : 0u);
            tc_mma_tf32(dn, ad + 2, bd + 2, idesc, 1u);
            tc_mma_tf32(dn, ad + 4, bd + 4, idesc, 1u);
            tc_mma_tf32(dn, ad + 6, bd + 6, idesc, 1u);
            tc_commit(&a2_free[k]);
            if (++issued == nC) tc_commit(&norm_full[b]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== two epilogue warpgroups =====================
    // Backward epilogues (eight serial saved-chunk round trips per item): BOTH groups work on every item, each on half
    // of the 32-channel chunks, so an item's TMEM buffer is handed back after half the epilogue latency and the issuer's
    // next main loop (other buffer) overlaps it; with alternate items per group a group's own next main loop could never
    // start before its epilogue had ended (g_a.2 dgrad: 20 us per item and group, of which 6 us main loop; 4.08 -> 3.47 ms).
    // Forward / linear / col2im epilogues keep alternate items: they have no load latency to hide and two items in
    // flight overlap their normalisation wait and store drain (split measured slower there: RGB layer 1.58 -> 1.96 ms).
    constexpr bool split = bwd;
    const int grp = (warp - 2) >> 2;           // 0: warps 2-5, 1: warps 6-9
    const int q = warp & 3;                    // TMEM lane quadrant this warp may touch
    const int row = q * 32 + lane;
    const bool leader = (row == 0);
    const uint32_t bar_id = 1 + grp;
    uint8_t* gbuf = grp0 + grp * p.grp_bytes;  // two 16 KB slots: normalisation A operand / store staging (col2im: Z tile)
    const int c_half = (nC + 1) >> 1;
    const int c_lo = split ? (grp == 0 ? 0 : c_half) : 0;        // this group's chunks [c_lo, c_hi)
    const int c_hi = split ? (grp == 0 ? c_half : nC) : nC;
    uint32_t acc_bits = 0, norm_bits = 0;      // acc_full / norm_full parity per TMEM buffer
    uint32_t slot_par[2] = {1, 1};             // a2_free parity per slot (first use: free)
    // Backward epilogues stage the saved y / scale of the item one 32-channel chunk at a time in this group's slot pair
    // [y | scale] (TMA, swizzled rows).  Everything that follows a chunk happens IN PLACE: pass 1 overwrites the y
    // rows with the normalisation operand (each thread reads and writes only its own row), pass 2 overwrites them
    // with the output chunk, which is TMA-stored from there.  The next chunk is requested once the tensor core
    // (pass 1) or the store engine (pass 2) has finished reading the pair.
    uint32_t ld_par = 0, a2c_par = 0;
    uint8_t* bufY = gbuf;
    uint8_t* bufS = gbuf + kABytes;
    auto fetch = [&](const TcpItem& t, int c) {
      if (leader) {
        mbar_arrive_expect_tx(&ld_full[grp], 2 * kABytes);
        tma_load_4d(bufY, &p.yprev_map[t.cls], &ld_full[grp], c * 32, t.j0, t.i0, t.img);
        tma_load_4d(bufS, &p.sc_map[t.cls], &ld_full[grp], c * 32, t.j0, t.i0, t.img);
      }
      __syncwarp();
    };
    const int item0 = blockIdx.x + (split ? 0 : grp * gridDim.x);
    const int item_step = (split ? 1 : 2) * gridDim.x;
    if constexpr (bwd) {
      if (item0 < total && c_lo < c_hi) fetch(tcp_decode(p, item0), c_lo);
    }

    int b = split ? 0 : grp;                   // TMEM buffer of the current item
    for (int item = item0; item < total; item += item_step, b ^= (split ? 1 : 0)) {
      const TcpItem it = tcp_decode(p, item);
      const int o_a = p.cls[it.cls].o_a, o_b = p.cls[it.cls].o_b;
      const int gi = it.i0 + row / kTW, gj = it.j0 + row % kTW;
      const bool px_ok = gi < p.t_h && gj < p.t_w;
      const int64_t pix = (((int64_t)it.img * p.o_h + p.o_s * gi + o_a) * p.o_w + p.o_s * gj + o_b) * p.n_ch;
      (void)px_ok; (void)pix;
      const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16) + b * 256;
      const bool eprof = p.dbg != nullptr && leader;
      const long long e0 = eprof ? clock64() : 0;
      mbar_wait(&acc_full[b], (acc_bits >> b) & 1u);
      acc_bits ^= 1u << b;
      tc_fence_after_sync();
      const long long e1 = eprof ? clock64() : 0;
      long long e2 = e1, e3 = e1;

      if constexpr (EPI == kEpiCol2im) {
        // Z[128 px][taps * nch] -> this group's shared-memory tile, TMEM handed back at once, then the transposed-conv
        // outputs of the tile interior are gathered from Z (same arithmetic as conv_tc_kernel<kEpiCol2im>)
        float* Zs = reinterpret_cast<float*>(gbuf);
        const int zcols = 25 * p.c2i_nch;
        const int OH = 2 * p.c2i_in_h, OW = 2 * p.c2i_in_w, nch = p.c2i_nch;
        constexpr int kIy = kTH - 2, kIx = kTW - 2;
        if (nch == 3) {
          // RGB fast path.  The scalar form below measured 5600 cycles per item and group for the gather (27 LDS.32 with
          // their address arithmetic per output pixel, 128 threads) against ~1900 cycles of HBM time per item: Z rows are
          // written as 25 x float4 (one STS.128 per tap) and every tap of an output pixel is ONE LDS.128 at a
          // compile-time offset from a per-thread base.  Same summation order (u outer, v inner) = same bits.
          {
            float v[96];
            tmem_ld32(t_lane, v); tmem_ld32(t_lane + 32, v + 32); tmem_ld32(t_lane + 64, v + 64);
            tmem_ld_wait();
            float4* zr = reinterpret_cast<float4*>(Zs + row * kZStride4);
#pragma unroll
            for (int t = 0; t < 25; ++t) zr[t] = make_float4(v[3 * t], v[3 * t + 1], v[3 * t + 2], 0.f);
          }
          tc_fence_before_sync();
          named_bar_sync(bar_id, 128);
          if (leader) mbar_arrive(&tmem_free[b]);
          if (eprof) e2 = clock64();
          const int a = (row >> 1) & 1, bb = row & 1;      // o = row + 128 k: the output parity is the thread's own
          const float b0 = sbias[0], b1 = sbias[1], b2 = sbias[2];
#pragma unroll
          for (int k = 0; k < (kIy * kIx * 4 + 127) / 128; ++k) {
            const int o = row + k * 128;
            if (o >= kIy * kIx * 4) break;
            const int cell = o >> 2;
            const int ti = 1 + cell / kIx, tj = 1 + cell % kIx;
            const int yi = it.i0 + ti, xj = it.j0 + tj;
            if (yi >= p.c2i_in_h || xj >= p.c2i_in_w) continue;
            const float* zb = Zs + ((ti + 1) * kTW + (tj + 1)) * kZStride4 + (a * 5 + bb) * 4;
            float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
#pragma unroll
              for (int v = 0; v < 3; ++v) {
                if ((u < 2 || a == 0) && (v < 2 || bb == 0)) {   // kh = a + 2u < 5, kw = b + 2v < 5
                  const float4 z = *reinterpret_cast<const float4*>(zb - (u * kTW + v) * kZStride4 + (10 * u + 2 * v) * 4);
                  ax += z.x; ay += z.y; az += z.z;
                }
              }
            }
            float* dst = p.c2i_out + (((int64_t)it.img * OH + 2 * yi + a) * OW + 2 * xj + bb) * 3;
            dst[0] = ax + b0; dst[1] = ay + b1; dst[2] = az + b2;
          }
          named_bar_sync(bar_id, 128);   // Z is rewritten by this group's next item
          if (eprof) {
            long long* q = p.dbg + (int64_t)blockIdx.x * 16 + 9 + grp * 3;   // per group: wait acc, Z tile, gather + stores
            const long long e4 = clock64();
            q[0] += e1 - e0; q[1] += e2 - e1; q[2] += e4 - e2;
          }
          continue;
        }
        for (int c = 0; c < nC; ++c) {
          float v[32];
          tmem_ld32(t_lane + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < zcols) Zs[row * kZStride + c * 32 + j] = v[j];
        }
        tc_fence_before_sync();
        named_bar_sync(bar_id, 128);
        if (leader) mbar_arrive(&tmem_free[b]);
        if (eprof) e2 = clock64();
        for (int o = row; o < kIy * kIx * 4; o += 128) {
          const int cell = o >> 2, a = (o >> 1) & 1, b = o & 1;
          const int ti = 1 + cell / kIx, tj = 1 + cell % kIx;
          const int yi = it.i0 + ti, xj = it.j0 + tj;
          if (yi >= p.c2i_in_h || xj >= p.c2i_in_w) continue;
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int u = 0; u < 3; ++u) {
#pragma unroll
            for (int v = 0; v < 3; ++v) {
              const int kh = a + 2 * u, kw = b + 2 * v;
              const bool ok = kh < 5 && kw < 5;
              const float* z = Zs + ((ti + 1 - u) * kTW + (tj + 1 - v)) * kZStride + (ok ? (kh * 5 + kw) * nch : 0);
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < nch) acc[c] += ok ? z[c] : 0.f;
            }
          }
          float* dst = p.c2i_out + (((int64_t)it.img * OH + 2 * yi + a) * OW + 2 * xj + b) * nch;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < nch) dst[c] = acc[c] + sbias[c];
        }
        named_bar_sync(bar_id, 128);   // Z is rewritten by this group's next item
        if (eprof) {
          long long* q = p.dbg + (int64_t)blockIdx.x * 16 + 9 + grp * 3;   // per group: wait acc, Z tile, gather + stores
          const long long e4 = clock64();
          q[0] += e1 - e0; q[1] += e2 - e1; q[2] += e4 - e2;
        }
        continue;
      }

      auto load_acc = [&](int c, float* v) {
        tmem_ld32(t_lane + c * 32, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(sbias + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
      };
      int stores = 0;
      // double-buffered store of one output chunk through slot (stores & 1): a slot is rewritten once the store issued
      // from it two stores ago has finished reading (each store is its own bulk group)
      auto store_one = [&](int c, const float* o, const CUtensorMap* map, bool round) {
        uint8_t* buf = gbuf + (stores & 1) * kABytes;
        if (stores >= 2) {
          if (leader) tma_store_wait_read1();
          __syncwarp();
          named_bar_sync(bar_id, 128);
        }
        if (round) {
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = round_tf32(o[j]);
          write_row32(buf, row, r);
        } else {
          write_row32(buf, row, o);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) { tma_store_4d(map, buf, c * 32, it.j0, it.i0, it.img); tma_store_commit(); }
        __syncwarp();
        ++stores;
      };

      if constexpr (!gdn) {
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32];
          load_acc(c, v);
          if (p.act == ICADV_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.act == ICADV_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
          } else if (p.act == ICADV_ACT_ABS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fabsf(v[j]);
          }
          store_one(c, v, &p.out_map[it.cls], p.round_out != 0);
        }
      } else {
        // ---- pass 1: A operand of the normalisation GEMM, this group's chunk c -> slot (c - c_lo) & 1 ----
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32], a2[32];
          load_acc(c, v);
          if constexpr (!bwd) {
#pragma unroll
            for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * v[j]);
          } else {
            float yv[32], sv[32];
            mbar_wait(&ld_full[grp], ld_par);
            ld_par ^= 1;
            read_row32(bufY, row, yv);
            read_row32(bufS, row, sv);
            if constexpr (EPI == ICADV_EPI_GDN_BWD) {
#pragma unroll
              for (int j = 0; j < 32; ++j) a2[j] = round_tf32(v[j] * yv[j] * sv[j] * sv[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float s2 = sv[j] * sv[j];
                a2[j] = s2 > 0.f ? round_tf32(fast_div(v[j] * yv[j], s2)) : 0.f;
              }
            }
            write_row32(bufY, row, a2);                 // in place: this thread's own row
            fence_proxy_async_smem();
            mbar_arrive(&a2_ready[grp * 2]);
            // the pair is refilled once the tensor core has read the operand (all 128 threads arrived before that)
            if (leader) mbar_wait(&a2_free[grp * 2], a2c_par);
            a2c_par ^= 1;
            if (c + 1 < c_hi) fetch(it, c + 1);
            continue;
          }
          const int k = (c - c_lo) & 1;
          mbar_wait(&a2_free[grp * 2 + k], slot_par[k]);   // the MMAs that last read this slot are done
          slot_par[k] ^= 1;
          write_row32(gbuf + k * kABytes, row, a2);
          fence_proxy_async_smem();
          mbar_arrive(&a2_ready[grp * 2 + k]);
        }
        // ---- pass 2: normalise ----
        if (eprof) e2 = clock64();
        mbar_wait(&norm_full[b], (norm_bits >> b) & 1u);
        norm_bits ^= 1u << b;
        tc_fence_after_sync();
        if (eprof) e3 = clock64();
        if constexpr (bwd) { if (c_lo < c_hi) fetch(it, c_lo); }   // every normalisation MMA has completed: the pair is free again
        for (int c = c_lo; c < c_hi; ++c) {
          float v[32], w[32];
          load_acc(c, v);
          tmem_ld32(t_lane + p.n_ch + c * 32, w);
          tmem_ld_wait();
          if constexpr (!bwd) {
            float sc[32];
            const float4* be4 = reinterpret_cast<const float4*>(sbeta + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 be = be4[j];
              const float nn[4] = {be.x + w[4 * j], be.y + w[4 * j + 1], be.z + w[4 * j + 2], be.w + w[4 * j + 3]};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float r = rsqrtf(nn[t]);
                sc[4 * j + t] = (EPI == ICADV_EPI_GDN_FWD) ? r : nn[t] * r;
                v[4 * j + t] *= sc[4 * j + t];
              }
            }
            // the two outputs of the chunk leave through alternating slots, each store its own bulk group
            store_one(c, v, &p.out_map[it.cls], p.round_out != 0);
            store_one(c, sc, &p.sc_map[it.cls], false);
          } else {
            float yv[32], sv[32];
            mbar_wait(&ld_full[grp], ld_par);
            ld_par ^= 1;
            read_row32(bufY, row, yv);
            read_row32(bufS, row, sv);
            constexpr float sign = (EPI == ICADV_EPI_GDN_BWD) ? -1.f : 1.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float xs = sv[j] > 0.f ? fast_div(yv[j], sv[j]) : 0.f;
              v[j] = v[j] * sv[j] + sign * xs * w[j];
              if (p.round_out) v[j] = round_tf32(v[j]);
            }
            write_row32(bufY, row, v);                  // in place, then stored from here
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 128);
            if (leader) {
              tma_store_4d(&p.out_map[it.cls], bufY, c * 32, it.j0, it.i0, it.img);
              tma_store_commit();
              if (c + 1 < c_hi) tma_store_wait_read0();   // the store engine has read the pair: refill it
            }
            __syncwarp();
            if (c + 1 < c_hi) fetch(it, c + 1);
          }
        }
      }
      // this item's TMEM reads are over: hand the buffer back to the MMA issuer; the staging slots are reused by
      // the next item (as normalisation operands or staging) only after every store has finished READING them
      tc_fence_before_sync();
      if (leader) tma_store_wait_read0();
      __syncwarp();
      named_bar_sync(bar_id, 128);
      if (leader) mbar_arrive(&tmem_free[b]);
      if (eprof) {
        long long* q = p.dbg + (int64_t)blockIdx.x * 16 + 9 + grp * 3;   // per group: wait acc, pass 1 (+ wait norm), pass 2
        const long long e4 = clock64();
        q[0] += e1 - e0; q[1] += e3 - e1; q[2] += e4 - e3;
        if (grp == 0) p.dbg[(int64_t)blockIdx.x * 16 + 3] += e3 - e2;  // of which: waiting for the last normalisation MMA
      }
      if constexpr (bwd) {   // request the first saved chunk of this group's next item while its main loop runs
        const int nxt = item + item_step;
        if (nxt < total && c_lo < c_hi) fetch(tcp_decode(p, nxt), c_lo);
      }
    }
    if (leader) tma_store_wait0();
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// =====================================================================================================
// Streaming backward variant (GDN / IGDN backward epilogues): conv_tcp_kernel's item walk, rings and TMEM double buffer,
// with the saved tensors of the matching forward treated as what they are -- the HBM stream that bounds these launches
// (y + scale in, gradient out: 3 x N x 4 bytes per pixel against a main loop that is short on the RGB layer and
// tensor-bound by a small margin on the 128-channel ones).
//   * the saved y / scale chunks of successive items stream through their own ring of [y | scale] slots, requested by a
//     producing thread as soon as a slot is free -- across item boundaries, so the epilogue of item i finds its chunks
//     already in shared memory and the ring keeps HBM requests in flight while the epilogue computes (the per-tile and
//     the persistent kernel above fetch one pair per epilogue group on demand: eight serial round trips per tile);
//   * pass 1 leaves its per-element products in TMEM instead of re-reading y and scale in pass 2: g*sc overwrites the
//     accumulator chunk in place and y/sc goes to a stash region (tcgen05.st), so every saved byte crosses the SM once;
//     TMEM: [acc 0 | acc 1 | norm | stash] x 128 columns;
//   * a slot is handed back as soon as its 128 consumer threads hold the pair in registers (the normalisation operand
//     goes to the group's staging tile, not in place), so the next pair of the stream is in flight during the arithmetic;
//   * gamma^T is streamed (two 16 KB stages, one chunk per normalisation K-block, L2-resident) instead of held resident:
//     the 32 KB that frees pay for one 16 KB output staging tile per epilogue group -- the output leaves through TMA
//     stores (per-thread 128-byte rows written with plain stores were tried: the LSU needs 32 line transactions per
//     instruction, 2 us per item).
// Patches, saved chunks and gamma^T chunks are produced by ONE polling thread (three cursors, mbarrier.test_wait, no
// blocking wait: a blocked patch request must not hold back the saved-tensor stream and vice versa).
// Arithmetic and accumulation order are those of conv_tcp_kernel (bit-identical results).
// =====================================================================================================
constexpr int kMaxYs = 4;

template <int EPI>
__global__ void __launch_bounds__(kPThreads, 1) conv_tcb_kernel(const __grid_constant__ TcpParams p) {
  static_assert(EPI == ICADV_EPI_GDN_BWD || EPI == ICADV_EPI_IGDN_BWD, "backward epilogues only");
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int P = p.num_patch, S = p.num_stages, nC = p.n_chunks, R = p.ys_slots;
  const int half = nC >> 1;                                          // chunks per epilogue group (nC is even)
  const uint32_t b_bytes = static_cast<uint32_t>(p.n_ch) * 128u;
  uint8_t* wring = smem + P * p.patch_bytes;
  uint8_t* gring = wring + S * b_bytes;                              // 2 stages [32 x n_ch] of gamma^T chunks (streamed from L2)
  uint8_t* ysr = gring + 2 * b_bytes;                                // R slots x [y chunk 16 KB | scale chunk 16 KB]
  uint8_t* ostage = ysr + R * 2 * kABytes;                           // one 16 KB output staging tile per epilogue group
  uint64_t* wfull = reinterpret_cast<uint64_t*>(ostage + 2 * kABytes);
  uint64_t* wempty = wfull + kMaxStages;
  uint64_t* pfull = wempty + kMaxStages;
  uint64_t* pempty = pfull + kMaxPatch;
  uint64_t* g_full = pempty + kMaxPatch;      // [2]
  uint64_t* g_empty = g_full + 2;             // [2]
  uint64_t* acc_full = g_empty + 2;           // [2]
  uint64_t* tmem_free = acc_full + 2;         // [2]  both epilogue groups have finished with accumulator b
  uint64_t* norm_full = tmem_free + 2;        // [1]
  uint64_t* norm_free = norm_full + 1;        // [1]  both groups have read the norm region
  uint64_t* ys_full = norm_free + 1;          // [kMaxYs]  TMA arrival of a [y | scale] pair
  uint64_t* ys_empty = ys_full + kMaxYs;      // [kMaxYs]  the 128 threads of the consuming group hold the pair in registers
  uint64_t* a2_ready = ys_empty + kMaxYs;     // [2]  per group: 128 threads have written the normalisation operand
  uint64_t* a2_free = a2_ready + 2;           // [2]  per group: the MMAs that read the operand have completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a2_free + 2);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(wfull) + kBarBlock);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int s = 0; s < P; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1); mbar_init(&tmem_free[b], 2);
      mbar_init(&g_full[b], 1); mbar_init(&g_empty[b], 1);
    }
    mbar_init(norm_full, 1); mbar_init(norm_free, 2);
    for (int k = 0; k < R; ++k) { mbar_init(&ys_full[k], 1); mbar_init(&ys_empty[k], 128); }
    for (int g = 0; g < 2; ++g) { mbar_init(&a2_ready[g], 128); mbar_init(&a2_free[g], 1); }
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]); tma_prefetch_desc(&p.w_map); tma_prefetch_desc(&p.g_map);
    tma_prefetch_desc(&p.yprev_map[0]); tma_prefetch_desc(&p.sc_map[0]); tma_prefetch_desc(&p.out_map[0]);
  }
  for (int i = threadIdx.x; i < p.n_ch; i += kPThreads) sbias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  const int n_slots = p.n_active != nullptr ? *p.n_active : p.n_img;
  const int total = n_slots * p.tiles_x * p.tiles_y * p.n_class;
  // TMEM columns: accumulator b at b * 128, norm at 256, stash (y / sc) at 384

  if (warp == 0) {
    // ===================== TMA producer: weights =====================
    if (elect_one_sync()) {
      int s = 0;
      uint32_t s_par = 1;
      uint8_t* wdst = wring;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const TcpItem it = tcp_decode(p, item);
        const TcpClass cl = p.cls[it.cls];
        const int t_begin = p.groups[cl.g_begin].tap_begin, t_end = p.groups[cl.g_end - 1].tap_end;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          const int c0 = kc * 32;
          int wrow = p.tap_wtap[t_begin] * p.n_total;
          for (int t = t_begin; t < t_end; ++t) {
            mbar_wait(&wempty[s], s_par);
            if (p.dbg_skip_w) {   // developer experiment (ICADV_TC_DBG bit 16): stage handed over without loading the weights
              mbar_arrive(&wfull[s]);
            } else {
              mbar_arrive_expect_tx(&wfull[s], b_bytes);
              tma_load_2d(wdst, &p.w_map, &wfull[s], c0, wrow);
            }
            wrow = p.tap_wtap[t + 1] * p.n_total;   // one slack entry
            if (++s == S) { s = 0; s_par ^= 1; wdst = wring; } else { wdst += b_bytes; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kPPatchWarp) {
    // ===================== TMA producer: input halo patches AND the saved-tensor stream (two cursors, polled) ==========
    if (elect_one_sync()) {
      int p_item = blockIdx.x, p_kc = 0, p_g = 0, p_gend = 0;
      TcpItem pit = {0, 0, 0, 0};
      bool p_live = p_item < total;
      if (p_live) { pit = tcp_decode(p, p_item); p_g = p.cls[pit.cls].g_begin; p_gend = p.cls[pit.cls].g_end; }
      int ps = 0;
      uint32_t p_par = 1;
      int y_item = blockIdx.x, y_q = 0;
      TcpItem yit = {0, 0, 0, 0};
      bool y_live = y_item < total;
      if (y_live) yit = tcp_decode(p, y_item);
      int ys = 0;
      uint32_t y_par = 1;
      // third cursor: the gamma^T chunk of every normalisation K-block, in the stream's chunk order (L2-resident, 64 KB)
      int g_item = blockIdx.x, g_q = 0, gs = 0;
      uint32_t g_par = 1;
      bool g_live = g_item < total;
      while (p_live || y_live || g_live) {
        bool idle = true;
        if (g_live && mbar_test_wait(&g_empty[gs], g_par)) {
          idle = false;
          const int c = (g_q & 1) * half + (g_q >> 1);
          mbar_arrive_expect_tx(&g_full[gs], b_bytes);
          tma_load_2d(gring + gs * b_bytes, &p.g_map, &g_full[gs], c * 32, 0);
          if (++gs == 2) { gs = 0; g_par ^= 1; }
          if (++g_q == nC) { g_q = 0; g_item += gridDim.x; g_live = g_item < total; }
        }
        if (y_live && mbar_test_wait(&ys_empty[ys], y_par)) {
          idle = false;
          const int c = (y_q & 1) * half + (y_q >> 1);               // the two groups' chunks alternate (fixed order)
          uint8_t* dst = ysr + ys * 2 * kABytes;
          mbar_arrive_expect_tx(&ys_full[ys], 2 * kABytes);
          tma_load_4d(dst, &p.yprev_map[yit.cls], &ys_full[ys], c * 32, yit.j0, yit.i0, yit.img);
          tma_load_4d(dst + kABytes, &p.sc_map[yit.cls], &ys_full[ys], c * 32, yit.j0, yit.i0, yit.img);
          if (++ys == R) { ys = 0; y_par ^= 1; }
          if (++y_q == nC) {
            y_q = 0;
            y_item += gridDim.x;
            y_live = y_item < total;
            if (y_live) yit = tcp_decode(p, y_item);
          }
        }
        if (p_live && mbar_test_wait(&pempty[ps], p_par)) {
          idle = false;
          const int cx = pit.j0 + p.groups[p_g].dx0, cy = pit.i0 + p.groups[p_g].dy0;
          uint8_t* pdst = smem + ps * p.patch_bytes;
          mbar_arrive_expect_tx(&pfull[ps], p.groups[p_g].bytes);
          if (p.a_rank5)
            tma_load_5d(pdst, &p.a_map[0], &pfull[ps], 0, cx, cy, p.groups[p_g].plane5, pit.img);
          else
            tma_load_4d(pdst, &p.a_map[p.groups[p_g].map], &pfull[ps], p_kc * 32, cx, cy, pit.img);
          if (++ps == P) { ps = 0; p_par ^= 1; }
          if (++p_g == p_gend) {
            if (++p_kc == p.k_chunks) {
              p_kc = 0;
              p_item += gridDim.x;
              p_live = p_item < total;
              if (p_live) { pit = tcp_decode(p, p_item); p_gend = p.cls[pit.cls].g_end; }
            }
            if (p_live) p_g = p.cls[pit.cls].g_begin;
          }
        }
        if (idle) __nanosleep(64);   // both rings full: leave the issue slots of this scheduler to its epilogue warps
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== main-loop MMA issuer (as conv_tcp_kernel) =====================
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      const uint32_t hi_dense = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t w_lo0 = ((smem_u32(wring) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t p_lo0 = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_step = b_bytes >> 4, p_step = static_cast<uint32_t>(p.patch_bytes) >> 4;
      int s = 0, ps = 0;
      uint32_t s_par = 0, p_par = 0, w_lo = w_lo0, p_lo = p_lo0;
      uint32_t free_bits = 3;
      int b = 0;
      bool w_ok = false;
      const bool prof = p.dbg != nullptr;
      long long c_free = 0, c_ring = 0, c_ring_p = 0, n_items = 0;
      const long long c_start = prof ? clock64() : 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, b ^= 1) {
        const TcpItem it = tcp_decode(p, item);
        const TcpClass cl = p.cls[it.cls];
        ++n_items;
        if (!mbar_test_wait(&tmem_free[b], (free_bits >> b) & 1u)) {
          const long long t0 = clock64();
          mbar_wait(&tmem_free[b], (free_bits >> b) & 1u);
          c_free += clock64() - t0;
        }
        free_bits ^= 1u << b;
        tc_fence_after_sync();
        const uint32_t d = tmem + b * 128;
        uint32_t acc = 0;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          for (int g = cl.g_begin; g < cl.g_end; ++g) {
            const uint32_t a_hi = (static_cast<uint32_t>(p.groups[g].sbo_bytes) >> 4) | (1u << 14) | (2u << 29);
            const int t_end = p.groups[g].tap_end;
            int t = p.groups[g].tap_begin;
            uint32_t aoff = static_cast<uint32_t>(p.tap_aoff[t]) >> 4;
            if (!mbar_test_wait(&pfull[ps], p_par)) {
              const long long t0 = clock64();
              mbar_wait(&pfull[ps], p_par);
              c_ring_p += clock64() - t0;
            }
            for (; t < t_end; ++t) {
              const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (p_lo + aoff);
              const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | w_lo;
              if (!w_ok) {
                const long long t0 = clock64();
                mbar_wait(&wfull[s], s_par);
                c_ring += clock64() - t0;
              }
              uint64_t* wdone = &wempty[s];
              if (++s == S) { s = 0; s_par ^= 1; w_lo = w_lo0; } else { w_lo += w_step; }
              tc_fence_after_sync();
              tc_mma_tf32(d, ad, bd, idesc, acc);
              tc_mma_tf32(d, ad + 2, bd + 2, idesc, 1u);
              tc_mma_tf32(d, ad + 4, bd + 4, idesc, 1u);
              tc_mma_tf32(d, ad + 6, bd + 6, idesc, 1u);
              tc_commit(wdone);
              acc = 1u;
              aoff = static_cast<uint32_t>(p.tap_aoff[t + 1]) >> 4;
              w_ok = mbar_test_wait(&wfull[s], s_par);
            }
            tc_commit(&pempty[ps]);
            if (++ps == P) { ps = 0; p_par ^= 1; p_lo = p_lo0; } else { p_lo += p_step; }
          }
        }
        tc_commit(&acc_full[b]);
      }
      if (prof) {
        long long* q = p.dbg + (int64_t)blockIdx.x * 16;
        q[0] = clock64() - c_start; q[1] = c_ring + c_ring_p; q[2] = c_free; q[4] = c_ring_p; q[7] = n_items; q[8] = q[0];
      }
    }
    __syncwarp();
  } else if (warp == kPNormWarp) {
    // ===================== normalisation MMA issuer =====================
    // per item: nC K-blocks [128 x 32] (operand written in place over the y rows of a ring slot) x gamma^T chunk -> norm
    // region, in the fixed chunk order of the stream; each K-block's commit hands its slot back to the producer
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(kTileM, p.n_ch);
      const uint32_t hi_dense = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t g_lo0 = ((smem_u32(gring) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t o_lo0 = ((smem_u32(ostage) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_step = b_bytes >> 4;
      const uint32_t dn = tmem + 256;
      int gs = 0;
      uint32_t g_par = 0, nf_par = 1, a2_par = 0;   // a2_par: bit g = parity of group g's next operand hand-off
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        mbar_wait(norm_free, nf_par);            // both groups have read the previous item's norm region
        nf_par ^= 1;
        for (int q = 0; q < nC; ++q) {
          const int g = q & 1;                   // the stream alternates the two groups' chunks
          const uint64_t ad = (static_cast<uint64_t>(hi_dense) << 32) | (o_lo0 + g * (kABytes >> 4));
          const uint64_t bd = (static_cast<uint64_t>(hi_dense) << 32) | (g_lo0 + gs * w_step);
          mbar_wait(&g_full[gs], g_par);
          mbar_wait(&a2_ready[g], (a2_par >> g) & 1u);
          a2_par ^= 1u << g;
          tc_fence_after_sync();
          tc_mma_tf32(dn, ad, bd, idesc, q > 0 ? 1u : 0u);
          tc_mma_tf32(dn, ad + 2, bd + 2, idesc, 1u);
          tc_mma_tf32(dn, ad + 4, bd + 4, idesc, 1u);
          tc_mma_tf32(dn, ad + 6, bd + 6, idesc, 1u);
          tc_commit(&a2_free[g]);
          tc_commit(&g_empty[gs]);
          if (q == nC - 1) tc_commit(norm_full);
          if (++gs == 2) { gs = 0; g_par ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== two epilogue warpgroups: both work on every item, half of the chunks each =====================
    const int grp = (warp - 2) >> 2;           // 0: warps 2-5, 1: warps 6-9
    const int q4 = warp & 3;                   // TMEM lane quadrant this warp may touch
    const int row = q4 * 32 + lane;
    const bool leader = (row == 0);
    const uint32_t bar_id = 1 + grp;
    uint8_t* obuf = ostage + grp * kABytes;
    constexpr float sign = (EPI == ICADV_EPI_GDN_BWD) ? -1.f : 1.f;
    uint32_t acc_bits = 0, nfull_par = 0, a2f_par = 1;   // a2_free: first use free
    int b = 0;
    uint32_t seq0 = 0;                         // stream position of this item's first chunk
    const uint32_t r_log2 = R == 4 ? 2u : 1u;
    for (int item = blockIdx.x; item < total; item += gridDim.x, b ^= 1, seq0 += nC) {
      const TcpItem it = tcp_decode(p, item);
      const int o_a = p.cls[it.cls].o_a, o_b = p.cls[it.cls].o_b;
      const int gi = it.i0 + row / kTW, gj = it.j0 + row % kTW;
      const bool px_ok = gi < p.t_h && gj < p.t_w;
      (void)o_a; (void)o_b; (void)px_ok;
      const uint32_t t_lane = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
      const uint32_t t_acc = t_lane + b * 128, t_norm = t_lane + 256, t_stash = t_lane + 384;
      const bool eprof = p.dbg != nullptr && leader;
      const long long e0 = eprof ? clock64() : 0;
      long long e_ys = 0;
      mbar_wait(&acc_full[b], (acc_bits >> b) & 1u);
      acc_bits ^= 1u << b;
      tc_fence_after_sync();
      const long long e1 = eprof ? clock64() : 0;
      long long e_drain = 0;
      // ---- pass 1: this group's chunks, in stream order
      for (int j = 0; j < half; ++j) {
        const int c = grp * half + j;
        const uint32_t seq = seq0 + 2 * j + grp;
        const uint32_t slot = seq & (R - 1), par = (seq >> r_log2) & 1u;    // R is a power of two
        uint8_t* bufY = ysr + slot * 2 * kABytes;
        uint8_t* bufS = bufY + kABytes;
        float v[32], yv[32], sv[32];
        tmem_ld32(t_acc + c * 32, v);
        tmem_ld_wait();
        if (p.bias != nullptr) {                    // an input-gradient contraction has no bias: uniform, normally skipped
          const float4* b4 = reinterpret_cast<const float4*>(sbias + c * 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 bb = b4[k];
            v[4 * k] += bb.x; v[4 * k + 1] += bb.y; v[4 * k + 2] += bb.z; v[4 * k + 3] += bb.w;
          }
        }
        if (eprof) { const long long t0 = clock64(); mbar_wait(&ys_full[slot], par); e_ys += clock64() - t0; }
        else mbar_wait(&ys_full[slot], par);
        read_row32(bufY, row, yv);
        read_row32(bufS, row, sv);
        {
          // The pair is in registers: hand the slot back to the producer NOW, so that the next pair of the stream is in
          // flight while this one is being processed.  Every shared-memory load above must have RETURNED before the
          // arrival (the slot is overwritten through the async proxy): one word of each 16-byte load feeds the address
          // of the arrive instruction, so the scoreboard holds it until the data is here.
          uint32_t dep = 0;
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) dep |= __float_as_uint(yv[4 * k4]) | __float_as_uint(sv[4 * k4]);
          fence_proxy_async_smem();
          asm volatile("{\n.reg .b32 t;\nand.b32 t, %1, 0;\nadd.u32 t, t, %0;\nmbarrier.arrive.shared::cta.b64 _, [t];\n}"
                       ::"r"(smem_u32(&ys_empty[slot])), "r"(dep) : "memory");
        }
        mbar_wait(&a2_free[grp], a2f_par);          // the MMAs that read this group's previous operand are done
        a2f_par ^= 1;
        if (j == 0) {
          // the staging tile doubles as the operand buffer: the last TMA store of the previous item must have read it
          // (waited for here, after this chunk's loads and hand-back, so the drain overlaps them)
          const long long t0 = eprof ? clock64() : 0;
          if (leader) tma_store_wait_read0();
          __syncwarp();
          named_bar_sync(bar_id, 128);
          if (eprof) e_drain = clock64() - t0;
        }
        // operand of the normalisation GEMM -> this group's staging tile (16 bytes at a time)
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          float a2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * k4 + e;
            // same expression order as conv_tc_kernel / conv_tcp_kernel: the operand is re-rounded to TF32, so a last-bit
            // difference here would show up as a TF32 ulp in the normalisation GEMM
            if constexpr (EPI == ICADV_EPI_GDN_BWD) {
              a2[e] = round_tf32(v[k] * yv[k] * sv[k] * sv[k]);
            } else {
              const float s2 = sv[k] * sv[k];
              a2[e] = s2 > 0.f ? round_tf32(fast_div(v[k] * yv[k], s2)) : 0.f;
            }
          }
          *reinterpret_cast<float4*>(obuf + sw128_off(row, k4)) = make_float4(a2[0], a2[1], a2[2], a2[3]);
        }
        fence_proxy_async_smem();
        // the products pass 2 needs: g * sc over the accumulator chunk, y / sc in the stash
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float xs = sv[k] > 0.f ? fast_div(yv[k], sv[k]) : 0.f;
          v[k] = v[k] * sv[k];
          yv[k] = xs;
        }
        tmem_st32(t_acc + c * 32, v);
        tmem_st32(t_stash + c * 32, yv);
        tmem_st_wait();
        mbar_arrive(&a2_ready[grp]);
      }
      // ---- pass 2: out = g sc -+ (y / sc) (gamma^T t), from TMEM only
      const long long e2 = eprof ? clock64() : 0;
      mbar_wait(norm_full, nfull_par);
      nfull_par ^= 1;
      tc_fence_after_sync();
      const long long e3 = eprof ? clock64() : 0;
      for (int j = 0; j < half; ++j) {
        const int c = grp * half + j;
        float v[32], xs[32], w[32];
        tmem_ld32(t_acc + c * 32, v);
        tmem_ld32(t_stash + c * 32, xs);
        tmem_ld32(t_norm + c * 32, w);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          v[k] = v[k] + sign * xs[k] * w[k];
          if (p.round_out) v[k] = round_tf32(v[k]);
        }
        // out through this group's staging tile and one TMA store (a per-thread 128-byte row written with plain
        // stores costs the LSU 32 line transactions per instruction: measured 2 us per item)
        if (leader) tma_store_wait_read0();        // the previous chunk's store has finished reading the tile
        __syncwarp();
        named_bar_sync(bar_id, 128);
        write_row32(obuf, row, v);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (leader) { tma_store_4d(&p.out_map[it.cls], obuf, c * 32, it.j0, it.i0, it.img); tma_store_commit(); }
        __syncwarp();
      }
      // this item's TMEM reads are over: accumulator b goes back to the main issuer, the norm region to its issuer
      tc_fence_before_sync();
      named_bar_sync(bar_id, 128);
      if (leader) { mbar_arrive(&tmem_free[b]); mbar_arrive(norm_free); }
      __syncwarp();
      if (eprof) {   // per group: wait acc, pass 1 (+ norm wait), pass 2; slots 5 / 6: of pass 1, blocked on the saved-tensor ring
        long long* q = p.dbg + (int64_t)blockIdx.x * 16;
        const long long e4 = clock64();
        q[9 + grp * 3] += e1 - e0; q[10 + grp * 3] += e3 - e1; q[11 + grp * 3] += e4 - e3;
        q[5 + grp] += e_ys;
        if (grp == 0) q[15] += e_drain;
        if (grp == 0) q[3] += e3 - e2;
      }
    }
    if (leader) tma_store_wait0();
    __syncwarp();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

typedef void (*TcpKernelFn)(const TcpParams);

static TcpKernelFn pick_persistent(int epi) {
  switch (epi) {
    case ICADV_EPI_LINEAR: return conv_tcp_kernel<ICADV_EPI_LINEAR>;
    case ICADV_EPI_GDN_FWD: return conv_tcp_kernel<ICADV_EPI_GDN_FWD>;
    case ICADV_EPI_IGDN_FWD: return conv_tcp_kernel<ICADV_EPI_IGDN_FWD>;
    case ICADV_EPI_GDN_BWD: return conv_tcp_kernel<ICADV_EPI_GDN_BWD>;
    case ICADV_EPI_IGDN_BWD: return conv_tcp_kernel<ICADV_EPI_IGDN_BWD>;
    case kEpiCol2im: return conv_tcp_kernel<kEpiCol2im>;
    default: return nullptr;
  }
}

static TcpKernelFn pick_streaming(int epi) {
  switch (epi) {
    case ICADV_EPI_GDN_BWD: return conv_tcb_kernel<ICADV_EPI_GDN_BWD>;
    case ICADV_EPI_IGDN_BWD: return conv_tcb_kernel<ICADV_EPI_IGDN_BWD>;
    default: return nullptr;
  }
}

typedef void (*TcKernelFn)(const TcParams);

static TcKernelFn pick_kernel(int epi, int from_in) {
  switch (epi * 2 + (from_in ? 1 : 0)) {
    case ICADV_EPI_LINEAR * 2 + 0: return conv_tc_kernel<ICADV_EPI_LINEAR, false>;
    case ICADV_EPI_LINEAR * 2 + 1: return conv_tc_kernel<ICADV_EPI_LINEAR, true>;
    case ICADV_EPI_GDN_FWD * 2 + 0: return conv_tc_kernel<ICADV_EPI_GDN_FWD, false>;
    case ICADV_EPI_GDN_FWD * 2 + 1: return conv_tc_kernel<ICADV_EPI_GDN_FWD, true>;
    case ICADV_EPI_IGDN_FWD * 2 + 0: return conv_tc_kernel<ICADV_EPI_IGDN_FWD, false>;
    case ICADV_EPI_IGDN_FWD * 2 + 1: return conv_tc_kernel<ICADV_EPI_IGDN_FWD, true>;
    case ICADV_EPI_GDN_BWD * 2 + 0: return conv_tc_kernel<ICADV_EPI_GDN_BWD, false>;
    case ICADV_EPI_GDN_BWD * 2 + 1: return conv_tc_kernel<ICADV_EPI_GDN_BWD, true>;
    case ICADV_EPI_IGDN_BWD * 2 + 0: return conv_tc_kernel<ICADV_EPI_IGDN_BWD, false>;
    case ICADV_EPI_IGDN_BWD * 2 + 1: return conv_tc_kernel<ICADV_EPI_IGDN_BWD, true>;
    case kEpiCol2im * 2 + 0: return conv_tc_kernel<kEpiCol2im, false>;
    default: return nullptr;
  }
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

// 4-D channels-last view: dims (C, W, H, N) with arbitrary pixel strides (in floats); box [32, box_w, box_h, 1]
// (the tile itself for outputs, the tile plus its halo for input patches)
static int encode_nhwc(CUtensorMap* m, const float* base, int C, int W, int H, int N, int64_t sw, int64_t sh,
                       int64_t sn, int box_w = kTW, int box_h = kTH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
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
static int encode_plane(CUtensorMap* m, const float* base, int C, int W, int H, int N, int s, int a, int b,
                        int box_w = kTW, int box_h = kTH) {
  const int Wp = (W - b + s - 1) / s, Hp = (H - a + s - 1) / s;
  if (Wp <= 0 || Hp <= 0) { set_error("empty parity plane"); return ICADV_EINVAL; }
  return encode_nhwc(m, base + ((int64_t)a * W + b) * C, C, Wp, Hp, N, (int64_t)s * C, (int64_t)s * W * C,
                     (int64_t)H * W * C, box_w, box_h);
}

// padded RGB0 input of the first-layer form: [n_img][in_h + 4][in_w + 8][4] floats, pixel (h, w) at (h+2, w+2)
static int encode_pad4(CUtensorMap* m, const float* base, int W, int H, int N, int box_h = kTH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return ICADV_ECUDA; }
  const int64_t Wp = W + 8, Hp = H + 4;
  // dims: 32 floats (8 px x 4 ch, overlapping windows) | output column (2 px apart) | output row (2 rows apart)
  //       | row parity | image
  cuuint64_t dims[5] = {32, (cuuint64_t)(W / 2), (cuuint64_t)(Hp / 2), 2, (cuuint64_t)N};
  cuuint64_t strides[4] = {32, (cuuint64_t)(2 * Wp * 16), (cuuint64_t)(Wp * 16), (cuuint64_t)(Hp * Wp * 16)};
  cuuint32_t box[5] = {32, (cuuint32_t)kTW, (cuuint32_t)box_h, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(5d pad4 W=%d H=%d N=%d) failed: %d", W, H, N, (int)r);
    return ICADV_ECUDA;
  }
  return ICADV_OK;
}

}  // namespace icadv

using namespace icadv;

struct icadv_conv_plan {
  int n_launch;
  TcKernelFn fn;
  TcParams params[4];
  dim3 grid[4];
  int smem_bytes[4];
  // persistent variant (one launch covers every parity class)
  int persistent;
  TcpKernelFn pfn;
  TcpParams pp;
  int pgrid, psmem;
};

// largest multiple of 32 that divides n and is <= 256 (0 if none)
static int n_tile_of(int n) {
  for (int t = 256; t >= 32; t -= 32)
    if (n % t == 0) return t;
  return 0;
}

enum TcMode { kModeNone = 0, kModeGeneric = 1, kModeRgbIn = 2, kModeCol2im = 3 };

static int tc_mode(const icadv_conv_desc* d, bool report) {
#define TC_REQ(cond, msg)                     \
  do {                                        \
    if (!(cond)) {                            \
      if (report) set_error("conv_tc: " msg); \
      return kModeNone;                       \
    }                                         \
  } while (0)
  TC_REQ(d != nullptr, "null descriptor");
  TC_REQ(d->epi >= ICADV_EPI_LINEAR && d->epi <= ICADV_EPI_IGDN_BWD, "bad epilogue");
  if (d->in_pad4) {   // first-layer form: 3 input channels in the padded RGB0 layout, 5x5 stride 2
    TC_REQ(d->form == ICADV_FORM_SCONV && d->ksize == 5 && d->stride == 2 && d->k_ch == 3, "in_pad4 needs a 5x5/2 SCONV with k_ch = 3");
    TC_REQ(d->in_h % 2 == 0 && d->in_w % 2 == 0, "in_pad4 needs even image sizes");
    TC_REQ(d->n_ch % 32 == 0 && d->n_ch >= 32 && d->n_ch <= 256, "n_ch must be a multiple of 32 in [32,256]");
    if (d->epi != ICADV_EPI_LINEAR) TC_REQ(2 * d->n_ch <= 512, "GDN epilogue needs 2*n_ch <= 512 TMEM columns");
    TC_REQ(!d->acc_from_in, "in_pad4 excludes acc_from_in");
    return kModeRgbIn;
  }
  if (d->form == ICADV_FORM_TCONV && d->ksize == 5 && d->stride == 2 && d->n_ch <= 4 && d->k_ch % 32 == 0 &&
      d->k_ch >= 32 && d->epi == ICADV_EPI_LINEAR && d->act == ICADV_ACT_NONE && !d->acc_from_in)
    return kModeCol2im;   // narrow-output transposed conv (RGB end layer): 1x1 GEMM + col2im epilogue
  TC_REQ(d->k_ch % 32 == 0 && d->k_ch >= 32, "k_ch must be a multiple of 32");
  TC_REQ(d->n_ch % 32 == 0 && d->n_ch >= 32, "n_ch must be a multiple of 32");
  if (d->n_ch > 256) {   // N is tiled over blockIdx.z: linear epilogue only
    TC_REQ(d->epi == ICADV_EPI_LINEAR && !d->acc_from_in && n_tile_of(d->n_ch) > 0, "n_ch > 256 needs a linear epilogue and a tile size");
  }
  if (d->epi != ICADV_EPI_LINEAR) TC_REQ(2 * d->n_ch <= 512, "GDN epilogue needs 2*n_ch <= 512 TMEM columns");
  if (d->acc_from_in) TC_REQ(d->k_ch == d->n_ch && d->ksize == 1 && d->stride == 1, "acc_from_in needs a 1x1 identity geometry");
  TC_REQ(d->ksize * d->ksize <= kMaxTaps, "too many taps");
#undef TC_REQ
  return kModeGeneric;
}

// ---- persistent variant: eligibility + parameter block.  Returns 1 if built, 0 if this shape stays on the per-tile
//      kernel, < 0 on error.
static int build_persistent(const icadv_conv_desc* d, int mode, const Geometry& g, icadv_conv_plan* plan) {
  // ICADV_TC_PERSIST: 0 = never, 2 = every eligible shape (developer), unset / 1 = where it measured faster on B200
  // (profiles/r1_step_breakdown_v*.json): stride-2 transposed convs (four short parity classes in one launch, epilogue
  // overlapped: forward 3.96 -> 3.45 ms on g_s.4, backward 5.07 -> 4.3 ms on the g_a.2 dgrad), linear epilogues, the
  // col2im end layers (2.1 -> 1.27 ms) and the forward RGB first layer (2.12 -> 1.68 ms).  A single MMA stream per SM
  // runs at ~2x the tensor floor for N = 128, so long stride-2 convs (forward and backward) and the RGB dgrad layer
  // stay on the two-CTA-per-SM kernel.
  const char* env = getenv("ICADV_TC_PERSIST");
  const int level = env != nullptr ? atoi(env) : 1;
  if (level == 0) return 0;
  if (!(mode == kModeGeneric || mode == kModeRgbIn || mode == kModeCol2im) || d->acc_from_in) return 0;
  const bool c2i = mode == kModeCol2im;
  const bool gdn = d->epi != ICADV_EPI_LINEAR;
  const bool bwd = d->epi == ICADV_EPI_GDN_BWD || d->epi == ICADV_EPI_IGDN_BWD;
  // col2im: Z has 25 * n_ch columns; 80 (the next multiple of 16 above 75) for the RGB case keeps the resident weights
  // at 10 KB per K-block, which pays for a fifth patch slot
  const int N = c2i ? (25 * d->n_ch <= 80 && 25 * d->n_ch > 64 ? 80 : 96) : d->n_ch, K = d->k_ch, s = d->stride;
  if (N > 256 || (gdn && N > 128)) return 0;            // 2 TMEM buffers x [acc N | norm N] must fit 512 columns
  const bool tconv2 = !c2i && d->form == ICADV_FORM_TCONV && s == 2;
  // ICADV_TC_STREAM_BWD: 0 = never; unset / 1 = where it measured faster on B200 (profiles/r2_stream_bwd_ab.txt): the
  // HBM-bound backward epilogues -- the RGB first-layer form (g_s.6 input gradient + IGDN backward: 2.84 -> 2.3 ms at 64
  // images) and stride-2 transposed convs (g_a.2 / g_a.4 input gradients + GDN backward: 3.29 -> 2.9 ms); 2 = every
  // GDN / IGDN backward epilogue it takes (stride-2 convs with their long main loops stay faster on the two-CTA kernel)
  const char* senv = getenv("ICADV_TC_STREAM_BWD");
  const int stream_level = senv != nullptr ? atoi(senv) : 1;
  const bool stream = bwd && stream_level > 0 && N <= 128 && (N / 32) % 2 == 0 &&
                      (mode == kModeGeneric || mode == kModeRgbIn) &&
                      (stream_level >= 2 || mode == kModeRgbIn || tconv2);
  if (level == 1 && !stream && !(tconv2 || (!gdn && mode == kModeGeneric) || c2i || (mode == kModeRgbIn && !bwd))) return 0;
  TcpParams& p = plan->pp;
  memset(&p, 0, sizeof(p));
  p.n_class = tconv2 ? g.n_launch : 1;
  if (!tconv2 && !c2i && g.n_launch != 1) return 0;
  int rc = 0;
  int n_groups = 0, n_taps = 0, max_patch = kABytes;
  if (tconv2) {
    // one dense input map shared by the classes: a uniform box (the largest halo), per-class origin and tap offsets
    int pw = kTW, ph = kTH;
    for (int l = 0; l < p.n_class; ++l) {
      int dy0 = 1 << 14, dy1 = -(1 << 14), dx0 = 1 << 14, dx1 = -(1 << 14);
      for (int t = 0; t < g.n_taps[l]; ++t) {
        const Tap& tp = g.taps[l][t];
        dy0 = tp.dy < dy0 ? tp.dy : dy0; dy1 = tp.dy > dy1 ? tp.dy : dy1;
        dx0 = tp.dx < dx0 ? tp.dx : dx0; dx1 = tp.dx > dx1 ? tp.dx : dx1;
      }
      pw = kTW + dx1 - dx0 > pw ? kTW + dx1 - dx0 : pw;
      ph = kTH + dy1 - dy0 > ph ? kTH + dy1 - dy0 : ph;
    }
    for (int l = 0; l < p.n_class; ++l) {
      int dy0 = 1 << 14, dx0 = 1 << 14;
      for (int t = 0; t < g.n_taps[l]; ++t) {
        dy0 = g.taps[l][t].dy < dy0 ? g.taps[l][t].dy : dy0;
        dx0 = g.taps[l][t].dx < dx0 ? g.taps[l][t].dx : dx0;
      }
      if (n_groups == kMaxGroups || n_taps + g.n_taps[l] > kMaxTaps) return 0;
      Group& gr = p.groups[n_groups];
      gr.map = 0; gr.plane5 = 0; gr.dy0 = (int16_t)dy0; gr.dx0 = (int16_t)dx0;
      gr.sbo_bytes = pw * 128; gr.bytes = pw * ph * 128;
      gr.tap_begin = (int16_t)n_taps;
      for (int t = 0; t < g.n_taps[l]; ++t) {
        p.tap_aoff[n_taps] = ((g.taps[l][t].dy - dy0) * pw + (g.taps[l][t].dx - dx0)) * 128;
        p.tap_wtap[n_taps] = g.taps[l][t].wtap;
        ++n_taps;
      }
      gr.tap_end = (int16_t)n_taps;
      p.cls[l].g_begin = (int16_t)n_groups; p.cls[l].g_end = (int16_t)(n_groups + 1);
      p.cls[l].o_a = (int16_t)g.out_a[l]; p.cls[l].o_b = (int16_t)g.out_b[l];
      max_patch = gr.bytes > max_patch ? gr.bytes : max_patch;
      ++n_groups;
    }
    rc = encode_nhwc(&p.a_map[0], d->in, K, d->in_w, d->in_h, d->n_img, K, (int64_t)d->in_w * K,
                     (int64_t)d->in_h * d->in_w * K, pw, ph);
    if (rc) return rc;
    p.a_map[1] = p.a_map[2] = p.a_map[3] = p.a_map[0];
  } else {
    // single class: reuse the grouping of the per-tile plan (already built in plan->params[0])
    const TcParams& q = plan->params[0];
    n_groups = q.num_groups; n_taps = q.num_taps;
    for (int i = 0; i < n_groups; ++i) { p.groups[i] = q.groups[i]; max_patch = q.groups[i].bytes > max_patch ? q.groups[i].bytes : max_patch; }
    for (int i = 0; i < n_taps; ++i) { p.tap_aoff[i] = q.tap_aoff[i]; p.tap_wtap[i] = q.tap_wtap[i]; }
    for (int i = 0; i < 4; ++i) p.a_map[i] = q.a_map[i];
    p.cls[0].g_begin = 0; p.cls[0].g_end = (int16_t)n_groups; p.cls[0].o_a = 0; p.cls[0].o_b = 0;
    p.a_rank5 = q.a_rank5;
  }
  // ---- weights, gamma, per-class output maps
  if (mode == kModeRgbIn) rc = encode_mat(&p.w_map, d->wpack, 32, 5 * N, N);
  else if (c2i) rc = encode_mat(&p.w_map, d->wpack, K, d->ksize * d->ksize * d->n_ch, N);   // rows >= 25*n_ch read as zeros
  else rc = encode_mat(&p.w_map, d->wpack, K, d->ksize * d->ksize * N, N);
  if (!rc && gdn) rc = encode_mat(&p.g_map, d->gmat, N, N, N);
  if (rc) return rc;
  if (!gdn) p.g_map = p.w_map;
  for (int l = 0; l < 4; ++l) {
    const int ll = l < p.n_class ? l : 0;
    if (c2i) { p.out_map[l] = p.sc_map[l] = p.yprev_map[l] = p.w_map; continue; }   // col2im writes with plain stores
    if (tconv2) {
      rc = encode_plane(&p.out_map[l], d->out, N, g.out_w, g.out_h, d->n_img, 2, g.out_a[ll], g.out_b[ll]);
      if (!rc && gdn && !bwd) rc = encode_plane(&p.sc_map[l], d->out_scale, N, g.out_w, g.out_h, d->n_img, 2, g.out_a[ll], g.out_b[ll]);
      if (!rc && bwd) rc = encode_plane(&p.sc_map[l], d->sc_prev, N, g.out_w, g.out_h, d->n_img, 2, g.out_a[ll], g.out_b[ll]);
      if (!rc && bwd) rc = encode_plane(&p.yprev_map[l], d->y_prev, N, g.out_w, g.out_h, d->n_img, 2, g.out_a[ll], g.out_b[ll]);
    } else {
      rc = encode_nhwc(&p.out_map[l], d->out, N, g.out_w, g.out_h, d->n_img, N, (int64_t)g.out_w * N,
                       (int64_t)g.out_h * g.out_w * N);
      if (!rc && gdn && !bwd) rc = encode_nhwc(&p.sc_map[l], d->out_scale, N, g.out_w, g.out_h, d->n_img, N,
                                                (int64_t)g.out_w * N, (int64_t)g.out_h * g.out_w * N);
      if (!rc && bwd) rc = encode_nhwc(&p.sc_map[l], d->sc_prev, N, g.out_w, g.out_h, d->n_img, N,
                                       (int64_t)g.out_w * N, (int64_t)g.out_h * g.out_w * N);
      if (!rc && bwd) rc = encode_nhwc(&p.yprev_map[l], d->y_prev, N, g.out_w, g.out_h, d->n_img, N,
                                       (int64_t)g.out_w * N, (int64_t)g.out_h * g.out_w * N);
    }
    if (rc) return rc;
    if (!gdn) p.sc_map[l] = p.out_map[l];
    if (!bwd) p.yprev_map[l] = p.out_map[l];
  }
  p.k_chunks = mode == kModeRgbIn ? 1 : K / 32;
  p.n_ch = N; p.n_chunks = N / 32; p.n_total = N;
  p.tile_step_y = kTH; p.tile_step_x = kTW; p.tile_off = 0;
  if (c2i) {
    const TcParams& q = plan->params[0];
    p.tile_step_y = q.tile_step_y; p.tile_step_x = q.tile_step_x; p.tile_off = q.tile_off;
    p.c2i_in_h = q.c2i_in_h; p.c2i_in_w = q.c2i_in_w; p.c2i_nch = q.c2i_nch; p.c2i_out = q.c2i_out;
  }
  p.tiles_x = (g.tile_w + p.tile_step_x - 1) / p.tile_step_x; p.tiles_y = (g.tile_h + p.tile_step_y - 1) / p.tile_step_y;
  p.n_img = d->n_img;
  p.act = d->act; p.round_out = d->round_out_tf32;
  p.t_h = g.tile_h; p.t_w = g.tile_w; p.o_h = g.out_h; p.o_w = g.out_w; p.o_s = tconv2 ? 2 : 1;
  p.yprev = d->y_prev; p.scprev = d->sc_prev; p.bias = d->bias; p.beta = d->beta;
  p.active = d->active; p.n_active = d->n_active;
  p.out = d->out;
  p.dbg_skip_w = (getenv("ICADV_TC_DBG") != nullptr && (atoi(getenv("ICADV_TC_DBG")) & 16)) ? 1 : 0;
  if (stream) {
    // ---- streaming backward kernel: [patch ring | weight ring | gamma^T ring | saved-tensor ring R x 32 KB | staging | barriers + bias]
    p.patch_bytes = (max_patch + 1023) & ~1023;
    const int wbytes = N * 128, pair = 2 * kABytes;
    const int fixed = 1024 + kBarBlock + 2 * N * 4 + 2 * wbytes + 2 * kABytes;   // + gamma^T ring (2 stages) + 2 staging tiles
    int per_item = 0;   // patches of one item
    for (int l = 0; l < p.n_class; ++l) {
      const int n = (p.cls[l].g_end - p.cls[l].g_begin) * p.k_chunks;
      per_item = n > per_item ? n : per_item;
    }
    int P = 2, S = per_item <= 8 ? 2 : 3;
    int R = (kSmemLimit - fixed - P * p.patch_bytes - S * wbytes) / pair;
    R = R >= 4 ? 4 : (R >= 2 ? 2 : 0);   // power of two (slot = sequence number & (R - 1) in the kernel)
    if (R >= 2) {
      // left-over shared memory: more weight stages (long main loops), then a third patch slot
      while (S < 4 && fixed + P * p.patch_bytes + (S + 1) * wbytes + R * pair <= kSmemLimit) ++S;
      if (fixed + (P + 1) * p.patch_bytes + S * wbytes + R * pair <= kSmemLimit) ++P;
      p.num_patch = P; p.num_stages = S; p.ys_slots = R; p.grp_bytes = 0;
      plan->psmem = fixed + P * p.patch_bytes + S * wbytes + R * pair;
      if (plan->psmem < 120 * 1024) plan->psmem = 120 * 1024;
      int sms = 148, dev = 0;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const long long items = (long long)d->n_img * p.tiles_x * p.tiles_y * p.n_class;
      plan->pgrid = (int)(items < sms ? items : sms);
      plan->pfn = pick_streaming(d->epi);
      static std::once_flag sonce;
      static cudaError_t serr = cudaSuccess;
      std::call_once(sonce, [] {
        serr = cudaFuncSetAttribute(conv_tcb_kernel<ICADV_EPI_GDN_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
        if (serr == cudaSuccess)
          serr = cudaFuncSetAttribute(conv_tcb_kernel<ICADV_EPI_IGDN_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      });
      if (serr != cudaSuccess) { set_error("cudaFuncSetAttribute(streaming) failed: %s", cudaGetErrorString(serr)); return ICADV_ECUDA; }
      plan->persistent = 1;
      return 1;
    }
    if (level == 1 && !(tconv2 || (mode == kModeRgbIn && !bwd))) return 0;   // not enough shared memory: previous policy
  }
  // ---- shared memory: [patch ring | weight ring | gamma | 2 groups x 2 x 16 KB | barriers + bias/beta]
  p.patch_bytes = (max_patch + 1023) & ~1023;
  const int wbytes = N * 128;
  const int gbytes = gdn ? N * N * 4 : 0;
  p.grp_bytes = c2i ? ((128 * kZStride4 * 4 + 1023) & ~1023) : 2 * kABytes;
  const int fixed = 1024 + kBarBlock + 2 * N * 4 + gbytes + 2 * p.grp_bytes;
  int groups_per_item = 0;
  for (int l = 0; l < p.n_class; ++l) {
    const int n = (p.cls[l].g_end - p.cls[l].g_begin) * p.k_chunks;
    groups_per_item = n > groups_per_item ? n : groups_per_item;
  }
  int P, S;
  if (groups_per_item <= 8) {
    // short main loops (RGB end layers, 1x1): the input is the DRAM stream -> deep patch ring, 3 weight stages
    S = 3;
    P = (kSmemLimit - fixed - S * wbytes) / p.patch_bytes;
    if (P > kMaxPatch) P = kMaxPatch;
    if (P < 2) { P = 2; S = (kSmemLimit - fixed - P * p.patch_bytes) / wbytes; }
  } else {
    P = 2;
    S = (kSmemLimit - fixed - P * p.patch_bytes) / wbytes;
    if (S > kMaxStages) S = kMaxStages;
    while (P < 3 && S > 4 && fixed + (P + 1) * p.patch_bytes + (S - 1) * wbytes <= kSmemLimit) { ++P; --S; }
  }
  if (S > kMaxStages) S = kMaxStages;
  if (S < 2) return 0;
  if (c2i) {
    // the weights of the 1x1 GEMM (k_chunks boxes) stay resident; everything else goes to the patch ring
    if (n_groups != 1 || n_taps != 1 || p.k_chunks > kMaxStages) return 0;
    S = p.k_chunks;
    P = (kSmemLimit - fixed - S * wbytes) / p.patch_bytes;
    if (P > kMaxPatch) P = kMaxPatch;
    if (P < 3) return 0;
    p.w_resident = 1;
  }
  p.num_patch = P; p.num_stages = S;
  plan->psmem = fixed + P * p.patch_bytes + S * wbytes;
  if (plan->psmem < 120 * 1024) plan->psmem = 120 * 1024;   // one CTA per SM: each allocates all 512 TMEM columns
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)d->n_img * p.tiles_x * p.tiles_y * p.n_class;
  plan->pgrid = (int)(items < sms ? items : sms);
  plan->pfn = pick_persistent(c2i ? kEpiCol2im : d->epi);
  if (plan->pfn == nullptr) return 0;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    for (int e = 0; e <= kEpiCol2im && err == cudaSuccess; ++e)
      err = cudaFuncSetAttribute(pick_persistent(e), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
  });
  if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(persistent) failed: %s", cudaGetErrorString(err)); return ICADV_ECUDA; }
  plan->persistent = 1;
  return 1;
}

extern "C" {

int icadv_conv_tc_supported(const icadv_conv_desc* d) { return tc_mode(d, false) != kModeNone ? 1 : 0; }

int icadv_conv_plan_create(const icadv_conv_desc* d, icadv_conv_plan** out_plan) {
  ICADV_REQUIRE(out_plan != nullptr, "null plan pointer");
  *out_plan = nullptr;
  const int mode = tc_mode(d, true);
  if (mode == kModeNone) return ICADV_EINVAL;
  ICADV_REQUIRE(d->in && d->out && (d->acc_from_in || d->wpack), "null tensor pointer");
  const bool gdn = d->epi != ICADV_EPI_LINEAR;
  const bool bwd = d->epi == ICADV_EPI_GDN_BWD || d->epi == ICADV_EPI_IGDN_BWD;
  if (gdn) ICADV_REQUIRE(d->gmat != nullptr, "GDN epilogue needs gmat");
  if (gdn && !bwd) ICADV_REQUIRE(d->beta && d->out_scale, "GDN forward needs beta and out_scale");
  if (bwd) ICADV_REQUIRE(d->y_prev && d->sc_prev, "GDN backward needs y_prev and sc_prev");
  ICADV_REQUIRE(d->n_img <= 65535, "conv_tc: n_img too large");
  Geometry g;
  int rc = make_geometry(d, &g);
  if (rc) return rc;
  if (d->form == ICADV_FORM_TCONV && !d->acc_from_in && g.n_launch > 1) {
    // Output-parity classes no tap reaches (the input gradient of a strided 1x1 conv: 3 of its 4 classes) are dropped:
    // their pixels are NOT written (the caller pre-fills them with the bias / zero, see icadv.h).  Left in, such a class
    // would run an epilogue on an accumulator no MMA ever wrote.
    int keep = 0;
    for (int l = 0; l < g.n_launch; ++l) {
      if (g.n_taps[l] == 0) continue;
      if (keep != l) {
        g.n_taps[keep] = g.n_taps[l]; g.out_a[keep] = g.out_a[l]; g.out_b[keep] = g.out_b[l];
        for (int t = 0; t < g.n_taps[l]; ++t) g.taps[keep][t] = g.taps[l][t];
      }
      ++keep;
    }
    ICADV_REQUIRE(keep > 0, "conv_tc: no output class has a tap");
    for (int l = keep; l < 4; ++l) g.n_taps[l] = 0;
    g.n_launch = keep;
  }
  int dev_rc = icadv_check_device();
  if (dev_rc) return dev_rc;

  icadv_conv_plan* plan = new (std::nothrow) icadv_conv_plan();
  if (!plan) { set_error("out of host memory"); return ICADV_ENOMEM; }
  const int K = d->k_ch, taps_total = d->ksize * d->ksize;
  const int s = d->stride;
  const int n_total = mode == kModeCol2im ? 96 : d->n_ch;
  const int N = n_total > 256 ? n_tile_of(n_total) : n_total;   // MMA N (one N-tile)
  const int n_tiles = n_total / N;
  plan->n_launch = mode == kModeCol2im ? 1 : g.n_launch;
  for (int l = 0; l < plan->n_launch; ++l) {
    TcParams& p = plan->params[l];
    memset(&p, 0, sizeof(p));
    p.tile_step_y = kTH; p.tile_step_x = kTW; p.tile_off = 0;
    // ---- taps of this launch, grouped into halo patches: one group per input parity plane (stride-2 SCONV),
    //      per kernel row (RGB first-layer form), else a single group
    Tap taps[kMaxTaps];
    int n_taps = 0;
    if (mode == kModeRgbIn) {
      // one K-block per kernel row kh: padded input row 2*oh + kh -> (parity kh & 1, half-row oh + (kh >> 1)).
      // Kernel rows of the same parity read the same half-rows shifted by kh >> 1: ONE halo patch per row parity
      // (2 boxes of 18 half-rows per tile instead of 5 boxes of 16)
      for (int kh = 0; kh < 5; ++kh) {
        Tap t; t.plane = (int16_t)(kh & 1); t.dy = (int16_t)(kh >> 1); t.dx = 0; t.wtap = (int16_t)kh;
        taps[n_taps++] = t;
      }
    } else if (mode == kModeCol2im) {
      Tap t; t.plane = 0; t.dy = 0; t.dx = 0; t.wtap = 0;
      taps[n_taps++] = t;
      p.tile_step_y = kTH - 2; p.tile_step_x = kTW - 2; p.tile_off = -1;   // 1-pixel halo, interior 14 x 6
      p.c2i_in_h = d->in_h; p.c2i_in_w = d->in_w; p.c2i_nch = d->n_ch; p.c2i_out = d->out;
    } else if (!d->acc_from_in) {
      n_taps = g.n_taps[l];
      for (int t = 0; t < n_taps; ++t) taps[t] = g.taps[l][t];
    }
    int box_w[kMaxGroups], box_h[kMaxGroups];
    p.num_groups = 0; p.num_taps = 0;
    int max_patch = kABytes;
    for (int t = 0; t < n_taps; ++t) {
      bool seen = false;
      for (int u = 0; u < t; ++u) seen = seen || taps[u].plane == taps[t].plane;
      if (seen) continue;
      if (p.num_groups == kMaxGroups) { delete plan; set_error("conv_tc: too many tap groups"); return ICADV_EINVAL; }
      const int plane = taps[t].plane;
      int dy0 = 1 << 14, dy1 = -(1 << 14), dx0 = 1 << 14, dx1 = -(1 << 14);
      for (int u = t; u < n_taps; ++u)
        if (taps[u].plane == plane) {
          dy0 = taps[u].dy < dy0 ? taps[u].dy : dy0; dy1 = taps[u].dy > dy1 ? taps[u].dy : dy1;
          dx0 = taps[u].dx < dx0 ? taps[u].dx : dx0; dx1 = taps[u].dx > dx1 ? taps[u].dx : dx1;
        }
      Group& gr = p.groups[p.num_groups];
      const int pw = kTW + dx1 - dx0, ph = kTH + dy1 - dy0;
      gr.map = (int16_t)(mode == kModeRgbIn ? 0 : plane);
      gr.plane5 = (int16_t)(plane & 1);
      gr.dy0 = (int16_t)dy0; gr.dx0 = (int16_t)dx0;
      gr.sbo_bytes = pw * 128; gr.bytes = pw * ph * 128;
      gr.tap_begin = (int16_t)p.num_taps;
      for (int u = t; u < n_taps; ++u)
        if (taps[u].plane == plane) {
          p.tap_aoff[p.num_taps] = ((taps[u].dy - dy0) * pw + (taps[u].dx - dx0)) * 128;
          p.tap_wtap[p.num_taps] = taps[u].wtap;
          ++p.num_taps;
        }
      gr.tap_end = (int16_t)p.num_taps;
      box_w[p.num_groups] = pw; box_h[p.num_groups] = ph;
      if (pw > 256 || ph > 256) { delete plan; set_error("conv_tc: patch exceeds the TMA box limit"); return ICADV_EINVAL; }
      max_patch = gr.bytes > max_patch ? gr.bytes : max_patch;
      ++p.num_groups;
    }
    // ---- input maps: one per group, box = tile + halo
    if (mode == kModeRgbIn) {
      int ph_max = kTH;   // both row parities go through one 5-D map: a uniform box (the taller halo)
      for (int gi = 0; gi < p.num_groups; ++gi) ph_max = box_h[gi] > ph_max ? box_h[gi] : ph_max;
      for (int gi = 0; gi < p.num_groups; ++gi) p.groups[gi].bytes = kTW * ph_max * 128;
      max_patch = kTW * ph_max * 128 > max_patch ? kTW * ph_max * 128 : max_patch;
      rc = encode_pad4(&p.a_map[0], d->in, d->in_w, d->in_h, d->n_img, ph_max);
      if (rc) { delete plan; return rc; }
      p.a_map[1] = p.a_map[2] = p.a_map[3] = p.a_map[0];
      p.a_rank5 = 1;
    } else if (d->acc_from_in) {
      // no contraction: the maps are placeholders (set below)
    } else if (d->form == ICADV_FORM_SCONV && s == 2) {
      bool have[4] = {false, false, false, false};
      for (int gi = 0; gi < p.num_groups; ++gi) {
        const int plane = p.groups[gi].map, a = plane >> 1, b = plane & 1;
        if (d->in_h <= a || d->in_w <= b) {   // this parity plane is empty: its taps only ever read zeros
          rc = encode_plane(&p.a_map[plane], d->in, K, d->in_w, d->in_h, d->n_img, 2, 0, 0, box_w[gi], box_h[gi]);
          p.groups[gi].dy0 = (int16_t)(1 << 13);   // far outside: the whole patch is out-of-bounds fill
        } else {
          rc = encode_plane(&p.a_map[plane], d->in, K, d->in_w, d->in_h, d->n_img, 2, a, b, box_w[gi], box_h[gi]);
        }
        if (rc) { delete plan; return rc; }
        have[plane] = true;
      }
      for (int q = 0; q < 4; ++q)
        if (!have[q]) p.a_map[q] = p.a_map[p.groups[0].map];
    } else {
      rc = encode_nhwc(&p.a_map[0], d->in, K, d->in_w, d->in_h, d->n_img, K, (int64_t)d->in_w * K,
                       (int64_t)d->in_h * d->in_w * K, box_w[0], box_h[0]);
      if (rc) { delete plan; return rc; }
      p.a_map[1] = p.a_map[2] = p.a_map[3] = p.a_map[0];
    }
    // ---- output-side maps (dense for SCONV, parity plane for TCONV); col2im writes with plain stores
    auto out_side = [&](CUtensorMap* m, const float* base) -> int {
      if (d->form == ICADV_FORM_TCONV && s == 2)
        return encode_plane(m, base, n_total, g.out_w, g.out_h, d->n_img, 2, g.out_a[l], g.out_b[l]);
      return encode_nhwc(m, base, n_total, g.out_w, g.out_h, d->n_img, n_total, (int64_t)g.out_w * n_total,
                         (int64_t)g.out_h * g.out_w * n_total);
    };
    if (mode == kModeCol2im) {
      rc = encode_mat(&p.w_map, d->wpack, K, taps_total * d->n_ch, N);   // rows >= 25*n_ch read as zeros
      if (rc) { delete plan; return rc; }
      p.out_map = p.sc_map = p.yprev_map = p.scprev_map = p.g_map = p.w_map;
    } else {
      rc = out_side(&p.out_map, d->out);
      if (!rc && gdn && !bwd) rc = out_side(&p.sc_map, d->out_scale);
      if (!rc && bwd) rc = out_side(&p.yprev_map, d->y_prev);
      if (!rc && bwd) rc = out_side(&p.scprev_map, d->sc_prev);
      if (!rc && !d->acc_from_in) {
        if (mode == kModeRgbIn) rc = encode_mat(&p.w_map, d->wpack, 32, 5 * N, N);   // [5 kh][N][32]
        else rc = encode_mat(&p.w_map, d->wpack, K, taps_total * n_total, N);
      }
      if (!rc && gdn) rc = encode_mat(&p.g_map, d->gmat, N, N, N);
      if (rc) { delete plan; return rc; }
      if (!(gdn && !bwd)) p.sc_map = p.out_map;
      if (!bwd) { p.yprev_map = p.out_map; p.scprev_map = p.out_map; }
      if (d->acc_from_in) { p.w_map = p.out_map; p.a_map[0] = p.a_map[1] = p.a_map[2] = p.a_map[3] = p.out_map; }
      if (!gdn) p.g_map = p.out_map;
    }
    p.k_chunks = mode == kModeRgbIn ? 1 : K / 32;
    p.n_ch = N; p.n_chunks = N / 32; p.n_total = n_total;
    p.tiles_x = (g.tile_w + p.tile_step_x - 1) / p.tile_step_x;
    p.tiles_y = (g.tile_h + p.tile_step_y - 1) / p.tile_step_y;
    p.epi = mode == kModeCol2im ? kEpiCol2im : d->epi;
    p.act = d->act; p.acc_from_in = d->acc_from_in; p.round_out = d->round_out_tf32;
    p.t_h = g.tile_h; p.t_w = g.tile_w; p.o_h = g.out_h; p.o_w = g.out_w;
    if (d->form == ICADV_FORM_TCONV && s == 2) { p.o_s = 2; p.o_a = g.out_a[l]; p.o_b = g.out_b[l]; }
    else { p.o_s = 1; p.o_a = 0; p.o_b = 0; }
    p.yprev = d->y_prev; p.scprev = d->sc_prev; p.xin = d->in;
    p.dbg_flags = getenv("ICADV_TC_DBG") ? atoi(getenv("ICADV_TC_DBG")) : 0;
    p.bwd_lookahead = bwd ? (getenv("ICADV_TC_LOOKAHEAD") ? atoi(getenv("ICADV_TC_LOOKAHEAD")) : 0) : 0;
    // ---- shared memory: [patch ring | weight ring | BWD staging | barriers + bias/beta].  Aim for two CTAs per SM
    //      (their prologue / epilogue overlap each other's main loop).
    p.patch_bytes = (max_patch + 1023) & ~1023;
    const int wbytes = N * 128;
    // long backward main loops: four weight stages matter more than prefetching the first saved chunk, so the staging
    // pair aliases the tail of the weight ring (measured: g_s.4 dgrad main loop 0.47 -> 0.33 us per K-block)
    const bool want_alias = bwd && !d->acc_from_in && mode == kModeGeneric && N * 128 >= kABytes &&
                            K / 32 * g.n_taps[l] >= 16 && getenv("ICADV_TC_NO_ALIAS") == nullptr;
    p.ld_alias = want_alias ? 1 : 0;
    p.ld_bufs = bwd && !want_alias ? 2 : 0;
    int fixed = 1024 + kBarBlock + 2 * N * 4 + p.ld_bufs * kABytes;
    // bytes of ring the epilogue aliases for its store staging (16 KB regions) / the col2im scatter tile
    const int need_ring = mode == kModeCol2im ? 128 * kZStride * 4 : ((gdn && !bwd) ? 4 : 2) * kABytes;
    // a short main loop (the RGB end layers: 4-5 K-blocks) is one DRAM round trip if every patch is in flight at once
    const int main_patches = p.k_chunks * p.num_groups;
    const int short_p = (main_patches >= 3 && main_patches <= 8) ? (main_patches > kMaxPatch ? kMaxPatch : main_patches) : 2;
    auto fit = [&](int budget, int max_p, int max_s) -> bool {
      int P = 2, S = (budget - fixed - P * p.patch_bytes) / wbytes;
      if (S > max_s) S = max_s;
      if (S < 2) return false;
      if (short_p > 2) {   // trade weight stages for patch slots
        while (P < short_p && fixed + (P + 1) * p.patch_bytes + 2 * wbytes <= budget) ++P;
        S = (budget - fixed - P * p.patch_bytes) / wbytes;
        if (S > max_s) S = max_s;
      }
      while (P < max_p && fixed + (P + 1) * p.patch_bytes + S * wbytes <= budget) ++P;
      while (fixed + P * p.patch_bytes + S * wbytes - fixed < need_ring) {   // grow the rings for the staging
        if (S < kMaxStages && fixed + P * p.patch_bytes + (S + 1) * wbytes <= budget) ++S;
        else return false;
      }
      p.num_patch = P; p.num_stages = S;
      return true;
    };
    // TMEM columns: accumulator N (+ normalisation product N); a stand-alone (I)GDN launch (accumulator read from global
    // memory) keeps only the normalisation product there, so N = 192 still leaves room for a second CTA on the SM
    const int tm_cols = d->acc_from_in ? N : (gdn ? 2 * N : N);
    const bool want_two = tm_cols <= 256 && getenv("ICADV_TC_ONE_CTA") == nullptr;
    const int env_s = getenv("ICADV_TC_S") ? atoi(getenv("ICADV_TC_S")) : 0;   // developer override: weight stages
    bool ok = (want_two && fit(kSmemTwoCta, 2, env_s > 0 ? env_s : 4)) || fit(kSmemLimit, 3, env_s > 0 ? env_s : kMaxStages);
    if (ok && p.ld_alias && p.num_stages < 4) {   // not enough stages to give two away: separate staging pair instead
      p.ld_alias = 0; p.ld_bufs = 2;
      fixed += 2 * kABytes;
      ok = (want_two && fit(kSmemTwoCta, 2, env_s > 0 ? env_s : 4)) || fit(kSmemLimit, 3, env_s > 0 ? env_s : kMaxStages);
    }
    if (!ok) {
      delete plan; set_error("conv_tc: not enough shared memory for n_ch=%d", N); return ICADV_EINVAL;
    }
    int cols = tm_cols, pow2 = 32;
    while (pow2 < cols) pow2 <<= 1;
    p.tmem_cols = pow2;
    p.bias = d->bias; p.beta = d->beta; p.active = d->active; p.n_active = d->n_active;
    plan->grid[l] = dim3(p.tiles_x * p.tiles_y, d->n_img, n_tiles);
    plan->smem_bytes[l] = fixed + p.num_patch * p.patch_bytes + p.num_stages * wbytes;
  }
  plan->fn = pick_kernel(plan->params[0].epi, d->acc_from_in);
  if (plan->fn == nullptr) { delete plan; set_error("conv_tc: no kernel for epi=%d from_in=%d", d->epi, d->acc_from_in); return ICADV_EINVAL; }
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (int e = 0; e <= kEpiCol2im && attr_err == cudaSuccess; ++e)
      for (int f = 0; f < 2 && attr_err == cudaSuccess; ++f) {
        TcKernelFn fn = pick_kernel(e, f);
        if (fn) attr_err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
      }
  });
  if (attr_err != cudaSuccess) {
    delete plan;
    set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
    return ICADV_ECUDA;
  }
  plan->persistent = 0;
  rc = build_persistent(d, mode, g, plan);
  if (rc < 0) { delete plan; return rc; }
  *out_plan = plan;
  return ICADV_OK;
}

int icadv_conv_plan_launch(const icadv_conv_plan* plan, icadv_stream_t stream) {
  ICADV_REQUIRE(plan != nullptr, "null plan");
  if (plan->persistent) {
    plan->pfn<<<plan->pgrid, kPThreads, plan->psmem, as_stream(stream)>>>(plan->pp);
    ICADV_CUDA_TRY(cudaGetLastError());
    return ICADV_OK;
  }
  for (int l = 0; l < plan->n_launch; ++l) {
    plan->fn<<<plan->grid[l], kThreads, plan->smem_bytes[l], as_stream(stream)>>>(plan->params[l]);
    ICADV_CUDA_TRY(cudaGetLastError());
  }
  return ICADV_OK;
}

int icadv_conv_plan_num_launches(const icadv_conv_plan* plan) {
  return plan ? (plan->persistent ? 1 : plan->n_launch) : 0;
}

/* developer profiling: per-CTA phase timestamps (16 x int64 per CTA of launch 0), NULL to disable */
int icadv_conv_plan_set_debug(icadv_conv_plan* plan, long long* dbg) {
  ICADV_REQUIRE(plan != nullptr, "null plan");
  plan->params[0].dbg = dbg;
  plan->pp.dbg = dbg;   // persistent variant: 16 counters per CTA (148 CTAs)
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
