// Developer probe: what does ONE MMA-issuing thread cost per K-block (four 128x128x8 tf32 MMAs = 256 tensor cycles) as a
// function of the ORDER of waits / descriptor formation / commits in its loop?  Same mini pipeline as tma_rate_probe.cu
// (producing threads fill a ring of 16 KB stages by TMA from an L2-resident set, the issuing thread consumes them).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o issue_probe issue_probe.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../imagecompression_adversarial_b200/csrc/icadv_ptx.cuh"
using namespace icadv;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define MMA4(AD, BD)                                                                   \
  tc_mma_tf32(tmem, (AD), (BD), idesc, 1u); tc_mma_tf32(tmem, (AD) + 2, (BD) + 2, idesc, 1u); \
  tc_mma_tf32(tmem, (AD) + 4, (BD) + 4, idesc, 1u); tc_mma_tf32(tmem, (AD) + 6, (BD) + 6, idesc, 1u)

// V: 0 naive (early test_wait of the next stage + conditional wait), 1 blocking wait per K-block, 2 pipelined order with one
// register set, 3 pipelined order with two register sets, 4 no waits (producers off) but commits, 5 bare (no waits, no
// commits), 6 two K-blocks per pass (two waits, eight MMAs, two commits), 7 as 1 with the wait AFTER the descriptors are
// in uniform registers (asm barrier), 8 as 6 with four K-blocks per pass
template <int V, int ONEWARP = 0>
__global__ void __launch_bounds__(160) probe(const __grid_constant__ CUtensorMap map, int kblocks, int S, int np, int n_boxes_total,
                                             long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* R = smem + 16384;
  uint64_t* full = reinterpret_cast<uint64_t*>(R + S * 16384);
  uint64_t* empty = full + 8;
  uint64_t* done = empty + 8;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (S + 1) * 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(done, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  constexpr bool producers = (V != 4 && V != 5);
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, 128);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_lo0 = ((smem_u32(R) >> 4) & 0x3FFFu) | (1u << 16);
      const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | a_lo;
      int s = 0;
      uint32_t par = 0, w_lo = w_lo0;
      auto advance = [&]() { if (++s == S) { s = 0; par ^= 1; w_lo = w_lo0; } else { w_lo += 1024u; } };
      const long long t0 = clock64();
      if constexpr (V == 0) {
        bool ok = false;
        for (int k = 0; k < kblocks; ++k) {
          const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
          if (!ok) mbar_wait(&full[s], par);
          uint64_t* wdone = &empty[s];
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd);
          tc_commit(wdone);
          ok = mbar_test_wait(&full[s], par);
        }
      } else if constexpr (V == 1 || V == 4 || V == 5) {
        for (int k = 0; k < kblocks; ++k) {
          const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
          if (V == 1) mbar_wait(&full[s], par);
          uint64_t* wdone = &empty[s];
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd);
          if (V != 5) tc_commit(wdone);
        }
      } else if constexpr (V == 2) {
        mbar_wait(&full[s], par);
        uint64_t bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
        uint64_t* wdone = &empty[s];
        advance();
        for (int k = 0; k < kblocks; ++k) {
          if (k + 1 < kblocks) mbar_wait(&full[s], par);
          const uint64_t bdn = (static_cast<uint64_t>(hi) << 32) | w_lo;
          uint64_t* wdn = &empty[s];
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd);
          tc_commit(wdone);
          bd = bdn; wdone = wdn;
        }
      } else if constexpr (V == 3) {
        mbar_wait(&full[s], par);
        uint64_t bd0 = (static_cast<uint64_t>(hi) << 32) | w_lo;
        uint64_t* done0 = &empty[s];
        advance();
        for (int k = 0; k < kblocks; k += 2) {
          mbar_wait(&full[s], par);
          const uint64_t bd1 = (static_cast<uint64_t>(hi) << 32) | w_lo;
          uint64_t* done1 = &empty[s];
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd0);
          tc_commit(done0);
          if (k + 2 < kblocks) mbar_wait(&full[s], par);
          bd0 = (static_cast<uint64_t>(hi) << 32) | w_lo;
          done0 = &empty[s];
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd1);
          tc_commit(done1);
        }
      } else if constexpr (V == 6) {
        for (int k = 0; k < kblocks; k += 2) {
          const uint64_t bd0 = (static_cast<uint64_t>(hi) << 32) | w_lo;
          uint64_t* d0 = &empty[s];
          mbar_wait(&full[s], par);
          advance();
          const uint64_t bd1 = (static_cast<uint64_t>(hi) << 32) | w_lo;
          uint64_t* d1 = &empty[s];
          mbar_wait(&full[s], par);
          advance();
          tc_fence_after_sync();
          MMA4(ad, bd0);
          tc_commit(d0);
          MMA4(ad, bd1);
          tc_commit(d1);
        }
      } else if constexpr (V == 8) {
        for (int k = 0; k < kblocks; k += 4) {
          uint64_t bd[4]; uint64_t* dn[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            bd[j] = (static_cast<uint64_t>(hi) << 32) | w_lo; dn[j] = &empty[s];
            mbar_wait(&full[s], par);
            advance();
          }
          tc_fence_after_sync();
#pragma unroll
          for (int j = 0; j < 4; ++j) { MMA4(ad, bd[j]); tc_commit(dn[j]); }
        }
      } else if constexpr (V == 7) {
        // the wait of stage k sits between the commit of k - 1 and the MMAs of k, as in 1, but is a bare try_wait spin
        for (int k = 0; k < kblocks; ++k) {
          const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
          while (!mbar_try_wait(&full[s], par)) {}
          uint64_t* wdone = &empty[s];
          advance();
          MMA4(ad, bd);
          tc_commit(wdone);
        }
      }
      tc_commit(done);
      mbar_wait(done, 0);
      out[blockIdx.x * 2] = clock64() - t0;
      out[blockIdx.x * 2 + 1] = 0;
    }
  } else if (ONEWARP && warp == 1 && producers) {
    // both producing threads live in ONE warp (lanes 0 .. np - 1 on divergent paths)
    const int me = threadIdx.x & 31;
    if (me < np) {
      int s = 0, owner = 0;
      uint32_t par = 1;
      for (int k = 0; k < kblocks; ++k) {
        if (owner == me) {
          mbar_wait(&empty[s], par);
          mbar_arrive_expect_tx(&full[s], 16384);
          const int box = (blockIdx.x * kblocks + k) & (n_boxes_total - 1);
          tma_load_2d(R + s * 16384, &map, &full[s], 0, box * 128);
        }
        if (++s == S) { s = 0; par ^= 1; }
        if (++owner == np) owner = 0;
      }
    }
  } else if (!ONEWARP && warp <= np && producers) {
    if (elect_one_sync()) {
      const int me = warp - 1;
      int s = 0, owner = 0;
      uint32_t par = 1;
      for (int k = 0; k < kblocks; ++k) {
        if (owner == me) {
          mbar_wait(&empty[s], par);
          mbar_arrive_expect_tx(&full[s], 16384);
          const int box = (blockIdx.x * kblocks + k) & (n_boxes_total - 1);
          tma_load_2d(R + s * 16384, &map, &full[s], 0, box * 128);
        }
        if (++s == S) { s = 0; par ^= 1; }
        if (++owner == np) owner = 0;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int V, int ONEWARP = 0>
void run(const CUtensorMap& map, int n_boxes, long long* d_out, const char* name) {
  for (int S : {3, 5, 8})
    for (int np : {1, 2}) {
      const int smem_p = (S + 1) * 16384 + 256 + 1024, kblocks = 4096;
      const int smem_use = smem_p < 120 * 1024 ? 120 * 1024 : smem_p;
      cudaFuncSetAttribute(probe<V, ONEWARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_use);
      for (int rep = 0; rep < 2; ++rep) probe<V, ONEWARP><<<148, 160, smem_use>>>(map, kblocks, S, np, n_boxes, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
      std::vector<long long> h(296);
      cudaMemcpy(h.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int b = 0; b < 148; ++b) cyc += (double)h[2 * b];
      printf("V%d %-58s S=%d producers=%d : %6.1f cycles per K-block\n", V, name, S, np, cyc / 148 / kblocks);
      if (V == 4 || V == 5) return;
    }
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  const size_t set = (size_t)64 << 20;
  float* buf;
  cudaMalloc(&buf, set);
  cudaMemset(buf, 0, set);
  long long* d_out;
  cudaMalloc(&d_out, 296 * 8);
  cuuint64_t dims[2] = {32, set / 512};
  cuuint64_t strd[1] = {512};
  cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
  CUtensorMap map;
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strd, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int nb = (int)(set / 512 / 128);
  run<5>(map, nb, d_out, "bare: no waits, no commits");
  run<4>(map, nb, d_out, "no waits, commit per K-block");
  run<0>(map, nb, d_out, "naive: early test_wait + conditional wait");
  run<1>(map, nb, d_out, "blocking wait per K-block");
  run<7>(map, nb, d_out, "bare try_wait spin per K-block, no fence");
  run<2>(map, nb, d_out, "pipelined order, one register set");
  run<3>(map, nb, d_out, "pipelined order, two register sets");
  run<6>(map, nb, d_out, "two K-blocks per pass (2 waits, 8 MMAs, 2 commits)");
  run<0, 1>(map, nb, d_out, "naive, producers = lanes of ONE warp");
  run<7, 1>(map, nb, d_out, "bare try_wait spin, producers = lanes of ONE warp");
  return 0;
}
