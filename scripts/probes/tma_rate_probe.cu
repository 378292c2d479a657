// Developer probe: how fast does ONE SM's TMA unit move 16 KB boxes [32 floats x 128 rows] (SWIZZLE_128B) from L2 / HBM into
// shared memory, as a function of the global-memory stride between the 128-byte box rows?
//   stride 128 B : the box is one contiguous 16 KB run            (a "blocked" [C/32][H][W][32] activation layout)
//   stride 512 B : rows 512 B apart, as the channels-last [H][W][128] activations and the K-major weights are read today
// One elected thread per CTA keeps `depth` loads in flight over a ring of shared-memory slots; nobody reads the data.
// Prints bytes per clock per SM and the aggregate TB/s for an L2-resident and an HBM-sized working set.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_rate_probe tma_rate_probe.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../imagecompression_adversarial_b200/csrc/icadv_ptx.cuh"
using namespace icadv;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, int boxes_per_cta, int depth, int n_boxes_total,
                                             int box_rows, int issuers, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * depth * 32768);
  if (threadIdx.x == 0) { for (int i = 0; i < depth * 2; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w < issuers) {
    if (elect_one_sync()) {
      bar += w * depth;
      smem += w * depth * 32768;
      boxes_per_cta /= issuers;
      const long long t0 = clock64();
      for (int i = 0; i < boxes_per_cta + depth; ++i) {
        const int s = i % depth;
        if (i >= depth) mbar_wait(&bar[s], ((i / depth) - 1) & 1);
        if (i < boxes_per_cta) {
          const int box = ((blockIdx.x * 2 + w) * boxes_per_cta + i) & (n_boxes_total - 1);   // power of two: no division in the loop
          mbar_arrive_expect_tx(&bar[s], box_rows * 128);
          tma_load_2d(smem + s * 32768, &map, &bar[s], 0, box * box_rows);
        }
      }
      if (w == 0) out[blockIdx.x] = clock64() - t0;
    }
  }
}


// Shared-memory contention between the tensor core's operand reads and TMA fills: warp 0 issues the bare tcgen05.mma stream
// (kind::tf32, M = 128, N = 128: 8 KB of operand reads per 64-cycle MMA = 128 B/clk), warps 1..nw stream 16 KB boxes into a
// separate ring of the same shared memory (each issuing thread ~33 B/clk, see above).
__global__ void __launch_bounds__(160) probe_mix(const __grid_constant__ CUtensorMap map, int iters, int nw, int n_boxes_total, int nmma,
                                                 long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                       // 2 x 16 KB
  uint8_t* B = smem + 2 * 16384;           // 2 x 32 KB (N up to 256)
  uint8_t* R = smem + 6 * 16384;           // TMA rings: 4 warps x 2 slots x 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(R + 8 * 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 12);
  volatile int* stop = reinterpret_cast<volatile int*>(tptr + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, nmma);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
        const uint32_t sel = static_cast<uint32_t>((i >> 2) & 1);
        const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (a_lo0 + sel * 1024u), bd = (static_cast<uint64_t>(hi) << 32) | (b_lo0 + sel * 2048u);
        tc_mma_tf32(tmem, ad, bd, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
      }
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      out[blockIdx.x * 2] = clock64() - t0;
      *stop = 1;
    }
  } else if (warp <= nw) {
    if (elect_one_sync()) {
      uint64_t* mybar = bar + 2 + (warp - 1) * 2;
      uint8_t* ring = R + (warp - 1) * 2 * 16384;
      long long n = 0;
      int i = 0;
      while (!*stop) {
        const int s = i & 1;
        if (i >= 2) mbar_wait(&mybar[s], ((i >> 1) - 1) & 1);
        const int box = ((blockIdx.x * 4 + warp) * 4096 + i) & (n_boxes_total - 1);
        mbar_arrive_expect_tx(&mybar[s], 16384);
        tma_load_2d(ring + s * 16384, &map, &mybar[s], 0, box * 128);
        ++i; ++n;
      }
      // drain the loads still in flight before the CTA exits
      for (int k = i > 2 ? i - 2 : 0; k < i; ++k) mbar_wait(&mybar[k & 1], (k >> 1) & 1);
      atomicAdd(reinterpret_cast<unsigned long long*>(&out[blockIdx.x * 2 + 1]), (unsigned long long)n);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}


// The kernels' core pipeline in miniature: `np` producing threads (alternate stages) fill a ring of S 16 KB stages with
// weight-like boxes, one issuing thread waits for a stage, issues four 128x128x8 MMAs (static A tile, B = the stage) and
// commits the stage back.  Reports cycles per K-block (floor 256).
__global__ void __launch_bounds__(160) probe_pipe(const __grid_constant__ CUtensorMap map, int kblocks, int S, int np, int n_boxes_total,
                                                  int mode, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                       // 16 KB static
  uint8_t* R = smem + 16384;               // S x 16 KB ring
  uint64_t* full = reinterpret_cast<uint64_t*>(R + S * 16384);
  uint64_t* empty = full + 8;   // up to 8 stages
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
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, 128);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_lo0 = ((smem_u32(R) >> 4) & 0x3FFFu) | (1u << 16);
      int s = 0;
      uint32_t par = 0, w_lo = w_lo0;
      bool ok = false;
      long long waited = 0;
      const long long t0 = clock64();
      for (int k = 0; k < kblocks; ++k) {
        const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | a_lo, bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
        if (!ok) { const long long w0 = clock64(); mbar_wait(&full[s], par); waited += clock64() - w0; }
        uint64_t* wdone = &empty[s];
        if (++s == S) { s = 0; par ^= 1; w_lo = w_lo0; } else { w_lo += 1024u; }
        tc_fence_after_sync();
        if (!(mode & 4)) {   // mode 4: no MMAs at all (pure TMA ring with a consumer)
          tc_mma_tf32(tmem, ad, bd, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
          tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
          if (mode & 8) {   // mode 8: eight MMAs per stage (floor 512 per stage)
            tc_mma_tf32(tmem, ad, bd, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
            tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
          }
        }
        if (mode & 2) mbar_arrive(wdone); else tc_commit(wdone);   // mode 2: plain arrive (stage handed back at once: WRONG for real use)
        ok = mbar_test_wait(&full[s], par);
      }
      tc_commit(done);
      mbar_wait(done, 0);
      out[blockIdx.x * 2] = clock64() - t0;
      out[blockIdx.x * 2 + 1] = waited;
    }
  } else if (warp <= np) {
    if (elect_one_sync()) {
      const int me = warp - 1;
      int s = 0;
      uint32_t par = 1;
      int owner = 0;
      for (int k = 0; k < kblocks; ++k) {
        if (owner == me) {
          if (!(mode & 1)) mbar_wait(&empty[s], par);             // mode 1: never wait for the stage to be free
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


// Same mini pipeline, issuing loop software-pipelined by hand: the NEXT stage's barrier is resolved and its descriptors
// are formed BEFORE the current stage's MMAs are issued (two K-blocks per pass so the two descriptor sets can live in
// different registers), so nothing but the MMAs themselves sits between two MMA groups.
__global__ void __launch_bounds__(160) probe_pipe2(const __grid_constant__ CUtensorMap map, int kblocks, int S, int np, int n_boxes_total,
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
      // prologue: stage 0 resolved, descriptor set 0 formed
      mbar_wait(&full[s], par);
      uint64_t bd0 = (static_cast<uint64_t>(hi) << 32) | w_lo;
      uint64_t* done0 = &empty[s];
      advance();
      for (int k = 0; k < kblocks; k += 2) {
        // resolve stage k + 1 and form its descriptors BEFORE issuing block k
        mbar_wait(&full[s], par);
        const uint64_t bd1 = (static_cast<uint64_t>(hi) << 32) | w_lo;
        uint64_t* done1 = &empty[s];
        advance();
        tc_fence_after_sync();
        tc_mma_tf32(tmem, ad, bd0, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd0 + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd0 + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd0 + 6, idesc, 1u);
        tc_commit(done0);
        // resolve stage k + 2 and form its descriptors BEFORE issuing block k + 1
        if (k + 2 < kblocks) mbar_wait(&full[s], par);
        bd0 = (static_cast<uint64_t>(hi) << 32) | w_lo;
        done0 = &empty[s];
        advance();
        tc_fence_after_sync();
        tc_mma_tf32(tmem, ad, bd1, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd1 + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd1 + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd1 + 6, idesc, 1u);
        tc_commit(done1);
      }
      tc_commit(done);
      mbar_wait(done, 0);
      out[blockIdx.x * 2] = clock64() - t0;
      out[blockIdx.x * 2 + 1] = 0;
    }
  } else if (warp <= np) {
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

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  int dev_clock_khz = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  const size_t big = (size_t)6 << 30;   // 6 GiB buffer
  float* buf;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 0, big);
  long long* d_out;
  cudaMalloc(&d_out, 296 * 8);
  {
    const size_t set = (size_t)64 << 20;
    cuuint64_t dims[2] = {32, set / 512};
    cuuint64_t strd[1] = {512};
    cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    CUtensorMap map;
    enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strd, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int S : {4, 6})
      for (int np : {1, 2}) {
        const int smem_p = (S + 1) * 16384 + 256 + 1024, kblocks = 4096;
        const int smem_use = smem_p < 120 * 1024 ? 120 * 1024 : smem_p;
        cudaFuncSetAttribute(probe_pipe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_use);
        for (int rep = 0; rep < 2; ++rep) probe_pipe2<<<148, 160, smem_use>>>(map, kblocks, S, np, (int)(set / 512 / 128), d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0;
        for (int b = 0; b < 148; ++b) cyc += (double)h[2 * b];
        printf("mini pipeline, software-pipelined issuer: %d stages, %d producing thread(s): %6.1f cycles per K-block (floor 256)\n",
               S, np, cyc / 148 / kblocks);
      }
    for (int mode : {0, 8, 6})
    for (int S : {3, 6})
      for (int np : {1, 2}) {
        const int smem_p = (S + 1) * 16384 + 256 + 1024, kblocks = 4096;
        const int smem_use = smem_p < 120 * 1024 ? 120 * 1024 : smem_p;   // one CTA per SM
        cudaFuncSetAttribute(probe_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_use);
        for (int rep = 0; rep < 2; ++rep) probe_pipe<<<148, 160, smem_use>>>(map, kblocks, S, np, (int)(set / 512 / 128), mode, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0, wt = 0;
        for (int b = 0; b < 148; ++b) { cyc += (double)h[2 * b]; wt += (double)h[2 * b + 1]; }
        printf("mini pipeline (%s): %d stages, %d producing thread(s): %6.1f cycles per K-block (floor 256), of which issuer blocked on the ring %6.1f\n",
               mode == 0 ? "commit -> empty -> producer" : mode == 8 ? "EIGHT MMAs per stage (floor 512)" : mode == 2 ? "plain arrive instead of commit" : mode == 6 ? "NO MMAs, plain arrive" : "producers never wait for empty", S, np, cyc / 148 / kblocks, wt / 148 / kblocks);
      }
    const int smem = 14 * 16384 + 256 + 1024, iters = 8192;
    cudaFuncSetAttribute(probe_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int nmma : {128, 256})
      for (int nw = 0; nw <= 4; ++nw) {
        for (int rep = 0; rep < 2; ++rep) { cudaMemset(d_out, 0, 296 * 8); probe_mix<<<148, 160, smem>>>(map, iters, nw, (int)(set / 512 / 128), nmma, d_out); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        e = cudaGetLastError();
        if (e != cudaSuccess) { printf("launch error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), d_out, 296 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0, nb = 0;
        for (int b = 0; b < 148; ++b) { cyc += (double)h[2 * b]; nb += (double)h[2 * b + 1]; }
        cyc /= 148; nb /= 148;
        printf("MMA N=%3d stream + %d TMA-issuing warps: %6.1f cycles per MMA (alone: %3d) | TMA fills %5.1f B/clk, operand reads %5.1f B/clk\n",
               nmma, nw, cyc / iters, nmma == 128 ? 64 : 128, nb * 16384.0 / cyc, (4096.0 + nmma * 32.0) * iters / cyc);
      }
  }
  const int strides[] = {512};
  const size_t sets[] = {(size_t)64 << 20};   // L2-resident
  for (size_t set : sets)
    for (int stride : strides)
     for (int box_rows : {64, 128, 256})
     for (int ctas : {1})
     for (int issuers : {1, 2})
      for (int depth : {2, 4}) {
        // 2-D view: dim0 = 32 floats, dim1 = rows `stride` bytes apart
        const cuuint64_t rows = set / stride;
        cuuint64_t dims[2] = {32, rows};
        cuuint64_t strd[1] = {(cuuint64_t)stride};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
        CUtensorMap map;
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strd, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int n_boxes_total = (int)(rows / box_rows);
        const int boxes_per_cta = 2048;
        const int smem = 2 * depth * 32768 + 256 + 1024;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        float best_ms = 1e9f;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 3; ++rep) {
          cudaEventRecord(e0);
          probe<<<148 * ctas, 128, smem>>>(map, boxes_per_cta, depth, n_boxes_total, box_rows, issuers, d_out);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        std::vector<long long> h(148 * ctas);
        cudaMemcpy(h.data(), d_out, 148 * ctas * 8, cudaMemcpyDeviceToHost);
        double cyc = 0;
        for (int b = 0; b < 148 * ctas; ++b) cyc += (double)h[b];
        cyc /= 148 * ctas;
        const double bytes = 128.0 * box_rows * boxes_per_cta;
        printf("row stride %4d B  box %3d rows  %d CTA/SM  %d issuing warps  depth %d : %6.1f B/clk/SM  (%5.2f TB/s aggregate, %6.1f cycles per box per CTA, %5.2f cycles per row per SM)\n",
               stride, box_rows, ctas, issuers, depth, ctas * bytes / cyc, 148.0 * ctas * bytes / (best_ms * 1e-3) / 1e12, cyc / boxes_per_cta,
               cyc / boxes_per_cta / box_rows / ctas);
      }
  return 0;
}
