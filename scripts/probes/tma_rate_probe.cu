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
          const int box = ((blockIdx.x * 2 + w) * boxes_per_cta + i) % n_boxes_total;   // every issuer streams its own range
          mbar_arrive_expect_tx(&bar[s], box_rows * 128);
          tma_load_2d(smem + s * 32768, &map, &bar[s], 0, box * box_rows);
        }
      }
      if (w == 0) out[blockIdx.x] = clock64() - t0;
    }
  }
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
  const int strides[] = {128, 512};
  const size_t sets[] = {(size_t)64 << 20};   // L2-resident
  for (size_t set : sets)
    for (int stride : strides)
     for (int box_rows : {128, 256})
     for (int ctas : {1})
     for (int issuers : {1, 2})
      for (int depth : {2}) {
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
