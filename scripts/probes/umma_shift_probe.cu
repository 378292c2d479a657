// Developer probe: does a SWIZZLE_128B K-major UMMA shared-memory descriptor accept (a) a start address that is not
// 1024-byte aligned (row offset inside the 8-row swizzle atom), (b) a stride between 8-row groups (SBO) that is not a
// multiple of 1024 bytes, and which `base_offset` value makes it read the rows a TMA box would have written?
// D[m][n] = A[row(m)][n] through an identity B; the fill makes every (row, column) recognisable.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe umma_shift_probe.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../imagecompression_adversarial_b200/csrc/icadv_ptx.cuh"
using namespace icadv;

constexpr int kRows = 448;   // patch rows of 128 B

struct Variant { int off_rows, sbo_bytes, base_off, fill; };

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int sbo_bytes, int base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(Variant v, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                       // kRows x 128 B
  uint8_t* B = smem + kRows * 128;         // 32 x 128 B  (1024-aligned: kRows*128 = 57344)
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 32 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill as TMA SWIZZLE_128B would: 16-byte chunk j of linear row L at L*128 + ((j ^ (L & 7)) << 4)
  for (int L = threadIdx.x; L < kRows; L += 128)
    for (int c = 0; c < 32; ++c) {
      const float val = v.fill == 0 ? (float)L : (float)(c + 32 * (L & 7));
      *reinterpret_cast<float*>(A + sw128_off(L, c >> 2) + (c & 3) * 4) = val;
    }
  for (int n = threadIdx.x; n < 32; n += 128)
    for (int k = 0; k < 32; ++k)
      *reinterpret_cast<float*>(B + sw128_off(n, k >> 2) + (k & 3) * 4) = (n == k) ? 1.f : 0.f;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(tptr, 32); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, 32);
    const uint32_t a_addr = smem_u32(A) + v.off_rows * 128;
    const uint64_t ad = make_desc(a_addr, v.sbo_bytes, v.base_off), bd = make_desc(smem_u32(B), 1024, 0);
    for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, k > 0 ? 1u : 0u);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  float r[32];
  tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = r[j];
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 32 * 4);
  const int smem = kRows * 128 + 32 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int offs[] = {0, 3, 13, 19, 37};
  const int sbos[] = {1024, 1280, 2048, 2304};
  std::vector<float> h(128 * 32);
  for (int off : offs)
    for (int sbo : sbos)
      for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
        const int bo = bo_mode ? (off & 7) : 0;
        if (bo_mode && bo == 0) continue;
        int ok_rows[2] = {0, 0};
        int first_bad[2] = {-1, -1};
        float got_bad[2][4] = {};
        for (int fill = 0; fill < 2; ++fill) {
          Variant v{off, sbo, bo, fill};
          probe<<<1, 128, smem>>>(v, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("off=%d sbo=%d bo=%d: CUDA error %s\n", off, sbo, bo, cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m) {
            const int L = off + (m / 8) * (sbo / 128) + (m % 8);
            bool ok = true;
            for (int c = 0; c < 32; ++c) {
              const float want = fill == 0 ? (float)L : (float)(c + 32 * (L & 7));
              if (h[m * 32 + c] != want) ok = false;
            }
            if (ok) ++ok_rows[fill];
            else if (first_bad[fill] < 0) { first_bad[fill] = m; for (int c = 0; c < 4; ++c) got_bad[fill][c] = h[m * 32 + c * 4]; }
          }
        }
        printf("off=%2d sbo=%4d base_off=%d : rows ok (row-id fill) %3d/128, (column fill) %3d/128", off, sbo, bo, ok_rows[0], ok_rows[1]);
        if (first_bad[0] >= 0) printf("  first bad row-id m=%d got %.0f", first_bad[0], got_bad[0][0]);
        if (first_bad[1] >= 0) printf("  first bad col m=%d got [%.0f %.0f %.0f %.0f]", first_bad[1], got_bad[1][0], got_bad[1][1], got_bad[1][2], got_bad[1][3]);
        printf("\n");
      }
  return 0;
}
