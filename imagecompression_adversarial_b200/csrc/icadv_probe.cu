// Measured denominators for the roofline lines of bench.py: the tensor core's own kind::tf32 rate.
// A bare tcgen05.mma loop on static shared-memory operands (no TMA, no epilogue): what one SM's tensor core retires
// when nothing else limits it.  N = 256 instructions are math-bound (128 cycles each, scripts/probes/mma_rate_probe.cu),
// so FLOP / time of this launch is the TF32 peak the contraction kernels are held against -- measured on the GPU and
// under the clocks the benchmark itself runs at, instead of "half the cuBLAS bf16 number".
#include "icadv_common.cuh"
#include "icadv_ptx.cuh"

namespace icadv {

__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, int n) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  // operands: A [128 x 32 fp32] and B [n x 32 fp32], SWIZZLE_128B K-major tiles; zeros (values do not change timing)
  for (int i = threadIdx.x; i < (128 + n) * 32; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, n);
      const uint64_t ad = umma_desc_sw128(smem_u32(smem));
      const uint64_t bd = umma_desc_sw128(smem_u32(smem + 128 * 128));
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + ((it & 1) ? 256u : 0u);      // two accumulators: no dependence between K-blocks
        tc_mma_tf32(d, ad, bd, idesc, 1u);
        tc_mma_tf32(d, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(d, ad + 4, bd + 4, idesc, 1u);
        tc_mma_tf32(d, ad + 6, bd + 6, idesc, 1u);
      }
      tc_commit(&done);
    }
    __syncwarp();
  }
  mbar_wait(&done, 0);
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_probe_tf32_peak(int iters, int n, double* flops_out, icadv_stream_t stream) {
  ICADV_REQUIRE(iters > 0 && n >= 32 && n <= 256 && n % 16 == 0, "probe_tf32_peak: iters > 0, n in [32,256], n %% 16 == 0");
  int rc = icadv_check_device();
  if (rc) return rc;
  int dev = 0, sms = 148;
  ICADV_CUDA_TRY(cudaGetDevice(&dev));
  ICADV_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // 120 KB of dynamic shared memory keeps it at one CTA per SM (each CTA allocates all 512 TMEM columns)
  const int smem = 120 * 1024;
  static cudaError_t attr = cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ICADV_CUDA_TRY(attr);
  tf32_peak_kernel<<<sms, 128, smem, as_stream(stream)>>>(iters, n);
  ICADV_CUDA_TRY(cudaGetLastError());
  if (flops_out) *flops_out = (double)sms * iters * 4.0 * 2.0 * 128.0 * n * 8.0;
  return ICADV_OK;
}

}  // extern "C"
