// Developer probe: sustained issue rate of tcgen05.mma kind::tf32 (M = 128 per CTA, K = 8) from ONE issuing thread,
// as a function of N, of the number of resident CTAs per SM, of the number of accumulators / issuing warps, and of
// cta_group (1 = one SM, 2 = a CTA pair sharing the B operand).  Operands are static shared-memory tiles (no TMA),
// rotated over `nbuf` buffers so that consecutive instructions read different addresses, as the conv kernels do.
// Prints cycles per MMA per SM and the fraction of the 4096 FLOP/cycle/SM TF32 floor.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_rate_probe mma_rate_probe.cu
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../imagecompression_adversarial_b200/csrc/icadv_ptx.cuh"
using namespace icadv;

struct Cfg {
  int n;          // MMA N
  int nbuf;       // operand buffers rotated through
  int iters;      // MMAs per issuing warp
  int issuers;    // 1 or 2 issuing warps (each with its own accumulator)
  int accs;       // accumulators an issuer alternates between (1 or 2)
  int cg;         // cta_group
  int same_a;     // 1: every MMA reads the same A tile
  int kb;         // > 0: per `kb` MMAs one mbarrier wait (already complete) + fence before and one tcgen05.commit after
  int same_acc;   // 1: both issuers accumulate into the SAME TMEM region
  int kbmode;     // which per-K-block extras run when kb > 0: 1 = mbarrier wait, 2 = tcgen05.fence::after_thread_sync, 4 = tcgen05.commit (0 = all)
};

__device__ __forceinline__ void mma_cg2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_cg2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CG>
__global__ void __launch_bounds__(128) probe(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = c.n / CG;                      // rows of B this CTA holds
  const int a_bytes = 128 * 128, b_bytes = b_rows * 128;
  uint8_t* A = smem;
  uint8_t* B = smem + c.nbuf * a_bytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + c.nbuf * b_bytes);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = c.nbuf * (a_bytes + b_bytes) / 4;
  for (int i = threadIdx.x; i < total; i += 128) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  int cols = 32;
  while (cols < c.issuers * c.accs * c.n) cols <<= 1;
  if (warp == 0) {
    if (CG == 1) { tmem_alloc(tptr, cols); tmem_relinquish(); }
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  const bool leader_cta = CG == 1 || cluster_rank() == 0;
  if (leader_cta && warp < c.issuers) {
    const uint32_t idesc = umma_idesc_tf32(128 * CG, c.n);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
    const int acc_cols = c.n;   // accumulators side by side
    __syncwarp();
    const long long t0 = clock64();
    int buf = 0;
    uint64_t* rdy = &bar[2 + warp];       // a barrier whose phase 0 completes at once: the "full" wait of a K-block
    uint64_t* done = &bar[4 + warp];      // receives the per-K-block commits (nobody waits on it)
    if (c.kb > 0 && lane == 0) mbar_arrive(rdy);
    __syncwarp();
    for (int i = 0; i < c.iters; i += 4) {
      if (c.kb > 0 && (i % c.kb) == 0) {
        if (c.kbmode == 0 || (c.kbmode & 1)) mbar_wait(rdy, 0);
        if (c.kbmode == 0 || (c.kbmode & 2)) tc_fence_after_sync();
      }
      const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (a_lo0 + (c.same_a ? 0 : buf) * (a_bytes >> 4));
      const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | (b_lo0 + buf * (b_bytes >> 4));
      const uint32_t d = tmem + ((c.same_acc ? 0 : warp) * c.accs + ((i >> 2) % c.accs)) * acc_cols;
      if (elect_one_sync()) {
        if (CG == 1) {
          tc_mma_tf32(d, ad, bd, idesc, 1u); tc_mma_tf32(d, ad + 2, bd + 2, idesc, 1u);
          tc_mma_tf32(d, ad + 4, bd + 4, idesc, 1u); tc_mma_tf32(d, ad + 6, bd + 6, idesc, 1u);
        } else {
          mma_cg2(d, ad, bd, idesc, 1u); mma_cg2(d, ad + 2, bd + 2, idesc, 1u);
          mma_cg2(d, ad + 4, bd + 4, idesc, 1u); mma_cg2(d, ad + 6, bd + 6, idesc, 1u);
        }
      }
      if (c.kb > 0 && (c.kbmode == 0 || (c.kbmode & 4)) && ((i + 4) % c.kb) == 0 && elect_one_sync()) { if (CG == 1) tc_commit(done); else commit_cg2(done); }
      if (++buf == c.nbuf) buf = 0;
    }
    if (elect_one_sync()) { if (CG == 1) tc_commit(&bar[warp]); else commit_cg2(&bar[warp]); }
    __syncwarp();
    const long long t_issue = clock64();
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    if (lane == 0) {
      out[(blockIdx.x * 2 + warp) * 2 + 0] = t1 - t0;
      out[(blockIdx.x * 2 + warp) * 2 + 1] = t_issue - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) tmem_dealloc(tmem, cols);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
  }
}

// Minimal issue loops (cg1, N = 128, one accumulator): U MMAs per elect.sync block, or MODE 1 = only lane 0 runs the loop
template <int U, int MODE, int NN = 128>
__global__ void __launch_bounds__(128) probe_min(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* B = smem + 4 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 4 * 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, NN);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
    __syncwarp();
    const long long t0 = clock64();
    if (MODE == 0) {
      for (int i = 0; i < iters; i += U) {
        if (elect_one_sync()) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t off = ((u >> 2) & 3) * 1024u + (u & 3) * 2u;
            tc_mma_tf32(tmem, (static_cast<uint64_t>(hi) << 32) | (a_lo0 + off), (static_cast<uint64_t>(hi) << 32) | (b_lo0 + off), idesc, 1u);
          }
        }
      }
      if (elect_one_sync()) tc_commit(&bar[0]);
    } else if (lane == 0) {
      for (int i = 0; i < iters; i += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t off = ((u >> 2) & 3) * 1024u + (u & 3) * 2u;
          tc_mma_tf32(tmem, (static_cast<uint64_t>(hi) << 32) | (a_lo0 + off), (static_cast<uint64_t>(hi) << 32) | (b_lo0 + off), idesc, 1u);
        }
      }
      tc_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int U, int MODE, int NN = 128>
static void run_min(const char* name) {
  const int iters = 2048, smem = 8 * 16384 + 128 + 1024;
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_min<U, MODE, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe_min<U, MODE, NN><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int b = 0; b < 148; ++b) sum += (double)h[b];
  printf("%-58s %7.1f cyc/MMA/SM  (floor 64.0)\n", name, sum / 148 / iters);
  cudaFree(d_out);
}


// Shared-memory contention: warp 0 issues the bare N = 128 MMA stream (64 cycles per MMA alone, all of it operand reads at
// 128 B/clk); `lsw` other warps stream 16-byte ld.shared (mode 1) or st.shared (mode 2) over a private 32 KB buffer.
__global__ void __launch_bounds__(288) probe_contend(int iters, int lsw, int mode, long long* out, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* B = smem + 4 * 16384;
  uint8_t* X = B + 4 * 16384;                       // 32 KB private buffer of the load/store warps
  uint64_t* bar = reinterpret_cast<uint64_t*>(X + 32768);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(tptr + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_fence_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_tf32(128, 128);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
      if (elect_one_sync()) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const uint32_t off = ((u >> 2) & 3) * 1024u + (u & 3) * 2u;
          tc_mma_tf32(tmem, (static_cast<uint64_t>(hi) << 32) | (a_lo0 + off), (static_cast<uint64_t>(hi) << 32) | (b_lo0 + off), idesc, 1u);
        }
      }
    }
    if (elect_one_sync()) tc_commit(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; *stop = 1; }
  } else if (warp <= lsw) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long n = 0;
    float4* xb = reinterpret_cast<float4*>(X) + (warp - 1) * 256 + lane;   // 4 KB per warp, conflict-free 512 B per instruction
    while (!*stop) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (mode == 1) { const float4 v = xb[k * 32]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        else { xb[k * 32] = acc; acc.x += 1.f; }
      }
      n += 8;
    }
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&out[blockIdx.x * 2 + 1]), (unsigned long long)n);
    if (acc.x == 123.456f) sink[0] = acc.y + acc.z + acc.w;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

static void run_contend(const char* name, int lsw, int mode) {
  const int iters = 4096, smem = 8 * 16384 + 32768 + 256 + 1024;
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
  cudaMalloc(&d_sink, 16);
  cudaFuncSetAttribute(probe_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(d_out, 0, 148 * 2 * sizeof(long long)); probe_contend<<<148, 288, smem>>>(iters, lsw, mode, d_out, d_sink); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148 * 2);
  cudaMemcpy(h.data(), d_out, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0, ins = 0;
  for (int b = 0; b < 148; ++b) { cyc += (double)h[2 * b]; ins += (double)h[2 * b + 1]; }
  cyc /= 148; ins /= 148;
  printf("%-58s %7.1f cyc/MMA  | other warps: %6.1f B/clk of %s (MMA operand reads: %5.1f B/clk)\n", name, cyc / iters,
         ins * 512.0 / cyc, mode == 1 ? "ld.shared" : "st.shared", 8192.0 * iters / cyc);
  cudaFree(d_out); cudaFree(d_sink);
}

// Issue loops of ONE thread (the whole loop inside one elect.sync block, as the kernels run it) with runtime descriptors:
// FEAT bit 0 = descriptors depend on a per-iteration register value (R2UR), bit 1 = tcgen05.commit per 4 MMAs,
// bit 2 = mbarrier.test_wait probe per 4 MMAs, bit 3 = tcgen05.fence::after_thread_sync per 4 MMAs.
template <int FEAT>
__global__ void __launch_bounds__(128) probe_thread(int iters, int stride_rt, int a_sbo, int a_shift_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* B = smem + 4 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 4 * 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
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
      const uint32_t a_hi = (static_cast<uint32_t>(a_sbo) >> 4) | (1u << 14) | (2u << 29);   // halo patch: tile rows a_sbo bytes apart
      const uint32_t a_lo0 = (((smem_u32(A) + a_shift_rows * 128) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
      mbar_arrive(&bar[1]);                       // bar[1]: phase 0 complete -> the probes below succeed at once
      int s = 0;
      uint32_t seen = 0;
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
        const uint32_t off = (FEAT & 1) ? static_cast<uint32_t>(s) * static_cast<uint32_t>(stride_rt) : 0u;
        const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + ((FEAT & 16) ? static_cast<uint32_t>(s) * 8u : off));
        const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | (b_lo0 + off);
        if (FEAT & 8) tc_fence_after_sync();
        tc_mma_tf32(tmem, ad, bd, idesc, 1u);
        tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u);
        tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
        if (FEAT & 2) tc_commit(&bar[2 + (s & 1)]);
        if (FEAT & 4) seen += mbar_test_wait(&bar[1], 0) ? 1u : 0u;
        if (++s == 4) s = 0;
      }
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      out[blockIdx.x] = clock64() - t0 + (seen == 0xffffffffu ? 1 : 0);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int FEAT>
static void run_thread(const char* name, int a_sbo = 1024, int a_shift = 0) {
  const int iters = 2048, smem = 8 * 16384 + 128 + 1024;
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_thread<FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe_thread<FEAT><<<148, 128, smem>>>(iters, 1024, a_sbo, a_shift, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int b = 0; b < 148; ++b) sum += (double)h[b];
  printf("%-58s %7.1f cyc/MMA/SM  (floor 64.0)\n", name, sum / 148 / iters);
  cudaFree(d_out);
}

// TMEM contention: warp 0 issues the bare N = 128 MMA stream into columns [0,128); warps 4-7 (one per lane quadrant) read
// columns [128,256) with tcgen05.ld 32x32b.x32 in a loop, as the epilogue warpgroups do while the next item's main loop runs.
__global__ void __launch_bounds__(256) probe_tmem(int iters, int readers, long long* out, float* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* B = smem + 4 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 4 * 16384);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 8);
  volatile int* stop = reinterpret_cast<volatile int*>(tptr + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_fence_init(); *stop = 0; }
  if (warp == 0) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, 128);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
        const uint32_t off = static_cast<uint32_t>((i >> 2) & 3) * 1024u;
        const uint64_t ad = (static_cast<uint64_t>(hi) << 32) | (a_lo0 + off), bd = (static_cast<uint64_t>(hi) << 32) | (b_lo0 + off);
        tc_mma_tf32(tmem, ad, bd, idesc, 1u); tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u); tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
      }
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      out[blockIdx.x * 2] = clock64() - t0;
      *stop = 1;
    }
  } else if (warp >= 4 && warp < 4 + readers) {
    const int q = warp & 3;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(q * 32) << 16) + 128;
    float acc = 0.f;
    long long n = 0;
    while (!*stop) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(t_lane + c * 32, v);
        tmem_ld_wait();
        acc += v[0] + v[31];
      }
      n += 4;
    }
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&out[blockIdx.x * 2 + 1]), (unsigned long long)n);
    if (acc == 123.456f) sink[0] = acc;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static void run_tmem(const char* name, int readers) {
  const int iters = 4096, smem = 8 * 16384 + 256 + 1024;
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
  cudaMalloc(&d_sink, 16);
  cudaFuncSetAttribute(probe_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(d_out, 0, 148 * 2 * sizeof(long long)); probe_tmem<<<148, 256, smem>>>(iters, readers, d_out, d_sink); }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148 * 2);
  cudaMemcpy(h.data(), d_out, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0, ins = 0;
  for (int b = 0; b < 148; ++b) { cyc += (double)h[2 * b]; ins += (double)h[2 * b + 1]; }
  cyc /= 148; ins /= 148;
  printf("%-58s %7.1f cyc/MMA  | readers: %6.1f B/clk of tcgen05.ld (32 lanes x 32 columns x 4 B per instruction)\n", name,
         cyc / iters, ins * 4096.0 / cyc);
  cudaFree(d_out); cudaFree(d_sink);
}

// The kernels' issue loop, re-created feature by feature on pre-armed barriers (nobody produces or consumes):
// FEAT bit 0 = tap offsets from a kernel-parameter table with a register index (LDC), bit 1 = ring bookkeeping with wrap
// (stage index, parity, descriptor base), bit 2 = a commit to a ROTATING barrier per K-block + early test_wait probe of the
// next stage (barriers re-armed by the commits themselves), bit 3 = per-group patch wait + commit every 6 K-blocks.
struct ProbeTable { int32_t aoff[40]; int32_t S; int32_t pad[7]; };
template <int FEAT, int NN = 128>
__global__ void __launch_bounds__(128) probe_loop(const __grid_constant__ ProbeTable tb, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;
  uint8_t* B = smem + 4 * 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + 4 * 16384);   // [0] done, [1..8] "full" ring, [9..16] "empty" ring, [17] patch
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 20);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1e-3f * (float)((i * 37) & 255);
  if (threadIdx.x == 0) { for (int i = 0; i < 20; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(tptr, 128); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_tf32(128, NN);
      const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_hi = (1280u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_u32(A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t w_lo0 = ((smem_u32(B) >> 4) & 0x3FFFu) | (1u << 16);
      const int S = tb.S;
      uint64_t* full = bar + 1;
      uint64_t* empty = bar + 9;
      // "full" barriers: the commit of iteration k on full[s] re-arms it for the probe of iteration k + S (self-feeding)
      for (int i = 0; i < S; ++i) mbar_arrive(&full[i]);
      int s = 0, t = 0;
      uint32_t s_par = 0, w_lo = w_lo0, aoff = (FEAT & 1) ? static_cast<uint32_t>(tb.aoff[0]) >> 4 : 0u;
      bool w_ok = false;
      const long long t0 = clock64();
      for (int i = 0; i < iters; i += 4) {
        if ((FEAT & 8) && t == 0) mbar_wait(&full[(s + 1 == S) ? 0 : s + 1], s_par ^ ((s + 1 == S) ? 1u : 0u) ^ 0u) ;
        const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + aoff);
        const uint64_t bd = (static_cast<uint64_t>(hi) << 32) | w_lo;
        if ((FEAT & 4) && !w_ok) mbar_wait(&full[s], s_par);
        uint64_t* wdone = (FEAT & 4) ? &full[s] : &empty[s];
        if (FEAT & 2) { if (++s == S) { s = 0; s_par ^= 1; w_lo = w_lo0; } else { w_lo += 1024u; } }
        tc_fence_after_sync();
        tc_mma_tf32(tmem, ad, bd, idesc, 1u);
        tc_mma_tf32(tmem, ad + 2, bd + 2, idesc, 1u);
        tc_mma_tf32(tmem, ad + 4, bd + 4, idesc, 1u);
        tc_mma_tf32(tmem, ad + 6, bd + 6, idesc, 1u);
        tc_commit(wdone);
        if (++t == 6) { t = 0; if (FEAT & 8) tc_commit(&bar[17]); }
        if (FEAT & 1) aoff = static_cast<uint32_t>(tb.aoff[t + 1]) >> 4;
        if (FEAT & 4) w_ok = mbar_test_wait(&full[s], s_par);
      }
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

template <int FEAT, int NN = 128>
static void run_loop(const char* name) {
  const int iters = 2048, smem = 8 * 16384 + 256 + 1024;
  ProbeTable tb;
  for (int i = 0; i < 40; ++i) tb.aoff[i] = ((i % 3) * 10 + (i % 2)) * 128;
  tb.S = 3;
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe_loop<FEAT, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) probe_loop<FEAT, NN><<<148, 128, smem>>>(tb, iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(148);
  cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0;
  for (int b = 0; b < 148; ++b) sum += (double)h[b];
  printf("%-58s %7.1f cyc/MMA/SM  (floor 64.0)\n", name, sum / 148 / iters);
  cudaFree(d_out);
}

static void run(const char* name, Cfg c, int ctas_per_sm) {
  const int b_rows = c.n / c.cg;
  int smem = c.nbuf * (128 * 128 + b_rows * 128) + 128 + 1024;
  if (ctas_per_sm == 1 && smem < 120 * 1024) smem = 120 * 1024;   // force one CTA per SM
  const int grid = 148 * ctas_per_sm;
  long long* d_out;
  cudaMalloc(&d_out, grid * 4 * sizeof(long long));
  cudaMemset(d_out, 0, grid * 4 * sizeof(long long));
  cudaError_t e;
  if (c.cg == 1) {
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) probe<1><<<grid, 128, smem>>>(c, d_out);
  } else {
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, probe<2>, c, d_out);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-58s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid * 4);
  cudaMemcpy(h.data(), d_out, grid * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
  double sum = 0, sum_i = 0; int cnt = 0;
  for (int b = 0; b < grid; ++b)
    for (int w = 0; w < 2; ++w)
      if (h[(b * 2 + w) * 2] > 0) { sum += (double)h[(b * 2 + w) * 2]; sum_i += (double)h[(b * 2 + w) * 2 + 1]; ++cnt; }
  const double cyc = sum / cnt, cyc_issue = sum_i / cnt;
  // MMAs retired per SM while one issuer runs: issuers * ctas_per_sm streams in parallel (cg2: one stream feeds 2 SMs)
  const double streams_per_sm = (c.cg == 1) ? (double)c.issuers * ctas_per_sm : (double)c.issuers * ctas_per_sm;
  const double per_mma_per_sm = cyc / c.iters / streams_per_sm;   // cg2: each instruction = 128 x N x 8 on EACH SM
  const double floor_cyc = 128.0 * c.n * 8 / 2048.0;                  // 128 x N x 8 MACs at 2048 MAC/cycle/SM
  printf("%-58s %7.1f cyc/MMA/SM (issue loop %6.1f)  floor %5.1f  -> %5.1f %% of the tensor floor\n", name, per_mma_per_sm,
         cyc_issue / c.iters / streams_per_sm, floor_cyc, 100.0 * floor_cyc / per_mma_per_sm);
  cudaFree(d_out);
}

int main() {
  const int L = 2048;
  run_loop<7, 16>("kernel-like loop, N=16 MMAs (tensor time 39 cycles): the issue path itself");
  run_loop<15, 16>("kernel-like loop + per-group wait/commit, N=16 MMAs");
  run_loop<0>("kernel-like loop: bare (commit per K-block)");
  run_loop<1>("kernel-like loop: + LDC tap offsets");
  run_loop<3>("kernel-like loop: + LDC + ring bookkeeping");
  run_loop<7>("kernel-like loop: + LDC + ring + barrier wait/probe");
  run_loop<6>("kernel-like loop: ring + barrier wait/probe (no LDC)");
  run_tmem("MMA stream, no TMEM readers", 0);
  run_tmem("MMA stream + 1 warp tcgen05.ld", 1);
  run_tmem("MMA stream + 2 warps tcgen05.ld", 2);
  run_tmem("MMA stream + 4 warps tcgen05.ld (one epilogue group)", 4);
  run_thread<31>("one thread: halo-patch A (SBO 1280 B, start +13 rows, shifting)", 1280, 13);
  run_thread<31>("one thread: halo-patch A (SBO 2304 B, start +19 rows, shifting)", 2304, 19);
  run_thread<15>("one thread: A with SBO 1280 B, start +0", 1280, 0);
  run_thread<15>("one thread: A dense (SBO 1024), start +3 rows", 1024, 3);
  run_thread<0>("one thread: constant descriptors");
  run_thread<1>("one thread: runtime descriptors (R2UR)");
  run_thread<3>("one thread: runtime descriptors + commit per 4");
  run_thread<5>("one thread: runtime descriptors + test_wait per 4");
  run_thread<9>("one thread: runtime descriptors + fence per 4");
  run_thread<15>("one thread: runtime descriptors + commit + test_wait + fence");
  run_thread<2>("one thread: constant descriptors + commit per 4");
  run("cg1 N=128 1 issuer, per 4 MMAs: wait only", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 1}, 1);
  run("cg1 N=128 1 issuer, per 4 MMAs: fence only", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 2}, 1);
  run("cg1 N=128 1 issuer, per 4 MMAs: commit only", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 4}, 1);
  run("cg1 N=128 1 issuer, per 4 MMAs: wait + fence", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 3}, 1);
  run("cg1 N=128 1 issuer, per 4 MMAs: none of them (loop only)", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 8}, 1);
  run_contend("MMA stream alone", 0, 1);
  run_contend("MMA stream + 1 warp ld.shared", 1, 1);
  run_contend("MMA stream + 2 warps ld.shared", 2, 1);
  run_contend("MMA stream + 4 warps ld.shared", 4, 1);
  run_contend("MMA stream + 8 warps ld.shared", 8, 1);
  run_contend("MMA stream + 1 warp st.shared", 1, 2);
  run_contend("MMA stream + 2 warps st.shared", 2, 2);
  run_contend("MMA stream + 4 warps st.shared", 4, 2);
  run_contend("MMA stream + 8 warps st.shared", 8, 2);
  run_min<4, 0, 64>("min loop N=64: elect, 4 MMAs per block (floor 32)");
  run_min<4, 0, 32>("min loop N=32: elect, 4 MMAs per block (floor 16)");
  run_min<16, 0, 32>("min loop N=32: elect, 16 MMAs per block (floor 16)");
  run_min<4, 0, 16>("min loop N=16: elect, 4 MMAs per block (floor 8)");
  run_min<4, 0>("min loop: elect, 4 MMAs per block");
  run_min<8, 0>("min loop: elect, 8 MMAs per block");
  run_min<16, 0>("min loop: elect, 16 MMAs per block");
  run_min<64, 0>("min loop: elect, 64 MMAs per block");
  run_min<4, 1>("min loop: lane 0 only, 4 MMAs per iteration");
  run_min<16, 1>("min loop: lane 0 only, 16 MMAs per iteration");
  run("cg1 N=128 1 CTA/SM, 1 issuer, 4 bufs", Cfg{128, 4, L, 1, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=128 1 CTA/SM, 1 issuer, same A", Cfg{128, 4, L, 1, 1, 1, 1, 0, 0, 0}, 1);
  run("cg1 N=128 1 CTA/SM, 1 issuer, 1 buf", Cfg{128, 1, L, 1, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=128 1 CTA/SM, 1 issuer, 2 accumulators", Cfg{128, 4, L, 1, 2, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=128 1 CTA/SM, 2 issuers", Cfg{128, 4, L, 2, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=128 2 CTA/SM, 1 issuer each", Cfg{128, 3, L, 1, 1, 1, 0, 0, 0, 0}, 2);
  run("cg1 N=64  1 CTA/SM, 1 issuer", Cfg{64, 4, L, 1, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=192 1 CTA/SM, 1 issuer", Cfg{192, 4, L, 1, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=256 1 CTA/SM, 1 issuer", Cfg{256, 4, L, 1, 1, 1, 0, 0, 0, 0}, 1);
  run("cg1 N=256 1 CTA/SM, 1 issuer, 2 accumulators", Cfg{256, 4, L, 1, 2, 1, 0, 0, 0, 0}, 1);
  run("cg2 M=256 N=128 (B split 64+64), 1 issuer per pair", Cfg{128, 4, L, 1, 1, 2, 0, 0, 0, 0}, 1);
  run("cg2 M=256 N=256 (B split 128+128), 1 issuer per pair", Cfg{256, 4, L, 1, 1, 2, 0, 0, 0, 0}, 1);
  run("cg2 M=256 N=128, 2 accumulators", Cfg{128, 4, L, 1, 2, 2, 0, 0, 0, 0}, 1);
  run("cg1 N=128 1 issuer, wait+commit per 4 MMAs", Cfg{128, 4, L, 1, 1, 1, 0, 4, 0, 0}, 1);
  run("cg1 N=128 1 issuer, wait+commit per 8 MMAs", Cfg{128, 4, L, 1, 1, 1, 0, 8, 0, 0}, 1);
  run("cg1 N=128 2 issuers, wait+commit per 4 MMAs", Cfg{128, 4, L, 2, 1, 1, 0, 4, 0, 0}, 1);
  run("cg1 N=128 2 issuers SAME accumulator", Cfg{128, 4, L, 2, 1, 1, 0, 0, 1, 0}, 1);
  run("cg1 N=128 2 issuers SAME accumulator, wait+commit per 4", Cfg{128, 4, L, 2, 1, 1, 0, 4, 1, 0}, 1);
  run("cg1 N=256 1 issuer, wait+commit per 4 MMAs", Cfg{256, 4, L, 1, 1, 1, 0, 4, 0, 0}, 1);
  run("cg1 N=128 2 CTA/SM, wait+commit per 4 MMAs", Cfg{128, 3, L, 1, 1, 1, 0, 4, 0, 0}, 2);
  run("cg2 N=128 2 issuers", Cfg{128, 4, L, 2, 1, 2, 0, 0, 0, 0}, 1);
  return 0;
}
