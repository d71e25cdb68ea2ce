// tcgen05.mma issue/execution rate probe (perf triage, sm_100a): one CTA per SM issues `reps` x 8 MMAs with the operand
// layout of conv_tc_kernel (K-major, SWIZZLE_128B, M = 128, K = 16 slices of a 64-wide stage, bf16 -> fp32) into one
// TMEM accumulator, commits, waits, and reports cycles per MMA.  Variants: N, whether all 148 SMs run, and an optional
// background warp set that keeps writing shared memory with cp.async-like traffic (plain stores) to load the port.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/mma_rate_probe tools/mma_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  constexpr uint64_t sbo = (8 * 128) >> 4;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(256, 1) probe(int reps, int bg, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + 16384, b_hi = a_hi + 32768, b_lo = b_hi + N * 128;
  if (warp == 0 && lane == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t adv = (uint64_t)(ks * 2);
        umma(tm, desc_sw128(a_hi) + adv, desc_sw128(b_hi) + adv, idesc_bf16(128, N), (r | ks) ? 1u : 0u);
        umma(tm, desc_sw128(a_hi) + adv, desc_sw128(b_lo) + adv, idesc_bf16(128, N), 1u);
        umma(tm, desc_sw128(a_lo) + adv, desc_sw128(b_hi) + adv, idesc_bf16(128, N), 1u);
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    ((volatile uint32_t*)smem)[24 * 1024 + 7] = 1;            // stop flag for the background warps
  } else if (bg && warp >= 4) {
    // background shared-memory write traffic into a scratch region (like the producers' cp.async landing)
    uint4* dst = (uint4*)(smem + 100 * 1024) + (threadIdx.x - 128);
    volatile uint32_t* stop = ((volatile uint32_t*)smem) + 24 * 1024 + 7;
    uint4 v = make_uint4(1, 2, 3, 4);
    while (*stop != 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dst[i * 128] = v;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

template <int N> void run(int grid, int bg) {
  long long* d; cudaMalloc(&d, 16);
  const int reps = 2000, smem = 160 * 1024;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N><<<grid, 256, smem>>>(200, bg, d);
  probe<N><<<grid, 256, smem>>>(reps, bg, d);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d grid=%3d bg=%d: issue %.1f cycles/MMA, issue+drain %.1f cycles/MMA (nominal %d)  %s\n", N, grid, bg,
         (double)h[0] / (reps * 12), (double)h[1] / (reps * 12), N / 2, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<128>(1, 0); run<128>(148, 0); run<128>(148, 1);
  run<64>(1, 0); run<64>(148, 0); run<64>(148, 1);
  run<32>(148, 0); run<16>(148, 0); run<256>(148, 0);
  return 0;
}
