// Measurement helper (not part of the documented ABI): the L2 -> SM gather roofline of the A producers of conv_tc_kernel.
// A kernel that does nothing else issues their access pattern -- 16-byte cp.async copies of 128-byte row pieces (hi + lo
// plane of a 64-channel split-bf16 row) at rulebook-like indices of a buffer that fits L2, into a shared-memory ring,
// indices prefetched one stage ahead -- and the achieved bytes per second are returned.  bench.py reports it next to
// the conv family's L2 traffic; tools/l2_gather_probe.cu is the standalone sweep over ring depths and row sizes.
#include <vector>

#include "common.cuh"

namespace fd {

__device__ __forceinline__ void probe_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

constexpr int PROBE_DEPTH = 4;          // 32 KB stages in flight per CTA (the conv kernel's ring holds 4-5)

__global__ void __launch_bounds__(256, 1)
l2_gather_probe_kernel(const char* __restrict__ buf, const int* __restrict__ idx, unsigned mask, int row_bytes, int stages) {
  extern __shared__ __align__(128) unsigned char probe_smem[];
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(probe_smem);
  const int tid = threadIdx.x, chunk = tid & 7, r0 = tid >> 3;          // 32 rows per pass, 4 passes, 2 planes
  unsigned pos = (blockIdx.x * 7919u * 128u) & mask;
  int nxt[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) nxt[q] = __ldg(idx + ((pos + r0 + 32 * q) & mask));
  for (int s = 0; s < stages; ++s) {
    const uint32_t dst = s0 + (uint32_t)(s % PROBE_DEPTH) * 32768u;
    int cur[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    pos = (pos + 128u * 148u) & mask;
#pragma unroll
    for (int q = 0; q < 4; ++q) nxt[q] = __ldg(idx + ((pos + r0 + 32 * q) & mask));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + 32 * q;
#pragma unroll
      for (int p = 0; p < 2; ++p)
        probe_cp_async16(dst + (uint32_t)((p * 128 + r) * 128 + chunk * 16), buf + (size_t)cur[q] * row_bytes + p * 128 + chunk * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(PROBE_DEPTH - 1) : "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace fd

extern "C" {

/* bytes per second through L2 into shared memory for the conv producers' gather pattern (rows of `rows` x 256 bytes,
 * neighbour-like indices); returns 0 and *out_bytes_per_s on success */
int fd_debug_l2_gather_probe(int rows, double* out_bytes_per_s, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(rows >= 1024 && out_bytes_per_s, "fd_debug_l2_gather_probe: bad argument");
  const int row_bytes = 256, n_idx = 1 << 20, stages = 1500;
  char* buf = nullptr;
  int* idx = nullptr;
  FD_CUDA(cudaMalloc(&buf, (size_t)rows * row_bytes));
  FD_CUDA(cudaMalloc(&idx, (size_t)n_idx * 4));
  FD_CUDA(cudaMemsetAsync(buf, 1, (size_t)rows * row_bytes, stream));
  std::vector<int> h(n_idx);
  unsigned s = 12345;
  for (int i = 0; i < n_idx; ++i) {
    s = s * 1664525u + 1013904223u;
    h[i] = (int)(((size_t)i * 3 / 4 + ((s >> 8) % 600)) % (size_t)rows);       // a sorted sparse level: nearby rows, with gaps
  }
  FD_CUDA(cudaMemcpyAsync(idx, h.data(), (size_t)n_idx * 4, cudaMemcpyHostToDevice, stream));
  const size_t smem = (size_t)PROBE_DEPTH * 32768;
  FD_CUDA(cudaFuncSetAttribute(l2_gather_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  FD_CUDA(cudaEventCreate(&e0));
  FD_CUDA(cudaEventCreate(&e1));
  l2_gather_probe_kernel<<<kNumSMs, 256, smem, stream>>>(buf, idx, (unsigned)n_idx - 1u, row_bytes, 200);      // warm-up
  FD_CUDA(cudaEventRecord(e0, stream));
  l2_gather_probe_kernel<<<kNumSMs, 256, smem, stream>>>(buf, idx, (unsigned)n_idx - 1u, row_bytes, stages);
  FD_CUDA(cudaEventRecord(e1, stream));
  FD_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  FD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *out_bytes_per_s = (double)kNumSMs * stages * 32768.0 / ((double)ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(idx);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
