// L2 -> SM gather roofline probe (perf triage, sm_100a): every CTA gathers random 128-byte row pieces from a buffer that
// fits L2 with 16-byte cp.async into a shared-memory ring, exactly like the A producers of conv_tc_kernel, with nothing
// else going on.  Prints achieved TB/s for several amounts of bytes in flight per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2_gather_probe tools/l2_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// stage = 128 rows x `pieces` 128-byte pieces; DEPTH stages in flight per CTA; 256 threads: thread -> (row, 16-byte chunk)
template <int DEPTH>
__global__ void __launch_bounds__(256, 1)
gather_kernel(const char* __restrict__ buf, const int* __restrict__ idx, int n_idx, int row_bytes, int pieces, int stages) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const int tid = threadIdx.x, chunk = tid & 7, r0 = tid >> 3;          // 32 rows per pass, 4 passes
  const int stage_bytes = 128 * pieces * 128;
  // n_idx is a power of two; the next stage's row indices are loaded before the current stage is issued (as the
  // producers of conv_tc_kernel do), so no index round trip sits on the issue path
  const unsigned mask = (unsigned)n_idx - 1u;
  unsigned pos = (blockIdx.x * 7919u * 128u) & mask;
  int nxt[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) nxt[q] = __ldg(idx + ((pos + r0 + 32 * q) & mask));
  for (int s = 0; s < stages; ++s) {
    const uint32_t dst = s0 + (uint32_t)(s % DEPTH) * stage_bytes;
    int cur[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
    pos = (pos + 128u * 148u) & mask;
#pragma unroll
    for (int q = 0; q < 4; ++q) nxt[q] = __ldg(idx + ((pos + r0 + 32 * q) & mask));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + 32 * q;
      for (int p = 0; p < pieces; ++p)
        cp_async16(dst + (uint32_t)((p * 128 + r) * 128 + chunk * 16), buf + (size_t)cur[q] * row_bytes + p * 128 + chunk * 16);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    wait_group<DEPTH - 1>();
  }
  wait_group<0>();
}

template <int DEPTH>
static void run(const char* buf, const int* idx, int n_idx, int row_bytes, int pieces, const char* what) {
  const int stages = 2000;
  const size_t smem = (size_t)DEPTH * 128 * pieces * 128;
  cudaFuncSetAttribute(gather_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_kernel<DEPTH><<<148, 256, smem>>>(buf, idx, n_idx, row_bytes, pieces, 200);
  cudaEventRecord(e0);
  gather_kernel<DEPTH><<<148, 256, smem>>>(buf, idx, n_idx, row_bytes, pieces, stages);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 148.0 * stages * 128 * pieces * 128;
  printf("%-28s depth %d (%3zu KB in flight/SM): %.2f TB/s  (%.1f B/clk/SM at 1.965 GHz)  err=%s\n", what, DEPTH, smem >> 10,
         bytes / ms / 1e9, bytes / ms / 1e6 / 148 / 1965.0 * 1e3 / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows = 212740, row_bytes = 256;                 // a 64-channel split-bf16 level: 54 MB
  char* buf; int* idx;
  cudaMalloc(&buf, (size_t)rows * row_bytes);
  cudaMemset(buf, 1, (size_t)rows * row_bytes);
  {
    // ~2 s of load first: a process that starts measuring right away sees the idle clocks (1.3 GHz instead of 1.965 GHz
    // on this pool -- the first version of this tool reported 11.7 TB/s for what is 17.3 TB/s at the clocks the
    // forward pass runs at)
    int* widx;
    std::vector<int> w(1 << 20);
    for (size_t i = 0; i < w.size(); ++i) w[i] = (int)(i % rows);
    cudaMalloc(&widx, w.size() * 4);
    cudaMemcpy(widx, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(gather_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * 2 * 128);
    for (int i = 0; i < 150; ++i) gather_kernel<4><<<148, 256, 4 * 128 * 2 * 128>>>(buf, widx, (int)w.size(), row_bytes, 2, 2000);
    cudaDeviceSynchronize();
    cudaFree(widx);
  }
  std::vector<int> h(1 << 22);
  // (a) random rows  (b) neighbour-like: row + small offsets (what a sorted sparse level looks like)
  for (int mode = 0; mode < 2; ++mode) {
    unsigned s = 12345;
    for (size_t i = 0; i < h.size(); ++i) {
      s = s * 1664525u + 1013904223u;
      h[i] = mode == 0 ? (int)((s >> 8) % rows) : (int)((i * 3 / 4 + ((s >> 8) % 600)) % rows);
    }
    cudaMalloc(&idx, h.size() * 4);
    cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    const char* w = mode == 0 ? "random rows, hi+lo (2x128B)" : "local rows, hi+lo (2x128B)";
    run<1>(buf, idx, (int)h.size(), row_bytes, 2, w);
    run<2>(buf, idx, (int)h.size(), row_bytes, 2, w);
    run<4>(buf, idx, (int)h.size(), row_bytes, 2, w);
    run<6>(buf, idx, (int)h.size(), row_bytes, 2, w);
    const char* w1 = mode == 0 ? "random rows, 128B" : "local rows, 128B";
    run<4>(buf, idx, (int)h.size(), row_bytes, 1, w1);
    run<8>(buf, idx, (int)h.size(), row_bytes, 1, w1);
    run<12>(buf, idx, (int)h.size(), row_bytes, 1, w1);
    cudaFree(idx);
  }
  return 0;
}
