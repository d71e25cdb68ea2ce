// Shared host/device helpers for the futuredet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/futuredet_b200.h"

namespace fd {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char* fmt, ...);

#define FD_REQUIRE(cond, ...)                       \
  do {                                              \
    if (!(cond)) return fd::set_error(-1, __VA_ARGS__); \
  } while (0)

#define FD_CUDA(expr)                                                            \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess)                                                      \
      return fd::set_error((int)e__, "%s failed: %s (%s:%d)", #expr,             \
                           cudaGetErrorString(e__), __FILE__, __LINE__);         \
  } while (0)

// Count + check a kernel launch (the count is what bench.py reports as gpu_launches).
#define FD_LAUNCHED()                                                            \
  do {                                                                           \
    fd::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess)                                                      \
      return fd::set_error((int)e__, "kernel launch failed: %s (%s:%d)",         \
                           cudaGetErrorString(e__), __FILE__, __LINE__);         \
  } while (0)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// Grid for a grid-stride kernel: a whole number of waves of `per_sm` CTAs on 148 SMs.
inline int persistent_grid(int64_t work_blocks, int per_sm) {
  int64_t full = (int64_t)kNumSMs * per_sm;
  if (work_blocks <= 0) return 1;
  return (int)(work_blocks < full ? work_blocks : full);
}

// ---- device-wide exclusive scan of int32 (3 phases, deterministic) ------------
// tmp: fd_scan_tmp_bytes(n).  out may alias in.  total (device int) may be null.
int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total,
                       void* d_tmp, cudaStream_t stream);
// same, input = popcount of 32-bit words
int exclusive_scan_popc(const uint32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total,
                        void* d_tmp, cudaStream_t stream);

__device__ __forceinline__ uint32_t hash64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return (uint32_t)k;
}

constexpr long long kEmptyKey = -1LL;

}  // namespace fd
