// Device-wide exclusive scan (3 phases, deterministic), templated on a load functor.
#pragma once
#include "common.cuh"

namespace fd {

// ------------------------------------------------------------------ scan ------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct LoadI32 {
  const int32_t* p;
  __device__ __forceinline__ int operator()(int64_t i) const { return p[i]; }
};
struct LoadPopc {
  const uint32_t* p;
  __device__ __forceinline__ int operator()(int64_t i) const { return __popc(p[i]); }
};

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread across the block; returns block total via *total.
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_sums[THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int s = lane < THREADS / 32 ? warp_sums[lane] : 0;
    int si = warp_incl_scan(s, lane);
    if (lane < THREADS / 32) warp_sums[lane] = si - s;
    if (lane == 31) block_total = si;
  }
  __syncthreads();
  int r = incl - v + warp_sums[warp];
  *total = block_total;
  __syncthreads();
  return r;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) scan_block_sums(Load ld, int64_t n, int32_t* sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int acc = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
    if (i < n) acc += ld(i);
  }
  // block reduce
  __shared__ int ws[kScanThreads / 32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) t += ws[w];
    sums[blockIdx.x] = t;
  }
}

static __global__ void __launch_bounds__(1024) scan_sums_inplace(int32_t* sums, int nblocks, int32_t* total) {
  int carry = 0;
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblocks ? sums[i] : 0;
    int t;
    int e = block_excl_scan<1024>(v, &t);
    if (i < nblocks) sums[i] = e + carry;
    carry += t;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) scan_apply(Load ld, int64_t n, const int32_t* sums,
                                                            int32_t* out) {
  // blocked arrangement: thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int tsum = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    v[j] = i < n ? ld(i) : 0;
    tsum += v[j];
  }
  int t;
  int e = block_excl_scan<kScanThreads>(tsum, &t) + sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    int64_t i = base + j;
    if (i < n) out[i] = e;
    e += v[j];
  }
}

template <class Load>
inline int scan_impl(Load ld, int32_t* d_out, int64_t n, int32_t* d_total, void* d_tmp,
                     cudaStream_t stream) {
  if (n <= 0) {
    if (d_total) FD_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), stream));
    return 0;
  }
  int nblocks = ceil_div(n, kScanTile);
  int32_t* sums = (int32_t*)d_tmp;
  scan_block_sums<<<nblocks, kScanThreads, 0, stream>>>(ld, n, sums);
  FD_LAUNCHED();
  scan_sums_inplace<<<1, 1024, 0, stream>>>(sums, nblocks, d_total);
  FD_LAUNCHED();
  scan_apply<<<nblocks, kScanThreads, 0, stream>>>(ld, n, sums, d_out);
  FD_LAUNCHED();
  return 0;
}


}  // namespace fd
