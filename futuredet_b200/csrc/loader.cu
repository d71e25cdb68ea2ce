// Multi-sweep point-cloud assembly on the GPU: the step in front of the voxelizer.
//
// Replaces the per-sweep numpy work of det3d/datasets/pipelines/loading.py:
//   :24-33   read_file            raw .bin records (x, y, z, intensity, ring) -> first num_feat columns
//   :36-45   remove_close         drop points with |x| < r and |y| < r (sweeps only, r = 1.0; the key frame keeps all)
//   :48-60   read_sweep           xyz <- (T . [x y z 1])[:3] with the 4x4 float64 sweep-to-keyframe transform
//                                 (numpy computes in float64 and stores float32), time-lag column
//   :120-140 LoadPointCloudFromFile.__call__  concatenation key frame + sweeps, `combined = hstack([points, times])`
// so that raw sweeps are the wire format: one H2D copy of the untouched .bin payloads, then one pass here.
// Output order is the reference's (sweep order, point order inside a sweep): a keep bit per record, an exclusive
// popcount scan for the output row, one scatter.  Rows past the kept count are filled with NaN, which the voxelizer
// rejects, so the whole chain needs no host synchronisation (batch offsets are produced on the device).
#include "common.cuh"
#include "scan.cuh"

namespace fd {

struct SweepArgs {
  const float* raw; int raw_stride; int num_feat;
  const int32_t* sweep_off;        // [S+1] record offsets of every sweep in `raw`
  const double* xform;             // [S,16] row-major 4x4, used when flags & 1
  const int32_t* flags;            // [S] bit0: apply transform, bit1: remove close points
  const float* time_lag;           // [S]
  const int32_t* sweep_scene;      // [S] scene (batch element) of every sweep, non-decreasing
  int S, B; long long total; float radius;
  uint32_t* bits; int32_t* wordprefix; const int32_t* count;
  float* out; int32_t* batch_off;  // [total, num_feat+1], [B+1]
};

__device__ __forceinline__ int find_sweep(const int32_t* off, int S, long long i) {
  int lo = 0, hi = S;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
sweep_keep_kernel(const SweepArgs a) {
  extern __shared__ int32_t s_off[];
  for (int j = threadIdx.x; j <= a.S; j += blockDim.x) s_off[j] = a.sweep_off[j];
  __syncthreads();
  const long long words = (a.total + 31) / 32;
  const int lane = threadIdx.x & 31;
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words;
       w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long i = w * 32 + lane;
    bool keep = false;
    if (i < a.total) {
      keep = true;
      const int s = find_sweep(s_off, a.S, i);
      if (a.flags[s] & 2) {
        const float* p = a.raw + (size_t)i * a.raw_stride;
        keep = !(fabsf(p[0]) < a.radius && fabsf(p[1]) < a.radius);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) a.bits[w] = m;
  }
}

__global__ void __launch_bounds__(256)
sweep_emit_kernel(const SweepArgs a) {
  extern __shared__ int32_t s_off[];
  for (int j = threadIdx.x; j <= a.S; j += blockDim.x) s_off[j] = a.sweep_off[j];
  __syncthreads();
  const int nf = a.num_feat, ostride = a.num_feat + 1;
  const long long kept = *a.count;
  const float qnan = __int_as_float(0x7fc00000);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.total;
       i += (long long)gridDim.x * blockDim.x) {
    if (i >= kept) {                                   // tail of the capacity-sized output: rejected by the voxelizer
      float* o = a.out + (size_t)i * ostride;
      for (int c = 0; c < ostride; ++c) o[c] = qnan;
    }
    const unsigned m = a.bits[i >> 5];
    if (!((m >> (i & 31)) & 1u)) continue;
    const long long row = a.wordprefix[i >> 5] + __popc(m & ((1u << (i & 31)) - 1u));
    const int s = find_sweep(s_off, a.S, i);
    const float* p = a.raw + (size_t)i * a.raw_stride;
    float* o = a.out + (size_t)row * ostride;
    float x = p[0], y = p[1], z = p[2];
    if (a.flags[s] & 1) {
      // float64 like numpy's `transform_matrix.dot(vstack(xyz, ones))` (loading.py:55-57), stored back as float32.
      // numpy hands the [4,4]x[4,n] product to BLAS dgemm, whose micro-kernel accumulates over k = 0..3 with one FMA per
      // term starting from zero: rn(T0*x), fma(T1,y,.), fma(T2,z,.), fma(T3,1,.) = rn(. + T3).  The explicit intrinsics
      // pin exactly that order (a bare expression lets nvcc pick which product feeds the FMA, which flips the last
      // float64 bit -- and now and then the float32 rounding -- of some coordinates).
      const double* T = a.xform + (size_t)s * 16;
      const double dx = x, dy = y, dz = z;
      x = (float)__dadd_rn(__fma_rn(T[2], dz, __fma_rn(T[1], dy, __dmul_rn(T[0], dx))), T[3]);
      y = (float)__dadd_rn(__fma_rn(T[6], dz, __fma_rn(T[5], dy, __dmul_rn(T[4], dx))), T[7]);
      z = (float)__dadd_rn(__fma_rn(T[10], dz, __fma_rn(T[9], dy, __dmul_rn(T[8], dx))), T[11]);
    }
    o[0] = x; o[1] = y; o[2] = z;
    for (int c = 3; c < nf; ++c) o[c] = p[c];
    o[nf] = a.time_lag[s];
  }
}

__global__ void sweep_batch_offsets_kernel(const SweepArgs a) {
  // first output row of every scene = rank of the first record of its first sweep
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b <= a.B; b += gridDim.x * blockDim.x) {
    if (b == a.B) { a.batch_off[b] = *a.count; continue; }
    int s = 0;
    while (s < a.S && a.sweep_scene[s] < b) ++s;
    const long long i = s < a.S ? a.sweep_off[s] : a.total;
    int r;
    if (i >= a.total) r = *a.count;
    else r = a.wordprefix[i >> 5] + __popc(a.bits[i >> 5] & ((1u << (i & 31)) - 1u));
    a.batch_off[b] = r;
  }
}

}  // namespace fd

extern "C" {

size_t fd_sweeps_workspace_bytes(int64_t total_records) {
  if (total_records < 0) return 0;
  const int64_t words = (total_records + 31) / 32 + 1;
  return (size_t)words * 8 + fd_scan_tmp_bytes(words) + 1024;
}

int fd_assemble_sweeps(const float* d_raw, int64_t total_records, int raw_stride, int num_feat,
                       const int32_t* d_sweep_offsets, const double* d_transforms, const int32_t* d_flags,
                       const float* d_time_lag, const int32_t* d_sweep_scene, int S, int B, float close_radius,
                       float* d_points, int32_t* d_batch_offsets, int32_t* d_count, void* d_workspace,
                       size_t workspace_bytes, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(total_records >= 0 && total_records < 0x7f000000LL, "fd_assemble_sweeps: bad record count");
  FD_REQUIRE(S >= 1 && S <= 8192 && B >= 1 && B <= S, "fd_assemble_sweeps: need 1 <= B <= S <= 8192");
  FD_REQUIRE(num_feat >= 3 && raw_stride >= num_feat, "fd_assemble_sweeps: need 3 <= num_feat <= raw_stride");
  FD_REQUIRE(d_sweep_offsets && d_transforms && d_flags && d_time_lag && d_sweep_scene && d_points && d_batch_offsets &&
                 d_count && d_workspace, "fd_assemble_sweeps: null argument");
  FD_REQUIRE(d_raw || total_records == 0, "fd_assemble_sweeps: null records");
  FD_REQUIRE(workspace_bytes >= fd_sweeps_workspace_bytes(total_records), "fd_assemble_sweeps: workspace too small");
  const int64_t words = (total_records + 31) / 32;
  SweepArgs a{};
  a.raw = d_raw; a.raw_stride = raw_stride; a.num_feat = num_feat;
  a.sweep_off = d_sweep_offsets; a.xform = d_transforms; a.flags = d_flags; a.time_lag = d_time_lag;
  a.sweep_scene = d_sweep_scene; a.S = S; a.B = B; a.total = total_records; a.radius = close_radius;
  a.bits = (uint32_t*)d_workspace;
  a.wordprefix = (int32_t*)((char*)d_workspace + (size_t)(words + 1) * 4);
  void* scan_tmp = (char*)d_workspace + (size_t)(words + 1) * 8;
  a.count = d_count; a.out = d_points; a.batch_off = d_batch_offsets;
  const size_t sh = (size_t)(S + 1) * sizeof(int32_t);
  if (total_records > 0) {
    sweep_keep_kernel<<<persistent_grid(ceil_div(total_records, 256), 8), 256, sh, stream>>>(a);
    FD_LAUNCHED();
  }
  int rc = exclusive_scan_popc(a.bits, a.wordprefix, words, d_count, scan_tmp, stream);
  if (rc) return rc;
  if (total_records > 0) {
    sweep_emit_kernel<<<persistent_grid(ceil_div(total_records, 256), 8), 256, sh, stream>>>(a);
    FD_LAUNCHED();
  }
  sweep_batch_offsets_kernel<<<1, 128, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
