// Fused, deterministic voxelize + mean-VFE for sm_100a.
//
// Reproduces the *sequential* semantics of the reference loop
// (det3d/ops/point_cloud/point_cloud_ops.py:7-55) with data-parallel passes:
//   P1  every point -> float32 voxel coordinate (true IEEE sub/div/floor, :36),
//       claim a hash slot for its voxel and atomicMin the voxel's first point index;
//   P2  rank voxels by first point index = exclusive scan over "is first point"
//       flags (voxel id == order of first appearance, :44-50);
//   P3  voxels whose per-scene rank >= max_voxels are dropped (:46-47); every
//       surviving point bubbles its index into the voxel's sorted list of the
//       max_points smallest indices (first max_points points in input order, :51-54);
//   P4  one thread per voxel sums those points in input order and divides by the
//       count (det3d/models/readers/voxel_encoder.py:20-22), writes features,
//       (b,z,y,x) coordinates (collate.py:199-206) and num_points.
// All indices are bit-exact with the reference; no atomics on floats anywhere.
#include "common.cuh"
#include "scan.cuh"

namespace fd {

struct VoxGeom {
  float lo[3];
  float vs[3];
  int grid[3];  // gx, gy, gz
};

constexpr int kSlotEmpty = 0x7f7f7f7f;  // memset(0x7f) pattern; > any point index we accept

__device__ __forceinline__ int find_scene(const int32_t* s_off, int B, int i) {
  int lo = 0, hi = B;  // offsets[lo] <= i < offsets[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (s_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- P1 -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vox_hash_points(const float* __restrict__ pts, int n, int pstride, const int32_t* __restrict__ boff,
                int B, VoxGeom g, long long* __restrict__ keys, int* __restrict__ first,
                uint32_t mask, uint32_t scene_cap, int* __restrict__ pslot) {
  extern __shared__ int32_t s_off[];
  for (int j = threadIdx.x; j <= B; j += blockDim.x) s_off[j] = boff[j];
  __syncthreads();
  const long long cells = (long long)g.grid[0] * g.grid[1] * g.grid[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = pts + (size_t)i * pstride;
    int c[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float f = floorf(__fdiv_rn(__fsub_rn(p[j], g.lo[j]), g.vs[j]));
      ok = ok && (f >= 0.0f) && (f < (float)g.grid[j]);
      c[j] = (int)f;
    }
    int slot = -1;
    if (ok) {
      int b = find_scene(s_off, B, i);
      long long key = (long long)b * cells + ((long long)c[2] * g.grid[1] + c[1]) * g.grid[0] + c[0];
      // every scene hashes into its own region of the table (probing may spill into the next one): blocks run in
      // launch order over scene-contiguous points, so the live part of the table stays L2 resident in a big batch
      uint32_t h = ((uint32_t)b * scene_cap + (hash64((uint64_t)key) & (scene_cap - 1))) & mask;
      while (true) {
        long long prev = atomicCAS((unsigned long long*)&keys[h], (unsigned long long)kEmptyKey,
                                   (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) break;
        h = (h + 1) & mask;
      }
      atomicMin(&first[h], i);
      slot = (int)h;
    }
    pslot[i] = slot;
  }
}

// ---- P2: flag functor for the scan --------------------------------------------
struct LoadIsFirst {
  const int* pslot;
  const int* first;
  __device__ __forceinline__ int operator()(int64_t i) const {
    int s = pslot[i];
    return (s >= 0 && first[s] == (int)i) ? 1 : 0;
  }
};

// ---- per-scene bookkeeping (tiny) ----------------------------------------------
__global__ void vox_scene_counts(const int32_t* __restrict__ rank, const int32_t* __restrict__ total_first,
                                 const int32_t* __restrict__ boff, int B, int n, int max_voxels,
                                 int32_t* __restrict__ scene_rank0, int32_t* __restrict__ out_base,
                                 int32_t* __restrict__ nvox, int32_t* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int run = 0;
  for (int b = 0; b < B; ++b) {
    int o0 = boff[b], o1 = boff[b + 1];
    int r0 = o0 < n ? rank[o0] : *total_first;
    int r1 = o1 < n ? rank[o1] : *total_first;
    int cnt = r1 - r0;
    if (cnt > max_voxels) cnt = max_voxels;
    scene_rank0[b] = r0;
    out_base[b] = run;
    nvox[b] = cnt;
    run += cnt;
  }
  out_base[B] = run;
  *total = run;
}

// ---- P3 -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vox_assign_points(int n, const int32_t* __restrict__ boff, int B, VoxGeom g,
                  const long long* __restrict__ keys, const int* __restrict__ first,
                  const int* __restrict__ pslot, const int32_t* __restrict__ rank,
                  const int32_t* __restrict__ scene_rank0, const int32_t* __restrict__ out_base,
                  int max_voxels, int max_points, int* __restrict__ slots, int32_t* __restrict__ coords) {
  extern __shared__ int32_t s_off[];
  int32_t* s_r0 = s_off + (B + 1);
  int32_t* s_base = s_r0 + B;
  for (int j = threadIdx.x; j <= B; j += blockDim.x) s_off[j] = boff[j];
  for (int j = threadIdx.x; j < B; j += blockDim.x) { s_r0[j] = scene_rank0[j]; s_base[j] = out_base[j]; }
  __syncthreads();
  const long long cells = (long long)g.grid[0] * g.grid[1] * g.grid[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int ps = pslot[i];
    if (ps < 0) continue;
    int f = first[ps];
    int b = find_scene(s_off, B, i);
    int local = rank[f] - s_r0[b];
    if (local >= max_voxels) continue;
    int vid = s_base[b] + local;
    if (f == i) {
      long long lin = keys[ps] - (long long)b * cells;
      int x = (int)(lin % g.grid[0]);
      long long t = lin / g.grid[0];
      int y = (int)(t % g.grid[1]);
      int z = (int)(t / g.grid[1]);
      reinterpret_cast<int4*>(coords)[vid] = make_int4(b, z, y, x);
    }
    // bubble i into the ascending list of the max_points smallest indices of this voxel
    int* sl = slots + (size_t)vid * max_points;
    int cur = i;
    for (int s = 0; s < max_points; ++s) {
      int old = atomicMin(&sl[s], cur);
      if (old == kSlotEmpty) break;
      cur = old > cur ? old : cur;
    }
  }
}

// ---- P4 -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vox_reduce_mean(const float* __restrict__ pts, int pstride, int num_feat, const int* __restrict__ slots,
                int max_points, const int32_t* __restrict__ total, float* __restrict__ feat,
                int feat_stride, int32_t* __restrict__ npts, float* __restrict__ voxels) {
  const int nv = *total;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const int* sl = slots + (size_t)v * max_points;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    int cnt = 0;
    for (int s = 0; s < max_points; ++s) {
      int idx = sl[s];
      if (idx == kSlotEmpty) break;
      const float* p = pts + (size_t)idx * pstride;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < num_feat) {
          float q = p[c];
          acc[c] = __fadd_rn(acc[c], q);
          if (voxels) voxels[((size_t)v * max_points + s) * num_feat + c] = q;
        }
      ++cnt;
    }
    if (voxels)
      for (int e = cnt * num_feat; e < max_points * num_feat; ++e) voxels[(size_t)v * max_points * num_feat + e] = 0.f;
    const float fc = (float)cnt;
    float* o = feat + (size_t)v * feat_stride;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < feat_stride) o[c] = c < num_feat ? __fdiv_rn(acc[c], fc) : 0.f;
    for (int c = 8; c < feat_stride; ++c) o[c] = 0.f;
    npts[v] = cnt;
  }
}

struct VoxWorkspace {
  long long* keys; int* first; int* pslot; int32_t* rank; int* slots;
  int32_t* scene_rank0; int32_t* out_base; int32_t* total_first; void* scan_tmp;
  int64_t cap; size_t bytes;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static VoxWorkspace carve(void* base, int64_t n, int B, int max_voxels, int max_points) {
  VoxWorkspace w{};
  int64_t cap = 1024;
  while (cap < 4 * n) cap <<= 1;      // load factor <= 0.25 overall, <= 0.5 inside a scene region for any B
  w.cap = cap;
  char* p = (char*)base;
  size_t off = 0;
  auto take = [&](size_t bytes) { char* r = p ? p + off : nullptr; off += align_up(bytes); return (void*)r; };
  // [keys] is memset to 0xff; [first | slots] are contiguous and memset to 0x7f
  w.keys = (long long*)take(sizeof(long long) * cap);
  w.first = (int*)take(sizeof(int) * cap);
  w.slots = (int*)take(sizeof(int) * (size_t)B * max_voxels * max_points);
  w.pslot = (int*)take(sizeof(int) * (n > 0 ? n : 1));
  w.rank = (int32_t*)take(sizeof(int32_t) * (n > 0 ? n : 1));
  w.scene_rank0 = (int32_t*)take(sizeof(int32_t) * (B + 1));
  w.out_base = (int32_t*)take(sizeof(int32_t) * (B + 1));
  w.total_first = (int32_t*)take(sizeof(int32_t));
  w.scan_tmp = take(fd_scan_tmp_bytes(n));
  w.bytes = off;
  return w;
}

// VoxelFeatureExtractorV3 on padded voxels: one thread per (voxel, channel)
__global__ void __launch_bounds__(256)
vfe_mean_kernel(const float* __restrict__ voxels, const int32_t* __restrict__ npts, long long M, int S, int F,
                int num_feat, float* __restrict__ mean) {
  const long long total = M * num_feat;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long v = e / num_feat;
    int c = (int)(e - v * num_feat);
    const float* p = voxels + (size_t)v * S * F + c;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = __fadd_rn(acc, p[(size_t)s * F]);
    mean[e] = __fdiv_rn(acc, (float)npts[v]);
  }
}

}  // namespace fd

extern "C" {

int fd_vfe_mean(const float* d_voxels, const int32_t* d_npts, int64_t M, int S, int F, int num_feat, float* d_mean,
                void* stream) {
  using namespace fd;
  FD_REQUIRE(M >= 0 && S >= 1 && F >= 1 && num_feat >= 1 && num_feat <= F, "fd_vfe_mean: bad shape");
  if (M == 0) return 0;
  FD_REQUIRE(d_voxels && d_npts && d_mean, "fd_vfe_mean: null argument");
  vfe_mean_kernel<<<persistent_grid(ceil_div(M * num_feat, 256), 8), 256, 0, (cudaStream_t)stream>>>(
      d_voxels, d_npts, M, S, F, num_feat, d_mean);
  FD_LAUNCHED();
  return 0;
}


size_t fd_voxelize_workspace_bytes(int64_t total_points, int B, int max_voxels, int max_points) {
  if (total_points < 0 || B < 1 || max_voxels < 1 || max_points < 1) return 0;
  return fd::carve(nullptr, total_points, B, max_voxels, max_points).bytes;
}

int fd_voxelize_vfe(const float* d_points, int64_t total_points, int point_stride, int num_feat,
                    const int32_t* d_batch_offsets, int B, const float* range6,
                    const float* voxel_size3, const int32_t* grid3, int max_points, int max_voxels,
                    float* d_feat, int feat_stride, int32_t* d_coords, int32_t* d_npts,
                    int32_t* d_nvox, int32_t* d_total, float* d_voxels, void* d_workspace,
                    size_t workspace_bytes, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(total_points >= 0 && total_points < 0x7f000000LL, "fd_voxelize_vfe: total_points %lld out of range",
             (long long)total_points);
  FD_REQUIRE(B >= 1 && B <= 4096, "fd_voxelize_vfe: B=%d out of range [1,4096]", B);
  FD_REQUIRE(num_feat >= 3 && num_feat <= 8 && point_stride >= num_feat,
             "fd_voxelize_vfe: need 3 <= num_feat(%d) <= 8 and point_stride(%d) >= num_feat", num_feat, point_stride);
  FD_REQUIRE(feat_stride >= num_feat, "fd_voxelize_vfe: feat_stride %d < num_feat %d", feat_stride, num_feat);
  FD_REQUIRE(max_points >= 1 && max_points <= 64 && max_voxels >= 1, "fd_voxelize_vfe: bad max_points/max_voxels");
  FD_REQUIRE(range6 && voxel_size3 && grid3 && d_batch_offsets && d_feat && d_coords && d_npts && d_nvox && d_total,
             "fd_voxelize_vfe: null argument");
  FD_REQUIRE(d_points != nullptr || total_points == 0, "fd_voxelize_vfe: null points");
  FD_REQUIRE(((uintptr_t)d_coords & 15) == 0, "fd_voxelize_vfe: d_coords must be 16-byte aligned");
  FD_REQUIRE((int64_t)B * max_voxels * max_points < 0x7fffffffLL, "fd_voxelize_vfe: B*max_voxels*max_points overflows");
  VoxWorkspace w = carve(d_workspace, total_points, B, max_voxels, max_points);
  FD_REQUIRE(d_workspace && workspace_bytes >= w.bytes, "fd_voxelize_vfe: workspace %zu < required %zu",
             workspace_bytes, w.bytes);
  VoxGeom g;
  for (int j = 0; j < 3; ++j) {
    g.lo[j] = range6[j]; g.vs[j] = voxel_size3[j]; g.grid[j] = grid3[j];
    FD_REQUIRE(grid3[j] > 0 && voxel_size3[j] > 0.f, "fd_voxelize_vfe: bad grid/voxel size");
  }
  const int n = (int)total_points;
  FD_CUDA(cudaMemsetAsync(w.keys, 0xff, sizeof(long long) * w.cap, stream));
  // first[] and slots[] are adjacent in the workspace: one memset covers both
  FD_CUDA(cudaMemsetAsync(w.first, 0x7f, (char*)w.pslot - (char*)w.first, stream));
  const int threads = 256;
  const int grid_pts = ceil_div(n, threads) > 0 ? ceil_div(n, threads) : 1;   // one block per 256 consecutive points
  uint32_t scene_cap = 1024;                                                   // per-scene region: pow2 >= 2 * mean scene size
  while ((int64_t)scene_cap * B < w.cap) scene_cap <<= 1;
  if ((int64_t)scene_cap * B > w.cap) scene_cap >>= 1;
  if (n > 0) {
    vox_hash_points<<<grid_pts, threads, (B + 1) * sizeof(int32_t), stream>>>(
        d_points, n, point_stride, d_batch_offsets, B, g, w.keys, w.first, (uint32_t)(w.cap - 1), scene_cap, w.pslot);
    FD_LAUNCHED();
  }
  // rank[i] = number of voxel-first points before i
  {
    int rc = scan_impl(LoadIsFirst{w.pslot, w.first}, w.rank, n, w.total_first, w.scan_tmp, stream);
    if (rc) return rc;
  }
  vox_scene_counts<<<1, 32, 0, stream>>>(w.rank, w.total_first, d_batch_offsets, B, n, max_voxels,
                                         w.scene_rank0, w.out_base, d_nvox, d_total);
  FD_LAUNCHED();
  if (n > 0) {
    vox_assign_points<<<grid_pts, threads, (3 * B + 1) * sizeof(int32_t), stream>>>(
        n, d_batch_offsets, B, g, w.keys, w.first, w.pslot, w.rank, w.scene_rank0, w.out_base,
        max_voxels, max_points, w.slots, d_coords);
    FD_LAUNCHED();
    const int64_t vcap = (int64_t)B * max_voxels;
    vox_reduce_mean<<<persistent_grid(ceil_div(vcap < n ? vcap : n, threads), 8), threads, 0, stream>>>(
        d_points, point_stride, num_feat, w.slots, max_points, d_total, d_feat, feat_stride, d_npts, d_voxels);
    FD_LAUNCHED();
  }
  return 0;
}

}  // extern "C"
