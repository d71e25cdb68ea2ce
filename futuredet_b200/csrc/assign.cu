// CenterPoint target assignment on the GPU: the step that feeds CenterHead.loss.
//
// Replaces the per-object Python loops of det3d/datasets/pipelines/preprocess.py:464-546 (AssignLabel.__call__, one
// task, one timestep: Gaussian radius, heat-map splat, ind / mask / cat, anno_box encoding) together with
//   det3d/core/utils/center_utils.py:17-39   gaussian_radius
//   det3d/core/utils/center_utils.py:41-64   gaussian2D + draw_umich_gaussian (element-wise max into the heat map)
//   det3d/core/bbox/box_np_ops.py limit_period on rot / rrot (preprocess.py:449-456).
// One CTA per (sample, object): thread 0 reproduces the reference's float32 scalar arithmetic (numpy keeps float32
// through every step of the radius / centre computation, so the integer radius and cell are bit-exact), all threads
// splat the (2r+1)^2 window in float64 like gaussian2D and merge with an integer atomicMax on the float bits (the
// reference's np.maximum is order independent, so the result is deterministic).
#include "common.cuh"

namespace fd {

struct AssignArgs {
  const float* boxes; const int32_t* cls; const int32_t* num;   // [B,n_max,box_dim], [B,n_max] (1-based), [B]
  int B, n_max, box_dim, num_cls, W, H;
  float x0, y0, vx, vy, osf, overlap, overlap_p1, overlap_m1;
  int min_radius, radius_mult, timestep, max_objs;
  float* hm; float* anno; long long* ind; unsigned char* mask; long long* cat;   // anno [B,max_objs,14]
};

__device__ __forceinline__ float limit_period_f32(float v) {
  const float P = 6.283185307179586f;                 // float32(2*pi): numpy converts the python scalar first
  const float t = floorf(__fadd_rn(__fdiv_rn(v, P), 0.5f));
  return __fsub_rn(v, __fmul_rn(t, P));
}

// gaussian_radius((height, width) = (l, w), min_overlap) in float32, operation by operation
__device__ float gaussian_radius_f32(float height, float width, float ov, float ov_p1, float ov_m1) {
  const float hw = __fadd_rn(height, width), wh = __fmul_rn(width, height);
  const float c1 = __fdiv_rn(__fmul_rn(wh, ov_m1), ov_p1);                          // w*h*(1-ov)/(1+ov)
  const float sq1 = sqrtf(__fsub_rn(__fmul_rn(hw, hw), __fmul_rn(4.f, c1)));
  const float r1 = __fdiv_rn(__fadd_rn(hw, sq1), 2.f);
  const float b2 = __fmul_rn(2.f, hw);
  const float c2 = __fmul_rn(__fmul_rn(ov_m1, width), height);                      // (1-ov)*w*h
  const float sq2 = sqrtf(__fsub_rn(__fmul_rn(b2, b2), __fmul_rn(16.f, c2)));
  const float r2 = __fdiv_rn(__fadd_rn(b2, sq2), 2.f);
  const float a3x4 = (float)(4.0 * (4.0 * (double)ov));                              // python: 4 * a3, a3 = 4 * ov
  const float b3 = __fmul_rn((float)(-2.0 * (double)ov), hw);
  const float c3 = __fmul_rn(__fmul_rn((float)((double)ov - 1.0), width), height);
  const float sq3 = sqrtf(__fsub_rn(__fmul_rn(b3, b3), __fmul_rn(a3x4, c3)));
  const float r3 = __fdiv_rn(__fadd_rn(b3, sq3), 2.f);
  return fminf(r1, fminf(r2, r3));
}

__global__ void __launch_bounds__(128)
assign_targets_kernel(const AssignArgs a) {
  __shared__ int s_ok, s_r, s_cx, s_cy, s_cls;
  const int b = blockIdx.y, k = blockIdx.x;
  const int n = min(min(a.num[b], a.n_max), a.max_objs);
  if (k >= n) return;
  const float* box = a.boxes + ((size_t)b * a.n_max + k) * a.box_dim;
  if (threadIdx.x == 0) {
    s_ok = 0;
    const int cls_id = a.cls[(size_t)b * a.n_max + k] - 1;
    const float w = __fdiv_rn(__fdiv_rn(box[3], a.vx), a.osf), l = __fdiv_rn(__fdiv_rn(box[4], a.vy), a.osf);
    if (cls_id >= 0 && cls_id < a.num_cls && w > 0.f && l > 0.f) {
      float mult = 1.f;
      if (a.radius_mult) {
        const float vn = sqrtf(__fadd_rn(__fmul_rn(box[6], box[6]), __fmul_rn(box[7], box[7])));
        mult = fminf(fmaxf(1.f, __fdiv_rn(__fmul_rn(vn, (float)(1 + a.timestep)), 2.f)), 4.f);
      }
      const float rad = __fmul_rn(mult, gaussian_radius_f32(l, w, a.overlap, a.overlap_p1, a.overlap_m1));
      const int radius = max(a.min_radius, (int)rad);
      const float cxf = __fdiv_rn(__fdiv_rn(__fsub_rn(box[0], a.x0), a.vx), a.osf);
      const float cyf = __fdiv_rn(__fdiv_rn(__fsub_rn(box[1], a.y0), a.vy), a.osf);
      const int cx = (int)cxf, cy = (int)cyf;                              // astype(int32): truncation
      if (cx >= 0 && cx < a.W && cy >= 0 && cy < a.H) {
        s_ok = 1; s_r = radius; s_cx = cx; s_cy = cy; s_cls = cls_id;
        const size_t o = (size_t)b * a.max_objs + k;
        a.cat[o] = cls_id;
        a.ind[o] = (long long)cy * a.W + cx;
        a.mask[o] = 1;
        float* ab = a.anno + o * 14;
        ab[0] = __fsub_rn(cxf, (float)cx);
        ab[1] = __fsub_rn(cyf, (float)cy);
        ab[2] = box[2];
        ab[3] = logf(box[3]); ab[4] = logf(box[4]); ab[5] = logf(box[5]);
        ab[6] = box[6]; ab[7] = box[7]; ab[8] = box[8]; ab[9] = box[9];
        const float rot = limit_period_f32(box[10]), rrot = limit_period_f32(box[11]);   // box_dim == 12: columns -2, -1
        ab[10] = sinf(rot); ab[11] = cosf(rot); ab[12] = sinf(rrot); ab[13] = cosf(rrot);
      }
    }
  }
  __syncthreads();
  if (!s_ok) return;
  const int r = s_r, cx = s_cx, cy = s_cy;
  const int left = min(cx, r), right = min(a.W - cx, r + 1), top = min(cy, r), bottom = min(a.H - cy, r + 1);
  const int ww = left + right, wh = top + bottom;
  if (ww <= 0 || wh <= 0) return;
  const double sigma = (double)(2 * r + 1) / 6.0;
  const double inv = 1.0 / (2.0 * sigma * sigma);
  int* plane = reinterpret_cast<int*>(a.hm + ((size_t)b * a.num_cls + s_cls) * a.H * a.W);
  for (int e = threadIdx.x; e < ww * wh; e += blockDim.x) {
    const int dy = e / ww - top, dx = e - (e / ww) * ww - left;
    double g = exp(-(double)(dx * dx + dy * dy) * inv);
    if (g < 2.220446049250313e-16) g = 0.0;                               // gaussian2D: h[h < eps * h.max()] = 0
    const float gf = (float)g;
    atomicMax(plane + (size_t)(cy + dy) * a.W + (cx + dx), __float_as_int(gf));     // values >= 0: int order == float order
  }
}

}  // namespace fd

extern "C" {

int fd_assign_center_targets(const float* d_boxes, const int32_t* d_classes, const int32_t* d_num, int B, int n_max,
                             int box_dim, int num_cls, int W, int H, float pc_x0, float pc_y0, float voxel_x,
                             float voxel_y, float out_size_factor, float gaussian_overlap, int min_radius,
                             int radius_mult, int timestep, int max_objs, float* d_hm, float* d_anno_box,
                             int64_t* d_ind, uint8_t* d_mask, int64_t* d_cat, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_boxes && d_classes && d_num && d_hm && d_anno_box && d_ind && d_mask && d_cat,
             "fd_assign_center_targets: null argument");
  FD_REQUIRE(B >= 1 && B <= 65535 && n_max >= 1 && box_dim == 12 && num_cls >= 1 && W >= 1 && H >= 1 && max_objs >= 1,
             "fd_assign_center_targets: bad shape (nuScenes boxes: x,y,z,w,l,h,vx,vy,rvx,rvy,rot,rrot)");
  FD_CUDA(cudaMemsetAsync(d_hm, 0, sizeof(float) * (size_t)B * num_cls * H * W, stream));
  FD_CUDA(cudaMemsetAsync(d_anno_box, 0, sizeof(float) * (size_t)B * max_objs * 14, stream));
  FD_CUDA(cudaMemsetAsync(d_ind, 0, sizeof(int64_t) * (size_t)B * max_objs, stream));
  FD_CUDA(cudaMemsetAsync(d_mask, 0, (size_t)B * max_objs, stream));
  FD_CUDA(cudaMemsetAsync(d_cat, 0, sizeof(int64_t) * (size_t)B * max_objs, stream));
  AssignArgs a{};
  a.boxes = d_boxes; a.cls = d_classes; a.num = d_num; a.B = B; a.n_max = n_max; a.box_dim = box_dim;
  a.num_cls = num_cls; a.W = W; a.H = H; a.x0 = pc_x0; a.y0 = pc_y0; a.vx = voxel_x; a.vy = voxel_y;
  a.osf = out_size_factor; a.overlap = gaussian_overlap;
  a.overlap_p1 = (float)(1.0 + (double)gaussian_overlap);      // python computes 1 +/- min_overlap in float64 first
  a.overlap_m1 = (float)(1.0 - (double)gaussian_overlap);
  a.min_radius = min_radius; a.radius_mult = radius_mult; a.timestep = timestep; a.max_objs = max_objs;
  a.hm = d_hm; a.anno = d_anno_box; a.ind = (long long*)d_ind; a.mask = d_mask; a.cat = (long long*)d_cat;
  const int n = n_max < max_objs ? n_max : max_objs;
  assign_targets_kernel<<<dim3(n, B), 128, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
