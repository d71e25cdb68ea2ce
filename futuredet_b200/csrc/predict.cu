// CenterHead.predict (standard mode) for sm_100a: fused decode + score/range mask + top-k + rotated BEV NMS on device.
//
// Replaces  det3d/models/bbox_heads/center_head.py:541-747   (predict + post_processing: NHWC permute, sigmoid / exp /
//                                                              atan2, meshgrid decode, masks, per-timestep deepcopy)
//           det3d/core/bbox/box_torch_ops.py:248-276          (rotate_nms_pcdet: pcdet box convention, score sort,
//                                                              pre_max / post_max truncation)
//           det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu:104-311 and iou3d_nms.cpp:90-136
//                                                             (64x64 bitmask IoU kernel, D2H copy, host greedy sweep,
//                                                              cudaMalloc/cudaFree per call).
// Design: in the standard mode every forecast timestep shares hm / reg / height / dim / rot (only `vel` differs,
// center_head.py:561-570) and the NMS boxes carry no velocity, so candidate selection and NMS run ONCE per sample and
// the kept boxes are emitted once per timestep.  Four launches per batch: (1) every BEV cell -> a sortable 64-bit key
// (score bits | inverted cell index; 0 = rejected); (2) one 1024-thread CTA per sample: exact radix select of the
// `pre_max` largest keys, bitonic sort in shared memory, decode of the NMS boxes; (3) the upper-triangular IoU bit
// matrix in 64 x 64 tiles over the whole GPU, far-apart pairs decided by a circumscribed-circle test; (4) one warp per
// sample sweeps greedily in score order with the suppressed set in registers, then the CTA emits -- no host round trip,
// no allocation.
// The polygon-clipping arithmetic follows the reference kernel operation by operation (fp32, same evaluation order),
// so the keep / suppress decisions agree with it.
#include "common.cuh"

namespace fd {

constexpr int PRED_MAX_PRE = 1024;     // candidates entering NMS (reference config: 1000)
constexpr int PRED_WORDS = PRED_MAX_PRE / 64;
constexpr int PRED_THREADS = 1024;

struct PredictArgs {
  const float* out; int row_stride;            // head output, channels last: [B, H*W, row_stride]
  int c_reg, c_height, c_dim, c_rot, c_hm, num_cls;
  int c_vel[16]; int T;                         // velocity channel pair of every emitted timestep
  int B, H, W;
  float score_thr; float range[6];              // post_center_limit_range
  float osf, vx, vy, x0, y0;                    // out_size_factor, voxel size, pc_range origin
  float iou_thr; int pre_max, post_max;
  unsigned long long* keys;                     // [B, H*W]
  float* cand_box; unsigned long long* cand_key; int* cand_n;   // [B,1024,8] NMS boxes (+half diagonal), [B,1024], [B]
  unsigned long long* mask;                     // [B, 1024, 16] IoU bit matrix (upper triangle)
  float* boxes; float* scores; int* labels; int* cells; int* count;   // [B,T,post_max,9] [B,T,post_max] x2, [B,post_max], [B]
};

__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.f / (1.f + expf(-x)); }

// box_preds row of the reference (center_head.py:640-665): x, y, height, exp(dim) x3, atan2(rot_s, rot_c); vel added later
__device__ __forceinline__ void decode_cell(const PredictArgs& a, const float* row, int cell, float* b7) {
  const int y = cell / a.W, x = cell - y * a.W;
  const float xs = __fadd_rn((float)x, row[a.c_reg]);
  const float ys = __fadd_rn((float)y, row[a.c_reg + 1]);
  b7[0] = __fadd_rn(__fmul_rn(__fmul_rn(xs, a.osf), a.vx), a.x0);
  b7[1] = __fadd_rn(__fmul_rn(__fmul_rn(ys, a.osf), a.vy), a.y0);
  b7[2] = row[a.c_height];
  b7[3] = expf(row[a.c_dim]);
  b7[4] = expf(row[a.c_dim + 1]);
  b7[5] = expf(row[a.c_dim + 2]);
  b7[6] = atan2f(row[a.c_rot], row[a.c_rot + 1]);
}

__global__ void __launch_bounds__(256)
predict_keys_kernel(const PredictArgs a) {
  const int HW = a.H * a.W;
  const long long total = (long long)a.B * HW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int cell = (int)(e % HW);
    const float* row = a.out + (size_t)e * a.row_stride;
    float best = sigmoidf_ref(row[a.c_hm]);
    for (int c = 1; c < a.num_cls; ++c) best = fmaxf(best, sigmoidf_ref(row[a.c_hm + c]));   // torch.max(dim=-1)
    float b7[7];
    decode_cell(a, row, cell, b7);
    bool ok = best > a.score_thr;
#pragma unroll
    for (int j = 0; j < 3; ++j) ok = ok && b7[j] >= a.range[j] && b7[j] <= a.range[3 + j];
    // positive floats order like their bit patterns; ties resolved towards the lower cell index (stable sort)
    a.keys[e] = ok ? (((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(0xffffffffu - (unsigned)cell)) : 0ULL;
  }
}

// ---- rotated BEV IoU: the arithmetic of iou3d_nms_kernel.cu:20-228 restated (fp32, same operation order) -----------
struct P2 { float x, y; };
__device__ __forceinline__ float cross3(const P2& p1, const P2& p2, const P2& p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ float cross2(const P2& a, const P2& b) { return a.x * b.y - a.y * b.x; }

__device__ __forceinline__ bool segment_hit(const P2& p1, const P2& p0, const P2& q1, const P2& q0, P2& ans) {
  const bool boxes_touch = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                           fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
  if (!boxes_touch) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0.f && s3 * s4 > 0.f)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > 1e-8f) {
    ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    ans.x = (b0 * c1 - b1 * c0) / D;
    ans.y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

__device__ __forceinline__ bool inside_box(const float* box, const P2& p) {
  const float ac = cosf(-box[6]), as = sinf(-box[6]);
  const float rx = (p.x - box[0]) * ac + (p.y - box[1]) * (-as);
  const float ry = (p.x - box[0]) * as + (p.y - box[1]) * ac;
  return fabsf(rx) < box[3] / 2 + 1e-2f && fabsf(ry) < box[4] / 2 + 1e-2f;
}

__device__ void oriented_corners(const float* box, P2* c) {   // 5 entries, last = first
  const float hx = box[3] / 2, hy = box[4] / 2;
  const float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
  const float ac = cosf(box[6]), as = sinf(box[6]);
  const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    c[k].x = (px[k] - box[0]) * ac + (py[k] - box[1]) * (-as) + box[0];
    c[k].y = (px[k] - box[0]) * as + (py[k] - box[1]) * ac + box[1];
  }
  c[4] = c[0];
}

__device__ float overlap_area(const float* A, const float* Bx) {
  P2 ca[5], cb[5];
  oriented_corners(A, ca);
  oriented_corners(Bx, cb);
  P2 pts[16];
  P2 ctr = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (segment_hit(ca[i + 1], ca[i], cb[j + 1], cb[j], pts[cnt])) {
        ctr.x = ctr.x + pts[cnt].x; ctr.y = ctr.y + pts[cnt].y;
        ++cnt;
      }
  for (int k = 0; k < 4; ++k) {
    if (inside_box(A, cb[k])) { ctr.x = ctr.x + cb[k].x; ctr.y = ctr.y + cb[k].y; pts[cnt++] = cb[k]; }
    if (inside_box(Bx, ca[k])) { ctr.x = ctr.x + ca[k].x; ctr.y = ctr.y + ca[k].y; pts[cnt++] = ca[k]; }
  }
  ctr.x /= cnt; ctr.y /= cnt;
  // bubble sort by polar angle around the centroid (same comparison and order of swaps as the reference)
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
        const P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k) {
    const P2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
    area += cross2(u, v);
  }
  return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_bev_ref(const float* A, const float* Bx) {
  const float sa = A[3] * A[4], sb = Bx[3] * Bx[4];
  const float so = overlap_area(A, Bx);
  return so / fmaxf(sa + sb - so, 1e-8f);
}

// standalone pairwise IoU (API parity with boxes_iou_bev_gpu, iou3d_nms.cpp:61-88; also used by the tests)
__global__ void __launch_bounds__(256)
iou_bev_pairs_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, float* __restrict__ out) {
  const long long total = (long long)na * nb;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / nb), j = (int)(e - (long long)i * nb);
    out[e] = iou_bev_ref(a + (size_t)i * 7, b + (size_t)j * 7);
  }
}

// ---- stage A: one CTA per sample -- exact top-k selection, sort, decode of the NMS boxes ---------------------------
struct SelSmem {
  unsigned long long key[PRED_MAX_PRE];
  int hist[256];
  unsigned long long prefix; int remaining; int n_valid; int n_sel;
};

__global__ void __launch_bounds__(PRED_THREADS)
predict_select_kernel(const PredictArgs a) {
  __shared__ SelSmem s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int HW = a.H * a.W;
  const unsigned long long* keys = a.keys + (size_t)b * HW;
  if (tid == 0) { s.n_valid = 0; s.n_sel = 0; s.prefix = 0ULL; }
  __syncthreads();
  int local = 0;
  for (int i = tid; i < HW; i += PRED_THREADS) local += keys[i] != 0ULL;
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((tid & 31) == 0 && local) atomicAdd(&s.n_valid, local);
  __syncthreads();
  const int k = min(min(a.pre_max, PRED_MAX_PRE), s.n_valid);
  if (tid == 0) a.cand_n[b] = k;
  if (k == 0) return;
  // exact k-th largest key: 8 passes of an 8-bit radix select (keys are unique)
  if (tid == 0) s.remaining = k;
  for (int pass = 7; pass >= 0; --pass) {
    if (tid < 256) s.hist[tid] = 0;
    __syncthreads();
    const unsigned long long prefix = s.prefix;
    const int hs = 8 * (pass + 1);
    for (int i = tid; i < HW; i += PRED_THREADS) {
      const unsigned long long v = keys[i];
      if (v != 0ULL && (pass == 7 || (v >> hs) == prefix)) atomicAdd(&s.hist[(int)((v >> (8 * pass)) & 255ULL)], 1);
    }
    __syncthreads();
    if (tid == 0) {
      int rem = s.remaining, d = 255;
      while (d > 0 && s.hist[d] < rem) { rem -= s.hist[d]; --d; }
      s.remaining = rem;
      s.prefix = (prefix << 8) | (unsigned long long)d;
    }
    __syncthreads();
  }
  const unsigned long long kth = s.prefix;
  // gather the k winners, pad to a power of two, bitonic sort (descending)
  for (int i = tid; i < PRED_MAX_PRE; i += PRED_THREADS) s.key[i] = 0ULL;
  __syncthreads();
  for (int i = tid; i < HW; i += PRED_THREADS) {
    const unsigned long long v = keys[i];
    if (v >= kth && v != 0ULL) s.key[atomicAdd(&s.n_sel, 1)] = v;
  }
  __syncthreads();
  for (int size = 2; size <= PRED_MAX_PRE; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < PRED_MAX_PRE / 2; i += PRED_THREADS) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long x = s.key[lo], y = s.key[hi];
        if ((x < y) == desc) { s.key[lo] = y; s.key[hi] = x; }
      }
      __syncthreads();
    }
  // boxes of the sorted candidates in the NMS convention (box_torch_ops.py:256-257)
  const float* out_b = a.out + (size_t)b * HW * a.row_stride;
  for (int i = tid; i < k; i += PRED_THREADS) {
    const unsigned long long key = s.key[i];
    const int cell = (int)(0xffffffffu - (unsigned)(key & 0xffffffffULL));
    float b7[7];
    decode_cell(a, out_b + (size_t)cell * a.row_stride, cell, b7);
    float* cb = a.cand_box + ((size_t)b * PRED_MAX_PRE + i) * 8;
    cb[0] = b7[0]; cb[1] = b7[1]; cb[2] = b7[2];
    cb[3] = b7[4]; cb[4] = b7[3]; cb[5] = b7[5];
    cb[6] = __fsub_rn(-b7[6], 1.5707963267948966f);
    cb[7] = 0.5f * sqrtf(b7[3] * b7[3] + b7[4] * b7[4]);          // half diagonal: cheap far-apart test in stage B
    a.cand_key[(size_t)b * PRED_MAX_PRE + i] = key;
  }
}

// ---- stage B: the IoU bit matrix, one 64 x 64 tile per CTA (same tiling as iou3d_nms_kernel.cu:264-311) -------------
// Pairs whose circumscribed circles (plus the kernel's 1e-2 containment margin) do not touch have an empty
// intersection in the reference arithmetic too (no edge crossing, no contained corner -> area 0), so they are decided
// without running the polygon clip.
__global__ void __launch_bounds__(64)
predict_mask_kernel(const PredictArgs a) {
  __shared__ float cbox[64][8];
  const int b = blockIdx.z, rb = blockIdx.y, cw = blockIdx.x, tid = threadIdx.x;
  const int k = a.cand_n[b];
  if (rb * 64 >= k || cw * 64 >= k || cw < rb) return;            // outside / strictly lower triangle
  const float* boxes = a.cand_box + (size_t)b * PRED_MAX_PRE * 8;
  const int j0 = cw * 64;
  if (j0 + tid < k) {
#pragma unroll
    for (int c = 0; c < 8; ++c) cbox[tid][c] = boxes[(size_t)(j0 + tid) * 8 + c];
  }
  __syncthreads();
  const int i = rb * 64 + tid;
  if (i >= k) return;
  float mine[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) mine[c] = boxes[(size_t)i * 8 + c];
  unsigned long long bits = 0ULL;
  const int jbeg = max(j0, i + 1), jend = min(j0 + 64, k);
  for (int j = jbeg; j < jend; ++j) {
    const float* o = cbox[j - j0];
    const float dx = mine[0] - o[0], dy = mine[1] - o[1], reach = mine[7] + o[7] + 0.05f;
    if (dx * dx + dy * dy > reach * reach) continue;
    if (iou_bev_ref(mine, o) > a.iou_thr) bits |= 1ULL << (j & 63);
  }
  a.mask[((size_t)b * PRED_MAX_PRE + i) * PRED_WORDS + cw] = bits;
}

// ---- stage C: greedy sweep in score order (iou3d_nms.cpp:116-131) by one warp, then emit -------------------------------
__global__ void __launch_bounds__(256)
predict_sweep_emit_kernel(const PredictArgs a) {
  __shared__ int keep[PRED_MAX_PRE];
  __shared__ int n_keep;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int HW = a.H * a.W;
  const int k = a.cand_n[b];
  if (k == 0) {
    if (tid == 0) a.count[b] = 0;
    return;
  }
  const int words = (k + 63) / 64;
  const unsigned long long* mask = a.mask + (size_t)b * PRED_MAX_PRE * PRED_WORDS;
  if (tid < 32) {
    unsigned long long remv = 0ULL;      // lane w holds word w of the suppressed set
    int nk = 0;
    for (int i = 0; i < k && nk < a.post_max; ++i) {
      const unsigned long long word = __shfl_sync(0xffffffffu, remv, i >> 6);
      if (!((word >> (i & 63)) & 1ULL)) {
        if (tid == 0) keep[nk] = i;
        ++nk;
        // words left of the diagonal tile were never written (lower triangle): only read w >= i / 64
        if (tid < words && tid >= (i >> 6)) remv |= mask[(size_t)i * PRED_WORDS + tid];
      }
    }
    if (tid == 0) { n_keep = nk; a.count[b] = nk; }
  }
  __syncthreads();
  // emit: box3d_lidar (x, y, z, w, l, h, vx, vy, rot), score, label for every forecast timestep
  const int nk = n_keep;
  const float* out_b = a.out + (size_t)b * HW * a.row_stride;
  for (int e = tid; e < nk * a.T; e += blockDim.x) {
    const int t = e / nk, q = e - t * nk;
    const unsigned long long key = a.cand_key[(size_t)b * PRED_MAX_PRE + keep[q]];
    const int cell = (int)(0xffffffffu - (unsigned)(key & 0xffffffffULL));
    const float* row = out_b + (size_t)cell * a.row_stride;
    float b7[7];
    decode_cell(a, row, cell, b7);
    float* ob = a.boxes + (((size_t)b * a.T + t) * a.post_max + q) * 9;
    ob[0] = b7[0]; ob[1] = b7[1]; ob[2] = b7[2]; ob[3] = b7[3]; ob[4] = b7[4]; ob[5] = b7[5];
    ob[6] = row[a.c_vel[t]]; ob[7] = row[a.c_vel[t] + 1]; ob[8] = b7[6];
    int lab = 0;
    float best = row[a.c_hm];
    for (int c = 1; c < a.num_cls; ++c)
      if (row[a.c_hm + c] > best) { best = row[a.c_hm + c]; lab = c; }     // first maximum, as torch.max
    a.scores[((size_t)b * a.T + t) * a.post_max + q] = __uint_as_float((unsigned)(key >> 32));
    a.labels[((size_t)b * a.T + t) * a.post_max + q] = lab;
    if (t == 0) a.cells[(size_t)b * a.post_max + q] = cell;
  }
}

}  // namespace fd

extern "C" {

static size_t pred_align(size_t x) { return (x + 255) & ~(size_t)255; }

size_t fd_center_predict_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  return pred_align(sizeof(unsigned long long) * (size_t)B * H * W) +
         pred_align(sizeof(float) * (size_t)B * fd::PRED_MAX_PRE * 8) +
         pred_align(sizeof(unsigned long long) * (size_t)B * fd::PRED_MAX_PRE) + pred_align(sizeof(int) * (size_t)B) +
         pred_align(sizeof(unsigned long long) * (size_t)B * fd::PRED_MAX_PRE * fd::PRED_WORDS);
}

int fd_center_predict(const float* d_out, int row_stride, int c_reg, int c_height, int c_dim, int c_rot,
                      const int32_t* c_vel, int T, int c_hm, int num_cls, int B, int H, int W, float score_threshold,
                      const float* post_center_range6, float out_size_factor, float voxel_x, float voxel_y,
                      float pc_x0, float pc_y0, float nms_iou_threshold, int pre_max, int post_max, float* d_boxes,
                      float* d_scores, int32_t* d_labels, int32_t* d_cells, int32_t* d_count, void* d_workspace,
                      void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_out && c_vel && post_center_range6 && d_boxes && d_scores && d_labels && d_cells && d_count && d_workspace,
             "fd_center_predict: null argument");
  FD_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && T >= 1 && T <= 16 && num_cls >= 1 && row_stride >= 1,
             "fd_center_predict: bad shape (T must be <= 16)");
  FD_REQUIRE(pre_max >= 1 && pre_max <= PRED_MAX_PRE && post_max >= 1 && post_max <= pre_max,
             "fd_center_predict: need 1 <= post_max <= pre_max <= %d", PRED_MAX_PRE);
  FD_REQUIRE((long long)H * W < 0x7fffffffLL, "fd_center_predict: grid too large");
  FD_REQUIRE(nms_iou_threshold >= 0.f, "fd_center_predict: negative IoU threshold");
  PredictArgs a{};
  a.out = d_out; a.row_stride = row_stride;
  a.c_reg = c_reg; a.c_height = c_height; a.c_dim = c_dim; a.c_rot = c_rot; a.c_hm = c_hm; a.num_cls = num_cls;
  for (int t = 0; t < T; ++t) a.c_vel[t] = c_vel[t];
  a.T = T; a.B = B; a.H = H; a.W = W;
  a.score_thr = score_threshold;
  for (int j = 0; j < 6; ++j) a.range[j] = post_center_range6[j];
  a.osf = out_size_factor; a.vx = voxel_x; a.vy = voxel_y; a.x0 = pc_x0; a.y0 = pc_y0;
  a.iou_thr = nms_iou_threshold; a.pre_max = pre_max; a.post_max = post_max;
  char* ws = (char*)d_workspace;
  a.keys = (unsigned long long*)ws;      ws += pred_align(sizeof(unsigned long long) * (size_t)B * H * W);
  a.cand_box = (float*)ws;               ws += pred_align(sizeof(float) * (size_t)B * PRED_MAX_PRE * 8);
  a.cand_key = (unsigned long long*)ws;  ws += pred_align(sizeof(unsigned long long) * (size_t)B * PRED_MAX_PRE);
  a.cand_n = (int*)ws;                   ws += pred_align(sizeof(int) * (size_t)B);
  a.mask = (unsigned long long*)ws;
  a.boxes = d_boxes; a.scores = d_scores; a.labels = d_labels; a.cells = d_cells; a.count = d_count;
  predict_keys_kernel<<<persistent_grid(ceil_div((int64_t)B * H * W, 256), 8), 256, 0, stream>>>(a);
  FD_LAUNCHED();
  predict_select_kernel<<<B, PRED_THREADS, 0, stream>>>(a);
  FD_LAUNCHED();
  const int tiles = ceil_div(pre_max, 64);
  predict_mask_kernel<<<dim3(tiles, tiles, B), 64, 0, stream>>>(a);
  FD_LAUNCHED();
  predict_sweep_emit_kernel<<<B, 256, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

int fd_boxes_iou_bev(const float* d_boxes_a, int na, const float* d_boxes_b, int nb, float* d_iou, void* stream) {
  using namespace fd;
  FD_REQUIRE(na >= 0 && nb >= 0 && (d_iou || (long long)na * nb == 0), "fd_boxes_iou_bev: bad argument");
  if (na == 0 || nb == 0) return 0;
  FD_REQUIRE(d_boxes_a && d_boxes_b, "fd_boxes_iou_bev: null boxes");
  iou_bev_pairs_kernel<<<persistent_grid(ceil_div((int64_t)na * nb, 256), 8), 256, 0, (cudaStream_t)stream>>>(
      d_boxes_a, na, d_boxes_b, nb, d_iou);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
