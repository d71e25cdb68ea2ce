// CenterHead.loss forward (standard branch) for sm_100a: CenterNet focal loss on the heat map + masked-L1
// regression loss on the boxes gathered at the object centres, for every forecast timestep.
//
// Replaces  det3d/models/bbox_heads/center_head.py:396-539 (standard branch), :392-394 (_sigmoid, in place),
//           det3d/models/losses/centernet_loss.py:18-25 (RegLoss), :75-95 (FastFocalLoss),
//           det3d/core/utils/center_utils.py:66-80 (_transpose_and_gather_feat).
// Two launches: a grid-wide pass over the heat map (sigmoid+clamp written back in place, per-block partial sums of
// the negative term) and one block that finishes every reduction in a fixed order (deterministic, no float atomics).
#include "common.cuh"

namespace fd {

constexpr int kLossThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < kLossThreads / 32; ++i) t += sh[i];
  __syncthreads();
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

// hm logits -> p = clamp(sigmoid(x), 1e-4, 1-1e-4) in place; partial[b] = sum log(1-p) * p^2 * (1-gt)^4
__global__ void __launch_bounds__(kLossThreads)
focal_neg_kernel(float* __restrict__ hm, long long sb, long long sc, long long ssp, const float* __restrict__ gt,
                 int B, int C, int HW, double* __restrict__ partial) {
  __shared__ double sh[kLossThreads / 32];
  const long long total = (long long)B * C * HW;
  double acc = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e % HW);
    const long long bc = e / HW;
    const int c = (int)(bc % C), b = (int)(bc / C);
    float* px = hm + b * sb + c * sc + s * ssp;
    float p = 1.f / (1.f + expf(-*px));
    p = fminf(fmaxf(p, 1e-4f), 1.f - 1e-4f);
    *px = p;
    const float g = 1.f - gt[e];
    const float g2 = g * g;
    acc += (double)(logf(1.f - p) * p * p * (g2 * g2));
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

struct LossFinalArgs {
  const float* hm; long long hm_sb, hm_sc, hm_ssp;     // (already sigmoid-clamped) heat map
  const double* partial; int n_partial;
  const long long* ind; const unsigned char* mask; const long long* cat;   // [B, M] targets of timestep 0
  const unsigned char* const* mask_t;                   // [T] masks of every timestep (num_positive)
  int B, M, T, NC;                                       // NC regression channels (10)
  const float* const* pred_ptr; const long long* pred_sb; const long long* pred_ssp;   // [T*NC] channel planes
  const float* const* tgt_ptr; int tgt_dim; const int* tgt_sel;                           // [T] targets [B,M,tgt_dim], [NC] columns
  const float* code_w; const float* code_w_forecast;     // [NC]
  float weight;
  float* out;   // [0] loss, [1] hm_loss, [2] num_positive, [3..3+T) loc_loss, then [T*NC] loc_loss_elem
};

__global__ void __launch_bounds__(kLossThreads)
loss_final_kernel(const LossFinalArgs a) {
  __shared__ double sh[kLossThreads / 32];
  const int BM = a.B * a.M;
  // negative term
  double v = 0.0;
  for (int i = threadIdx.x; i < a.n_partial; i += blockDim.x) v += a.partial[i];
  const double neg = block_sum(v, sh);
  // positive term and number of positives (timestep-0 mask)
  double pos = 0.0, npos = 0.0;
  for (int i = threadIdx.x; i < BM; i += blockDim.x) {
    const int b = i / a.M;
    const float m = (float)a.mask[i];
    const float p = a.hm[b * a.hm_sb + a.cat[i] * a.hm_sc + a.ind[i] * a.hm_ssp];
    const float om = 1.f - p;
    pos += (double)(logf(p) * om * om * m);
    npos += (double)m;
  }
  pos = block_sum(pos, sh);
  npos = block_sum(npos, sh);
  const double hm_loss = npos == 0.0 ? -neg : -(pos + neg) / npos;
  // num_positive of the reference: sum over timesteps of mask sums (center_head.py:527)
  double np_all = 0.0;
  for (int tt = 0; tt < a.T; ++tt)
    for (int i = threadIdx.x; i < BM; i += blockDim.x) np_all += (double)a.mask_t[tt][i];
  np_all = block_sum(np_all, sh);
  // regression: loss[t][c] = sum_{b,m} |pred*mask - tgt*mask| / (sum(mask) + 1e-4)
  double loc_total = 0.0;
  const float denom = (float)npos + 1e-4f;
  for (int tt = 0; tt < a.T; ++tt) {
    double loc_t = 0.0;
    for (int c = 0; c < a.NC; ++c) {
      const float* pp = a.pred_ptr[tt * a.NC + c];
      const long long sb = a.pred_sb[tt * a.NC + c], ssp = a.pred_ssp[tt * a.NC + c];
      const float* tg = a.tgt_ptr[tt];
      const int col = a.tgt_sel[c];
      double s = 0.0;
      for (int i = threadIdx.x; i < BM; i += blockDim.x) {
        const int b = i / a.M;
        const float m = (float)a.mask[i];
        const float pr = pp[b * sb + a.ind[i] * ssp] * m;
        const float tv = tg[(size_t)i * a.tgt_dim + col] * m;
        s += (double)(fabsf(pr - tv) / denom);
      }
      s = block_sum(s, sh);
      if (threadIdx.x == 0) a.out[3 + a.T + tt * a.NC + c] = (float)s;
      loc_t += (double)((float)s * (tt == 0 ? a.code_w[c] : a.code_w_forecast[c]));
    }
    if (threadIdx.x == 0) a.out[3 + tt] = (float)loc_t;
    loc_total += (double)(float)loc_t;
  }
  if (threadIdx.x == 0) {
    a.out[0] = (float)hm_loss + a.weight * (float)loc_total;
    a.out[1] = (float)hm_loss;
    a.out[2] = (float)np_all;
  }
}

}  // namespace fd

extern "C" {

size_t fd_center_loss_workspace_bytes(void) { return sizeof(double) * (size_t)fd::kNumSMs * 4; }

int fd_center_head_loss(float* d_hm, int64_t hm_sb, int64_t hm_sc, int64_t hm_ssp, const float* d_hm_target, int B,
                        int C, int H, int W, const int64_t* d_ind, const uint8_t* d_mask, const int64_t* d_cat,
                        const uint8_t* const* d_mask_t, int M, int T, int NC, const float* const* d_pred_ptr,
                        const int64_t* d_pred_sb, const int64_t* d_pred_ssp, const float* const* d_tgt_ptr,
                        int tgt_dim, const int32_t* d_tgt_sel, const float* d_code_w, const float* d_code_w_forecast,
                        float weight, float* d_out, void* d_workspace, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_hm && d_hm_target && d_ind && d_mask && d_cat && d_mask_t && d_pred_ptr && d_pred_sb && d_pred_ssp &&
                 d_tgt_ptr && d_tgt_sel && d_code_w && d_code_w_forecast && d_out && d_workspace,
             "fd_center_head_loss: null argument");
  FD_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1 && M >= 1 && T >= 1 && NC >= 1 && tgt_dim >= 1,
             "fd_center_head_loss: bad shape");
  const int grid = kNumSMs * 4;
  focal_neg_kernel<<<grid, kLossThreads, 0, stream>>>(d_hm, hm_sb, hm_sc, hm_ssp, d_hm_target, B, C, H * W,
                                                      (double*)d_workspace);
  FD_LAUNCHED();
  LossFinalArgs a{};
  a.hm = d_hm; a.hm_sb = hm_sb; a.hm_sc = hm_sc; a.hm_ssp = hm_ssp;
  a.partial = (const double*)d_workspace; a.n_partial = grid;
  a.ind = (const long long*)d_ind; a.mask = d_mask; a.cat = (const long long*)d_cat; a.mask_t = d_mask_t;
  a.B = B; a.M = M; a.T = T; a.NC = NC;
  a.pred_ptr = d_pred_ptr; a.pred_sb = (const long long*)d_pred_sb; a.pred_ssp = (const long long*)d_pred_ssp;
  a.tgt_ptr = d_tgt_ptr; a.tgt_dim = tgt_dim; a.tgt_sel = d_tgt_sel;
  a.code_w = d_code_w; a.code_w_forecast = d_code_w_forecast; a.weight = weight; a.out = d_out;
  loss_final_kernel<<<1, kLossThreads, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
