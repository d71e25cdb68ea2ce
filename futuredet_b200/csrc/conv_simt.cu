// fp32 CUDA-core gather -> implicit-GEMM convolution with fused epilogue.
//
// Exact-fp32 arm of fd_conv_forward (FD_PREC_FP32): the arithmetic of spconv 1.x
// `indice_conv` (gather -> SGEMM -> scatter-add per kernel offset, fp32) restated
// output-stationary, so it is deterministic and needs no atomics.  It is the on-device
// ground truth the tensor-core arm is validated against at full problem size.
//
// Tile: BM=128 output rows x BN output channels per CTA, BK=16 input channels per step,
// 256 threads, each owning a TM x 4 register block.  Kernel offsets whose gather list is
// empty for the whole tile are skipped (the common case on sparse 3-D data).
#include "conv_common.cuh"

namespace fd {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int TN = 4;
constexpr int kThreads = 256;

template <int BN>
__global__ void __launch_bounds__(kThreads)
conv_simt_kernel(const ConvArgs a) {
  constexpr int TCOLS = BN / TN;            // thread columns
  constexpr int TROWS = kThreads / TCOLS;   // thread rows
  constexpr int TM = BM / TROWS;            // rows per thread
  constexpr int APAD = 4;
  __shared__ __align__(16) float As[BK][BM + APAD];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ int s_idx[BM];

  const int tid = threadIdx.x;
  const int tx = tid % TCOLS, ty = tid / TCOLS;
  const int n = a.d_n ? min(*a.d_n, a.n_cap) : a.n_cap;
  const int n_tiles_m = (n + BM - 1) / BM;
  const int n_tiles_n = (a.cout + BN - 1) / BN;
  const bool split_in = a.in_fmt == FD_FMT_SPLIT_BF16;
  const bool vec_in = !split_in && (a.cin % 4 == 0) && (a.in_stride % 4 == 0) && (((uintptr_t)a.in & 15) == 0);
  const bool vec_w = (a.cout % 4 == 0) && (((uintptr_t)a.w & 15) == 0);

  for (int tile = blockIdx.x; tile < n_tiles_m * n_tiles_n; tile += gridDim.x) {
    const int tm = tile / n_tiles_n, tn = tile - tm * n_tiles_n;
    const int row0 = tm * BM, col0 = tn * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k = 0; k < a.K; ++k) {
      __syncthreads();  // previous users of s_idx / As / Bs are done
      int any = 0;
      if (tid < BM) {
        int o = row0 + tid;
        int r = o < n ? gather_row(a, o, k) : -1;
        s_idx[tid] = r;
        any = r >= 0;
      }
      if (!__syncthreads_or(any)) continue;
      const float* wk = a.w + (size_t)k * a.cin * a.cout;
      for (int c0 = 0; c0 < a.cin; c0 += BK) {
        // ---- A tile: BM rows x BK channels, gathered -------------------------------
        if (vec_in) {
#pragma unroll
          for (int it = 0; it < (BM * BK / 4) / kThreads; ++it) {
            int e = it * kThreads + tid;
            int r = e / (BK / 4), q = e % (BK / 4);
            int c = c0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int src = s_idx[r];
            if (src >= 0 && c < a.cin) v = __ldg(reinterpret_cast<const float4*>(a.in + (size_t)src * a.in_stride + c));
            As[q * 4 + 0][r] = v.x; As[q * 4 + 1][r] = v.y; As[q * 4 + 2][r] = v.z; As[q * 4 + 3][r] = v.w;
          }
        } else {
#pragma unroll
          for (int it = 0; it < (BM * BK) / kThreads; ++it) {
            int e = it * kThreads + tid;
            int r = e / BK, q = e % BK;
            int c = c0 + q;
            int src = s_idx[r];
            float v = 0.f;
            if (src >= 0 && c < a.cin) {
              const float* row = a.in + (size_t)src * a.in_stride;
              v = split_in ? split_load(row, c, a.in_ctot) : __ldg(row + c);
            }
            As[q][r] = v;
          }
        }
        // ---- B tile: BK channels x BN outputs ---------------------------------------
        if (vec_w) {
          for (int e = tid; e < BK * BN / 4; e += kThreads) {
            int r = e / (BN / 4), q = e % (BN / 4);
            int ci = c0 + r, co = col0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ci < a.cin && co < a.cout) v = __ldg(reinterpret_cast<const float4*>(wk + (size_t)ci * a.cout + co));
            *reinterpret_cast<float4*>(&Bs[r][q * 4]) = v;
          }
        } else {
          for (int e = tid; e < BK * BN; e += kThreads) {
            int r = e / BN, q = e % BN;
            int ci = c0 + r, co = col0 + q;
            Bs[r][q] = (ci < a.cin && co < a.cout) ? __ldg(wk + (size_t)ci * a.cout + co) : 0.f;
          }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float av[TM], bv[TN];
#pragma unroll
          for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
          const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
          bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
    }

    // ---- epilogue: BN(eval)/bias -> residual -> ReLU -> mapped store -------------------
    float sc[TN], sf[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int c = col0 + tx * TN + j;
      sc[j] = (a.scale && c < a.cout) ? a.scale[c] : 1.f;
      sf[j] = (a.shift && c < a.cout) ? a.shift[c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int o = row0 + ty * TM + i;
      if (o >= n) continue;
      if (a.row_perm) o = __ldg(a.row_perm + o);          // sorted table: tile position -> output row
      OutRow orow = map_out_row(a, o);
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int c = col0 + tx * TN + j;
        if (c >= a.cout) continue;
        float v = fmaf(acc[i][j], sc[j], sf[j]);
        if (a.residual) v += load_residual(a, o, c);
        if (a.relu) v = fmaxf(v, 0.f);
        store_out(a, orow, c, v);
      }
    }
  }
}

template <int BN>
static int launch_simt(const ConvArgs& a, cudaStream_t stream) {
  int tiles = ceil_div(a.n_cap, BM) * ceil_div(a.cout, BN);
  int grid = persistent_grid(tiles, 3);
  conv_simt_kernel<BN><<<grid, kThreads, 0, stream>>>(a);
  FD_LAUNCHED();
  return 0;
}

int conv_forward_simt(const ConvArgs& a, cudaStream_t stream) {
  if (a.n_cap <= 0) return 0;
  if (a.cout >= 48) return launch_simt<64>(a, stream);
  if (a.cout >= 24) return launch_simt<32>(a, stream);
  return launch_simt<16>(a, stream);
}

}  // namespace fd
