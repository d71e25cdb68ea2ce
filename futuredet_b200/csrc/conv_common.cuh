// Row gather / output mapping shared by the implicit-GEMM convolution kernels.
#pragma once
#include "common.cuh"

namespace fd {

enum { OUTMAP_UPSAMPLE = 100 };  // internal: one phase of ConvTranspose2d(k == stride)

struct ConvArgs {
  const float* in; int in_stride; int cin;
  const float* w; int cout; int K;
  const void* wp;            // packed bf16 hi/lo weights (tensor-core arm)
  const float* scale; const float* shift;
  const float* residual; int res_stride;
  int relu;
  float* out; int out_stride;
  const int32_t* d_n; int n_cap;
  int mode;
  const int32_t* nbr; int nbr_stride;
  int Hin, Win, Hout, Wout, kh, kw, sh, sw, ph, pw;
  int out_map;
  const int4* out_coords; int bevD, bevH, bevW;
  int up_s, up_dy, up_dx;  // OUTMAP_UPSAMPLE
};

// input row feeding output row `o` through kernel offset `k`, or -1
__device__ __forceinline__ int gather_row(const ConvArgs& a, int o, int k) {
  if (a.mode == FD_GATHER_TABLE) return a.nbr[(size_t)k * a.nbr_stride + o];
  // dense 2-D: o = (b, oy, ox)
  int hw = a.Hout * a.Wout;
  int b = o / hw;
  int r = o - b * hw;
  int oy = r / a.Wout, ox = r - oy * a.Wout;
  int ky = k / a.kw, kx = k - ky * a.kw;
  int iy = oy * a.sh - a.ph + ky, ix = ox * a.sw - a.pw + kx;
  if (iy < 0 || iy >= a.Hin || ix < 0 || ix >= a.Win) return -1;
  return (b * a.Hin + iy) * a.Win + ix;
}

// Where element (row o, channel c) of the result lives.  Returns base pointer and channel stride.
struct OutRow { float* p; int cstride; };
__device__ __forceinline__ OutRow map_out_row(const ConvArgs& a, int o) {
  if (a.out_map == FD_OUTMAP_IDENTITY) return {a.out + (size_t)o * a.out_stride, 1};
  if (a.out_map == FD_OUTMAP_BEV) {
    // SparseConvTensor.dense().view(N, C*D, H, W) (scn.py:165-168): channel = c*D + z, stored channels-last
    int4 c = a.out_coords[o];
    size_t pix = ((size_t)c.x * a.bevH + c.z) * a.bevW + c.w;
    return {a.out + pix * a.out_stride + c.y, a.bevD};
  }
  // OUTMAP_UPSAMPLE: o = (b, y, x) on the input grid -> (b, y*s+dy, x*s+dx) on the output grid
  int hw = a.Hin * a.Win;
  int b = o / hw;
  int r = o - b * hw;
  int y = r / a.Win, x = r - y * a.Win;
  size_t pix = ((size_t)b * a.Hin * a.up_s + (size_t)y * a.up_s + a.up_dy) * (a.Win * a.up_s) + (size_t)x * a.up_s + a.up_dx;
  return {a.out + pix * a.out_stride, 1};
}

int conv_forward_simt(const ConvArgs& a, cudaStream_t stream);
int conv_forward_tc(const ConvArgs& a, int precision, cudaStream_t stream);

}  // namespace fd
