// Row gather / output mapping shared by the implicit-GEMM convolution kernels.
#pragma once
#include "common.cuh"

namespace fd {

enum { OUTMAP_UPSAMPLE = 100 };  // internal: one phase of ConvTranspose2d(k == stride)

struct ConvArgs {
  const float* in; int in_stride; int cin;
  int in_fmt, in_ctot, out_fmt, out_ctot, res_fmt, res_ctot;
  const float* w; int cout; int K;
  const void* wp;            // packed bf16 hi/lo weights (tensor-core arm)
  const float* scale; const float* shift;
  const float* residual; int res_stride;
  int relu;
  float* out; int out_stride;
  const int32_t* d_n; int n_cap;
  int mode;
  const int32_t* nbr; int nbr_stride;
  const uint32_t* tile_mask;
  int Hin, Win, Hout, Wout, kh, kw, sh, sw, ph, pw;
  int out_map;
  const int4* out_coords; int bevD, bevH, bevW;
  int up_s, up_dy, up_dx;  // OUTMAP_UPSAMPLE
  int n_in_cap;            // rows of `in` (weight gradient only; 0 = unknown)
  const void* in_split;    // weight gradient only: dense FD_FMT_SPLIT_BF16 copies of `in` / dL/dy ([rows][C hi | C lo]) when
  const void* out_split;   // the caller already has them (NULL: the tcgen05 arm splits into its workspace)
  const int32_t* row_perm; // FD_GATHER_TABLE: nbr / tile_mask are sorted tables, tile position j is output row row_perm[j]
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
  if (a.mode == FD_GATHER_CONV2D_DGRAD) {
    // data gradient of Conv2d: rows are pixels of the conv's INPUT grid (Hout x Wout here), the gathered tensor is
    // dL/dy on the conv's OUTPUT grid (Hin x Win here); (y,x) receives dy[(y+p-ky)/s, (x+p-kx)/s] @ W[k]^T
    int ty = oy + a.ph - ky, tx = ox + a.pw - kx;
    if (ty < 0 || tx < 0 || ty % a.sh != 0 || tx % a.sw != 0) return -1;
    int gy = ty / a.sh, gx = tx / a.sw;
    if (gy >= a.Hin || gx >= a.Win) return -1;
    return (b * a.Hin + gy) * a.Win + gx;
  }
  int iy = oy * a.sh - a.ph + ky, ix = ox * a.sw - a.pw + kx;
  if (iy < 0 || iy >= a.Hin || ix < 0 || ix >= a.Win) return -1;
  return (b * a.Hin + iy) * a.Win + ix;
}

// ---- FD_FMT_SPLIT_BF16 helpers: a row is [ctot bf16 hi][ctot bf16 lo] in the same 4*ctot bytes ----------
__device__ __forceinline__ float split_load(const float* row_base, int c, int ctot) {
  const unsigned short* h = reinterpret_cast<const unsigned short*>(row_base);
  return __uint_as_float((unsigned)h[c] << 16) + __uint_as_float((unsigned)h[ctot + c] << 16);
}
__device__ __forceinline__ unsigned short bf16_bits_rn(float x) {
  unsigned u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) return (unsigned short)(u >> 16);      // inf / nan
  return (unsigned short)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
__device__ __forceinline__ void split_store(float* row_base, int c, int ctot, float v) {
  unsigned short* h = reinterpret_cast<unsigned short*>(row_base);
  unsigned short hi = bf16_bits_rn(v);
  h[c] = hi;
  h[ctot + c] = bf16_bits_rn(v - __uint_as_float((unsigned)hi << 16));
}

// Where the result row `o` lives: element c of the row is at (base, coff + c*cstride) -- in fp32 units for
// FD_FMT_FP32, in bf16 units inside the hi plane (lo plane out_ctot further) for FD_FMT_SPLIT_BF16.
struct OutRow { float* base; int coff; int cstride; };
__device__ __forceinline__ OutRow map_out_row(const ConvArgs& a, int o) {
  if (a.out_map == FD_OUTMAP_IDENTITY) return {a.out + (size_t)o * a.out_stride, 0, 1};
  if (a.out_map == FD_OUTMAP_BEV) {
    // SparseConvTensor.dense().view(N, C*D, H, W) (scn.py:165-168): channel = c*D + z, stored channels-last
    int4 c = a.out_coords[o];
    size_t pix = ((size_t)c.x * a.bevH + c.z) * a.bevW + c.w;
    return {a.out + pix * a.out_stride, c.y, a.bevD};
  }
  if (a.out_map == FD_OUTMAP_BEV_DMAJOR) {
    int4 c = a.out_coords[o];
    size_t pix = ((size_t)c.x * a.bevH + c.z) * a.bevW + c.w;
    return {a.out + pix * a.out_stride, c.y * a.cout, 1};
  }
  // OUTMAP_UPSAMPLE: o = (b, y, x) on the input grid -> (b, y*s+dy, x*s+dx) on the output grid
  int hw = a.Hin * a.Win;
  int b = o / hw;
  int r = o - b * hw;
  int y = r / a.Win, x = r - y * a.Win;
  size_t pix = ((size_t)b * a.Hin * a.up_s + (size_t)y * a.up_s + a.up_dy) * (a.Win * a.up_s) + (size_t)x * a.up_s + a.up_dx;
  return {a.out + pix * a.out_stride, 0, 1};
}
__device__ __forceinline__ void store_out(const ConvArgs& a, const OutRow& r, int c, float v) {
  const int e = r.coff + c * r.cstride;
  if (a.out_fmt == FD_FMT_SPLIT_BF16) split_store(r.base, e, a.out_ctot, v);
  else r.base[e] = v;
}
__device__ __forceinline__ float load_residual(const ConvArgs& a, int o, int c) {
  const float* row = a.residual + (size_t)o * a.res_stride;
  return a.res_fmt == FD_FMT_SPLIT_BF16 ? split_load(row, c, a.res_ctot) : row[c];
}

int conv_forward_simt(const ConvArgs& a, cudaStream_t stream);
bool wgrad_tc_supported(const ConvArgs& a);                              // wgrad_tc.cu
int conv_wgrad_tc_chunks(const ConvArgs& a);                                              // wgrad_tc.cu
int conv_wgrad_tc(const ConvArgs& a, float* dw, float* partial, cudaStream_t stream);    // wgrad_tc.cu
bool wgrad_os_supported(const ConvArgs& a);                                           // wgrad_tc.cu (output-stationary kernel)
int conv_wgrad_os_chunks(const ConvArgs& a);
size_t conv_wgrad_os_extra_bytes(const ConvArgs& a);
int conv_wgrad_os(const ConvArgs& a, float* partial, void* extra, cudaStream_t stream);
int conv_forward_tc(const ConvArgs& a, int precision, cudaStream_t stream);

}  // namespace fd
