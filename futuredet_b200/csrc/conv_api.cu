// C-ABI entry for the convolution family + SparseConvTensor.dense().
#include "conv_common.cuh"

namespace fd {

__global__ void __launch_bounds__(256)
sparse_to_dense_kernel(const float* __restrict__ feat, int feat_stride, int C, const int4* __restrict__ coords,
                       const int32_t* __restrict__ d_n, int n_cap, int D, int H, int W, float* __restrict__ dense) {
  const int n = d_n ? min(*d_n, n_cap) : n_cap;
  const long long total = (long long)n * C;
  const size_t plane = (size_t)D * H * W;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / C), c = (int)(e - (long long)row * C);
    int4 q = coords[row];  // (b, z, y, x)
    dense[((size_t)q.x * C + c) * plane + ((size_t)q.y * H + q.z) * W + q.w] = feat[(size_t)row * feat_stride + c];
  }
}

__global__ void __launch_bounds__(256)
convert_rows_kernel(const float* __restrict__ src, int sfmt, int sstride, int sctot, float* __restrict__ dst, int dfmt,
                    int dstride, int dctot, int C, const int32_t* __restrict__ d_n, long long n_cap) {
  const long long n = d_n ? min((long long)*d_n, n_cap) : n_cap;
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long r = e / C;
    int c = (int)(e - r * C);
    const float* srow = src + (size_t)r * sstride;
    float* drow = dst + (size_t)r * dstride;
    float v = sfmt == FD_FMT_SPLIT_BF16 ? split_load(srow, c, sctot) : srow[c];
    if (dfmt == FD_FMT_SPLIT_BF16) split_store(drow, c, dctot, v); else drow[c] = v;
  }
}

}  // namespace fd

extern "C" {

int fd_conv_forward(const fd_conv_desc* d, void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d != nullptr, "fd_conv_forward: null descriptor");
  FD_REQUIRE(d->d_in && d->d_w && d->d_out, "fd_conv_forward: null in/w/out pointer");
  FD_REQUIRE(d->cin >= 1 && d->cout >= 1 && d->K >= 1, "fd_conv_forward: bad cin/cout/K (%d,%d,%d)", d->cin,
             d->cout, d->K);
  FD_REQUIRE(d->in_stride >= d->cin, "fd_conv_forward: in_stride %d < cin %d", d->in_stride, d->cin);
  FD_REQUIRE(d->n_out_cap >= 0, "fd_conv_forward: negative n_out_cap");
  FD_REQUIRE(!d->d_residual || d->res_stride >= d->cout, "fd_conv_forward: res_stride < cout");
  ConvArgs a{};
  a.in = (const float*)d->d_in; a.in_stride = d->in_stride; a.cin = d->cin;
  a.in_fmt = d->in_format; a.in_ctot = d->in_ctot > 0 ? d->in_ctot : d->cin;
  a.out_fmt = d->out_format; a.out_ctot = d->out_ctot > 0 ? d->out_ctot : d->cout;
  a.res_fmt = d->res_format; a.res_ctot = d->res_ctot > 0 ? d->res_ctot : d->cout;
  for (int f : {d->in_format, d->out_format, d->res_format})
    FD_REQUIRE(f == FD_FMT_FP32 || f == FD_FMT_SPLIT_BF16, "fd_conv_forward: unknown row format %d", f);
  FD_REQUIRE(a.in_ctot >= d->cin && a.in_stride >= (d->in_format == FD_FMT_SPLIT_BF16 ? a.in_ctot : d->cin),
             "fd_conv_forward: in_ctot/in_stride inconsistent");
  a.w = d->d_w; a.cout = d->cout; a.K = d->K;
  a.wp = d->d_w_packed;
  a.scale = d->d_scale; a.shift = d->d_shift;
  a.residual = (const float*)d->d_residual; a.res_stride = d->res_stride;
  a.relu = d->relu;
  a.out = (float*)d->d_out; a.out_stride = d->out_stride;
  a.d_n = d->d_n_out; a.n_cap = d->n_out_cap;
  a.mode = d->mode; a.nbr = d->d_nbr; a.nbr_stride = d->nbr_stride; a.tile_mask = d->d_tile_mask;
  a.Hin = d->Hin; a.Win = d->Win; a.Hout = d->Hout; a.Wout = d->Wout;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->sh; a.sw = d->sw; a.ph = d->ph; a.pw = d->pw;
  a.row_perm = d->d_row_perm;
  FD_REQUIRE(!d->d_row_perm || d->mode == FD_GATHER_TABLE, "fd_conv_forward: d_row_perm needs FD_GATHER_TABLE");
  a.out_map = d->out_map;
  a.out_coords = (const int4*)d->d_out_coords4; a.bevD = d->bevD; a.bevH = d->bevH; a.bevW = d->bevW;

  if (d->out_map == FD_OUTMAP_BEV || d->out_map == FD_OUTMAP_BEV_DMAJOR) {
    FD_REQUIRE(d->d_out_coords4 && d->bevD >= 1 && d->bevH >= 1 && d->bevW >= 1,
               "fd_conv_forward: FD_OUTMAP_BEV needs out coords and bev dims");
    FD_REQUIRE(d->out_stride >= d->cout * d->bevD && a.out_ctot >= d->cout * d->bevD,
               "fd_conv_forward: BEV out_stride %d / out_ctot %d < cout*D %d", d->out_stride, a.out_ctot, d->cout * d->bevD);
    FD_REQUIRE(!d->d_residual, "fd_conv_forward: residual unsupported with FD_OUTMAP_BEV");
  } else {
    FD_REQUIRE(d->out_map == FD_OUTMAP_IDENTITY, "fd_conv_forward: unknown out_map %d", d->out_map);
    FD_REQUIRE(d->out_stride >= d->cout, "fd_conv_forward: out_stride %d < cout %d", d->out_stride, d->cout);
  }

  auto run = [&](const ConvArgs& args) -> int {
    if (d->precision == FD_PREC_FP32) return conv_forward_simt(args, stream);
    if (d->precision == FD_PREC_BF16X3 || d->precision == FD_PREC_BF16)
      return conv_forward_tc(args, d->precision, stream);
    return set_error(-1, "fd_conv_forward: unknown precision %d", d->precision);
  };

  switch (d->mode) {
    case FD_GATHER_TABLE:
      FD_REQUIRE(d->d_nbr && d->nbr_stride >= d->n_out_cap, "fd_conv_forward: bad neighbour table");
      return run(a);
    case FD_GATHER_CONV2D: {
      FD_REQUIRE(d->B >= 1 && d->Hin >= 1 && d->Win >= 1 && d->kh >= 1 && d->kw >= 1 && d->sh >= 1 && d->sw >= 1,
                 "fd_conv_forward: bad conv2d geometry");
      FD_REQUIRE(d->K == d->kh * d->kw, "fd_conv_forward: K != kh*kw");
      FD_REQUIRE(d->Hout == (d->Hin + 2 * d->ph - d->kh) / d->sh + 1 && d->Wout == (d->Win + 2 * d->pw - d->kw) / d->sw + 1,
                 "fd_conv_forward: Hout/Wout inconsistent with geometry");
      FD_REQUIRE(d->n_out_cap == d->B * d->Hout * d->Wout && !d->d_n_out, "fd_conv_forward: conv2d rows must be B*Hout*Wout");
      return run(a);
    }
    case FD_GATHER_CONV2D_DGRAD: {
      // training: data gradient of a Conv2d; rows = pixels of the conv's input grid (Hout x Wout of this descriptor)
      FD_REQUIRE(d->B >= 1 && d->Hin >= 1 && d->Win >= 1 && d->Hout >= 1 && d->Wout >= 1 && d->kh >= 1 && d->kw >= 1 &&
                     d->sh >= 1 && d->sw >= 1 && d->ph >= 0 && d->pw >= 0, "fd_conv_forward: bad conv2d-dgrad geometry");
      FD_REQUIRE(d->K == d->kh * d->kw, "fd_conv_forward: K != kh*kw");
      FD_REQUIRE(d->Hin == (d->Hout + 2 * d->ph - d->kh) / d->sh + 1 && d->Win == (d->Wout + 2 * d->pw - d->kw) / d->sw + 1,
                 "fd_conv_forward: conv2d-dgrad: dy grid inconsistent with the conv geometry");
      FD_REQUIRE(d->n_out_cap == d->B * d->Hout * d->Wout && !d->d_n_out, "fd_conv_forward: conv2d-dgrad rows must be B*H*W of the conv input");
      return run(a);
    }
    case FD_GATHER_CONVT2D: {
      // ConvTranspose2d with kernel == stride, no padding: each output pixel receives exactly one
      // (input pixel, kernel offset) product -> kh*kw independent 1x1 GEMMs with interleaved stores.
      FD_REQUIRE(d->kh == d->sh && d->kw == d->sw && d->kh == d->kw && d->ph == 0 && d->pw == 0,
                 "fd_conv_forward: convT2d supports kernel == stride, pad 0 only");
      FD_REQUIRE(d->K == d->kh * d->kw, "fd_conv_forward: K != kh*kw");
      FD_REQUIRE(d->Hout == d->Hin * d->sh && d->Wout == d->Win * d->sw, "fd_conv_forward: convT2d Hout/Wout mismatch");
      FD_REQUIRE(d->n_out_cap == d->B * d->Hin * d->Win && !d->d_n_out && d->out_map == FD_OUTMAP_IDENTITY,
                 "fd_conv_forward: convT2d rows must be B*Hin*Win (input pixels)");
      FD_REQUIRE(!d->d_residual, "fd_conv_forward: residual unsupported with convT2d");
      for (int k = 0; k < d->K; ++k) {
        ConvArgs p = a;
        p.mode = FD_GATHER_CONV2D;
        p.K = 1; p.kh = p.kw = 1; p.sh = p.sw = 1; p.ph = p.pw = 0;
        p.Hout = d->Hin; p.Wout = d->Win;
        p.w = (const float*)d->d_w + (size_t)k * d->cin * d->cout;
        if (d->d_w_packed) p.wp = (const char*)d->d_w_packed + (size_t)k * fd_conv_packed_bytes(1, d->cin, d->cout);
        p.out_map = OUTMAP_UPSAMPLE;
        p.up_s = d->sh; p.up_dy = k / d->kw; p.up_dx = k % d->kw;
        int rc = run(p);
        if (rc) return rc;
      }
      return 0;
    }
    default:
      return set_error(-1, "fd_conv_forward: unknown gather mode %d", d->mode);
  }
}

int fd_convert_rows(const void* d_src, int src_format, int src_stride, int src_ctot, void* d_dst, int dst_format,
                    int dst_stride, int dst_ctot, int C, const int32_t* d_n, int64_t n_cap, void* stream) {
  using namespace fd;
  FD_REQUIRE(d_src && d_dst && C >= 1 && src_ctot >= C && dst_ctot >= C && src_stride >= 1 && dst_stride >= 1,
             "fd_convert_rows: bad argument");
  FD_REQUIRE((src_format | dst_format) >> 1 == 0, "fd_convert_rows: unknown format");
  if (n_cap <= 0) return 0;
  convert_rows_kernel<<<persistent_grid(ceil_div(n_cap * C, 256), 8), 256, 0, (cudaStream_t)stream>>>(
      (const float*)d_src, src_format, src_stride, src_ctot, (float*)d_dst, dst_format, dst_stride, dst_ctot, C, d_n,
      n_cap);
  FD_LAUNCHED();
  return 0;
}

int fd_sparse_to_dense_ncdhw(const float* d_feat, int feat_stride, int C, const int32_t* d_coords4,
                             const int32_t* d_n, int n_cap, int B, int D, int H, int W, float* d_dense,
                             void* stream_) {
  using namespace fd;
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d_feat && d_coords4 && d_dense && C >= 1 && feat_stride >= C && B >= 1 && D >= 1 && H >= 1 && W >= 1,
             "fd_sparse_to_dense_ncdhw: bad argument");
  FD_CUDA(cudaMemsetAsync(d_dense, 0, sizeof(float) * (size_t)B * C * D * H * W, stream));
  if (n_cap <= 0) return 0;
  sparse_to_dense_kernel<<<persistent_grid(ceil_div((int64_t)n_cap * C, 256), 8), 256, 0, stream>>>(
      d_feat, feat_stride, C, (const int4*)d_coords4, d_n, n_cap, D, H, W, d_dense);
  FD_LAUNCHED();
  return 0;
}

}  // extern "C"
