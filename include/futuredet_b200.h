/*
 * futuredet_b200.h -- C ABI of the B200-native FutureDet LiDAR hot path.
 *
 * The reference (neeharperi/FutureDet) has no C ABI: its native code is reached
 * through numba JIT, pybind11 (iou3d_nms_cuda / deform_conv_cuda) and spconv's
 * torch.ops.  Every entry point below therefore cites the *reference Python/C++
 * interface it replaces* (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / pybind types.
 *   - every pointer named d_* is DEVICE memory owned by the caller; the library
 *     never allocates or frees in the hot loop.
 *   - every entry point takes the cudaStream_t (as void*) it must launch on and
 *     returns 0 on success, <0 for an invalid argument, >0 = cudaError_t.
 *     fd_last_error() returns a thread-local message.  Nothing ever exit()s.
 *   - row counts that are only known on the device (number of voxels, number of
 *     active sites after a strided sparse conv) are passed as `const int* d_n`
 *     together with a host-side capacity; kernels are persistent / grid-stride
 *     so no host synchronisation is needed anywhere in a forward pass.
 *   - all feature matrices are row-major fp32 [rows, stride] ("channels last").
 */
#ifndef FUTUREDET_B200_H_
#define FUTUREDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 4: fd_rulebook_sort_rows + fd_conv_desc.d_row_perm (tiles of rows sorted by neighbour pattern; appended, so ABI 3
 *    callers that zero the descriptor keep working).
 * 3: fd_conv_desc gained n_in_cap / d_in_split / d_out_split (weight gradient), fd_affine_act and fd_bn_backward gained
 *    the optional split-bf16 copy outputs (round 2, training step).  2: first published layout.                    */
#define FD_ABI_VERSION 4

/* ---- library ------------------------------------------------------------ */
int         fd_version(void);
const char* fd_last_error(void);
/* number of kernels this library has launched since load (bench "gpu_launches") */
int64_t     fd_launch_count(void);

/* ---- voxelize + VFE ------------------------------------------------------
 * Replaces  det3d/datasets/pipelines/preprocess.py:244-271  (Voxelization)
 *        -> det3d/core/input/voxel_generator.py:19-30       (VoxelGenerator.generate)
 *        -> det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184 (points_to_voxel)
 *        +  det3d/models/readers/voxel_encoder.py:17-24     (VoxelFeatureExtractorV3)
 *        +  det3d/torchie/parallel/collate.py:199-206       (batch index column)
 * Semantics are those of the sequential reference loop: voxel id = rank of the
 * voxel's first point in input order, only the first `max_points` points (input
 * order) of a voxel contribute, voxels first seen after `max_voxels` are dropped.
 * Coordinates use float32 IEEE sub/div/floor exactly as the reference does.
 *
 *   d_points        [total_points, point_stride] fp32, scenes concatenated
 *   d_batch_offsets [B+1] int32 row offsets of each scene in d_points
 *   range           host [6] xmin,ymin,zmin,xmax,ymax,zmax ; voxel_size host [3]
 *   grid            host [3] (gx,gy,gz) = round((hi-lo)/vs) computed by the caller in fp32
 * outputs (capacity B*max_voxels rows):
 *   d_feat   [*, feat_stride] fp32 mean of the first <=max_points points (cols >= num_feat zeroed)
 *   d_coords [*, 4] int32 (b,z,y,x) ;  d_npts [*] int32 ; d_nvox [B] int32 ; d_total [1] int32
 *   d_voxels optional (NULL to skip) [*, max_points, num_feat] fp32 zero-padded point lists, the
 *            `voxels` array of points_to_voxel (only needed by callers of the legacy padded API)
 */
size_t fd_voxelize_workspace_bytes(int64_t total_points, int B, int max_voxels, int max_points);
int fd_voxelize_vfe(const float* d_points, int64_t total_points, int point_stride, int num_feat,
                    const int32_t* d_batch_offsets, int B,
                    const float* range6, const float* voxel_size3, const int32_t* grid3,
                    int max_points, int max_voxels,
                    float* d_feat, int feat_stride, int32_t* d_coords, int32_t* d_npts,
                    int32_t* d_nvox, int32_t* d_total, float* d_voxels,
                    void* d_workspace, size_t workspace_bytes, void* stream);

/* VoxelFeatureExtractorV3.forward on padded voxels (det3d/models/readers/voxel_encoder.py:17-24):
 * d_mean[v,c] = sum_s d_voxels[v,s,c] / d_npts[v]   (d_voxels [M,S,F], first num_feat columns used) */
int fd_vfe_mean(const float* d_voxels, const int32_t* d_npts, int64_t M, int S, int F, int num_feat,
                float* d_mean, void* stream);

/* ---- rulebook ------------------------------------------------------------
 * Replaces spconv 1.x `ops.get_indice_pairs` (external, un-vendored; call sites
 * det3d/models/backbones/scn.py:99,105-106,110-146).  The rulebook is kept in
 * gather form: nbr[k*nbr_stride + o] = input row feeding output row o through
 * kernel offset k (row-major (kz,ky,kx)), or -1.  Cross-correlation:
 * in = out*stride - pad + k.
 *
 * A "coordinate index" is an open-addressing hash  key=((b*D+z)*H+y)*W+x -> row, one 64-bit word per
 * entry ((key << 32) | row, all-ones = empty): d_table [cap] uint64, cap = power of two >= 2*rows;
 * B*D*H*W must be < 2^32 - 1.
 */
int fd_coord_index_build(const int32_t* d_coords4, const int32_t* d_n, int n_cap, int B,
                         const int32_t* shape3, uint64_t* d_table, int64_t cap, void* stream);

/* Active output set of a regular (strided) SparseConv3d: every in-bounds out
 * coordinate hit by >=1 active input, ascending linear (b,z,y,x) order
 * (spconv-1.x GPU behaviour).  d_bitmap: >= ceil(B*oD*oH*oW/32)+1 int32 words,
 * d_wordprefix: same count, d_scan_tmp: fd_scan_tmp_bytes(words).           */
size_t fd_scan_tmp_bytes(int64_t n);
int fd_rulebook_out_coords(const int32_t* d_in_coords4, const int32_t* d_n_in, int n_in_cap,
                           int B, const int32_t* in_shape3, const int32_t* ksize3,
                           const int32_t* stride3, const int32_t* pad3, const int32_t* out_shape3,
                           uint32_t* d_bitmap, int32_t* d_wordprefix, void* d_scan_tmp,
                           int32_t* d_out_coords4, int n_out_cap, int32_t* d_n_out, void* stream);

/* Neighbour table for SubMConv3d (out == in coords, stride 1, pad = k/2) and for
 * SparseConv3d (out coords from fd_rulebook_out_coords).  d_pair_num (optional, [K] int32)
 * receives the number of pairs per kernel offset (spconv `indice_pair_num`); pass NULL on the hot path and call
 * fd_rulebook_count_pairs() only when the counts are needed (export, flop accounting).
 * d_tile_mask (optional, [ceil(n_out_cap/128)] uint32, K <= 32): bit k of word t is set when some row of
 * the 128-row tile t has a neighbour through offset k; fd_conv_forward skips the other offsets.
 * d_row_key (optional, [n_out_cap] uint16, K <= 32): the 9-bit neighbour-pattern signature of every row, the sort key
 * of fd_rulebook_sort_rows (the search has the pattern in registers; passing it on saves that call a pass over the table). */
int fd_rulebook_neighbors(const int32_t* d_out_coords4, const int32_t* d_n_out, int n_out_cap,
                          const uint64_t* d_in_table, int64_t in_cap,
                          const int32_t* in_shape3, const int32_t* ksize3, const int32_t* stride3,
                          const int32_t* pad3, int32_t* d_nbr, int nbr_stride, int32_t* d_pair_num,
                          uint32_t* d_tile_mask, uint16_t* d_row_key, void* stream);

int fd_rulebook_count_pairs(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap, int K,
                            int32_t* d_pair_num, void* stream);

/* Same search against the bitmap + per-word popcount prefix that fd_rulebook_out_coords left behind for a
 * strided conv's OUTPUT set (rows of that set are in ascending linear order, so row = rank): use it as the
 * coordinate index of every later layer at that resolution instead of building a hash.                    */
int fd_rulebook_neighbors_bitmap(const int32_t* d_out_coords4, const int32_t* d_n_out, int n_out_cap,
                                 const uint32_t* d_in_bitmap, const int32_t* d_in_wordprefix,
                                 const int32_t* in_shape3, const int32_t* ksize3, const int32_t* stride3,
                                 const int32_t* pad3, int32_t* d_nbr, int nbr_stride, int32_t* d_pair_num,
                                 uint32_t* d_tile_mask, uint16_t* d_row_key, void* stream);
/* Strided (SparseConv3d) rulebooks built from the INPUT side: every active input enumerates the <= prod(ceil(k/s))
 * outputs it reaches, finds their rows through the bitmap + popcount prefix of the OUTPUT set left by
 * fd_rulebook_out_coords, and writes nbr[k][out] = in (the call pre-fills the d_n_out live rows of d_nbr with -1 and
 * zeroes d_tile_mask; nbr_stride must be a multiple of 4 and d_nbr 16-byte aligned).
 * Same table, bit for bit, as fd_rulebook_neighbors over the output rows, with ~8x fewer lookups and no input index. */
int fd_rulebook_neighbors_scatter(const int32_t* d_in_coords4, const int32_t* d_n_in, int n_in_cap,
                                  const uint32_t* d_out_bitmap, const int32_t* d_out_wordprefix, const int32_t* out_shape3,
                                  const int32_t* ksize3, const int32_t* stride3, const int32_t* pad3,
                                  const int32_t* d_n_out, int n_out_cap, int32_t* d_nbr, int nbr_stride,
                                  uint32_t* d_tile_mask, void* stream);

/* Tile sorting (inference hot path).  The output-stationary convolution works on tiles of 128 consecutive rows and
 * skips the kernel offsets no row of a tile uses (d_tile_mask); in voxelizer / raster order nearly every tile of a LiDAR
 * level uses all 27.  This call reorders the ROWS OF THE TABLE, inside windows of `window` rows (rounded up to a multiple
 * of 1024; <= 0: 262144), by a 9-bit signature of each row's neighbour pattern, so that rows with the same pattern share
 * tiles:  d_nbr_sorted[k*nbr_stride + j] = d_nbr[k*nbr_stride + d_row_perm[j]],  d_tile_mask_sorted = masks of the sorted
 * tiles.  Pass the sorted table + masks + d_row_perm to fd_conv_forward: tile position j is then written to (and takes
 * its residual from) row d_row_perm[j], so inputs, outputs and their row order are unchanged and the results are
 * bit-identical to the unsorted call (a skipped offset only ever contributed exact zeros).  The order inside a bucket is
 * unspecified.  d_row_key: the keys fd_rulebook_neighbors[_bitmap] wrote for this table, or NULL (computed from the table).
 * Replaces nothing in the reference (spconv 1.x has no counterpart); K <= 32.                          */
size_t fd_rulebook_sort_workspace_bytes(int n_out_cap, int window);
int fd_rulebook_sort_rows(const int32_t* d_nbr, int nbr_stride, int K, const int32_t* d_n_out, int n_out_cap, int window,
                          const uint16_t* d_row_key, int32_t* d_row_perm, int32_t* d_nbr_sorted, uint32_t* d_tile_mask_sorted, void* d_workspace,
                          size_t workspace_bytes, void* stream);

/* Export to the spconv-1.x layout `indice_pairs [K,2,P_cap]` (pairs of offset k
 * listed in ascending output row), for parity checks and interop.            */
int fd_rulebook_to_pairs(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap,
                         int K, int32_t* d_pairs, int pair_cap, void* d_scan_tmp, void* stream);

/* ---- gather -> implicit GEMM convolution ----------------------------------
 * One kernel family serves
 *   spconv SubMConv3d / SparseConv3d forward (`indice_conv`, scn.py:11-34,98-146),
 *   nn.Conv2d / ZeroPad2d+Conv2d / ConvTranspose2d(k=s) of the RPN neck
 *   (det3d/models/necks/rpn.py:70-142) and of CenterHead / SepHead
 *   (det3d/models/bbox_heads/center_head.py:129-152,344-349),
 * with the eval-mode BatchNorm, conv bias, residual add and ReLU of
 * scn.py:64-80 / rpn.py:124-142 fused into the epilogue:
 *     y = act( (sum_k in[nbr(o,k)] @ W[k]) * scale + shift (+ residual) )
 * Weights are [K, Cin, Cout] fp32 (spconv-1.x layout [kD,kH,kW,Cin,Cout]).
 */
/* Row formats.  FD_FMT_FP32: row = C_tot fp32.  FD_FMT_SPLIT_BF16: row = [C_tot bf16 "hi"][C_tot bf16 "lo"]
 * with value = hi + lo (hi = bf16_rn(x), lo = bf16_rn(x - hi)); same 4*C_tot bytes per row as fp32, so a
 * buffer allocated for one format can hold the other.  It is the inter-layer format of the tensor-core arm:
 * the producing kernel's epilogue splits once, consuming kernels gather both planes with cp.async and feed
 * the bf16 tensor cores with no conversion work.  `*_ctot` = channels per row of the underlying buffer
 * (locates the lo plane; d_in/d_out/d_residual may point at a channel slice).  Strides stay in 4-byte units. */
enum { FD_FMT_FP32 = 0, FD_FMT_SPLIT_BF16 = 1 };

typedef struct fd_conv_desc {
  /* input rows */
  const void*    d_in;        int32_t in_stride;  int32_t cin;
  int32_t        in_format;   int32_t in_ctot;
  /* weights / epilogue vectors */
  const float*   d_w;         int32_t cout;       int32_t K;
  const void*    d_w_packed;  /* fd_conv_pack_weights output; required for FD_PREC_BF16X3 / FD_PREC_BF16 */
  const float*   d_scale;     /* [cout] or NULL (=1) */
  const float*   d_shift;     /* [cout] or NULL (=0) */
  const void*    d_residual;  int32_t res_stride; /* NULL: none */
  int32_t        res_format;  int32_t res_ctot;
  int32_t        relu;
  /* output rows */
  void*          d_out;       int32_t out_stride;
  int32_t        out_format;  int32_t out_ctot;
  const int32_t* d_n_out;     /* device row count, or NULL -> n_out_cap rows */
  int32_t        n_out_cap;
  /* gather mode */
  int32_t        mode;        /* FD_GATHER_* */
  const int32_t* d_nbr;       int32_t nbr_stride;          /* FD_GATHER_TABLE */
  const uint32_t* d_tile_mask;                             /* FD_GATHER_TABLE, optional (fd_rulebook_neighbors) */
  int32_t B, Hin, Win, Hout, Wout, kh, kw, sh, sw, ph, pw; /* FD_GATHER_CONV2D / _CONVT2D */
  /* output row mapping */
  int32_t        out_map;     /* FD_OUTMAP_* */
  const int32_t* d_out_coords4; int32_t bevD, bevH, bevW;  /* FD_OUTMAP_BEV */
  int32_t        precision;   /* FD_PREC_* */
  int32_t        n_in_cap;    /* fd_conv_wgrad* with FD_GATHER_TABLE: rows of d_in (lets the tcgen05 arm pre-split
                               * the layer input once); 0 = unknown (the per-offset tcgen05 kernel is used).  Ignored
                               * by fd_conv_forward.                                                              */
  const void*    d_in_split;  /* fd_conv_wgrad_det, optional: dense FD_FMT_SPLIT_BF16 copies ([rows][C hi | C lo]) of   */
  const void*    d_out_split; /* d_in / dL/dy that the caller already holds (fd_affine_act / fd_bn_backward write    */
                              /* them); NULL: the tcgen05 arm splits the fp32 rows into its workspace             */
  const int32_t* d_row_perm;  /* fd_conv_forward with FD_GATHER_TABLE, optional: d_nbr / d_tile_mask are the SORTED table
                               * of fd_rulebook_sort_rows and tile position j belongs to output row d_row_perm[j]
                               * (NULL: position j is row j).  Must be NULL for fd_conv_wgrad*.                       */
} fd_conv_desc;

enum { FD_GATHER_TABLE = 0, FD_GATHER_CONV2D = 1, FD_GATHER_CONVT2D = 2,
       /* data gradient of a Conv2d (training): rows = pixels of the conv's input grid (given as Hout x Wout), gathered
        * tensor = dL/dy on the conv's output grid (given as Hin x Win), kh/kw/sh/sw/ph/pw = the conv's own geometry,
        * weights = W[k]^T ([K, Cout, Cin]).                                                                      */
       FD_GATHER_CONV2D_DGRAD = 3 };
/* FD_OUTMAP_BEV: channel = c*D + z (the reference's dense().view() order).  FD_OUTMAP_BEV_DMAJOR: channel = z*C + c
 * (contiguous per row -> vector stores); only for consumers that permute their input channels accordingly. */
enum { FD_OUTMAP_IDENTITY = 0, FD_OUTMAP_BEV = 1, FD_OUTMAP_BEV_DMAJOR = 2 };
/* FD_PREC_FP32: CUDA-core fp32 FMA (exact reference arithmetic).
 * FD_PREC_BF16X3: tcgen05 tensor cores, 3-term bf16 split (fp32-class accuracy).
 * FD_PREC_BF16: tcgen05 single pass bf16 (fast mode, outside the 1e-3 contract). */
enum { FD_PREC_FP32 = 0, FD_PREC_BF16X3 = 1, FD_PREC_BF16 = 2 };

int fd_conv_forward(const fd_conv_desc* desc, void* stream);

/* Tensor-core weight pre-pack: fp32 [K,Cin,Cout] -> bf16 hi/lo planes, K-major, zero padded
 * ([2][Cout_pad][pad64(K*Cin)]).  For FD_GATHER_CONVT2D pack every kernel offset on its own
 * (K=1 per call, outputs fd_conv_packed_bytes(1,cin,cout) apart).                              */
size_t fd_conv_packed_bytes(int K, int cin, int cout);
int fd_conv_pack_weights(const float* d_w, int K, int cin, int cout, void* d_packed, void* stream);

/* SparseConvTensor.dense() for API parity (scn.py:165-168): scatter [N,C] rows
 * at (b,z,y,x) into a zeroed NCDHW fp32 tensor.                              */
int fd_sparse_to_dense_ncdhw(const float* d_feat, int feat_stride, int C, const int32_t* d_coords4,
                             const int32_t* d_n, int n_cap, int B, int D, int H, int W,
                             float* d_dense, void* stream);

/* Convert a row matrix between FD_FMT_FP32 and FD_FMT_SPLIT_BF16 (API boundaries, tests).
 * rows from d_n (device, may be NULL -> n_cap); src/dst may be channel slices (ctot locates lo planes). */
int fd_convert_rows(const void* d_src, int src_format, int src_stride, int src_ctot, void* d_dst,
                    int dst_format, int dst_stride, int dst_ctot, int C, const int32_t* d_n, int64_t n_cap,
                    void* stream);

/* ---- CenterHead.loss forward (standard branch) ----------------------------------------------------
 * Replaces det3d/models/bbox_heads/center_head.py:396-539 (+ _sigmoid :392-394, in place on d_hm),
 * det3d/models/losses/centernet_loss.py:18-25,75-95 and det3d/core/utils/center_utils.py:66-80.
 *   d_hm        heat-map logits of one task, element (b,c,s) at d_hm[b*hm_sb + c*hm_sc + s*hm_ssp], s = y*W + x;
 *               overwritten with clamp(sigmoid(x), 1e-4, 1-1e-4) as the reference does
 *   d_hm_target [B,C,H,W] fp32 ; d_ind / d_cat [B,M] int64 ; d_mask [B,M] uint8 (timestep 0, as the reference uses)
 *   d_mask_t    [T] device pointers to the per-timestep masks (only for `num_positive`)
 *   d_pred_ptr / d_pred_sb / d_pred_ssp  [T*NC]: channel c of timestep t of anno_box = cat(reg,height,dim,
 *               vel[2t:2t+2],rot) lives at ptr[b*sb + s*ssp]
 *   d_tgt_ptr   [T] pointers to anno_box targets [B,M,tgt_dim]; d_tgt_sel [NC] = columns [0..7,-2,-1]
 *   d_out       [3 + T + T*NC] fp32: loss, hm_loss, num_positive, loc_loss[T], loc_loss_elem[T][NC]            */
size_t fd_center_loss_workspace_bytes(void);
int fd_center_head_loss(float* d_hm, int64_t hm_sb, int64_t hm_sc, int64_t hm_ssp, const float* d_hm_target,
                        int B, int C, int H, int W, const int64_t* d_ind, const uint8_t* d_mask,
                        const int64_t* d_cat, const uint8_t* const* d_mask_t, int M, int T, int NC,
                        const float* const* d_pred_ptr, const int64_t* d_pred_sb, const int64_t* d_pred_ssp,
                        const float* const* d_tgt_ptr, int tgt_dim, const int32_t* d_tgt_sel,
                        const float* d_code_w, const float* d_code_w_forecast, float weight, float* d_out,
                        void* d_workspace, void* stream);

/* ---- training: backward kernels -----------------------------------------------------------------------
 * The reference trains through torch autograd over spconv's `indice_conv_backward`, cuDNN and ATen batch-norm
 * (det3d/torchie/trainer/trainer.py:317-344 -> loss.backward(); det3d/models/backbones/scn.py:37-176,
 * det3d/models/necks/rpn.py:124-159, det3d/models/bbox_heads/center_head.py:129-174,396-539).  Here every backward
 * op is an explicit kernel; all operands are FD_FMT_FP32 rows ("channels last") with explicit row strides.
 *
 * Data gradient of a sparse conv = fd_conv_forward over the TRANSPOSED rulebook with W[k]^T:
 *   SubMConv3d : the table is its own transpose with the offsets mirrored (nbrT[k] = nbr[K-1-k]);
 *   SparseConv3d: fd_rulebook_transpose builds nbrT[k][i] = o for every pair (i = nbr[k][o]).
 * Data gradient of Conv2d = fd_conv_forward in FD_GATHER_CONV2D_DGRAD mode; of ConvTranspose2d(k == s) = a
 * k x k stride-k FD_GATHER_CONV2D over dL/dy.                                                                */
int fd_rulebook_transpose(const int32_t* d_nbr, int nbr_stride, const int32_t* d_n_out, int n_out_cap, int K,
                          int32_t* d_nbr_t, int nbr_t_stride, int n_in_cap, void* stream);

/* Weight gradient of any convolution fd_conv_forward can run: dW[k] += gather_k(in)^T @ dy  (spconv
 * `indice_conv_backward` filter gradient / cuDNN wgrad).  `desc` describes the FORWARD convolution with
 * d_in = the layer input, d_out = dL/dy (read only) laid out as the forward output (out_map / out_stride honoured);
 * scale/shift/residual/relu/d_w are ignored.  Accumulates with fp32 atomics into d_dw [K,Cin,Cout]: the caller
 * zeroes it.  fp32 rows only.  desc->precision: FD_PREC_FP32 = exact fp32 on CUDA cores; FD_PREC_BF16X3 = tcgen05
 * (both operands MN-major, 3-term bf16 split, fp32 accumulation in TMEM) where the rows allow it (Cin, Cout multiples
 * of 8, 16-byte aligned rows, identity output map), the CUDA-core kernel otherwise.                            */
int fd_conv_wgrad(const fd_conv_desc* desc, float* d_dw, void* stream);
/* Same gradient, bit-reproducible: every row chunk of a (kernel offset, Cin x Cout) tile stores its partial tile into
 * its own slot of d_workspace and an ordered reduce adds the slots into d_dw (no floating-point atomics anywhere;
 * torch autograd over cuDNN / spconv makes no such promise).  fd_conv_wgrad_workspace_bytes(desc) sizes the buffer. */
size_t fd_conv_wgrad_workspace_bytes(const fd_conv_desc* desc);
int fd_conv_wgrad_det(const fd_conv_desc* desc, float* d_dw, void* d_workspace, size_t workspace_bytes, void* stream);

/* BatchNorm1d/2d in training mode over the first n rows of x [n, C] (det3d/models/utils/norm.py:59-64 ->
 * torch.nn.BatchNorm*: biased batch variance for normalisation, unbiased for the running estimate).
 * fd_bn_train_stats writes mean / invstd (saved for backward) and the folded scale = gamma*invstd,
 * shift = beta - mean*scale consumed by fd_affine_act (or by a fused conv epilogue), and updates the running
 * statistics with `momentum` (pass NULL to skip).  Deterministic two-stage reduction in fp64.               */
size_t fd_bn_workspace_bytes(int C);
int fd_bn_train_stats(const float* d_x, int x_stride, int C, const int32_t* d_n, int64_t n_cap, float eps,
                      float momentum, const float* d_gamma, const float* d_beta, float* d_running_mean,
                      float* d_running_var, float* d_mean, float* d_invstd, float* d_scale, float* d_shift,
                      void* d_workspace, void* stream);
/* y = act(x * scale[c] + shift[c] (+ res)), rows < n; scale/shift/res may be NULL (1 / 0 / none).  d_y_split
 * (optional): an FD_FMT_SPLIT_BF16 copy of y is written beside it -- rows of 2*split_ctot bf16, the pointer already
 * offset to the first channel of the slice -- so that the tensor-core convolutions that consume y (forward, weight
 * gradient) gather ready-made bf16 planes.                                                                      */
int fd_affine_act(const float* d_x, int x_stride, int C, const float* d_scale, const float* d_shift,
                  const float* d_res, int res_stride, int relu, float* d_y, int y_stride, void* d_y_split,
                  int split_ctot, const int32_t* d_n, int64_t n_cap, void* stream);
/* Backward of y = act(bn(x) (+ res)):  dz = dy * [y > 0] (relu) ;  dgamma = sum dz*xhat ; dbeta = sum dz ;
 * dx = gamma*invstd*(dz - dbeta/n - xhat*dgamma/n) ;  dres = dz (d_dres may be NULL).  d_dx_split (optional): dense
 * FD_FMT_SPLIT_BF16 copy of dx ([rows][C hi | C lo]) for the data- / weight-gradient convolutions.             */
int fd_bn_backward(const float* d_dy, int dy_stride, const float* d_y, int y_stride, int relu, const float* d_x,
                   int x_stride, int C, const int32_t* d_n, int64_t n_cap, const float* d_mean,
                   const float* d_invstd, const float* d_gamma, float* d_dx, int dx_stride, void* d_dx_split,
                   float* d_dres, int dres_stride, float* d_dgamma, float* d_dbeta, void* d_workspace, void* stream);
/* out[c] = sum_rows x[r, c]  (conv bias gradient). */
int fd_col_sum(const float* d_x, int x_stride, int C, const int32_t* d_n, int64_t n_cap, float* d_out,
               void* d_workspace, void* stream);
/* dst[r, c] += src[r, c]  (gradient accumulation where an activation feeds two consumers). */
int fd_add_rows(float* d_dst, int dst_stride, const float* d_src, int src_stride, int C, const int32_t* d_n,
                int64_t n_cap, void* stream);
/* SparseConvTensor.dense().view(N, C*D, H, W) (scn.py:165-168) as channels-last [B,H,W,C*D] (channel = c*D + z)
 * and its backward (gather of the BEV gradient back to the active rows).  d_bev is zero-filled by the call.  */
int fd_rows_to_bev(const float* d_rows, int row_stride, int C, const int32_t* d_coords4, const int32_t* d_n,
                   int n_cap, int B, int D, int H, int W, float* d_bev, void* stream);
int fd_bev_to_rows(const float* d_bev, int C, const int32_t* d_coords4, const int32_t* d_n, int n_cap, int B,
                   int D, int H, int W, float* d_rows, int row_stride, void* stream);

/* Backward of fd_center_head_loss (same argument meaning; d_hm holds the clamped probabilities the forward left
 * in place).  Writes dL/d(hm logits) densely into d_ghm (same strides as d_hm) and ADDS the masked-L1 gradients
 * at the object centres into the planes d_gpred_ptr [T*NC] (same strides as the predictions; the caller zeroes
 * them).  d_gscale: optional device scalar multiplying every gradient (upstream dL), NULL = 1.               */
int fd_center_head_loss_backward(const float* d_hm, float* d_ghm, int64_t hm_sb, int64_t hm_sc, int64_t hm_ssp,
                                 const float* d_hm_target, int B, int C, int H, int W, const int64_t* d_ind,
                                 const uint8_t* d_mask, const int64_t* d_cat, int M, int T, int NC,
                                 const float* const* d_pred_ptr, float* const* d_gpred_ptr,
                                 const int64_t* d_pred_sb, const int64_t* d_pred_ssp,
                                 const float* const* d_tgt_ptr, int tgt_dim, const int32_t* d_tgt_sel,
                                 const float* d_code_w, const float* d_code_w_forecast, float weight,
                                 const float* d_gscale, void* stream);

/* ---- CenterHead.predict: decode + masks + top-k + rotated BEV NMS, all on device ---------------------------
 * Replaces det3d/models/bbox_heads/center_head.py:541-747 (predict / post_processing, standard mode),
 * det3d/core/bbox/box_torch_ops.py:248-276 (rotate_nms_pcdet) and det3d/ops/iou3d_nms (nms_gpu:
 * src/iou3d_nms.cpp:90-136 + src/iou3d_nms_kernel.cu:264-311, which copies a bit mask to the host, sweeps it there
 * and cudaMalloc/cudaFrees per call).
 *   d_out       head output of one task, channels last [B, H*W, row_stride]; c_* = first channel of each head
 *               (reg 2, height 1, dim 3, rot 2 (sin, cos), hm num_cls); c_vel [T] = first velocity channel of every
 *               emitted forecast timestep (center_head.py:561-570: the timesteps share everything but `vel`, so
 *               selection and NMS run once per sample and the kept boxes are emitted T times)
 *   candidates  cells with sigmoid(hm).max > score_threshold and (x, y, z) inside post_center_range6; ordered by
 *               score (descending, ties by cell index), first pre_max (<= 1024) enter the NMS, first post_max kept
 *   outputs     d_boxes [B,T,post_max,9] (x,y,z,w,l,h,vx,vy,rot), d_scores / d_labels [B,T,post_max],
 *               d_cells [B,post_max] (BEV cell index of every kept box), d_count [B]                         */
size_t fd_center_predict_workspace_bytes(int B, int H, int W);
int fd_center_predict(const float* d_out, int row_stride, int c_reg, int c_height, int c_dim, int c_rot,
                      const int32_t* c_vel, int T, int c_hm, int num_cls, int B, int H, int W, float score_threshold,
                      const float* post_center_range6, float out_size_factor, float voxel_x, float voxel_y,
                      float pc_x0, float pc_y0, float nms_iou_threshold, int pre_max, int post_max, float* d_boxes,
                      float* d_scores, int32_t* d_labels, int32_t* d_cells, int32_t* d_count, void* d_workspace,
                      void* stream);
/* Pairwise rotated BEV IoU of boxes [n,7] (x,y,z,dx,dy,dz,heading) -> d_iou [na,nb]; replaces boxes_iou_bev_gpu
 * (det3d/ops/iou3d_nms/src/iou3d_nms.cpp:61-88, iou3d_nms_kernel.cu:230-262).                               */
int fd_boxes_iou_bev(const float* d_boxes_a, int na, const float* d_boxes_b, int nb, float* d_iou, void* stream);

/* ---- multi-sweep assembly (the step in front of the voxelizer) ------------------------------------------
 * Replaces det3d/datasets/pipelines/loading.py:24-60 (read_file column selection, remove_close, read_sweep) and
 * :120-140 (key frame + sweeps concatenation, `combined = hstack([points, times])`).
 *   d_raw            [total_records, raw_stride] fp32: the untouched .bin payloads of every sweep of every scene,
 *                    concatenated in output order (nuScenes: raw_stride 5 = x,y,z,intensity,ring; num_feat 4)
 *   d_sweep_offsets  [S+1] int32 record offsets;  d_sweep_scene [S] int32 scene of every sweep (non-decreasing)
 *   d_transforms     [S,16] float64 row-major 4x4 sweep-to-keyframe transforms, applied when d_flags[s] & 1
 *                    (float64 product stored as float32, as numpy does);  d_flags[s] & 2: remove_close with
 *                    |x| < close_radius and |y| < close_radius;  d_time_lag [S] -> last output column
 * outputs: d_points [total_records, num_feat+1] (kept rows first, reference order; rows >= count are NaN so that
 *          fd_voxelize_vfe rejects them), d_batch_offsets [B+1] int32 rows of every scene, d_count [1].            */
size_t fd_sweeps_workspace_bytes(int64_t total_records);
int fd_assemble_sweeps(const float* d_raw, int64_t total_records, int raw_stride, int num_feat,
                       const int32_t* d_sweep_offsets, const double* d_transforms, const int32_t* d_flags,
                       const float* d_time_lag, const int32_t* d_sweep_scene, int S, int B, float close_radius,
                       float* d_points, int32_t* d_batch_offsets, int32_t* d_count, void* d_workspace,
                       size_t workspace_bytes, void* stream);

/* ---- CenterPoint target assignment (the step that feeds CenterHead.loss) -----------------------------------
 * Replaces det3d/datasets/pipelines/preprocess.py:464-546 (AssignLabel, one task, one timestep) with
 * det3d/core/utils/center_utils.py:17-64 (gaussian_radius, gaussian2D, draw_umich_gaussian).
 *   d_boxes   [B, n_max, box_dim = 12] fp32 (x,y,z,w,l,h,vx,vy,rvx,rvy,rot,rrot: the last two columns are
 *             wrapped to [-pi, pi) as preprocess.py:449-456 does), d_classes [B, n_max] 1-based class within the task
 *             (<= 0: skip), d_num [B] objects per sample (at most max_objs are used, object k -> slot k)
 *   outputs (zero-filled by the call): d_hm [B,num_cls,H,W], d_anno_box [B,max_objs,14] (dx,dy,z,log w,log l,log h,
 *             vx,vy,rvx,rvy,sin rot,cos rot,sin rrot,cos rrot), d_ind / d_cat [B,max_objs] int64, d_mask uint8.
 *   radius_mult / timestep: mult = clamp(|v| * (1 + timestep) / 2, 1, 4) on the Gaussian radius (:487-491).       */
int fd_assign_center_targets(const float* d_boxes, const int32_t* d_classes, const int32_t* d_num, int B, int n_max,
                             int box_dim, int num_cls, int W, int H, float pc_x0, float pc_y0, float voxel_x,
                             float voxel_y, float out_size_factor, float gaussian_overlap, int min_radius,
                             int radius_mult, int timestep, int max_objs, float* d_hm, float* d_anno_box,
                             int64_t* d_ind, uint8_t* d_mask, int64_t* d_cat, void* stream);

/* small helpers used by the host layer */
int fd_fill_i32(int32_t* d_ptr, int64_t n, int32_t value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FUTUREDET_B200_H_ */
