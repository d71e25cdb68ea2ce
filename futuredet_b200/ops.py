"""Thin torch-tensor wrappers over the C ABI (include/futuredet_b200.h).

PyTorch is used here for device memory and streams only; every byte of arithmetic happens
inside libfuturedet_b200.so.  All functions raise RuntimeError when the library is missing or a
call fails -- there is no fallback path.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _req(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise RuntimeError("%s must be a contiguous CUDA %s tensor" % (name, dtype))
    return t


def _pow2_at_least(n):
    c = 1024
    while c < n:
        c <<= 1
    return c


def grid_size_of(point_cloud_range, voxel_size):
    """grid = round((hi - lo) / vs) in float32 (det3d/core/input/voxel_generator.py:10-11)."""
    r = np.asarray(point_cloud_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


# --------------------------------------------------------------------------- voxelize
def voxelize_vfe(points, batch_offsets, voxel_size, point_cloud_range, max_points, max_voxels,
                 num_feat=None, feat_stride=None, want_voxels=False):
    """Fused voxelize + mean VFE + batch-index column.

    points [total, P] fp32 CUDA (scenes concatenated), batch_offsets [B+1] int32 CUDA.
    Returns dict(features [cap, feat_stride], coords [cap,4] (b,z,y,x), num_points [cap],
    num_voxels [B], total [1]) with cap = B*max_voxels; rows >= total are undefined.
    want_voxels adds voxels [cap, max_points, num_feat] (the padded array of the legacy API).
    """
    lib = L.load()
    _req(points, torch.float32, "points")
    _req(batch_offsets, torch.int32, "batch_offsets")
    if points.dim() != 2:
        raise RuntimeError("points must be [N, P]")
    total, pstride = points.shape
    B = batch_offsets.numel() - 1
    num_feat = pstride if num_feat is None else num_feat
    feat_stride = num_feat if feat_stride is None else feat_stride
    dev = points.device
    cap = B * max_voxels
    feat = torch.empty((cap, feat_stride), dtype=torch.float32, device=dev)
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    npts = torch.empty((cap,), dtype=torch.int32, device=dev)
    nvox = torch.empty((B,), dtype=torch.int32, device=dev)
    tot = torch.empty((1,), dtype=torch.int32, device=dev)
    voxels = torch.empty((cap, max_points, num_feat), dtype=torch.float32, device=dev) if want_voxels else None
    ws_bytes = lib.fd_voxelize_workspace_bytes(total, B, max_voxels, max_points)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    grid = grid_size_of(point_cloud_range, voxel_size)
    rc = lib.fd_voxelize_vfe(_ptr(points), total, pstride, num_feat, _ptr(batch_offsets), B,
                             L.f32(point_cloud_range), L.f32(voxel_size), L.i32x3(grid),
                             max_points, max_voxels, _ptr(feat), feat_stride, _ptr(coords), _ptr(npts),
                             _ptr(nvox), _ptr(tot), _ptr(voxels), _ptr(ws), ws_bytes, _stream())
    L.check(rc, "fd_voxelize_vfe")
    return dict(features=feat, coords=coords, num_points=npts, num_voxels=nvox, total=tot, voxels=voxels)


def vfe_mean(voxels, num_points, num_feat=None):
    """VoxelFeatureExtractorV3.forward: [M,S,F] padded voxels, [M] counts -> [M,num_feat] means."""
    lib = L.load()
    _req(voxels, torch.float32, "voxels")
    if num_points.dtype != torch.int32:
        num_points = num_points.int()
    _req(num_points, torch.int32, "num_points")
    M, S, F = voxels.shape
    num_feat = F if num_feat is None else num_feat
    mean = torch.empty((M, num_feat), dtype=torch.float32, device=voxels.device)
    L.check(lib.fd_vfe_mean(_ptr(voxels), _ptr(num_points), M, S, F, num_feat, _ptr(mean), _stream()), "fd_vfe_mean")
    return mean


# --------------------------------------------------------------------------- rulebook
class CoordIndex:
    """Hash (b,z,y,x) -> row for one set of active sites."""

    def __init__(self, coords, n_dev, n_cap, shape, batch_size):
        lib = L.load()
        _req(coords, torch.int32, "coords")
        self.cap = _pow2_at_least(2 * max(n_cap, 1))
        self.table = torch.empty((self.cap,), dtype=torch.int64, device=coords.device)   # (key << 32) | row
        self.shape = [int(s) for s in shape]
        rc = lib.fd_coord_index_build(_ptr(coords), _ptr(n_dev), n_cap, int(batch_size), L.i32x3(self.shape),
                                      _ptr(self.table), self.cap, _stream())
        L.check(rc, "fd_coord_index_build")


class BitmapIndex:
    """Coordinate index of a strided conv's output set: 1 bit per cell + exclusive popcount prefix per 32-bit word
    (left behind by fd_rulebook_out_coords); rows of the set are in ascending linear order, so row == rank."""

    def __init__(self, bitmap, prefix, shape):
        self.bitmap, self.prefix, self.shape = bitmap, prefix, [int(s) for s in shape]


SORT_TILES = os.environ.get("FD_SORT_TILES", "1") != "0"   # tensor-core inference convs over 3x3x3 rulebooks run on pattern-sorted tiles (Rulebook.sorted_tiles)
SORT_WINDOW = int(os.environ.get("FD_SORT_WINDOW", 256 * 1024))     # rows per sorting window (a few scenes of a level: the gathered rows stay L2 resident)
SORT_STRIDED = os.environ.get("FD_SORT_STRIDED", "0") != "0"   # also sort the strided tables that come without keys (the
                             # input-side scatter build of the first strided conv): one launch per table, and the copy plus
                             # the key pass cost more than the launch gains.  Strided tables whose search hands the keys over
                             # (levels 3, 4) are sorted: +0.25 ms per 16 scenes
SORT_MIN_BATCH = int(os.environ.get("FD_SORT_MIN_BATCH", 3))   # scenes per forward from which the sort pays for its launches
STEM_SPLIT = os.environ.get("FD_STEM_SPLIT", "1") != "0"
SORT_MIN_ROWS = 4096         # smaller levels are not worth three more launches


class Rulebook:
    """Gather-form rulebook: nbr [K, n_out_cap] int32 (input row or -1), pair_num [K]."""

    def __init__(self, nbr, pair_num, K, out_coords, n_out_dev, n_out_cap, out_shape, ksize, stride, padding,
                 tile_mask=None):
        self.nbr, self._pair_num, self.K, self.tile_mask = nbr, pair_num, K, tile_mask
        self._sorted = None
        self.row_key = None       # [n_out_cap] int16 neighbour-pattern keys left by the neighbour search (sort key)
        self.out_coords, self.n_out_dev, self.n_out_cap = out_coords, n_out_dev, n_out_cap
        self.out_shape, self.ksize, self.stride, self.padding = out_shape, ksize, stride, padding

    def sorted_tiles(self):
        """(row_perm [n_out_cap], nbr_sorted [K, stride], tile_mask_sorted) of fd_rulebook_sort_rows: the rows of the table
        regrouped, inside windows of SORT_WINDOW rows, by neighbour pattern, so that the 128-row tiles of the
        output-stationary convolution skip most kernel offsets.  Built once per rulebook, shared by its convolutions."""
        if self._sorted is None:
            lib = L.load()
            n, dev = self.n_out_cap, self.nbr.device
            stride = int(self.nbr.stride(0))
            perm = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
            nbr_s = torch.empty_like(self.nbr)
            tmask = torch.empty(((max(n, 1) + 127) // 128,), dtype=torch.int32, device=dev)
            ws_bytes = lib.fd_rulebook_sort_workspace_bytes(n, SORT_WINDOW)
            ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
            rc = lib.fd_rulebook_sort_rows(_ptr(self.nbr), stride, self.K, _ptr(self.n_out_dev), n, SORT_WINDOW,
                                           _ptr(self.row_key), _ptr(perm), _ptr(nbr_s), _ptr(tmask), _ptr(ws), ws_bytes, _stream())
            L.check(rc, "fd_rulebook_sort_rows")
            self._sorted = (perm, nbr_s, tmask)
        return self._sorted

    @property
    def pair_num(self):
        """[K] int32 pairs per kernel offset (spconv `indice_pair_num`); counted on first use, off the hot path."""
        if self._pair_num is None:
            lib = L.load()
            self._pair_num = torch.empty((self.K,), dtype=torch.int32, device=self.nbr.device)
            rc = lib.fd_rulebook_count_pairs(_ptr(self.nbr), self.nbr.stride(0), _ptr(self.n_out_dev), self.n_out_cap,
                                             self.K, _ptr(self._pair_num), _stream())
            L.check(rc, "fd_rulebook_count_pairs")
        return self._pair_num

    def to_pairs(self):
        """spconv-1.x layout: (indice_pairs [K,2,P] int32 padded with -1, indice_pair_num [K])."""
        lib = L.load()
        n = self.n_out_cap
        ns = int(self.nbr.stride(0))
        dev = self.nbr.device
        pairs = torch.empty((self.K, 2, max(n, 1)), dtype=torch.int32, device=dev)
        total = self.K * ns
        tmp_bytes = ((4 * total + 255) // 256) * 256 + lib.fd_scan_tmp_bytes(total) + 256
        tmp = torch.empty((tmp_bytes,), dtype=torch.uint8, device=dev)
        rc = lib.fd_rulebook_to_pairs(_ptr(self.nbr), ns, _ptr(self.n_out_dev), n, self.K, _ptr(pairs), max(n, 1),
                                      _ptr(tmp), _stream())
        L.check(rc, "fd_rulebook_to_pairs")
        return pairs, self.pair_num


def _neighbors(out_coords, n_out_dev, n_out_cap, index, ksize, stride, padding, want_keys=False):
    lib = L.load()
    K = int(ksize[0] * ksize[1] * ksize[2])
    dev = out_coords.device
    nbr = torch.empty((K, max(n_out_cap, 1)), dtype=torch.int32, device=dev)
    pair_num = None          # counted lazily (Rulebook.pair_num)
    tile_mask = torch.empty(((max(n_out_cap, 1) + 127) // 128,), dtype=torch.int32, device=dev) if K <= 32 else None
    row_key = torch.empty((max(n_out_cap, 1),), dtype=torch.int16, device=dev) if want_keys and K <= 32 else None
    if isinstance(index, BitmapIndex):
        rc = lib.fd_rulebook_neighbors_bitmap(_ptr(out_coords), _ptr(n_out_dev), n_out_cap, _ptr(index.bitmap),
                                              _ptr(index.prefix), L.i32x3(index.shape), L.i32x3(ksize), L.i32x3(stride),
                                              L.i32x3(padding), _ptr(nbr), max(n_out_cap, 1), _ptr(pair_num),
                                              _ptr(tile_mask), _ptr(row_key), _stream())
        L.check(rc, "fd_rulebook_neighbors_bitmap")
        return nbr, pair_num, K, tile_mask, row_key
    rc = lib.fd_rulebook_neighbors(_ptr(out_coords), _ptr(n_out_dev), n_out_cap, _ptr(index.table),
                                   index.cap, L.i32x3(index.shape), L.i32x3(ksize), L.i32x3(stride),
                                   L.i32x3(padding), _ptr(nbr), max(n_out_cap, 1), _ptr(pair_num), _ptr(tile_mask),
                                   _ptr(row_key), _stream())
    L.check(rc, "fd_rulebook_neighbors")
    return nbr, pair_num, K, tile_mask, row_key


def rulebook_subm(coords, n_dev, n_cap, shape, ksize, index=None, batch_size=None):
    """SubMConv3d rulebook: outputs == inputs (same order), centred window."""
    if index is None:
        if batch_size is None:
            raise RuntimeError("rulebook_subm needs batch_size (or a prebuilt CoordIndex)")
        index = CoordIndex(coords, n_dev, n_cap, shape, batch_size)
    pad = [k // 2 for k in ksize]
    nbr, pair_num, K, tmask, row_key = _neighbors(coords, n_dev, n_cap, index, ksize, [1, 1, 1], pad,
                                                  want_keys=SORT_TILES and n_cap >= SORT_MIN_ROWS)
    rb = Rulebook(nbr, pair_num, K, coords, n_dev, n_cap, list(shape), list(ksize), [1, 1, 1], pad, tmask)
    rb.row_key = row_key
    return rb, index


def conv_out_shape(shape, ksize, stride, padding):
    return [(int(s) + 2 * p - k) // st + 1 for s, k, st, p in zip(shape, ksize, stride, padding)]


SCATTER_STRIDED = True       # strided rulebooks from the input side (fd_rulebook_neighbors_scatter); False: output-side search
FORCE_SCATTER = False        # tests: use the scatter build on every level


def rulebook_conv(coords, n_dev, n_cap, batch_size, shape, ksize, stride, padding, n_out_cap=None, index=None):
    """Regular SparseConv3d rulebook: active output set (ascending linear order) + neighbour table."""
    lib = L.load()
    dev = coords.device
    out_shape = conv_out_shape(shape, ksize, stride, padding)
    cells = batch_size * out_shape[0] * out_shape[1] * out_shape[2]
    fan = 1
    for k, s in zip(ksize, stride):
        fan *= -(-k // s)  # outputs one input can reach along this axis
    bound = min(cells, n_cap * fan)
    n_out_cap = bound if n_out_cap is None else min(n_out_cap, bound)
    words = (cells + 31) // 32
    bitmap = torch.empty((words + 1,), dtype=torch.int32, device=dev)
    prefix = torch.empty((words + 1,), dtype=torch.int32, device=dev)
    tmp = torch.empty((lib.fd_scan_tmp_bytes(words) + 256,), dtype=torch.uint8, device=dev)
    out_coords = torch.empty((max(n_out_cap, 1), 4), dtype=torch.int32, device=dev)
    n_out = torch.empty((1,), dtype=torch.int32, device=dev)
    rc = lib.fd_rulebook_out_coords(_ptr(coords), _ptr(n_dev), n_cap, batch_size, L.i32x3(shape), L.i32x3(ksize),
                                    L.i32x3(stride), L.i32x3(padding), L.i32x3(out_shape), _ptr(bitmap),
                                    _ptr(prefix), _ptr(tmp), _ptr(out_coords), n_out_cap, _ptr(n_out), _stream())
    L.check(rc, "fd_rulebook_out_coords")
    if SCATTER_STRIDED and (FORCE_SCATTER or not isinstance(index, BitmapIndex)):
        # input-stationary build: every input writes the <= prod(ceil(k/s)) slots it feeds (no input index needed).
        # Measured (4 bench scenes): faster than 27 hash probes per output row (first strided conv: 0.37 vs 0.45 ms
        # including the output-set kernels), not faster than the bitmap search of the deeper levels, which keep it.
        K = int(ksize[0] * ksize[1] * ksize[2])
        stride_n = (max(n_out_cap, 1) + 3) // 4 * 4               # 16-byte aligned table rows
        nbr = torch.empty((K, stride_n), dtype=torch.int32, device=dev)
        pair_num = None
        tmask = torch.empty(((max(n_out_cap, 1) + 127) // 128,), dtype=torch.int32, device=dev) if K <= 32 else None
        rc = lib.fd_rulebook_neighbors_scatter(_ptr(coords), _ptr(n_dev), n_cap, _ptr(bitmap), _ptr(prefix),
                                               L.i32x3(out_shape), L.i32x3(ksize), L.i32x3(stride), L.i32x3(padding),
                                               _ptr(n_out), n_out_cap, _ptr(nbr), stride_n, _ptr(tmask), _stream())
        L.check(rc, "fd_rulebook_neighbors_scatter")
        row_key = None
    else:
        index = index or CoordIndex(coords, n_dev, n_cap, shape, batch_size)
        nbr, pair_num, K, tmask, row_key = _neighbors(out_coords, n_out, n_out_cap, index, ksize, stride, padding,
                                                      want_keys=SORT_TILES and n_out_cap >= SORT_MIN_ROWS)
    rb = Rulebook(nbr, pair_num, K, out_coords, n_out, n_out_cap, out_shape, list(ksize), list(stride),
                  list(padding), tmask)
    rb.row_key = row_key        # output-side (bitmap / hash) searches hand the sort keys over; the scatter build cannot
    rb.out_index = BitmapIndex(bitmap, prefix, out_shape)      # index of the OUTPUT set for the layers that follow
    return rb, index


# --------------------------------------------------------------------------- convolution
PROFILE = None   # bench.py sets this to a list to collect per-launch CUDA-event timings of the conv family


def _prof_begin():
    if PROFILE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _prof_end(e0, kind, flops, launches=1):
    if e0 is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    PROFILE.append(dict(kind=kind, start=e0, end=e1, flops=flops, launches=launches))


import collections

_PACK_CACHE = collections.OrderedDict()      # LRU: live layers are touched every pass, per-step temporaries age out
_PACK_CACHE_MAX = 192


def packed_weights(w, separate_offsets=False):
    """bf16 hi/lo K-major pack of w [K,Cin,Cout] for the tensor-core arm (fd_conv_pack_weights); cached until
    the weight tensor changes (data_ptr / version)."""
    lib = L.load()
    key = (w.data_ptr(), w._version, tuple(w.shape), bool(separate_offsets))
    hit = _PACK_CACHE.get(key)
    if hit is not None:
        _PACK_CACHE.move_to_end(key)
        return hit[1]
    K, cin, cout = w.shape
    if separate_offsets:
        per = lib.fd_conv_packed_bytes(1, cin, cout)
        buf = torch.empty((K * per,), dtype=torch.uint8, device=w.device)
        for k in range(K):
            L.check(lib.fd_conv_pack_weights(C.c_void_p(w[k].data_ptr()), 1, cin, cout,
                                             C.c_void_p(buf.data_ptr() + k * per), _stream()), "fd_conv_pack_weights")
    else:
        buf = torch.empty((lib.fd_conv_packed_bytes(K, cin, cout),), dtype=torch.uint8, device=w.device)
        L.check(lib.fd_conv_pack_weights(_ptr(w), K, cin, cout, _ptr(buf), _stream()), "fd_conv_pack_weights")
    while len(_PACK_CACHE) >= _PACK_CACHE_MAX:
        _PACK_CACHE.popitem(last=False)          # evict the least recently used pack (and the weight it pins)
    _PACK_CACHE[key] = (w, buf)      # keep w alive so the data_ptr key cannot be recycled
    return buf


def tc_supported(cin, K):
    """Shapes the tcgen05 arm accepts; anything else must be requested as fp32 explicitly by the caller."""
    return cin % 8 == 0 and (cin % 64 == 0 or 64 % cin == 0) and K <= 32


FMT = {"fp32": 0, "split": 1}


class Feat:
    """A row-matrix operand: `t` is an fp32-typed tensor [..., Ctot] with unit channel stride whose bytes hold
    either plain fp32 rows (fmt "fp32") or FD_FMT_SPLIT_BF16 rows [Ctot bf16 hi | Ctot bf16 lo] (fmt "split");
    (c0, c) selects a channel slice.  Split tensors must never be read with torch arithmetic -- use to_fp32()."""

    def __init__(self, t, fmt="fp32", c0=0, c=None):
        if t.dtype != torch.float32 or not t.is_cuda or t.stride(-1) != 1:
            raise RuntimeError("Feat needs a CUDA fp32-typed tensor with unit channel stride")
        self.t, self.fmt, self.c0 = t, fmt, int(c0)
        self.ctot = int(t.shape[-1])
        self.c = self.ctot - self.c0 if c is None else int(c)

    @property
    def ptr(self):
        return self.t.data_ptr() + self.c0 * (4 if self.fmt == "fp32" else 2)

    @property
    def row_stride(self):
        return int(self.t.stride(-2))

    @property
    def rows(self):
        return int(self.t.numel() // self.ctot)

    def slice(self, c0, c):
        return Feat(self.t, self.fmt, self.c0 + c0, c)

    def to_fp32(self, n_dev=None):
        """fp32 tensor [..., c] of this view (copy through fd_convert_rows when split, plain slice otherwise)."""
        if self.fmt == "fp32":
            return self.t[..., self.c0:self.c0 + self.c]
        lib = L.load()
        out = torch.empty(self.t.shape[:-1] + (self.c,), dtype=torch.float32, device=self.t.device)
        rc = lib.fd_convert_rows(C.c_void_p(self.ptr), 1, self.row_stride, self.ctot, _ptr(out), 0, self.c, self.c,
                                 self.c, _ptr(n_dev), self.rows, _stream())
        L.check(rc, "fd_convert_rows")
        return out


def as_feat(x):
    return x if isinstance(x, Feat) else Feat(x)


def to_split(x, n_dev=None):
    """fp32 tensor -> new split-format Feat (API boundary helper)."""
    lib = L.load()
    x = as_feat(x)
    out = torch.empty(x.t.shape[:-1] + (x.c,), dtype=torch.float32, device=x.t.device)
    rc = lib.fd_convert_rows(C.c_void_p(x.ptr), FMT[x.fmt], x.row_stride, x.ctot, _ptr(out), 1, x.c, x.c, x.c,
                             _ptr(n_dev), x.rows, _stream())
    L.check(rc, "fd_convert_rows")
    return Feat(out, "split")


def copy_rows(src, dst, n_dev=None):
    """dst[..., :] = src[..., :] between two Feat views (any format pair, channel slices allowed)."""
    lib = L.load()
    src, dst = as_feat(src), as_feat(dst)
    if src.c != dst.c or src.rows != dst.rows:
        raise RuntimeError("copy_rows: shape mismatch")
    rc = lib.fd_convert_rows(C.c_void_p(src.ptr), FMT[src.fmt], src.row_stride, src.ctot, C.c_void_p(dst.ptr),
                             FMT[dst.fmt], dst.row_stride, dst.ctot, src.c, _ptr(n_dev), src.rows, _stream())
    L.check(rc, "fd_convert_rows")
    return dst


def _conv_desc(x, cin, w, scale, shift, residual, relu, out, precision, separate=False):
    d = L.ConvDesc()
    d.d_in = x.ptr; d.in_stride = x.row_stride; d.cin = cin
    d.in_format = FMT[x.fmt]; d.in_ctot = x.ctot
    d.d_w = w.data_ptr(); d.K, _, d.cout = w.shape
    if (L.PRECISIONS[precision] if isinstance(precision, str) else int(precision)) != L.PREC_FP32:
        d.d_w_packed = packed_weights(w, separate).data_ptr()
    d.d_scale = scale.data_ptr() if scale is not None else None
    d.d_shift = shift.data_ptr() if shift is not None else None
    if residual is not None:
        d.d_residual = residual.ptr; d.res_stride = residual.row_stride
        d.res_format = FMT[residual.fmt]; d.res_ctot = residual.ctot
    d.relu = int(bool(relu))
    d.d_out = out.ptr; d.out_stride = out.row_stride
    d.out_format = FMT[out.fmt]; d.out_ctot = out.ctot
    d.precision = L.PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    return d


def sparse_conv(x, w, rb, scale=None, shift=None, residual=None, relu=False, out=None, precision="fp32",
                bev=None, out_fmt="fp32", bev_dmajor=False, sort_tiles=False):
    """out[o] = act((sum_k x[nbr[k,o]] @ w[k]) * scale + shift (+ residual[o])).

    x / residual / out: torch fp32 tensors or Feat views (fp32 or split rows); w [K, Cin, Cout] fp32; rb: Rulebook.
    With bev=(B,D,H,W) the result goes straight into a zero-initialised channels-last BEV buffer [B,H,W,Cout*D]
    (SparseConvTensor.dense().view(N, C*D, H, W) of scn.py:165-168).  Returns a tensor for out_fmt "fp32",
    a Feat for "split".
    """
    lib = L.load()
    _req(w, torch.float32, "w")
    K, cin, cout = w.shape
    if K != rb.K:
        raise RuntimeError("weight has %d kernel offsets, rulebook has %d" % (K, rb.K))
    x = as_feat(x)
    residual = as_feat(residual) if residual is not None else None
    n_cap = rb.n_out_cap
    if bev is not None:
        B, D, H, Wd = bev
        if out is None:
            out = Feat(torch.zeros((B, H, Wd, cout * D), dtype=torch.float32, device=x.t.device), out_fmt)
        out = as_feat(out)
        d = _conv_desc(x, cin, w, scale, shift, None, relu, out, precision)
        d.out_map = L.OUTMAP_BEV_DMAJOR if bev_dmajor else L.OUTMAP_BEV
        d.d_out_coords4 = rb.out_coords.data_ptr(); d.bevD, d.bevH, d.bevW = D, H, Wd
    else:
        if out is None:
            out = Feat(torch.empty((max(n_cap, 1), cout), dtype=torch.float32, device=x.t.device), out_fmt)
        out = as_feat(out)
        d = _conv_desc(x, cin, w, scale, shift, residual, relu, out, precision)
        d.out_map = L.OUTMAP_IDENTITY
    d.mode = L.GATHER_TABLE
    if sort_tiles and SORT_TILES and rb.K == 27 and n_cap >= SORT_MIN_ROWS:
        perm, nbr_s, tmask_s = rb.sorted_tiles()
        d.d_nbr = nbr_s.data_ptr(); d.nbr_stride = nbr_s.stride(0)
        d.d_tile_mask = tmask_s.data_ptr(); d.d_row_perm = perm.data_ptr()
    else:
        d.d_nbr = rb.nbr.data_ptr(); d.nbr_stride = rb.nbr.stride(0)
        d.d_tile_mask = rb.tile_mask.data_ptr() if rb.tile_mask is not None else None
    d.d_n_out = rb.n_out_dev.data_ptr() if rb.n_out_dev is not None else None
    d.n_out_cap = n_cap
    e0 = _prof_begin()
    L.check(lib.fd_conv_forward(C.byref(d), _stream()), "fd_conv_forward(sparse)")
    _prof_end(e0, "sparse3d_c%d" % cout, lambda: 2.0 * float(rb.pair_num.sum().item()) * cin * cout)
    return out.t if out.fmt == "fp32" and out.c0 == 0 and out.c == out.ctot else out


def conv2d_nhwc(x, w, ksize, stride, padding, scale=None, shift=None, relu=False, out=None, residual=None,
                precision="fp32", transposed=False, out_fmt="fp32"):
    """Dense 2-D convolution on channels-last activations.

    x: [B,H,W,Cs] fp32 tensor (channel-slice views allowed) or Feat; w: [kh*kw, Cin, Cout].  `out` may be a
    channel slice of a wider buffer (tensor view or Feat) -- this is how torch.cat(ups, dim=1) of rpn.py:156-157 is
    fused away.  transposed=True: ConvTranspose2d with kernel == stride.  Returns a tensor for fp32 outputs given
    as tensors/None, else a Feat.
    """
    lib = L.load()
    _req(w, torch.float32, "w")
    K, cin, cout = w.shape
    out_was_tensor = out is None or isinstance(out, torch.Tensor)
    if isinstance(x, torch.Tensor) and x.dim() == 4 and (x.storage_offset() or x.shape[-1] != x.stride(-2)):
        # channel-slice tensor view of a wider channels-last buffer
        base = x.as_strided((x.shape[0], x.shape[1], x.shape[2], x.stride(2)), (x.stride(0), x.stride(1), x.stride(2), 1),
                            x.storage_offset() - x.storage_offset() % x.stride(2))
        x = Feat(base, "fp32", x.storage_offset() % x.stride(2), x.shape[3])
    x = as_feat(x)
    B, H, Wd = x.t.shape[0], x.t.shape[1], x.t.shape[2]
    kh, kw = ksize
    sh, sw = stride
    ph, pw = padding
    if x.t.stride(2) * Wd != x.t.stride(1) or x.t.stride(1) * H != x.t.stride(0):
        raise RuntimeError("x must be a channels-last [B,H,W,C] buffer with dense pixels")
    if transposed:
        Ho, Wo = H * sh, Wd * sw
    else:
        Ho, Wo = (H + 2 * ph - kh) // sh + 1, (Wd + 2 * pw - kw) // sw + 1
    if out is None:
        out = Feat(torch.empty((B, Ho, Wo, cout), dtype=torch.float32, device=x.t.device), out_fmt)
    elif isinstance(out, torch.Tensor) and (out.storage_offset() or out.shape[-1] != out.stride(-2)):
        base = out.as_strided((out.shape[0], out.shape[1], out.shape[2], out.stride(2)),
                              (out.stride(0), out.stride(1), out.stride(2), 1),
                              out.storage_offset() - out.storage_offset() % out.stride(2))
        out = Feat(base, "fp32", out.storage_offset() % out.stride(2), out.shape[3])
    out = as_feat(out)
    if out.t.stride(2) * Wo != out.t.stride(1) or out.t.stride(1) * Ho != out.t.stride(0) or out.t.shape[1] != Ho:
        raise RuntimeError("out must be a channels-last [B,Ho,Wo,C] buffer with dense pixels")
    residual = as_feat(residual) if residual is not None else None
    d = _conv_desc(x, cin, w, scale, shift, residual, relu, out, precision, separate=transposed)
    d.mode = L.GATHER_CONVT2D if transposed else L.GATHER_CONV2D
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, Wd, Ho, Wo
    d.kh, d.kw, d.sh, d.sw, d.ph, d.pw = kh, kw, sh, sw, ph, pw
    d.out_map = L.OUTMAP_IDENTITY
    d.d_n_out = None
    d.n_out_cap = B * H * Wd if transposed else B * Ho * Wo
    e0 = _prof_begin()
    L.check(lib.fd_conv_forward(C.byref(d), _stream()), "fd_conv_forward(conv2d)")
    _prof_end(e0, "dense2d", 2.0 * d.n_out_cap * cin * cout * (1 if transposed else K), K if transposed else 1)
    if out_was_tensor and out.fmt == "fp32":
        return out.t[..., out.c0:out.c0 + out.c]
    return out


def sparse_to_dense(features, coords, n_dev, n_cap, batch_size, shape):
    """SparseConvTensor.dense(): [N,C] rows -> zero-filled NCDHW tensor."""
    lib = L.load()
    if isinstance(features, Feat):
        features = features.to_fp32(n_dev)
    Cc = features.shape[1]
    D, H, Wd = [int(s) for s in shape]
    dense = torch.empty((batch_size, Cc, D, H, Wd), dtype=torch.float32, device=features.device)
    rc = lib.fd_sparse_to_dense_ncdhw(_ptr(features), features.stride(0), Cc, _ptr(coords), _ptr(n_dev), n_cap,
                                      batch_size, D, H, Wd, _ptr(dense), _stream())
    L.check(rc, "fd_sparse_to_dense_ncdhw")
    return dense
