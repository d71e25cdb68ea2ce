"""Torch-tensor wrappers over the training (backward) entry points of the C ABI (include/futuredet_b200.h).

Same rules as ops.py: PyTorch provides device memory and streams, every byte of arithmetic happens inside
libfuturedet_b200.so, failures raise RuntimeError, there is no fallback.  All operands are fp32 "channels last"
row matrices `[rows, C]` with unit channel stride (channel-slice views of wider buffers are allowed).
"""
import ctypes as C

import torch

from . import lib as L
from . import ops
from .ops import _ptr, _stream

_WS = {}


def _workspace(dev, nbytes):
    """One grow-only scratch buffer per device for the two-stage reductions (stream ordered, so reuse is safe)."""
    buf = _WS.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, 1 << 20),), dtype=torch.uint8, device=dev)
        _WS[dev] = buf
    return buf


def _rows(t):
    """(tensor, row stride in floats, C) of a [..., C] fp32 CUDA tensor with unit channel stride and uniform row pitch."""
    if t.dtype != torch.float32 or not t.is_cuda or t.stride(-1) != 1:
        raise RuntimeError("expected a CUDA fp32 tensor with unit channel stride")
    rs = t.stride(-2) if t.dim() >= 2 else t.shape[-1]
    for d in range(t.dim() - 2):
        if t.stride(d) != t.stride(d + 1) * t.shape[d + 1]:
            raise RuntimeError("rows of the tensor are not uniformly strided")
    return t, int(rs), int(t.shape[-1])


def n_rows(t):
    return int(t.numel() // t.shape[-1])


# ------------------------------------------------------------------------------------------ rulebook
def rulebook_transpose(rb, n_in_cap):
    """nbrT [K, n_in_cap]: nbrT[k, i] = o for every pair i = nbr[k, o] (data gradient of a strided SparseConv3d)."""
    lib = L.load()
    nbr_t = torch.empty((rb.K, max(n_in_cap, 1)), dtype=torch.int32, device=rb.nbr.device)
    rc = lib.fd_rulebook_transpose(_ptr(rb.nbr), rb.nbr.stride(0), _ptr(rb.n_out_dev), rb.n_out_cap, rb.K,
                                   _ptr(nbr_t), nbr_t.stride(0), n_in_cap, _stream())
    L.check(rc, "fd_rulebook_transpose")
    return nbr_t


class TableView:
    """Minimal Rulebook-like view used to run fd_conv_forward over a transposed table."""

    def __init__(self, nbr, K, n_out_dev, n_out_cap):
        self.nbr, self.K, self.n_out_dev, self.n_out_cap = nbr, K, n_out_dev, n_out_cap
        self.tile_mask = None
        self._pair_num = None
        self.out_coords = None

    @property
    def pair_num(self):
        return Rulebook_pair_num(self)


def Rulebook_pair_num(rb):
    if rb._pair_num is None:
        lib = L.load()
        rb._pair_num = torch.empty((rb.K,), dtype=torch.int32, device=rb.nbr.device)
        rc = lib.fd_rulebook_count_pairs(_ptr(rb.nbr), rb.nbr.stride(0), _ptr(rb.n_out_dev), rb.n_out_cap, rb.K,
                                         _ptr(rb._pair_num), _stream())
        L.check(rc, "fd_rulebook_count_pairs")
    return rb._pair_num


# ------------------------------------------------------------------------------------------ weight gradients
DETERMINISTIC_WGRAD = True      # per-chunk partial tiles + ordered reduce (bit-reproducible); False: fp32 atomics


def _run_wgrad(lib, d, dw, what):
    if DETERMINISTIC_WGRAD:
        need = lib.fd_conv_wgrad_workspace_bytes(C.byref(d))
        ws = _workspace(dw.device, max(int(need), 256))
        L.check(lib.fd_conv_wgrad_det(C.byref(d), _ptr(dw), _ptr(ws), ws.numel(), _stream()), what)
    else:
        L.check(lib.fd_conv_wgrad(C.byref(d), _ptr(dw), _stream()), what)


def _split_ptr(t, like):
    """Pointer of a dense FD_FMT_SPLIT_BF16 copy of `like` (an fp32-typed tensor of the same shape), or None."""
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or t.shape != like.shape or not like.is_contiguous():
        raise RuntimeError("split copy must be a contiguous fp32-typed tensor shaped like its fp32 original")
    return t.data_ptr()


def sparse_conv_wgrad(x, dy, rb, dw, precision="fp32", x_split=None, dy_split=None):
    """dw [K,Cin,Cout] += sum over rulebook pairs of x[i]^T dy[o] (dw must be zeroed by the caller).
    x_split / dy_split: optional split-bf16 copies of x / dy (written by affine_act / bn_backward)."""
    lib = L.load()
    x, xs, cin = _rows(x)
    dy, dys, cout = _rows(dy)
    if dw.dtype != torch.float32 or not dw.is_contiguous() or dw.numel() != rb.K * cin * cout:
        raise RuntimeError("dw must be a contiguous fp32 [K,Cin,Cout] buffer")
    d = L.ConvDesc()
    d.d_in = x.data_ptr(); d.in_stride = xs; d.cin = cin; d.in_format = 0; d.in_ctot = cin
    d.cout = cout; d.K = rb.K
    d.d_out = dy.data_ptr(); d.out_stride = dys; d.out_format = 0; d.out_ctot = cout
    d.d_n_out = rb.n_out_dev.data_ptr() if rb.n_out_dev is not None else None
    d.n_out_cap = rb.n_out_cap
    d.mode = L.GATHER_TABLE
    d.d_nbr = rb.nbr.data_ptr(); d.nbr_stride = rb.nbr.stride(0)
    d.out_map = L.OUTMAP_IDENTITY
    d.precision = L.PRECISIONS[precision]
    d.n_in_cap = n_rows(x)
    d.d_in_split = _split_ptr(x_split, x); d.d_out_split = _split_ptr(dy_split, dy)
    _run_wgrad(lib, d, dw, "fd_conv_wgrad(sparse)")
    return dw


def conv2d_wgrad(x, dy, dw, ksize, stride, padding, transposed=False, precision="fp32", x_split=None, dy_split=None):
    """Weight gradient of conv2d_nhwc / its ConvTranspose2d(k == s) form.  x [B,H,W,Cin], dy [B,Ho,Wo,Cout] (channel
    slices allowed), dw [kh*kw, Cin, Cout] zeroed by the caller."""
    lib = L.load()
    x, xs, cin = _rows(x)
    dy, dys, cout = _rows(dy)
    B, H, W = x.shape[0], x.shape[1], x.shape[2]
    Ho, Wo = dy.shape[1], dy.shape[2]
    kh, kw = ksize
    d = L.ConvDesc()
    d.d_in = x.data_ptr(); d.in_stride = xs; d.cin = cin; d.in_format = 0; d.in_ctot = cin
    d.cout = cout; d.K = kh * kw
    d.d_out = dy.data_ptr(); d.out_stride = dys; d.out_format = 0; d.out_ctot = cout
    d.mode = L.GATHER_CONVT2D if transposed else L.GATHER_CONV2D
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, Ho, Wo
    d.kh, d.kw, d.sh, d.sw, d.ph, d.pw = kh, kw, stride[0], stride[1], padding[0], padding[1]
    d.out_map = L.OUTMAP_IDENTITY
    d.n_out_cap = B * H * W if transposed else B * Ho * Wo
    d.precision = L.PRECISIONS[precision]
    d.d_in_split = _split_ptr(x_split, x); d.d_out_split = _split_ptr(dy_split, dy)
    _run_wgrad(lib, d, dw, "fd_conv_wgrad(conv2d)")
    return dw


# ------------------------------------------------------------------------------------------ data gradients
def conv2d_dgrad(dy, w_t, in_hw, ksize, stride, padding, precision="fp32"):
    """dL/dx [B,H,W,Cin] of y = conv2d(x, w): gather of dy through FD_GATHER_CONV2D_DGRAD with w_t [K, Cout, Cin]."""
    lib = L.load()
    fmt = 0
    if isinstance(dy, ops.Feat):                     # split-bf16 copy of dL/dy (dense rows)
        fmt = 1 if dy.fmt == "split" else 0
        dy = dy.t
    dy, dys, cout = _rows(dy)
    B, Ho, Wo = dy.shape[0], dy.shape[1], dy.shape[2]
    H, W = in_hw
    K, co2, cin = w_t.shape
    if co2 != cout or not w_t.is_contiguous():
        raise RuntimeError("w_t must be contiguous [K, Cout, Cin]")
    dx = torch.empty((B, H, W, cin), dtype=torch.float32, device=dy.device)
    d = L.ConvDesc()
    d.d_in = dy.data_ptr(); d.in_stride = dys; d.cin = cout; d.in_format = fmt; d.in_ctot = cout
    d.d_w = w_t.data_ptr(); d.cout = cin; d.K = K
    p = L.PRECISIONS[precision]
    if p != L.PREC_FP32:
        d.d_w_packed = ops.packed_weights(w_t).data_ptr()
    d.d_out = dx.data_ptr(); d.out_stride = cin; d.out_format = 0; d.out_ctot = cin
    d.mode = L.GATHER_CONV2D_DGRAD
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, Ho, Wo, H, W
    d.kh, d.kw, d.sh, d.sw, d.ph, d.pw = ksize[0], ksize[1], stride[0], stride[1], padding[0], padding[1]
    d.out_map = L.OUTMAP_IDENTITY
    d.n_out_cap = B * H * W
    d.precision = p
    L.check(lib.fd_conv_forward(C.byref(d), _stream()), "fd_conv_forward(conv2d dgrad)")
    return dx


# ------------------------------------------------------------------------------------------ batch norm
class BNSaved:
    __slots__ = ("mean", "invstd", "scale", "shift")


def bn_train_stats(x, bn, n_dev=None, n_cap=None):
    """Batch statistics of x [rows, C] (first n rows), running-stat update with bn.momentum; returns BNSaved."""
    lib = L.load()
    x, xs, Cc = _rows(x)
    n_cap = n_rows(x) if n_cap is None else n_cap
    dev = x.device
    s = BNSaved()
    buf = torch.empty((4, Cc), dtype=torch.float32, device=dev)
    s.mean, s.invstd, s.scale, s.shift = buf[0], buf[1], buf[2], buf[3]
    ws = _workspace(dev, lib.fd_bn_workspace_bytes(Cc))
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    track = bn.track_running_stats and bn.running_mean is not None
    rc = lib.fd_bn_train_stats(_ptr(x), xs, Cc, _ptr(n_dev), n_cap, float(bn.eps), mom,
                               _ptr(bn.weight.detach()) if bn.affine else None,
                               _ptr(bn.bias.detach()) if bn.affine else None,
                               _ptr(bn.running_mean) if track else None, _ptr(bn.running_var) if track else None,
                               _ptr(s.mean), _ptr(s.invstd), _ptr(s.scale), _ptr(s.shift), _ptr(ws), _stream())
    L.check(rc, "fd_bn_train_stats")
    # the running statistics were updated through raw pointers: invalidate eval-mode folded-BN caches keyed on them
    bn.__dict__["_fd_stats_version"] = bn.__dict__.get("_fd_stats_version", 0) + 1
    if track and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return s


def affine_act(x, scale, shift, residual=None, relu=False, out=None, n_dev=None, n_cap=None, split=None):
    """split = (fp32-typed tensor [..., Ctot], c0): also write an FD_FMT_SPLIT_BF16 copy of the result into channels
    [c0, c0 + C) of that buffer (rows of Ctot bf16 hi | Ctot bf16 lo)."""
    lib = L.load()
    x, xs, Cc = _rows(x)
    n_cap = n_rows(x) if n_cap is None else n_cap
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    out, ys, _ = _rows(out)
    rs = 0
    if residual is not None:
        residual, rs, _ = _rows(residual)
    sp, sct = None, 0
    if split is not None:
        st, c0 = split
        if st.dtype != torch.float32 or not st.is_contiguous() or n_rows(st) != n_rows(x):
            raise RuntimeError("split buffer must be a contiguous fp32-typed tensor with the rows of x")
        sct = int(st.shape[-1])
        sp = C.c_void_p(st.data_ptr() + 2 * int(c0))
    rc = lib.fd_affine_act(_ptr(x), xs, Cc, _ptr(scale), _ptr(shift), _ptr(residual), rs, int(bool(relu)), _ptr(out), ys,
                           sp, sct, _ptr(n_dev), n_cap, _stream())
    L.check(rc, "fd_affine_act")
    return out


def bn_backward(dy, y, relu, x, saved, gamma, dgamma, dbeta, want_dres=False, n_dev=None, n_cap=None, want_split=False):
    """-> (dx, dres | None[, dx_split | None]); dgamma / dbeta are written in place.  want_split: also return a dense
    split-bf16 copy of dx (fp32-typed tensor shaped like dx) for the tensor-core data- / weight-gradient convolutions."""
    lib = L.load()
    dy, dys, Cc = _rows(dy)
    x, xs, _ = _rows(x)
    n_cap = n_rows(x) if n_cap is None else n_cap
    ys = 0
    if relu:
        y, ys, _ = _rows(y)
    dx = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    dres = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_dres else None
    dxs = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_split else None
    ws = _workspace(x.device, lib.fd_bn_workspace_bytes(Cc))
    rc = lib.fd_bn_backward(_ptr(dy), dys, _ptr(y) if relu else None, ys, int(bool(relu)), _ptr(x), xs, Cc, _ptr(n_dev),
                            n_cap, _ptr(saved.mean), _ptr(saved.invstd), _ptr(gamma), _ptr(dx), dx.stride(-2), _ptr(dxs),
                            _ptr(dres), dres.stride(-2) if want_dres else 0, _ptr(dgamma), _ptr(dbeta), _ptr(ws),
                            _stream())
    L.check(rc, "fd_bn_backward")
    if want_split:
        return dx, dres, dxs
    return dx, dres


def col_sum(x, out, n_dev=None, n_cap=None):
    lib = L.load()
    x, xs, Cc = _rows(x)
    n_cap = n_rows(x) if n_cap is None else n_cap
    ws = _workspace(x.device, lib.fd_bn_workspace_bytes(Cc))
    L.check(lib.fd_col_sum(_ptr(x), xs, Cc, _ptr(n_dev), n_cap, _ptr(out), _ptr(ws), _stream()), "fd_col_sum")
    return out


def add_rows_(dst, src, n_dev=None, n_cap=None):
    lib = L.load()
    dst, ds, Cc = _rows(dst)
    src, ss, _ = _rows(src)
    n_cap = n_rows(dst) if n_cap is None else n_cap
    L.check(lib.fd_add_rows(_ptr(dst), ds, _ptr(src), ss, Cc, _ptr(n_dev), n_cap, _stream()), "fd_add_rows")
    return dst


# ------------------------------------------------------------------------------------------ BEV scatter / gather
def rows_to_bev(rows, coords, n_dev, n_cap, B, D, H, W):
    lib = L.load()
    rows, rs, Cc = _rows(rows)
    bev = torch.empty((B, H, W, Cc * D), dtype=torch.float32, device=rows.device)
    rc = lib.fd_rows_to_bev(_ptr(rows), rs, Cc, _ptr(coords), _ptr(n_dev), n_cap, B, D, H, W, _ptr(bev), _stream())
    L.check(rc, "fd_rows_to_bev")
    return bev


def bev_to_rows(bev, Cc, coords, n_dev, n_cap, B, D, H, W):
    lib = L.load()
    if not bev.is_contiguous():
        raise RuntimeError("bev gradient must be contiguous [B,H,W,C*D]")
    rows = torch.empty((max(n_cap, 1), Cc), dtype=torch.float32, device=bev.device)
    rc = lib.fd_bev_to_rows(_ptr(bev), Cc, _ptr(coords), _ptr(n_dev), n_cap, B, D, H, W, _ptr(rows), Cc, _stream())
    L.check(rc, "fd_bev_to_rows")
    return rows
