"""Native training step of the VoxelNet/CenterPoint hot path: train-mode forward + hand-written backward.

The reference trains with torch autograd over spconv / cuDNN / ATen kernels (det3d/torchie/trainer/trainer.py:317-344:
`losses = model(example, return_loss=True)`, `loss = sum(losses["loss"])`, `loss.backward()`, optimizer step; DDP
gradient all-reduce in det3d/torchie/apis/train.py:311-317).  Here the same computation is an explicit tape of native
kernels (no autograd graph): every layer of

    SpMiddleResNetFHD (scn.py:83-176)  ->  RPN (rpn.py:124-159)  ->  CenterHead (center_head.py:375-539)

runs `conv (+bias) -> batch statistics -> normalise (+residual) -> ReLU` in training mode (batch statistics, running-stat
update with the module's momentum), records what its backward needs, and `Tape.backward()` replays the layers in
reverse: loss gradient -> ConvT/Conv2d dgrad+wgrad -> BEV gather -> sparse dgrad (transposed rulebooks) + wgrad.
Parameter gradients are written straight into flat bucket buffers (`GradBuckets`) so the data-parallel all-reduce
(`shard.GradSync`) needs no packing copy; `param.grad` are views of those buckets, so any torch optimizer applies.

Arithmetic is fp32 (`precision="fp32"`: exact CUDA-core kernels) or, with `precision="bf16x3"`, fp32-class on the
tensor cores: forward, data-gradient and weight-gradient convolutions with tensor-core friendly shapes run on tcgen05
with a 3-term bf16 split and fp32 accumulation in TMEM.
"""
import numpy as np
import torch
from torch import nn

from . import ops
from . import train_ops as T
from .loss import center_head_loss, center_head_loss_backward
from .neck import conv_weight_kio
from .sparse import SparseConvTensor, SparseSequential, _SparseConvBase


# ------------------------------------------------------------------------------------------------ tape
# Tensor-core arm: the kernels that PRODUCE an activation or a conv-output gradient (fd_affine_act, fd_bn_backward)
# also write an FD_FMT_SPLIT_BF16 copy of it, and the convolutions that CONSUME it (forward, data gradient, weight
# gradient) gather those ready-made bf16 planes with cp.async / TMA tiles instead of converting fp32 rows inside their
# producer warps (the values are the same hi / lo pairs either way: results are bit-identical).  fp32 stays the
# canonical copy for everything else (BatchNorm statistics and backward, residual sums, the loss).
SPLIT_COPIES = True


class Var:
    """An activation with a gradient slot.  `s` / `grad_s`: optional split-bf16 copies of `t` / the gradient (fp32-typed
    tensors of the same shape), see SPLIT_COPIES."""

    def __init__(self, t, needs_grad=True, n_dev=None, n_cap=None):
        self.t, self.needs_grad, self.n_dev, self.n_cap = t, needs_grad, n_dev, n_cap
        self._grad = None
        self.s = None
        self.grad_s = None

    @property
    def grad(self):
        return self._grad

    def accumulate(self, g, gs=None):
        if not self.needs_grad:
            return
        if self._grad is None:
            self._grad, self.grad_s = g, gs
        else:
            T.add_rows_(self._grad, g, self.n_dev, self.n_cap)
            self.grad_s = None               # the split copy no longer matches the sum


class SliceVar:
    """Channel slice [c0, c0+c) of a wider Var's buffer (torch.cat / multi-head outputs written in place)."""

    def __init__(self, parent, c0, c):
        self.parent, self.c0, self.c = parent, c0, c
        self.t = parent.t[..., c0:c0 + c]
        self.needs_grad = True
        self.n_dev = self.n_cap = None
        self.s = self.grad_s = None          # the split copy lives in parent.s (full width)

    @property
    def grad(self):
        g = self.parent.grad
        return None if g is None else g[..., self.c0:self.c0 + self.c]


class Tape:
    def __init__(self):
        self.nodes = []

    def add(self, fn):
        self.nodes.append(fn)

    def backward(self):
        for fn in reversed(self.nodes):
            fn()
        self.nodes = []


class GradBuckets:
    """Flat fp32 gradient storage: one contiguous buffer per bucket, `param.grad` = view.  Parameters are laid out
    in REVERSE registration order (the order their gradients become final during backward), as torch DDP does."""

    def __init__(self, params, bucket_bytes=25 << 20, attach=True):
        self.params = [p for p in params if p.requires_grad]
        self.attach = attach                 # False: leave param.grad to torch autograd (loss.backward() bridge)
        self.buckets, self.where = [], {}
        cur, cur_n = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self._close(cur)
                cur, cur_n = [], 0
        if cur:
            self._close(cur)
        self.pending = [0] * len(self.buckets)
        self.on_bucket_ready = None          # callback(bucket_index, flat_tensor), set by shard.GradSync

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros((n,), dtype=torch.float32, device=plist[0].device)
        off = 0
        for p in plist:
            view = flat[off:off + p.numel()].view(p.shape)
            self.where[id(p)] = (len(self.buckets), view)
            if self.attach:
                p.grad = view
            off += p.numel()
        self.buckets.append((flat, list(plist)))

    def zero(self):
        for i, (flat, plist) in enumerate(self.buckets):
            flat.zero_()
            self.pending[i] = len(plist)
            for p in plist:
                b, view = self.where[id(p)]
                if self.attach and (p.grad is None or p.grad.data_ptr() != view.data_ptr()):
                    p.grad = view

    def has(self, p):
        """False for frozen parameters (requires_grad=False): their gradient kernels are skipped."""
        return id(p) in self.where

    def grad(self, p):
        return self.where[id(p)][1]

    def scratch(self, p):
        """Gradient destination of `p`: its bucket view, or a throw-away buffer when `p` is frozen."""
        return self.where[id(p)][1] if id(p) in self.where else torch.empty_like(p, dtype=torch.float32)

    def done(self, p):
        """The gradient of `p` is final for this step."""
        if id(p) not in self.where:
            return
        b = self.where[id(p)][0]
        self.pending[b] -= 1
        if self.pending[b] == 0 and self.on_bucket_ready is not None:
            self.on_bucket_ready(b, self.buckets[b][0])

    def flush(self):
        """Report every bucket that still has pending parameters (unused parameters: zero gradient)."""
        for b in range(len(self.buckets)):
            if self.pending[b] > 0:
                self.pending[b] = 0
                if self.on_bucket_ready is not None:
                    self.on_bucket_ready(b, self.buckets[b][0])


# ------------------------------------------------------------------------------------------------ layers
def _prec_for(prec, cin, K, t=None):
    """Tensor-core arm only for the shapes / alignments it accepts; everything else is requested as exact fp32."""
    if prec == "fp32" or not ops.tc_supported(cin, K):
        return "fp32"
    if t is not None and (t.stride(-2) % 4 != 0 or t.data_ptr() % 16 != 0):
        return "fp32"
    return prec


def sparse_conv_train(tape, x, conv, xt, grads, prec):
    """y = conv(x) (+ bias) on a SparseConvTensor `xt` whose features are x.t; returns (Var, output SparseConvTensor)."""
    rb = conv.rulebook(xt)
    w = conv.weight_kio()
    K, cin, cout = w.shape
    bias = conv.bias.detach() if conv.bias is not None else None
    xin = x.t[..., :cin] if x.t.shape[-1] != cin else x.t
    p = _prec_for(prec, cin, K, x.t)
    xs = x.s if (p != "fp32" and x.s is not None and x.s.shape[-1] == cin) else None
    xf = ops.Feat(xs, "split", 0, cin) if xs is not None else ops.Feat(x.t, "fp32", 0, cin)
    # SubM layers of a multi-sample step run on pattern-sorted tiles like the inference path (same table for the data
    # gradient: it is its own transpose); bit-identical to the unsorted call
    sort = conv.subm and xt.batch_size >= ops.SORT_MIN_BATCH
    y_t = ops.sparse_conv(xf, w, rb, None, bias, None, False, precision=p, out_fmt="fp32", sort_tiles=sort and p != "fp32")
    y = Var(y_t, True, rb.n_out_dev, rb.n_out_cap)
    if conv.subm:
        out = xt._like(y_t)
    else:
        out = xt._like(y_t, rb.out_coords, rb.out_shape, rb.n_out_dev, rb.n_out_cap)
        out._index = getattr(rb, "out_index", None)
    n_in_dev, n_in_cap = xt.n_dev, xt.n_cap

    def backward():
        gy = y.grad
        if gy is None:
            return
        gys = y.grad_s
        if grads.has(conv.weight):
            T.sparse_conv_wgrad(xin, gy, rb, grads.grad(conv.weight).view(K, cin, cout), precision=prec,
                                x_split=xs, dy_split=gys)
            grads.done(conv.weight)
        if conv.bias is not None and grads.has(conv.bias):
            T.col_sum(gy, grads.grad(conv.bias), rb.n_out_dev, rb.n_out_cap)
            grads.done(conv.bias)
        if x.needs_grad:
            pd = _prec_for(prec, cout, K, gy)
            if conv.subm:       # the SubM table is its own transpose with mirrored offsets
                table, wt = rb, w.flip(0).transpose(1, 2).contiguous()
            else:
                table = T.TableView(T.rulebook_transpose(rb, n_in_cap), K, n_in_dev, n_in_cap)
                wt = w.transpose(1, 2).contiguous()
            gin = ops.Feat(gys, "split", 0, cout) if (gys is not None and pd != "fp32") else gy
            gx = ops.sparse_conv(gin, wt, table, precision=pd, out_fmt="fp32",
                                 sort_tiles=sort and pd != "fp32" and table is rb)
            x.accumulate(gx)

    tape.add(backward)
    return y, out


def bn_train(tape, x, bn, grads, residual=None, relu=True, out=None, prec="fp32"):
    """y = act(batch_norm_train(x) (+ residual)); x.t [rows, C] (rows < n_dev active)."""
    saved = T.bn_train_stats(x.t, bn, x.n_dev, x.n_cap)
    Cc = x.t.shape[-1]
    want_split = SPLIT_COPIES and prec != "fp32" and Cc % 8 == 0 and x.t.is_contiguous()
    split = None
    if want_split and out is None:
        split = (torch.empty(x.t.shape, dtype=torch.float32, device=x.t.device), 0)
    elif want_split and isinstance(out, SliceVar) and out.parent.s is not None:
        split = (out.parent.s, out.c0)
    y_t = T.affine_act(x.t, saved.scale, saved.shift, residual.t if residual is not None else None, relu,
                       out.t if out is not None else None, x.n_dev, x.n_cap, split=split)
    y = out if out is not None else Var(y_t, True, x.n_dev, x.n_cap)
    if split is not None and out is None:
        y.s = split[0]

    def backward():
        gy = y.grad
        if gy is None:
            return
        want_res = residual is not None and residual.needs_grad
        ws = want_split and isinstance(x, Var)
        r = T.bn_backward(gy, y.t, relu, x.t, saved, bn.weight.detach(), grads.scratch(bn.weight),
                          grads.scratch(bn.bias), want_res, x.n_dev, x.n_cap, want_split=ws)
        dx, dres, dxs = r if ws else (r[0], r[1], None)
        grads.done(bn.weight)
        grads.done(bn.bias)
        x.accumulate(dx, dxs)
        if want_res:
            residual.accumulate(dres)

    tape.add(backward)
    return y


DGRAD_AS_CONV = True       # data gradients of stride-1 "same" Conv2d layers through the forward convolution path


def conv2d_train(tape, x, conv, grads, prec, pad=None, out=None):
    """Conv2d / ConvTranspose2d(k == s) (+ bias) on channels-last x.t [B,H,W,Cin]."""
    transposed = isinstance(conv, nn.ConvTranspose2d)
    w = conv_weight_kio(conv)
    K, cin, cout = w.shape
    bias = conv.bias.detach() if conv.bias is not None else None
    padding = tuple(conv.padding) if pad is None else tuple(pad)
    ksize, stride = tuple(conv.kernel_size), tuple(conv.stride)
    B, H, W = x.t.shape[0], x.t.shape[1], x.t.shape[2]
    p = _prec_for(prec, cin, 1 if transposed else K, x.t)
    xs = x.s if (p != "fp32" and getattr(x, "s", None) is not None and x.s.shape[-1] == cin) else None
    y_t = ops.conv2d_nhwc(ops.Feat(xs, "split") if xs is not None else x.t, w, ksize, stride, padding, None, bias, False,
                          out=out.t if out is not None else None, precision=p, transposed=transposed, out_fmt="fp32")
    y = out if out is not None else Var(y_t)

    def backward():
        gy = y.grad
        if gy is None:
            return
        gys = y.grad_s if isinstance(y, Var) else None
        if grads.has(conv.weight):
            gw = torch.zeros((K, cin, cout), dtype=torch.float32, device=w.device)
            T.conv2d_wgrad(x.t, gy, gw, ksize, stride, padding, transposed, precision=prec, x_split=xs, dy_split=gys)
            g4 = gw.view(ksize[0], ksize[1], cin, cout)
            # back to the parameter layout: Conv2d [Cout,Cin,kh,kw], ConvTranspose2d [Cin,Cout,kh,kw]
            grads.grad(conv.weight).copy_(g4.permute(2, 3, 0, 1) if transposed else g4.permute(3, 2, 0, 1))
            grads.done(conv.weight)
        if conv.bias is not None and grads.has(conv.bias):
            T.col_sum(gy, grads.grad(conv.bias))
            grads.done(conv.bias)
        if x.needs_grad:
            wt = w.transpose(1, 2).contiguous()
            pd = _prec_for(prec, cout, K, gy)
            gin = ops.Feat(gys, "split") if (gys is not None and pd != "fp32") else gy
            if transposed:   # each input pixel fed k*k output pixels: a k x k stride-k conv over dL/dy
                gx = ops.conv2d_nhwc(gin, wt, ksize, stride, (0, 0), precision=pd, out_fmt="fp32")
                gx = gx.t if isinstance(gx, ops.Feat) else gx
            elif DGRAD_AS_CONV and stride == (1, 1) and all(2 * p_ == k_ - 1 for p_, k_ in zip(padding, ksize)):
                # a "same" convolution's data gradient is the same convolution over dL/dy with flipped, transposed taps:
                # gx[i] = sum_k gy[i - k + p] W[k]^T = sum_k' gy[i + k' - p] W[K-1-k']^T  -- it then takes the forward kernels'
                # dense fast path (TMA tile loads, tall stages for 3x3) instead of the per-thread gather of the dgrad mode
                gx = ops.conv2d_nhwc(gin, wt.flip(0).contiguous(), ksize, stride, padding, precision=pd, out_fmt="fp32")
                gx = gx.t if isinstance(gx, ops.Feat) else gx
            else:
                gx = T.conv2d_dgrad(gin, wt, (H, W), ksize, stride, padding, precision=pd)
            x.accumulate(gx)

    tape.add(backward)
    return y


# ------------------------------------------------------------------------------------------------ model
class NativeTrainer:
    """Train-mode forward + backward of a futuredet_b200 VoxelNet.  One instance per model / process."""

    def __init__(self, model, precision=None, bucket_bytes=25 << 20, attach_grads=True):
        from . import precision as _precision
        self.model = model
        self.precision = _precision.resolve(precision or getattr(model, "precision", None))
        h = model.bbox_head
        if h.forecast_feature or h.bev_map or h.two_stage or h.wide_head:
            raise NotImplementedError("NativeTrainer: the forecast_feature / bev_map / two_stage / wide_head head variants "
                                      "run forward (inference) only; training covers the standard and dense modes")
        self.grads = GradBuckets(list(model.parameters()), bucket_bytes, attach=attach_grads)
        self.tape = None
        self._loss_ctx = None

    # ---- backbone ------------------------------------------------------------------------------------
    def _sequential(self, tape, seq, x, xt):
        mods = list(seq._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, _SparseConvBase):
                x, xt = sparse_conv_train(tape, x, m, xt, self.grads, self.precision)
                i += 1
                if i < len(mods) and isinstance(mods[i], nn.BatchNorm1d):
                    relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                    x = bn_train(tape, x, mods[i], self.grads, None, relu, prec=self.precision)
                    i += 2 if relu else 1
            elif hasattr(m, "conv1") and hasattr(m, "bn2"):        # SparseBasicBlock (scn.py:64-80)
                if m.downsample is not None:
                    raise NotImplementedError("SparseBasicBlock.downsample is not used by the reference backbone")
                identity = x
                o, xt1 = sparse_conv_train(tape, x, m.conv1, xt, self.grads, self.precision)
                o = bn_train(tape, o, m.bn1, self.grads, None, True, prec=self.precision)
                o, xt = sparse_conv_train(tape, o, m.conv2, xt1, self.grads, self.precision)
                x = bn_train(tape, o, m.bn2, self.grads, identity, True, prec=self.precision)
                i += 1
            else:
                raise NotImplementedError("no native training kernel for %s" % type(m).__name__)
        return x, xt

    def _backbone(self, tape, bb, feats, coors, batch_size, input_shape, n_dev, n_cap):
        sparse_shape = np.array([int(v) for v in input_shape][::-1]) + [1, 0, 0]
        xt = SparseConvTensor(feats, coors, sparse_shape, batch_size, n_dev=n_dev, n_cap=n_cap)
        x = Var(feats, False, xt.n_dev, xt.n_cap)          # the voxel features need no gradient
        for name in ("conv_input", "conv1", "conv2", "conv3", "conv4", "extra_conv"):
            seq = getattr(bb, name)
            assert isinstance(seq, SparseSequential)
            x, xt = self._sequential(tape, seq, x, xt)
        D, H, W = xt.spatial_shape
        Cc = x.t.shape[-1]
        bev_t = T.rows_to_bev(x.t, xt.indices, xt.n_dev, xt.n_cap, batch_size, D, H, W)
        bev = Var(bev_t)
        coords, nd, nc = xt.indices, xt.n_dev, xt.n_cap

        def backward():
            if bev.grad is None:
                return
            x.accumulate(T.bev_to_rows(bev.grad, Cc, coords, nd, nc, batch_size, D, H, W))

        tape.add(backward)
        return bev

    # ---- neck ----------------------------------------------------------------------------------------
    def _neck(self, tape, neck, x):
        B = x.t.shape[0]
        out = None
        col = 0
        for i, block in enumerate(neck.blocks):
            mods = list(block)
            x = conv2d_train(tape, x, mods[1], self.grads, self.precision, pad=(1, 1))     # ZeroPad2d(1) + conv(pad 0)
            x = bn_train(tape, x, mods[2], self.grads, None, True, prec=self.precision)
            for k in range(4, len(mods), 3):
                x = conv2d_train(tape, x, mods[k], self.grads, self.precision)
                x = bn_train(tape, x, mods[k + 1], self.grads, None, True, prec=self.precision)
            j = i - neck._upsample_start_idx
            if j >= 0:
                up, bn = neck.deblocks[j][0], neck.deblocks[j][1]
                u = conv2d_train(tape, x, up, self.grads, self.precision)
                if out is None:
                    Ho, Wo = u.t.shape[1], u.t.shape[2]
                    out = Var(torch.empty((B, Ho, Wo, sum(neck._num_upsample_filters)), dtype=torch.float32,
                                          device=u.t.device))
                    if SPLIT_COPIES and self.precision != "fp32":
                        out.s = torch.empty_like(out.t)      # split copy of the concatenated map (the deblocks fill slices)
                c = neck._num_upsample_filters[j]
                bn_train(tape, u, bn, self.grads, None, True, out=SliceVar(out, col, c), prec=self.precision)
                col += c
        return out if out is not None else x

    # ---- head ----------------------------------------------------------------------------------------
    def _head(self, tape, head, x):
        s = conv2d_train(tape, x, head.shared_conv[0], self.grads, self.precision)
        s = bn_train(tape, s, head.shared_conv[1], self.grads, None, True, prec=self.precision)
        B, H, W = s.t.shape[0], s.t.shape[1], s.t.shape[2]
        preds, outs = [], []
        for task in head.tasks:
            total_c = sum(task.heads[h][0] for h in task.heads)
            out = Var(torch.empty((B, H, W, total_c), dtype=torch.float32, device=s.t.device))
            ret, col = {}, 0
            groups = task._stage_groups()
            for h in task.heads:
                y = s
                steps = groups[h]
                for si, (conv, bnm, relu) in enumerate(steps):
                    last = si == len(steps) - 1
                    c = conv.out_channels
                    if last and bnm is None and not relu:
                        y = conv2d_train(tape, y, conv, self.grads, self.precision, out=SliceVar(out, col, c))
                    else:
                        if last:
                            raise NotImplementedError("SepHead: final conv followed by BN/ReLU")
                        y = conv2d_train(tape, y, conv, self.grads, self.precision)
                        if bnm is not None:
                            y = bn_train(tape, y, bnm, self.grads, None, relu, prec=self.precision)
                        elif relu:
                            raise NotImplementedError("SepHead: conv + ReLU without BatchNorm")
                ret[h] = out.t[..., col:col + c].permute(0, 3, 1, 2)
                col += c
            preds.append(ret)
            outs.append(out)
        return preds, outs

    # ---- step ----------------------------------------------------------------------------------------
    def forward(self, example=None, points=None, batch_offsets=None):
        """Train-mode forward.  Either the collated `example` of det3d (voxels / num_points / coordinates / num_voxels /
        shape + targets) or raw `points`, `batch_offsets` (fused voxelizer) with the targets in `example`.
        Returns the reference's loss dict (per-task lists); call backward() next."""
        m = self.model
        tape = Tape()
        if points is not None:
            vox = m.voxelize(points, batch_offsets)
            B = batch_offsets.numel() - 1
            grid = ops.grid_size_of(m.voxel_cfg["range"], m.voxel_cfg["voxel_size"])
            feats, coors, n_dev, n_cap = vox["features"], vox["coords"], vox["total"], vox["coords"].shape[0]
        else:
            feats = m.reader(example["voxels"], example["num_points"])
            coors = example["coordinates"]
            B = len(example["num_voxels"])
            grid = example["shape"][0]
            n_dev, n_cap = None, None
        bev = self._backbone(tape, m.backbone, feats, coors, B, grid, n_dev, n_cap)
        x = self._neck(tape, m.neck, bev) if m.with_neck else bev
        preds, outs = self._head(tape, m.bbox_head, x)
        losses, ctxs = center_head_loss(m.bbox_head, example, preds, return_ctx=True)
        self.tape, self._loss_ctx = tape, (ctxs, outs)
        self.preds = preds
        self.activations = dict(bev=bev, neck_out=x, head_out=outs)      # Vars (value + gradient after backward)
        return losses

    def backward(self, gscales=None):
        """Backward of sum(losses["loss"]) (trainer.py:85); gscales: optional per-task device scalars."""
        ctxs, outs = self._loss_ctx
        self.grads.zero()
        for t_id, (ctx, out) in enumerate(zip(ctxs, outs)):
            g = torch.zeros_like(out.t)
            if gscales is None or gscales[t_id] is not None:      # None: this task's loss is not part of the objective
                center_head_loss_backward(ctx, out.t, g, None if gscales is None else gscales[t_id])
            out._grad = g
        self.tape.backward()
        self.grads.flush()
        self._last_loss_ctx = ctxs       # keeps the pinned argument tables alive (a captured graph re-reads them on replay)
        self.tape = self._loss_ctx = None

    def step(self, example=None, points=None, batch_offsets=None):
        losses = self.forward(example, points, batch_offsets)
        self.backward()
        return losses


class _LossBridge(torch.autograd.Function):
    """Makes `sum(losses["loss"]).backward()` of the reference trainer (trainer.py:85,317-344) drive the native tape:
    the parameters are the autograd inputs, the per-task losses the outputs, and backward() returns the gradients the
    native backward pass wrote into the buckets (torch then accumulates them into `param.grad` as usual)."""

    @staticmethod
    def forward(ctx, trainer, loss_list, *params):
        ctx.trainer = trainer
        return tuple(l.detach().clone() for l in loss_list)

    @staticmethod
    def backward(ctx, *gs):
        tr = ctx.trainer
        tr.backward([None if g is None else g.detach().reshape(1).float().contiguous() for g in gs])
        return (None, None) + tuple(tr.grads.grad(p) for p in tr.grads.params)


def bridged_losses(trainer, losses):
    """Replace losses["loss"] by autograd-connected scalars (see _LossBridge)."""
    outs = _LossBridge.apply(trainer, list(losses["loss"]), *trainer.grads.params)
    losses["loss"] = list(outs)
    return losses
