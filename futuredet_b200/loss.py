"""CenterHead.loss (standard branch) on the native loss kernels, forward and backward.

Mirrors det3d/models/bbox_heads/center_head.py:396-539 for the standard mode used by the n0 / n3 configs:
per task  hm_loss = FastFocalLoss(sigmoid-clamped hm, targets of timestep 0),  box_loss_t = RegLoss(cat(reg, height,
dim, vel[2t:2t+2], rot), mask/ind of timestep 0, anno_box_t[..., [0..7,-2,-1]]),  loc_loss_t = sum(box_loss_t *
code_weights (t = 0) | code_weights_forecast (t > 0)),  loss = hm_loss + weight * sum_t loc_loss_t.
Returns the reference's dict of per-task lists.  `center_head_loss_backward` is the hand-written gradient of the same
expression w.r.t. every head tensor (what autograd computes under the reference's `loss.backward()`).
"""
import ctypes as C
from collections import defaultdict

import torch

from . import lib as L
from .ops import _ptr, _stream

TGT_SEL = [0, 1, 2, 3, 4, 5, 6, 7, -2, -1]       # drop the velocity targets (center_head.py:468)


def _table(values, dtype, dev, keep):
    """Small host table -> device through pinned memory with an asynchronous copy (a memcpy node when the step is
    being captured into a CUDA graph; a pageable copy would synchronise and cannot be captured).  `keep` holds the
    pinned source alive for as long as the copy (or the graph) may read it."""
    host = torch.tensor(values, dtype=dtype).pin_memory()
    keep.append(host)
    return host.to(dev, non_blocking=True)


def _plane(v, c):
    """channel c of a logical [B,C,H,W] view -> (address, batch stride, spatial stride); needs y*W+x addressing."""
    B, Cc, H, W = v.shape
    if v.stride(2) != W * v.stride(3):
        raise RuntimeError("prediction tensors must have dense rows (stride_h == W * stride_w)")
    return v.data_ptr() + 4 * c * v.stride(1), v.stride(0), v.stride(3)


class _TaskArgs:
    """Device-side argument tables of one task (shared by the forward and backward kernels)."""

    def __init__(self, head, example, p, task_id):
        hm = p["hm"]
        dev = hm.device
        self.dev = dev
        self.B, self.Cc, self.H, self.W = hm.shape
        if not all(k in p for k in ("reg", "height", "dim", "vel", "rot")) or "rvel" in p:
            raise NotImplementedError("CenterHead.loss: only the vel+rot box encoding of the shipped configs is implemented")
        self.hm = hm
        if getattr(head, "dense", False):
            # dense mode (center_head.py:413-415,441-443,485-487,506-507): task i is the single-timestep head of
            # forecast timestep i, supervised by that timestep's targets of the one class group -> the same
            # expression with T = 1 and example[...][task_id][0]
            self.T = T = 1
            ts, g = [task_id], 0
        else:
            self.T = T = head.timesteps
            ts, g = list(range(T)), task_id
        self.hm_t = example["hm"][ts[0]][g].to(dev, torch.float32).contiguous()
        self.ind = example["ind"][ts[0]][g].to(dev, torch.int64).contiguous()
        self.mask = example["mask"][ts[0]][g].to(dev, torch.uint8).contiguous()
        self.cat = example["cat"][ts[0]][g].to(dev, torch.int64).contiguous()
        self.masks_t = [example["mask"][i][g].to(dev, torch.uint8).contiguous() for i in ts]
        self.tgts = [example["anno_box"][i][g].to(dev, torch.float32).contiguous() for i in ts]
        self.M = self.ind.shape[1]
        self.tgt_dim = self.tgts[0].shape[-1]
        self.keep = []
        self.sel = _table([s % self.tgt_dim for s in TGT_SEL], torch.int32, dev, self.keep)
        self.NC = len(TGT_SEL)
        self.planes = []            # (tensor, channel) per (t, c)
        ptrs, sbs, ssps = [], [], []
        for t in range(T):
            planes = ([(p["reg"], 0), (p["reg"], 1), (p["height"], 0), (p["dim"], 0), (p["dim"], 1), (p["dim"], 2),
                       (p["vel"], 2 * t), (p["vel"], 2 * t + 1), (p["rot"], 0), (p["rot"], 1)])
            for v, c in planes:
                a, sb, ssp = _plane(v, c)
                ptrs.append(a); sbs.append(sb); ssps.append(ssp)
            self.planes += planes
        i64 = lambda xs: _table(xs, torch.int64, dev, self.keep)
        self.ptrs = ptrs
        self.d_ptr, self.d_sb, self.d_ssp = i64(ptrs), i64(sbs), i64(ssps)
        self.d_tgt = i64([t_.data_ptr() for t_ in self.tgts])
        self.d_mask_t = i64([m.data_ptr() for m in self.masks_t])
        self.cw = _table([float(x) for x in head.code_weights], torch.float32, dev, self.keep)
        self.cwf = _table([float(x) for x in head.code_weights_forecast], torch.float32, dev, self.keep)
        self.weight = float(head.weight)
        self.hm_addr, self.hm_sb, self.hm_ssp = _plane(hm, 0)


def center_head_loss(head, example, preds_dicts, return_ctx=False):
    lib = L.load()
    if not (getattr(head, "standard", True) or getattr(head, "dense", False)) or getattr(head, "two_stage", False):
        raise NotImplementedError("CenterHead.loss: standard and dense modes are implemented (the modes of the shipped "
                                  "configs); reverse / sparse / classify / wide_head / two_stage are not")
    dense = getattr(head, "dense", False)
    rets, ctxs = [], []
    for task_id, p in enumerate(preds_dicts):
        a = _TaskArgs(head, example, p, task_id)
        T, NC = a.T, a.NC
        out = torch.empty((3 + T + T * NC,), dtype=torch.float32, device=a.dev)
        ws = torch.empty((lib.fd_center_loss_workspace_bytes(),), dtype=torch.uint8, device=a.dev)
        rc = lib.fd_center_head_loss(C.c_void_p(a.hm_addr), a.hm_sb, a.hm.stride(1), a.hm_ssp, _ptr(a.hm_t), a.B, a.Cc,
                                     a.H, a.W, _ptr(a.ind), _ptr(a.mask), _ptr(a.cat), _ptr(a.d_mask_t), a.M, T, NC,
                                     _ptr(a.d_ptr), _ptr(a.d_sb), _ptr(a.d_ssp), _ptr(a.d_tgt), a.tgt_dim, _ptr(a.sel),
                                     _ptr(a.cw), _ptr(a.cwf), a.weight, _ptr(out), _ptr(ws), _stream())
        L.check(rc, "fd_center_head_loss")
        elem = out[3 + T:].view(T, NC)
        le = [elem[t].detach().cpu() if not return_ctx else elem[t] for t in range(T)]
        rets.append({"loss": out[0], "hm_loss": out[1].detach().cpu() if not return_ctx else out[1],
                     "loc_loss": [out[3 + t] for t in range(T)],
                     "loc_loss_elem": le[0] if dense else le,          # dense: one tensor, not a list (:530-532)
                     "num_positive": out[2]})
        ctxs.append(a)
    merged = defaultdict(list)
    for r in rets:
        for k, v in r.items():
            merged[k].append(v)
    return (merged, ctxs) if return_ctx else merged


def center_head_loss_backward(ctx, out_base, gout_base, gscale=None):
    """Gradient of one task's loss w.r.t. its head tensors.  `ctx`: the _TaskArgs of the forward call (its `hm` now holds
    the clamped probabilities); every head tensor must be a view of `out_base`; the gradients are written into the
    same positions of `gout_base` (same shape/strides as out_base, zero-filled by the caller)."""
    lib = L.load()
    a = ctx
    delta = gout_base.data_ptr() - out_base.data_ptr()
    lo, hi = out_base.data_ptr(), out_base.data_ptr() + out_base.numel() * 4
    for ptr in a.ptrs + [a.hm_addr]:
        if not (lo <= ptr < hi):
            raise RuntimeError("center_head_loss_backward: head tensors must be views of the given output buffer")
    d_gptr = _table([q + delta for q in a.ptrs], torch.int64, a.dev, a.keep)
    rc = lib.fd_center_head_loss_backward(C.c_void_p(a.hm_addr), C.c_void_p(a.hm_addr + delta), a.hm_sb, a.hm.stride(1),
                                          a.hm_ssp, _ptr(a.hm_t), a.B, a.Cc, a.H, a.W, _ptr(a.ind), _ptr(a.mask),
                                          _ptr(a.cat), a.M, a.T, a.NC, _ptr(a.d_ptr), _ptr(d_gptr), _ptr(a.d_sb),
                                          _ptr(a.d_ssp), _ptr(a.d_tgt), a.tgt_dim, _ptr(a.sel), _ptr(a.cw), _ptr(a.cwf),
                                          a.weight, _ptr(gscale), _stream())
    L.check(rc, "fd_center_head_loss_backward")
    return gout_base
