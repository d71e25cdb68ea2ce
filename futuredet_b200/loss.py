"""CenterHead.loss (standard branch) on the native loss kernels.

Mirrors det3d/models/bbox_heads/center_head.py:396-539 for the standard mode used by the n0 / n3 configs:
per task  hm_loss = FastFocalLoss(sigmoid-clamped hm, targets of timestep 0),  box_loss_t = RegLoss(cat(reg, height,
dim, vel[2t:2t+2], rot), mask/ind of timestep 0, anno_box_t[..., [0..7,-2,-1]]),  loc_loss_t = sum(box_loss_t *
code_weights (t = 0) | code_weights_forecast (t > 0)),  loss = hm_loss + weight * sum_t loc_loss_t.
Returns the reference's dict of per-task lists.  Forward only in this round (no autograd graph).
"""
import ctypes as C
from collections import defaultdict

import torch

from . import lib as L
from .ops import _ptr, _stream

TGT_SEL = [0, 1, 2, 3, 4, 5, 6, 7, -2, -1]       # drop the velocity targets (center_head.py:468)


def _plane(v, c):
    """channel c of a logical [B,C,H,W] view -> (address, batch stride, spatial stride); needs y*W+x addressing."""
    B, Cc, H, W = v.shape
    if v.stride(2) != W * v.stride(3):
        raise RuntimeError("prediction tensors must have dense rows (stride_h == W * stride_w)")
    return v.data_ptr() + 4 * c * v.stride(1), v.stride(0), v.stride(3)


def center_head_loss(head, example, preds_dicts):
    lib = L.load()
    rets = []
    for task_id, p in enumerate(preds_dicts):
        hm = p["hm"]
        dev = hm.device
        B, Cc, H, W = hm.shape
        T = head.timesteps
        if not all(k in p for k in ("reg", "height", "dim", "vel", "rot")):
            raise NotImplementedError("CenterHead.loss: only the vel+rot box encoding of the n0/n3 configs is implemented")
        hm_t = example["hm"][0][task_id].to(dev, torch.float32).contiguous()
        ind = example["ind"][0][task_id].to(dev, torch.int64).contiguous()
        mask = example["mask"][0][task_id].to(dev, torch.uint8).contiguous()
        cat = example["cat"][0][task_id].to(dev, torch.int64).contiguous()
        masks_t = [example["mask"][i][task_id].to(dev, torch.uint8).contiguous() for i in range(T)]
        tgts = [example["anno_box"][i][task_id].to(dev, torch.float32).contiguous() for i in range(T)]
        M = ind.shape[1]
        tgt_dim = tgts[0].shape[-1]
        sel = torch.tensor([s % tgt_dim for s in TGT_SEL], dtype=torch.int32, device=dev)
        NC = len(TGT_SEL)
        ptrs, sbs, ssps = [], [], []
        for t in range(T):
            planes = ([(p["reg"], 0), (p["reg"], 1), (p["height"], 0), (p["dim"], 0), (p["dim"], 1), (p["dim"], 2),
                       (p["vel"], 2 * t), (p["vel"], 2 * t + 1), (p["rot"], 0), (p["rot"], 1)])
            for v, c in planes:
                a, sb, ssp = _plane(v, c)
                ptrs.append(a); sbs.append(sb); ssps.append(ssp)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        d_ptr, d_sb, d_ssp = i64(ptrs), i64(sbs), i64(ssps)
        d_tgt = i64([t_.data_ptr() for t_ in tgts])
        d_mask_t = i64([m.data_ptr() for m in masks_t])
        cw = torch.tensor(head.code_weights, dtype=torch.float32, device=dev)
        cwf = torch.tensor([float(x) for x in head.code_weights_forecast], dtype=torch.float32, device=dev)
        out = torch.empty((3 + T + T * NC,), dtype=torch.float32, device=dev)
        ws = torch.empty((lib.fd_center_loss_workspace_bytes(),), dtype=torch.uint8, device=dev)
        hm_addr, hm_sb, hm_ssp = _plane(hm, 0)
        rc = lib.fd_center_head_loss(C.c_void_p(hm_addr), hm_sb, hm.stride(1), hm_ssp, _ptr(hm_t), B, Cc, H, W, _ptr(ind),
                                     _ptr(mask), _ptr(cat), _ptr(d_mask_t), M, T, NC, _ptr(d_ptr), _ptr(d_sb), _ptr(d_ssp),
                                     _ptr(d_tgt), tgt_dim, _ptr(sel), _ptr(cw), _ptr(cwf), float(head.weight), _ptr(out),
                                     _ptr(ws), _stream())
        L.check(rc, "fd_center_head_loss")
        elem = out[3 + T:].view(T, NC)
        rets.append({"loss": out[0], "hm_loss": out[1].detach().cpu(), "loc_loss": [out[3 + t] for t in range(T)],
                     "loc_loss_elem": [elem[t].detach().cpu() for t in range(T)], "num_positive": out[2]})
    merged = defaultdict(list)
    for r in rets:
        for k, v in r.items():
            merged[k].append(v)
    return merged
