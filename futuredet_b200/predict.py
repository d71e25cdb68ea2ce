"""CenterHead.predict (standard mode) on the native decode + NMS kernels.

Mirrors det3d/models/bbox_heads/center_head.py:541-747: in the standard mode only `preds_dicts[0]` is decoded
(:560), once per forecast timestep with `vel[:, 2t:2t+2]` (a single-timestep head is replicated `target_timesteps`
times, :561-570); every copy goes through `post_processing` (score / range masks, rotated NMS with pre/post caps) and
the per-timestep results are concatenated per sample with `label_preds` offset by the timestep index (:693-713).
Here the T copies share one selection + NMS pass (they differ only in the two velocity columns, which the NMS never
sees): one fd_center_predict call per batch, one host read of the kept counts to build the variable-length lists.
"""
import ctypes as C

import torch

from . import lib as L
from .ops import _ptr, _stream

MAX_PRE = 1024


def _cfg(test_cfg, key, default=None):
    if isinstance(test_cfg, dict):
        return test_cfg.get(key, default)
    return getattr(test_cfg, key, default)


def _channels_last_layout(p, names):
    """(base tensor [B,H,W,S], {name: first channel}) when every head tensor is a channel slice of one channels-last
    buffer (as CenterHead.forward returns them); otherwise they are packed into one (copy)."""
    v0 = p[names[0]]
    B, _, H, W = v0.shape
    S = v0.stride(3)
    same = all(v.is_cuda and v.dtype == torch.float32 and v.shape[0] == B and v.shape[2:] == (H, W) and
               v.stride() == (H * W * S, 1, W * S, S) and
               v.untyped_storage().data_ptr() == v0.untyped_storage().data_ptr() for v in (p[n] for n in names))
    if same:
        lo = min(p[n].storage_offset() for n in names)
        lo -= lo % S
        base = v0.as_strided((B, H, W, S), (H * W * S, W * S, S, 1), lo)
        return base, {n: p[n].storage_offset() - lo for n in names}
    packed = torch.cat([p[n].permute(0, 2, 3, 1) for n in names], dim=-1).contiguous().float()
    offs, col = {}, 0
    for n in names:
        offs[n] = col
        col += p[n].shape[1]
    return packed, offs


def center_head_predict(head, example, preds_dicts, test_cfg):
    if getattr(head, "dense", False):
        # dense mode (center_head.py:606-607,693-713): every task is an independent single-timestep head; each is
        # decoded + NMS-ed on its own and the per-sample results are concatenated with label_preds offset by the
        # number of classes of the preceding tasks
        parts = [_predict_task(head, example, p, test_cfg, [0]) for p in preds_dicts]
        out, flag = [], 0
        flags = []
        for nc in head.num_classes:
            flags.append(flag)
            flag += nc
        for i in range(len(parts[0])):
            out.append({"box3d_lidar": torch.cat([pt[i]["box3d_lidar"] for pt in parts]),
                        "scores": torch.cat([pt[i]["scores"] for pt in parts]),
                        "label_preds": torch.cat([pt[i]["label_preds"] + f for pt, f in zip(parts, flags)]),
                        "metadata": parts[0][i]["metadata"], "cells": torch.cat([pt[i]["cells"] for pt in parts])})
        return out
    if not getattr(head, "standard", True):
        raise NotImplementedError("CenterHead.predict: standard and dense modes are implemented")
    T_head = head.timesteps
    vel_i = list(range(T_head))
    if len(vel_i) == 1:
        vel_i = vel_i * head.target_timesteps                          # :566-567
    return _predict_task(head, example, preds_dicts[0], test_cfg, vel_i)                  # center_head.py:560


def _predict_task(head, example, p, test_cfg, vel_i):
    lib = L.load()
    if _cfg(test_cfg, "circular_nms", False) or _cfg(test_cfg, "per_class_nms", False):
        raise NotImplementedError("CenterHead.predict: circular / per-class NMS are not used by the shipped configs")
    names = ["reg", "height", "dim", "rot", "vel", "hm"]
    if not all(n in p for n in names):
        raise NotImplementedError("CenterHead.predict: needs the reg/height/dim/rot/vel/hm heads of the shipped configs")
    base, off = _channels_last_layout(p, names)
    B, H, W, S = base.shape
    vel_c = [off["vel"] + 2 * i for i in vel_i]
    T = len(vel_c)
    nms = _cfg(test_cfg, "nms")
    pre = int(_cfg(nms, "nms_pre_max_size"))
    post = int(_cfg(nms, "nms_post_max_size"))
    if pre > MAX_PRE:
        raise NotImplementedError("nms_pre_max_size > %d" % MAX_PRE)
    rng = [float(v) for v in _cfg(test_cfg, "post_center_limit_range")]
    vs, pc = _cfg(test_cfg, "voxel_size"), _cfg(test_cfg, "pc_range")
    dev = base.device
    boxes = torch.empty((B, T, post, 9), dtype=torch.float32, device=dev)
    scores = torch.empty((B, T, post), dtype=torch.float32, device=dev)
    labels = torch.empty((B, T, post), dtype=torch.int32, device=dev)
    cells = torch.empty((B, post), dtype=torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = torch.empty((lib.fd_center_predict_workspace_bytes(B, H, W),), dtype=torch.uint8, device=dev)
    rc = lib.fd_center_predict(_ptr(base), S, off["reg"], off["height"], off["dim"], off["rot"],
                               (C.c_int32 * T)(*vel_c), T, off["hm"], p["hm"].shape[1], B, H, W,
                               float(_cfg(test_cfg, "score_threshold")), L.f32(rng),
                               float(_cfg(test_cfg, "out_size_factor")), float(vs[0]), float(vs[1]), float(pc[0]),
                               float(pc[1]), float(_cfg(nms, "nms_iou_threshold")), pre, post, _ptr(boxes),
                               _ptr(scores), _ptr(labels), _ptr(cells), _ptr(count), _ptr(ws), _stream())
    L.check(rc, "fd_center_predict")
    counts = count.tolist()                                            # the one host read (variable-length results)
    metas = example.get("metadata") if isinstance(example, dict) else None
    if not metas:
        metas = [None] * B
    flag = torch.arange(T, device=dev, dtype=torch.int64).view(T, 1)   # label offset per timestep (:700-705)
    ret_list = []
    for i in range(B):
        n = counts[i]
        ret_list.append({"box3d_lidar": boxes[i, :, :n].reshape(T * n, 9), "scores": scores[i, :, :n].reshape(T * n),
                         "label_preds": (labels[i, :, :n].long() + flag).reshape(T * n), "metadata": metas[i],
                         "cells": cells[i, :n]})
    return ret_list


def boxes_iou_bev(boxes_a, boxes_b):
    """Rotated BEV IoU [na, nb] of boxes [n,7] (x,y,z,dx,dy,dz,heading), as det3d.ops.iou3d_nms.boxes_iou_bev_gpu."""
    lib = L.load()
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    L.check(lib.fd_boxes_iou_bev(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()), "fd_boxes_iou_bev")
    return out
