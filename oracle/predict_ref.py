"""ORACLE (test infrastructure only): CenterHead.predict (standard mode) restated on CPU torch / numpy.

Follows, step by step:
  det3d/models/bbox_heads/center_head.py:541-718   predict: per-timestep velocity slices (:560-570), NHWC permute,
                                                   sigmoid / exp / atan2 decode (:610-665), merge with label offsets
  det3d/models/bbox_heads/center_head.py:720-770   post_processing: score / range masks, NMS box columns, gather
  det3d/core/bbox/box_torch_ops.py:248-276         rotate_nms_pcdet: pcdet convention, score sort, pre/post caps
  det3d/ops/iou3d_nms/src/iou3d_nms.cpp:90-136     nms_gpu: bit-mask matrix + host greedy sweep
The rotated IoU itself has two checkers:
  * `nms_reference_cuda`  -- the REFERENCE's own CUDA kernel (oracle/_ref/libiou3d_ref.so, compiled from
    det3d/ops/iou3d_nms/src/iou3d_nms_kernel.cu by oracle/Makefile), used on the GPU box;
  * `iou_bev_np`          -- an independent float64 polygon clip (Sutherland-Hodgman), used on the CPU and to
    cross-check both CUDA implementations.
Decode / masks / merge are pinned by tests/golden/predict.pt, produced by the reference `CenterHead.predict` itself with
only `rotate_nms_pcdet`'s CUDA call replaced by `rotate_nms_ref(..., nms_np)` (oracle/gen_golden.py).
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libiou3d_ref.so")


# ------------------------------------------------------------------------------------------------ IoU, float64
def _corners(box):
    x, y, dx, dy, a = float(box[0]), float(box[1]), float(box[3]), float(box[4]), float(box[6])
    c, s = np.cos(a), np.sin(a)
    pts = np.array([[-dx / 2, -dy / 2], [dx / 2, -dy / 2], [dx / 2, dy / 2], [-dx / 2, dy / 2]])
    rot = np.array([[c, -s], [s, c]])
    return pts @ rot.T + np.array([x, y])


def _clip(poly, a, b):
    """Keep the part of `poly` on the left of the directed edge a->b."""
    out = []
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        sp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        sq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
        if sp >= 0:
            out.append(p)
        if (sp >= 0) != (sq >= 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return out


def iou_bev_np(a, b):
    """Rotated BEV IoU of two boxes (x,y,z,dx,dy,dz,heading), float64 polygon clipping."""
    pa, pb = _corners(a), _corners(b)
    poly = [p for p in pa]
    for i in range(4):
        if not poly:
            break
        poly = _clip(poly, pb[i], pb[(i + 1) % 4])
    inter = 0.0
    if len(poly) >= 3:
        P = np.array(poly)
        inter = 0.5 * abs(np.sum(P[:, 0] * np.roll(P[:, 1], -1) - np.roll(P[:, 0], -1) * P[:, 1]))
    sa, sb = float(a[3]) * float(a[4]), float(b[3]) * float(b[4])
    return inter / max(sa + sb - inter, 1e-8)


def _greedy(n, suppresses):
    """iou3d_nms.cpp:116-131: walk the boxes in score order, keep what no kept box suppressed."""
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if not removed[i]:
            keep.append(i)
            removed |= suppresses(i)
    return np.array(keep, np.int64)


def nms_np(boxes, thresh):
    """boxes [n,7] sorted by score -> kept positions (float64 IoU)."""
    b = np.asarray(boxes, np.float64)
    n = len(b)

    def row(i):
        r = np.zeros(n, bool)
        for j in range(i + 1, n):
            # cheap reject: centres further apart than the two half diagonals
            if np.hypot(b[i, 0] - b[j, 0], b[i, 1] - b[j, 1]) > 0.5 * (np.hypot(b[i, 3], b[i, 4]) + np.hypot(b[j, 3], b[j, 4])):
                continue
            r[j] = iou_bev_np(b[i], b[j]) > thresh
        return r
    return _greedy(n, row)


# ------------------------------------------------------------------------------------------------ reference CUDA kernel
_ref = None


def reference_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_LIB):
            raise RuntimeError("oracle/_ref/libiou3d_ref.so missing: run `make -C oracle` where /root/reference exists")
        _ref = C.CDLL(REF_LIB)
        _ref._Z11nmsLauncherPKfPyif.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        _ref._Z19boxesioubevLauncheriPKfiS0_Pf.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return _ref


def nms_reference_cuda(boxes, thresh):
    """nms_gpu of the reference (iou3d_nms.cpp:90-136): its CUDA mask kernel on the current device + the host sweep."""
    lib = reference_lib()
    n = len(boxes)
    if n == 0:
        return np.zeros((0,), np.int64)
    bt = torch.as_tensor(np.asarray(boxes, np.float32)).contiguous().cuda()
    cols = (n + 63) // 64
    mask = torch.zeros((n, cols), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    lib._Z11nmsLauncherPKfPyif(C.c_void_p(bt.data_ptr()), C.c_void_p(mask.data_ptr()), n, C.c_float(thresh))   # default stream
    torch.cuda.synchronize()
    m = mask.cpu().numpy().view(np.uint64)
    remv = np.zeros(cols, np.uint64)
    keep = []
    for i in range(n):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            keep.append(i)
            remv |= m[i]
    return np.array(keep, np.int64)


def iou_reference_cuda(a, b):
    lib = reference_lib()
    at = torch.as_tensor(np.asarray(a, np.float32)).contiguous().cuda()
    bt = torch.as_tensor(np.asarray(b, np.float32)).contiguous().cuda()
    out = torch.zeros((len(a), len(b)), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    lib._Z19boxesioubevLauncheriPKfiS0_Pf(len(a), C.c_void_p(at.data_ptr()), len(b), C.c_void_p(bt.data_ptr()),
                                          C.c_void_p(out.data_ptr()))
    torch.cuda.synchronize()
    return out.cpu()


# ------------------------------------------------------------------------------------------------ predict
def rotate_nms_ref(boxes, scores, thresh, pre_maxsize=None, post_max_size=None, nms_fn=nms_np):
    """box_torch_ops.py:248-276 with the CUDA call replaced by `nms_fn(sorted boxes, thresh) -> kept positions`.
    Score order is made explicit: descending, ties towards the lower index (stable)."""
    boxes = boxes[:, [0, 1, 2, 4, 3, 5, -1]].clone()
    boxes[:, -1] = -boxes[:, -1] - np.pi / 2
    order = torch.sort(scores, dim=0, descending=True, stable=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = nms_fn(boxes.numpy(), thresh) if len(boxes) else np.zeros((0,), np.int64)
    selected = order[torch.from_numpy(keep)]
    if post_max_size is not None:
        selected = selected[:post_max_size]
    return selected


def _get(cfg, key, default=None):
    return cfg.get(key, default) if isinstance(cfg, dict) else getattr(cfg, key, default)


def predict_ref(preds, timesteps, test_cfg, target_timesteps=7, nms_fn=nms_np):
    """CenterHead.predict, standard mode, on CPU tensors.  preds: dict head -> [B,c,H,W] of task 0.
    Returns list over samples of dict(box3d_lidar, scores, label_preds, cells) -- `cells` (flat BEV index of every
    box) is extra, for index-exact comparisons."""
    vels = [preds["vel"][:, 2 * i:2 * i + 2] for i in range(timesteps)]
    if len(vels) == 1:
        vels = target_timesteps * vels
    rng = torch.tensor(_get(test_cfg, "post_center_limit_range"), dtype=torch.float32)
    nms = _get(test_cfg, "nms")
    rets = []
    for vel in vels:
        d = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in preds.items()}
        d["vel"] = vel.permute(0, 2, 3, 1).contiguous()
        hm = torch.sigmoid(d["hm"])
        dim = torch.exp(d["dim"])
        rot = torch.atan2(d["rot"][..., 0:1], d["rot"][..., 1:2])
        B, H, W, ncls = hm.shape
        reg = d["reg"].reshape(B, H * W, 2)
        ys, xs = torch.meshgrid([torch.arange(0, H), torch.arange(0, W)], indexing="ij")
        ys = ys.view(1, H, W).repeat(B, 1, 1).to(hm)
        xs = xs.view(1, H, W).repeat(B, 1, 1).to(hm)
        xs = xs.view(B, -1, 1) + reg[:, :, 0:1]
        ys = ys.view(B, -1, 1) + reg[:, :, 1:2]
        xs = xs * _get(test_cfg, "out_size_factor") * _get(test_cfg, "voxel_size")[0] + _get(test_cfg, "pc_range")[0]
        ys = ys * _get(test_cfg, "out_size_factor") * _get(test_cfg, "voxel_size")[1] + _get(test_cfg, "pc_range")[1]
        box = torch.cat([xs, ys, d["height"].reshape(B, H * W, 1), dim.reshape(B, H * W, 3), d["vel"].reshape(B, H * W, 2),
                         rot.reshape(B, H * W, 1)], dim=2)
        hm = hm.reshape(B, H * W, ncls)
        per_sample = []
        for i in range(B):
            scores, labels = torch.max(hm[i], dim=-1)
            mask = (scores > _get(test_cfg, "score_threshold")) & (box[i][..., :3] >= rng[:3]).all(1) & \
                   (box[i][..., :3] <= rng[3:]).all(1)
            cells = torch.nonzero(mask).flatten()
            bp, sc, lb = box[i][mask], scores[mask], labels[mask]
            sel = rotate_nms_ref(bp[:, [0, 1, 2, 3, 4, 5, -1]].float(), sc.float(), _get(nms, "nms_iou_threshold"),
                                 _get(nms, "nms_pre_max_size"), _get(nms, "nms_post_max_size"), nms_fn)
            per_sample.append(dict(box3d_lidar=bp[sel], scores=sc[sel], label_preds=lb[sel], cells=cells[sel]))
        rets.append(per_sample)
    out = []
    for i in range(len(rets[0])):
        out.append(dict(box3d_lidar=torch.cat([r[i]["box3d_lidar"] for r in rets]),
                        scores=torch.cat([r[i]["scores"] for r in rets]),
                        label_preds=torch.cat([r[i]["label_preds"] + j for j, r in enumerate(rets)]),
                        cells=torch.cat([r[i]["cells"] for r in rets])))
    return out


def synth_preds(B, H, W, timesteps, seed=0, num_cls=1, n_obj=40, background=-4.0):
    """Head tensors with object-like structure: Gaussian score blobs (several neighbouring cells above threshold with
    nearly identical car-sized boxes, so that the NMS has real work), everything else well below threshold."""
    g = torch.Generator().manual_seed(seed)
    hm = torch.full((B, num_cls, H, W), background) + 0.3 * torch.randn((B, num_cls, H, W), generator=g)
    ys, xs = torch.meshgrid([torch.arange(H), torch.arange(W)], indexing="ij")
    rot = torch.randn((B, 2, H, W), generator=g) * 0.05
    for b in range(B):
        for _ in range(n_obj):
            cy, cx = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
            amp = float(torch.rand((1,), generator=g)) * 5.0 + 1.0
            c = int(torch.randint(0, num_cls, (1,), generator=g))
            blob = (amp + 4.0) * torch.exp(-((ys - cy) ** 2 + (xs - cx) ** 2) / 3.0)
            hm[b, c] = torch.maximum(hm[b, c], background + blob)
            ang = float(torch.rand((1,), generator=g)) * 6.283
            y0, y1, x0, x1 = max(cy - 3, 0), min(cy + 4, H), max(cx - 3, 0), min(cx + 4, W)
            rot[b, 0, y0:y1, x0:x1] += np.sin(ang)
            rot[b, 1, y0:y1, x0:x1] += np.cos(ang)
    dim = torch.log(torch.tensor([1.95, 4.6, 1.7])).view(1, 3, 1, 1) + 0.05 * torch.randn((B, 3, H, W), generator=g)
    return dict(reg=torch.rand((B, 2, H, W), generator=g), height=torch.randn((B, 1, H, W), generator=g) * 0.5 - 1.0,
                dim=dim, rot=rot, vel=torch.randn((B, 2 * timesteps, H, W), generator=g), hm=hm)
