"""CenterPoint target assignment for a whole batch on the GPU (`fd_assign_center_targets`).

Mirrors det3d/datasets/pipelines/preprocess.py:336-570 (AssignLabel, standard sampler) and :573-897 (the extra
`*_trajectory` / `*_forecast` targets of the trajectory sampler): per forecast timestep and per
task the objects of the task's classes (grouped class by class, :417-447), rot / rrot wrapped to [-pi, pi) (:449-456),
then heat map / anno_box / ind / mask / cat (:464-546).  The reference does this per sample in DataLoader workers with
Python loops over objects and NumPy Gaussian patches; here the host only groups the annotation arrays and one launch per
(timestep, task) covers the batch.  The result is the `example` target dict in the collate layout
(det3d/torchie/parallel/collate.py:208-232): key -> [timestep][task] -> Tensor[B, ...].
"""
import numpy as np
import torch

from . import lib as L
from .ops import _ptr, _stream


def _get(cfg, key, default=None):
    return cfg.get(key, default) if isinstance(cfg, dict) else getattr(cfg, key, default)


TRAJECTORY_CLASS = {"static": 1, "linear": 2, "nonlinear": 3}        # preprocess.py:372-376


def _run_task(lib, per_sample, ncls, t, assigner_cfg, W, H, pc_range, voxel_size, osf, max_objs, device):
    """One (timestep, task): per_sample = [(boxes [n,12] grouped class by class, 1-based local classes [n])] -> tensors."""
    B = len(per_sample)
    n_max = max(1, max(len(b) for b, _ in per_sample))
    hb = np.zeros((B, n_max, 12), np.float32)
    hc = np.zeros((B, n_max), np.int32)
    hn = np.zeros((B,), np.int32)
    for i, (b, c) in enumerate(per_sample):
        hb[i, :len(b)], hc[i, :len(c)], hn[i] = b, c, len(b)
    db, dc, dn = (torch.from_numpy(x).to(device) for x in (hb, hc, hn))
    hm = torch.empty((B, ncls, H, W), dtype=torch.float32, device=device)
    anno = torch.empty((B, max_objs, 14), dtype=torch.float32, device=device)
    ind = torch.empty((B, max_objs), dtype=torch.int64, device=device)
    mask = torch.empty((B, max_objs), dtype=torch.uint8, device=device)
    cat = torch.empty((B, max_objs), dtype=torch.int64, device=device)
    rc = lib.fd_assign_center_targets(_ptr(db), _ptr(dc), _ptr(dn), B, n_max, 12, ncls, W, H, float(pc_range[0]),
                                      float(pc_range[1]), float(voxel_size[0]), float(voxel_size[1]), float(osf),
                                      float(_get(assigner_cfg, "gaussian_overlap")),
                                      int(_get(assigner_cfg, "min_radius")),
                                      int(bool(_get(assigner_cfg, "radius_mult", False))), t, max_objs, _ptr(hm),
                                      _ptr(anno), _ptr(ind), _ptr(mask), _ptr(cat), _stream())
    L.check(rc, "fd_assign_center_targets")
    return dict(hm=hm, anno_box=anno, ind=ind, mask=mask, cat=cat)


def _group(boxes, cls, ncls, flag=0):
    """Objects of classes flag+1 .. flag+ncls, class by class (preprocess.py:417-441), local 1-based class ids."""
    cls = np.asarray(cls)
    boxes = np.asarray(boxes, np.float32).reshape(-1, 12)
    sel = [np.where(cls == j + 1 + flag)[0] for j in range(ncls)]
    idx = np.concatenate(sel) if sel else np.zeros((0,), np.int64)
    return boxes[idx], (cls[idx] - flag).astype(np.int32)


def assign_targets(annotations, tasks, assigner_cfg, grid_size, pc_range, voxel_size, device):
    """annotations: list over samples of dict(gt_boxes=[T arrays [n,12] fp32], gt_classes=[T arrays [n] int], 1-based
    global class ids in task order; with sampler_type != "standard" also gt_trajectory=[T arrays [n] of "static" /
    "linear" / "nonlinear"]); tasks: list of dict(num_class, class_names).  Returns dict(hm, anno_box, ind, mask, cat)
    with ex[key][t][task] CUDA tensors; the trajectory / forecast samplers of the n3dtf / n3dtfm configs
    (preprocess.py:573-897) add `*_trajectory` (one task, 3 classes: static / linear / nonlinear) and `*_forecast` (one
    task, 7 classes: the boxes of all timesteps, class = timestep) -- the same per-task routine on regrouped classes."""
    lib = L.load()
    T = len(annotations[0]["gt_boxes"])
    osf = int(_get(assigner_cfg, "out_size_factor"))
    max_objs = int(_get(assigner_cfg, "max_objs"))
    W, H = int(grid_size[0]) // osf, int(grid_size[1]) // osf                       # feature_map_size = grid[:2] // osf
    common = (assigner_cfg, W, H, pc_range, voxel_size, osf, max_objs, device)
    keys = ("hm", "anno_box", "ind", "mask", "cat")
    ex = {k: [] for k in keys}
    for t in range(T):
        row = {k: [] for k in keys}
        flag = 0
        for task in tasks:
            ncls = len(list(_get(task, "class_names")))
            per_sample = [_group(a["gt_boxes"][t], a["gt_classes"][t], ncls, flag) for a in annotations]
            out = _run_task(lib, per_sample, ncls, t, *common)
            for k in keys:
                row[k].append(out[k])
            flag += ncls
        for k in keys:
            ex[k].append(row[k])
    if _get(assigner_cfg, "sampler_type", "standard") == "standard":
        return ex
    if len(tasks) != 1:
        raise NotImplementedError("the trajectory / forecast samplers index their single class group by task id "
                                  "(preprocess.py:628,817): one task only, as in the shipped configs")
    for suffix in ("_trajectory", "_forecast"):
        for k in keys:
            ex[k + suffix] = []
    # forecast task: identical object list for every timestep (only radius_mult's (1 + t) differs)
    fore = []
    for a in annotations:
        boxes = np.concatenate([np.asarray(b, np.float32).reshape(-1, 12) for b in a["gt_boxes"]])
        cls = np.concatenate([np.full(len(b), i + 1, np.int32) for i, b in enumerate(a["gt_boxes"])])
        fore.append(_group(boxes, cls, 7))
    for t in range(T):
        traj = [_group(a["gt_boxes"][t], np.array([TRAJECTORY_CLASS[str(s)] for s in a["gt_trajectory"][t]], np.int32), 3)
                for a in annotations]
        for suffix, per_sample, ncls in (("_trajectory", traj, 3), ("_forecast", fore, 7)):
            out = _run_task(lib, per_sample, ncls, t, *common)
            for k in keys:
                ex[k + suffix].append([out[k]])
    return ex
