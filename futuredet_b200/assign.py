"""CenterPoint target assignment for a whole batch on the GPU (`fd_assign_center_targets`).

Mirrors det3d/datasets/pipelines/preprocess.py:336-570 (AssignLabel, standard sampler): per forecast timestep and per
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


def assign_targets(annotations, tasks, assigner_cfg, grid_size, pc_range, voxel_size, device):
    """annotations: list over samples of dict(gt_boxes=[T arrays [n,12] fp32], gt_classes=[T arrays [n] int], 1-based
    global class ids in task order); tasks: list of dict(num_class, class_names).  Returns dict(hm, anno_box, ind, mask,
    cat) with ex[key][t][task] CUDA tensors."""
    lib = L.load()
    B = len(annotations)
    T = len(annotations[0]["gt_boxes"])
    osf = int(_get(assigner_cfg, "out_size_factor"))
    max_objs = int(_get(assigner_cfg, "max_objs"))
    W, H = int(grid_size[0]) // osf, int(grid_size[1]) // osf                       # feature_map_size = grid[:2] // osf
    ex = {k: [] for k in ("hm", "anno_box", "ind", "mask", "cat")}
    for t in range(T):
        row = {k: [] for k in ex}
        flag = 0
        for task in tasks:
            names = list(_get(task, "class_names"))
            ncls = len(names)
            per_sample = []
            for a in annotations:
                cls = np.asarray(a["gt_classes"][t])
                boxes = np.asarray(a["gt_boxes"][t], np.float32).reshape(-1, 12)
                sel = [np.where(cls == j + 1 + flag)[0] for j in range(ncls)]          # class by class (:417-441)
                idx = np.concatenate(sel) if sel else np.zeros((0,), np.int64)
                per_sample.append((boxes[idx], (cls[idx] - flag).astype(np.int32)))
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
            for k, v in zip(("hm", "anno_box", "ind", "mask", "cat"), (hm, anno, ind, mask, cat)):
                row[k].append(v)
            flag += ncls
        for k in ex:
            ex[k].append(row[k])
    return ex
