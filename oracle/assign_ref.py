"""ORACLE (test infrastructure only): CenterPoint target assignment restated in numpy, one task / one timestep.

Follows det3d/datasets/pipelines/preprocess.py:449-456 (limit_period on rot / rrot), :464-546 (radius, centre cell,
Gaussian splat, ind / mask / cat, anno_box) and det3d/core/utils/center_utils.py:17-64 (gaussian_radius, gaussian2D,
draw_umich_gaussian), keeping numpy's dtypes (float32 scalars through the radius / centre arithmetic, float64 Gaussian).
Pinned by tests/golden/assign.npz, produced by the reference `AssignLabel.__call__` itself (oracle/gen_golden.py assign)."""
import numpy as np


def gaussian_radius(det_size, min_overlap=0.5):
    height, width = det_size
    a1 = 1
    b1 = (height + width)
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * a1 * c1)) / 2
    a2 = 4
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 4 * a2 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def draw_gaussian(heatmap, center, radius):
    d = 2 * radius + 1
    y, x = np.ogrid[-radius:radius + 1, -radius:radius + 1]
    sigma = d / 6
    g = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    g[g < np.finfo(g.dtype).eps * g.max()] = 0
    cx, cy = int(center[0]), int(center[1])
    H, W = heatmap.shape
    left, right = min(cx, radius), min(W - cx, radius + 1)
    top, bottom = min(cy, radius), min(H - cy, radius + 1)
    mh = heatmap[cy - top:cy + bottom, cx - left:cx + right]
    mg = g[radius - top:radius + bottom, radius - left:radius + right]
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        np.maximum(mh, mg, out=mh)


def limit_period(val, offset=0.5, period=np.pi * 2):
    return val - np.floor(val / period + offset) * period


def assign_ref(boxes, classes, num_cls, fm_size, pc_range, voxel_size, osf, overlap, min_radius, max_objs,
               radius_mult=False, timestep=0):
    """boxes [n,12] float32 of one task (grouped by class), classes [n] 1-based -> hm, anno_box, ind, mask, cat."""
    boxes = np.array(boxes, np.float32)
    boxes[:, -1] = limit_period(boxes[:, -1])
    boxes[:, -2] = limit_period(boxes[:, -2])
    pc_range, voxel_size = np.asarray(pc_range, np.float32), np.asarray(voxel_size, np.float32)
    W, H = fm_size
    hm = np.zeros((num_cls, H, W), np.float32)
    anno = np.zeros((max_objs, 14), np.float32)
    ind, mask, cat = np.zeros(max_objs, np.int64), np.zeros(max_objs, np.uint8), np.zeros(max_objs, np.int64)
    for k in range(min(len(boxes), max_objs)):
        b = boxes[k]
        cls_id = int(classes[k]) - 1
        w, l = b[3] / voxel_size[0] / osf, b[4] / voxel_size[1] / osf
        if not (w > 0 and l > 0):
            continue
        mult = min(max(1, np.linalg.norm(b[6:8]) * (1 + timestep) / 2), 4) if radius_mult else 1.0
        radius = max(min_radius, int(mult * gaussian_radius((l, w), min_overlap=overlap)))
        ct = np.array([(b[0] - pc_range[0]) / voxel_size[0] / osf, (b[1] - pc_range[1]) / voxel_size[1] / osf], np.float32)
        ci = ct.astype(np.int32)
        if not (0 <= ci[0] < W and 0 <= ci[1] < H):
            continue
        draw_gaussian(hm[cls_id], ct, radius)
        cat[k], ind[k], mask[k] = cls_id, ci[1] * W + ci[0], 1
        anno[k] = np.concatenate((ct - (ci[0], ci[1]), b[2], np.log(b[3:6]), b[6], b[7], b[8], b[9], np.sin(b[10]),
                                  np.cos(b[10]), np.sin(b[11]), np.cos(b[11])), axis=None)
    return hm, anno, ind, mask, cat


TRAJECTORY_CLASS = {"static": 1, "linear": 2, "nonlinear": 3}        # preprocess.py:372-376 (per class name)


def synth_trajectories(seed, n):
    rng = np.random.default_rng(1000 + seed)
    return np.array(["static", "linear", "nonlinear"])[rng.integers(0, 3, n)]


def regroup(boxes, classes, num_cls):
    """Class-by-class grouping of one task (preprocess.py:417-441 and its trajectory / forecast copies)."""
    idx = np.concatenate([np.where(np.asarray(classes) == j + 1)[0] for j in range(num_cls)])
    return np.asarray(boxes, np.float32)[idx], np.asarray(classes)[idx]


def trajectory_task(boxes_t, trajectories):
    """`*_trajectory` targets of timestep t (:573-760): one task, classes static / linear / nonlinear."""
    cls = np.array([TRAJECTORY_CLASS[t] for t in trajectories], np.int32)
    return regroup(boxes_t, cls, 3)


def forecast_task(boxes_all):
    """`*_forecast` targets (:762-895): the boxes of ALL timesteps in one 7-class task, class = timestep + 1; the same
    list is used for every timestep (only radius_mult's (1 + i) factor differs)."""
    boxes = np.concatenate([np.asarray(b, np.float32) for b in boxes_all])
    cls = np.concatenate([np.full(len(b), i + 1, np.int32) for i, b in enumerate(boxes_all)])
    return regroup(boxes, cls, 7)


def synth_annotations(seed, n_obj=40, timesteps=3):
    """Car-like boxes on the nuScenes range, some outside it, some degenerate; later timesteps move along the velocity."""
    rng = np.random.default_rng(seed)
    b = np.zeros((n_obj, 12), np.float32)
    b[:, :2] = rng.uniform(-58, 58, (n_obj, 2))
    b[:, 2] = rng.uniform(-2, 0.5, n_obj)
    b[:, 3:6] = np.abs(rng.normal([1.95, 4.6, 1.7], [0.3, 0.8, 0.2], (n_obj, 3)))
    b[:, 6:8] = rng.normal(0, 3, (n_obj, 2))
    b[:, 8:10] = -b[:, 6:8]
    b[:, 10] = rng.uniform(-7, 7, n_obj)
    b[:, 11] = b[:, 10] + np.float32(np.pi)
    b[0, 3] = 0.0                                   # degenerate width: skipped
    b[1, :2] = [-54.2, 10.0]                        # cell index truncates to 0 from a negative coordinate: kept
    b[2, :2] = [53.99, -53.99]
    out = []
    for t in range(timesteps):
        bt = b.copy()
        bt[:, :2] += b[:, 6:8] * np.float32(0.5 * t)
        out.append(bt.astype(np.float32))
    return out
