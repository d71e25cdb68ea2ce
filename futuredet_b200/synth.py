"""Seeded synthetic nuScenes-shaped LiDAR scenes (the bench / parity workload).

Mirrors what the reference's loader produces (det3d/datasets/pipelines/loading.py:24-60,102-147):
`sweeps` sweeps of a 32-beam spinning LiDAR concatenated without shuffling, near-ego points removed,
columns (x, y, z, intensity, dt).  The RNG call order is part of the workload definition (it fixes the
point order, which the voxelizer is sensitive to); see SURVEY.md appendix A.3.
"""
import numpy as np

NUSC_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
NUSC_VOXEL = [0.075, 0.075, 0.2]


def synth_scene(n_target=300_000, sweeps=10, seed=0, beams=32):
    rng = np.random.default_rng(seed)
    per = n_target // sweeps
    az_n = max(per // beams, 1)
    elev = np.deg2rad(np.linspace(-30.67, 10.67, beams))
    pts = []
    for s in range(sweeps):
        az = np.linspace(-np.pi, np.pi, az_n, endpoint=False) + rng.uniform(0, 2 * np.pi / az_n)
        A, E = np.meshgrid(az, elev)
        A = A.ravel()
        E = E.ravel()
        r = np.where(E < -0.01, 1.84 / np.tan(-E), 80.0)          # ground plane at z = -1.84 m
        r_ob = rng.gamma(3.0, 7.0, size=r.shape) + 2
        hit = rng.random(r.shape) < 0.55
        r = np.where(hit, np.minimum(r, r_ob), r) + rng.normal(0, 0.02, r.shape)
        keep = r < 75
        x = r * np.cos(E) * np.cos(A) - 0.5 * s
        y = r * np.cos(E) * np.sin(A)
        z = r * np.sin(E)
        p = np.stack([x, y, z, rng.uniform(0, 255, r.shape), np.full(r.shape, 0.05 * s)], 1)[keep]
        pts.append(p[~((np.abs(p[:, 0]) < 1) & (np.abs(p[:, 1]) < 1))])   # remove_close, loading.py:36-45
    return np.concatenate(pts, 0).astype(np.float32)


def random_points(n=50_000, seed=0, snap_frac=0.0, pile=0):
    """BASELINE config 1 cloud: uniform box reaching past the range on every side (about 35 % rejected)."""
    rng = np.random.default_rng(seed)
    p = np.stack([rng.uniform(-60, 60, n), rng.uniform(-60, 60, n), rng.uniform(-6, 4, n),
                  rng.uniform(0, 255, n), rng.uniform(0, 0.5, n)], 1).astype(np.float32)
    if snap_frac > 0:            # snap x onto voxel boundaries: float32(-54) + k * float32(0.075)
        m = rng.random(n) < snap_frac
        k = rng.integers(-5, 1446, n).astype(np.float32)
        p[m, 0] = (np.float32(-54.0) + k * np.float32(0.075))[m]
    if pile > 0:                 # many points into few voxels (max_points stress)
        centres = p[rng.integers(0, n, 50), :3]
        sel = rng.integers(0, n, pile)
        p[sel, :3] = centres[rng.integers(0, 50, pile)] + rng.uniform(-0.03, 0.03, (pile, 3)).astype(np.float32)
    return p


def synth_targets(batch, H, W, timesteps, n_obj=50, max_objs=500, seed=0, sigma=2.0):
    """Seeded CenterPoint training targets in the collate layout `example[key][timestep][task] -> Tensor[B, ...]`
    (det3d/torchie/parallel/collate.py:208-232; produced in the reference by AssignLabel,
    det3d/datasets/pipelines/preprocess.py:464-546): `n_obj` random object centres per scene splatted as Gaussians
    into `hm [B,1,H,W]`, `ind / cat [B,max_objs] int64`, `mask [B,max_objs] uint8`, `anno_box [B,max_objs,14]`."""
    import torch
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:H, 0:W]
    ex = {k: [] for k in ("hm", "anno_box", "ind", "mask", "cat")}
    cy = rng.integers(0, H, (batch, n_obj))
    cx = rng.integers(0, W, (batch, n_obj))
    for t in range(timesteps):
        hm = np.zeros((batch, 1, H, W), np.float32)
        ind = np.zeros((batch, max_objs), np.int64)
        mask = np.zeros((batch, max_objs), np.uint8)
        box = np.zeros((batch, max_objs, 14), np.float32)
        for b in range(batch):
            for j in range(n_obj):
                g = np.exp(-((ys - cy[b, j]) ** 2 + (xs - cx[b, j]) ** 2) / (2 * sigma * sigma)).astype(np.float32)
                hm[b, 0] = np.maximum(hm[b, 0], g)
            ind[b, :n_obj] = cy[b] * W + cx[b]
            mask[b, :n_obj] = 1
            box[b, :n_obj] = rng.standard_normal((n_obj, 14)).astype(np.float32)
        ex["hm"].append([torch.from_numpy(hm)])
        ex["anno_box"].append([torch.from_numpy(box)])
        ex["ind"].append([torch.from_numpy(ind)])
        ex["mask"].append([torch.from_numpy(mask)])
        ex["cat"].append([torch.zeros((batch, max_objs), dtype=torch.int64)])
    return ex
