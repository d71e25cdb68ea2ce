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
