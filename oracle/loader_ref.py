"""ORACLE (test infrastructure only): multi-sweep assembly restated in numpy.

Follows det3d/datasets/pipelines/loading.py:24-33 (read_file: first 4 of the 5 columns of a .bin record), :36-45
(remove_close), :48-60 (read_sweep: float64 transform of float32 xyz, stored back as float32; time-lag column) and
:105-140 (LoadPointCloudFromFile: key frame with dt 0, then the sweeps, concatenated; `combined = hstack`).
Pinned by tests/golden/loader.npz, produced by the reference `LoadPointCloudFromFile.__call__` itself on synthetic
.bin files (oracle/gen_golden.py loader)."""
import numpy as np


def read_sweep_ref(records, transform, time_lag, radius=1.0):
    pts = np.asarray(records, np.float32).reshape(-1, 5)[:, :4].T.copy()            # [4, n]
    keep = ~((np.abs(pts[0]) < radius) & (np.abs(pts[1]) < radius))
    pts = pts[:, keep]
    n = pts.shape[1]
    if transform is not None:
        pts[:3, :] = np.asarray(transform).dot(np.vstack((pts[:3, :], np.ones(n))))[:3, :]
    return pts.T, time_lag * np.ones((n, 1))


def assemble_ref(key_records, sweeps):
    """key_records [n,5]; sweeps: list of (records, transform or None, time_lag) -> combined [N,5] float32."""
    pts = [np.asarray(key_records, np.float32).reshape(-1, 5)[:, :4]]
    times = [np.zeros((len(pts[0]), 1))]
    for rec, T, lag in sweeps:
        p, t = read_sweep_ref(rec, T, lag)
        pts.append(p)
        times.append(t)
    points = np.concatenate(pts, 0)
    return np.hstack([points, np.concatenate(times, 0).astype(points.dtype)])


def synth_sweeps(seed, n_sweeps=9, n_pts=3000):
    """Key frame + sweeps with ego-motion transforms, some points inside the 1 m ego box."""
    rng = np.random.default_rng(seed)

    def rec(n):
        r = rng.uniform(-40, 40, (n, 5)).astype(np.float32)
        r[:, 2] = rng.uniform(-3, 2, n)
        r[:, 3] = rng.uniform(0, 255, n)
        r[:, 4] = rng.integers(0, 32, n)
        close = rng.random(n) < 0.05
        r[close, :2] = rng.uniform(-1.2, 1.2, (int(close.sum()), 2)).astype(np.float32)
        return r
    sweeps = []
    for s in range(n_sweeps):
        yaw = 0.01 * (s + 1) * rng.uniform(0.5, 1.5)
        T = np.eye(4)
        T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
        T[:3, 3] = [0.5 * (s + 1), 0.03 * s, 0.001 * s]
        sweeps.append((rec(n_pts + 37 * s), T if s != 3 else None, 0.05 * (s + 1)))
    return rec(n_pts), sweeps
