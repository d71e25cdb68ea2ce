"""ORACLE (test infrastructure only): reference voxelizer + mean VFE on the CPU.

Two restatements of det3d/ops/point_cloud/point_cloud_ops.py:7-55,112-184 (+ VFE,
det3d/models/readers/voxel_encoder.py:17-24):
  * `*_c`  : plain C (oracle/voxelize_ref.c), loaded with ctypes;
  * `*_np` : vectorised numpy float32 (first-appearance order through np.unique).
Both are pinned against the reference numba function via tests/golden/voxel_*.npz.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libfd_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.fdo_voxelize_vfe.restype = C.c_int32
        _lib.fdo_voxelize_vfe.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def grid_size_of(coors_range, voxel_size):
    r = np.asarray(coors_range, np.float32)
    v = np.asarray(voxel_size, np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int32)   # voxel_generator.py:10-11


def points_to_voxel_c(points, voxel_size, coors_range, max_points, max_voxels, want_voxels=True, want_mean=True):
    """-> dict(voxels [M,max_points,F] | None, coors [M,3] (z,y,x), num_points [M], mean [M,F] | None)."""
    lib = _load()
    pts = np.ascontiguousarray(points, np.float32)
    n, f = pts.shape
    vs = np.ascontiguousarray(voxel_size, np.float32)
    rg = np.ascontiguousarray(coors_range, np.float32)
    grid = np.ascontiguousarray(grid_size_of(rg, vs), np.int32)
    voxels = np.empty((max_voxels, max_points, f), np.float32) if want_voxels else None
    coors = np.empty((max_voxels, 3), np.int32)
    npts = np.empty((max_voxels,), np.int32)
    mean = np.empty((max_voxels, f), np.float32) if want_mean else None
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    m = lib.fdo_voxelize_vfe(p(pts), n, f, p(vs), p(rg), p(grid), max_points, max_voxels, p(voxels), p(coors),
                             p(npts), p(mean))
    if m < 0:
        raise MemoryError("oracle voxelizer: allocation failed")
    return dict(voxels=voxels[:m] if want_voxels else None, coors=coors[:m], num_points=npts[:m],
                mean=mean[:m] if want_mean else None)


def points_to_voxel_np(points, voxel_size, coors_range, max_points, max_voxels):
    """Vectorised float32 restatement (coords, order, counts and mean)."""
    pts = np.asarray(points, np.float32)
    vs = np.asarray(voxel_size, np.float32)
    rg = np.asarray(coors_range, np.float32)
    grid = grid_size_of(rg, vs)
    q = np.floor((pts[:, :3] - rg[:3]) / vs)                      # float32 throughout (:36)
    ok = np.all((q >= 0) & (q < grid.astype(np.float32)), axis=1)  # :37
    idx = np.nonzero(ok)[0]
    qi = q[idx].astype(np.int64)
    key = (qi[:, 2] * grid[1] + qi[:, 1]) * grid[0] + qi[:, 0]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                       # voxel id = order of first appearance
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    vid = rank[inv]
    keep = vid < max_voxels                                        # :46-47
    m = int(min(order.size, max_voxels))
    vid_k, idx_k = vid[keep], idx[keep]
    # ordinal of each point inside its voxel (input order)
    srt = np.argsort(vid_k, kind="stable")
    vs_sorted = vid_k[srt]
    start = np.r_[0, np.nonzero(np.diff(vs_sorted))[0] + 1]
    counts = np.diff(np.r_[start, vs_sorted.size])
    ordinal = np.arange(vs_sorted.size) - np.repeat(start, counts)
    sel = ordinal < max_points                                     # :51
    f = pts.shape[1]
    voxels = np.zeros((m, max_points, f), np.float32)
    voxels[vs_sorted[sel], ordinal[sel]] = pts[idx_k[srt][sel]]
    num_points = np.minimum(np.bincount(vid_k, minlength=m), max_points).astype(np.int32)
    ukey = uniq[order][:m]
    coors = np.stack([ukey // (grid[0] * grid[1]), (ukey // grid[0]) % grid[1], ukey % grid[0]], 1).astype(np.int32)
    mean = np.zeros((m, f), np.float32)
    for s in range(max_points):                                    # sequential fp32 sum over the slot axis
        mean += voxels[:, s]
    mean = mean / num_points.astype(np.float32)[:, None]
    return dict(voxels=voxels, coors=coors, num_points=num_points, mean=mean)


def voxelize_batch_c(scenes, voxel_size, coors_range, max_points, max_voxels):
    """Batched oracle incl. the collate batch-index column (collate.py:199-206)."""
    feats, coords, npts, nvox = [], [], [], []
    for b, pts in enumerate(scenes):
        r = points_to_voxel_c(pts, voxel_size, coors_range, max_points, max_voxels, want_voxels=False)
        feats.append(r["mean"])
        coords.append(np.pad(r["coors"], ((0, 0), (1, 0)), mode="constant", constant_values=b))
        npts.append(r["num_points"])
        nvox.append(r["coors"].shape[0])
    return dict(features=np.concatenate(feats), coords=np.concatenate(coords).astype(np.int32),
                num_points=np.concatenate(npts), num_voxels=np.asarray(nvox, np.int32))
