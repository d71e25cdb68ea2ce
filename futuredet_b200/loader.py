"""GPU multi-sweep assembly: raw nuScenes sweeps -> the `[sum N, 5]` (x, y, z, intensity, dt) cloud the voxelizer eats.

Mirrors det3d/datasets/pipelines/loading.py:24-60,102-147 (`read_file` column selection, `remove_close` on the sweeps,
sweep-to-keyframe transform in float64, time-lag column, key frame first then the sweeps in the order given) with the
raw `.bin` payloads as the wire format: the host only concatenates the files' bytes; everything else is one native
pass (`fd_assemble_sweeps`).  The result (points, device-side batch offsets) plugs into `VoxelNet.forward_points`
without a host synchronisation.
"""
import numpy as np
import torch

from . import lib as L
from .ops import _ptr, _stream


class SweepBatch:
    """Host-side description of a batch: for every scene the key frame followed by its sweeps.

    add_scene(key_records, sweeps): `key_records` [n,5] float32 (np.fromfile(...).reshape(-1, 5));
    `sweeps` = list of (records [n,5] float32, transform_matrix 4x4 float64 or None, time_lag float)."""

    def __init__(self, raw_stride=5, num_feat=4, close_radius=1.0):
        self.raw_stride, self.num_feat, self.close_radius = raw_stride, num_feat, close_radius
        self.chunks, self.offsets, self.xforms, self.flags, self.lags, self.scene = [], [0], [], [], [], []
        self.n_scenes = 0

    def _add(self, rec, xform, flags, lag):
        rec = np.ascontiguousarray(rec, dtype=np.float32).reshape(-1, self.raw_stride)
        self.chunks.append(rec)
        self.offsets.append(self.offsets[-1] + len(rec))
        self.xforms.append(np.eye(4) if xform is None else np.asarray(xform, np.float64).reshape(4, 4))
        self.flags.append(flags | (0 if xform is None else 1))
        self.lags.append(float(lag))
        self.scene.append(self.n_scenes)

    def add_scene(self, key_records, sweeps):
        self._add(key_records, None, 0, 0.0)                       # key frame: no remove_close, dt = 0 (loading.py:112-113)
        for rec, xform, lag in sweeps:
            self._add(rec, xform, 2, lag)                          # read_sweep: remove_close(1.0) + transform + time lag
        self.n_scenes += 1
        return self

    def to_device(self, device):
        raw = np.concatenate(self.chunks, 0) if self.chunks else np.zeros((0, self.raw_stride), np.float32)
        t = lambda a, dt: torch.from_numpy(np.asarray(a, dt)).to(device, non_blocking=True)
        return dict(raw=torch.from_numpy(raw).pin_memory().to(device, non_blocking=True),
                    offsets=t(self.offsets, np.int32), xforms=t(np.stack(self.xforms).reshape(-1, 16), np.float64),
                    flags=t(self.flags, np.int32), lags=t(self.lags, np.float32), scene=t(self.scene, np.int32),
                    S=len(self.flags), B=self.n_scenes)


def assemble_sweeps(batch, device=None):
    """-> (points [total_records, num_feat+1] fp32 CUDA (rows >= count are NaN), batch_offsets [B+1] int32 CUDA,
    count [1] int32 CUDA).  `batch`: SweepBatch or the dict its to_device() returns."""
    lib = L.load()
    d = batch.to_device(device or torch.device("cuda", torch.cuda.current_device())) if isinstance(batch, SweepBatch) else batch
    raw = d["raw"]
    total, stride = raw.shape
    nf = batch.num_feat if isinstance(batch, SweepBatch) else d.get("num_feat", 4)
    radius = batch.close_radius if isinstance(batch, SweepBatch) else d.get("close_radius", 1.0)
    dev = raw.device
    pts = torch.empty((max(total, 1), nf + 1), dtype=torch.float32, device=dev)
    boff = torch.empty((d["B"] + 1,), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    ws_bytes = lib.fd_sweeps_workspace_bytes(total)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    rc = lib.fd_assemble_sweeps(_ptr(raw), total, stride, nf, _ptr(d["offsets"]), _ptr(d["xforms"]), _ptr(d["flags"]),
                                _ptr(d["lags"]), _ptr(d["scene"]), d["S"], d["B"], float(radius), _ptr(pts), _ptr(boff),
                                _ptr(count), _ptr(ws), ws_bytes, _stream())
    L.check(rc, "fd_assemble_sweeps")
    return pts[:total], boff, count
