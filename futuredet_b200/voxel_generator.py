"""VoxelGenerator (det3d/core/input/voxel_generator.py:5-46) backed by the CUDA voxelizer."""
import numpy as np
import torch

from . import ops


class VoxelGenerator:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels
        self._grid_size = ops.grid_size_of(point_cloud_range, voxel_size)

    def generate(self, points, max_voxels=-1, device="cuda"):
        """points [N, F] (numpy or tensor) -> (voxels [M,max_points,F], coordinates [M,3] (z,y,x), num_points [M])
        as numpy arrays, exactly what points_to_voxel returns (point_cloud_ops.py:181-184)."""
        if max_voxels == -1:
            max_voxels = self._max_voxels
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32) if isinstance(points, np.ndarray) else points)
        pts = pts.to(device=device, dtype=torch.float32).contiguous()
        off = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=pts.device)
        r = ops.voxelize_vfe(pts, off, self._voxel_size.tolist(), self._point_cloud_range.tolist(),
                             self._max_num_points, max_voxels, want_voxels=True)
        m = int(r["total"].item())
        return (r["voxels"][:m].cpu().numpy(), r["coords"][:m, 1:].contiguous().cpu().numpy(),
                r["num_points"][:m].cpu().numpy())

    voxel_size = property(lambda self: self._voxel_size)
    max_num_points_per_voxel = property(lambda self: self._max_num_points)
    point_cloud_range = property(lambda self: self._point_cloud_range)
    grid_size = property(lambda self: self._grid_size)
