"""`Voxelization` pipeline stage (det3d/datasets/pipelines/preprocess.py:226-271) on the CUDA voxelizer."""
import numpy as np

from .registry import PIPELINES
from .voxel_generator import VoxelGenerator


@PIPELINES.register_module
class Voxelization:
    def __init__(self, **kwargs):
        cfg = kwargs.get("cfg", None)
        self.range = cfg["range"]
        self.voxel_size = cfg["voxel_size"]
        self.max_points_in_voxel = cfg["max_points_in_voxel"]
        mv = cfg["max_voxel_num"]
        self.max_voxel_num = [mv, mv] if isinstance(mv, int) else mv
        self.double_flip = cfg.get("double_flip", False)
        if self.double_flip:
            raise NotImplementedError("double_flip test-time augmentation is outside the hot-path scope")
        self.voxel_generator = VoxelGenerator(voxel_size=self.voxel_size, point_cloud_range=self.range,
                                              max_num_points=self.max_points_in_voxel, max_voxels=self.max_voxel_num[0])

    def __call__(self, res, info):
        vg = self.voxel_generator
        max_voxels = self.max_voxel_num[0] if res["mode"] == "train" else self.max_voxel_num[1]   # preprocess.py:249-258
        voxels, coordinates, num_points = vg.generate(res["lidar"]["points"], max_voxels=max_voxels)
        res["lidar"]["voxels"] = dict(
            voxels=voxels, coordinates=coordinates, num_points=num_points,
            num_voxels=np.array([voxels.shape[0]], dtype=np.int64), shape=vg.grid_size,
            range=vg.point_cloud_range, size=vg.voxel_size)
        return res, info
