"""futuredet_b200 -- B200-native (sm_100a) implementation of FutureDet's LiDAR hot path.

voxelize + VFE -> sparse 3-D backbone -> BEV neck -> multi-timestep CenterHead, behind det3d's registry API.
All arithmetic runs in libfuturedet_b200.so (hand-written CUDA behind the C ABI of include/futuredet_b200.h);
this package is the ctypes / torch-tensor plumbing plus the det3d-compatible module classes.
"""
from . import lib  # noqa: F401
from .config import Config, ConfigDict, get_downsample_factor  # noqa: F401
from .registry import (BACKBONES, DETECTORS, HEADS, NECKS, PIPELINES, READERS, Registry, build_backbone,  # noqa: F401
                       build_detector, build_from_cfg, build_head, build_neck, build_reader)
from . import sparse, reader, backbone, neck, head, detector, voxel_generator, pipelines  # noqa: F401,E402
from .compat import install as install_det3d_aliases  # noqa: F401,E402
from .detector import VoxelNet  # noqa: F401,E402
from .precision import default_precision, set_default_precision, use_precision  # noqa: F401,E402

__version__ = "0.1.0"
