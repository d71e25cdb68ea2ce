"""`det3d` import aliases so the reference's configs and user code load unchanged.

configs/centerpoint/*.py start with `from det3d.utils.config_tool import get_downsample_factor`
(e.g. nusc_centerpoint_forecast_n0_detection.py:4); tools call `from det3d.torchie import Config`,
`from det3d.models import build_detector`.  install() registers lightweight module objects under those
names that forward to futuredet_b200.  It refuses to shadow a real det3d that is already imported.
"""
import sys
import types


def install():
    existing = sys.modules.get("det3d")
    if existing is not None and not getattr(existing, "__futuredet_b200__", False):
        raise RuntimeError("a different `det3d` package is already imported; cannot install the futuredet_b200 aliases")
    if existing is not None:
        return existing
    from . import config, registry
    from . import backbone, detector, head, neck, pipelines, reader, sparse, voxel_generator  # noqa: F401 (register)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__futuredet_b200__ = True
        m.__path__ = []
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, m)
        return m

    root = mod("det3d")
    mod("det3d.utils", Registry=registry.Registry, build_from_cfg=registry.build_from_cfg)
    mod("det3d.utils.registry", Registry=registry.Registry, build_from_cfg=registry.build_from_cfg)
    mod("det3d.utils.config_tool", get_downsample_factor=config.get_downsample_factor)
    mod("det3d.torchie", Config=config.Config, ConfigDict=config.ConfigDict)
    mod("det3d.torchie.utils", Config=config.Config, ConfigDict=config.ConfigDict)
    mod("det3d.torchie.utils.config", Config=config.Config, ConfigDict=config.ConfigDict)
    reg_names = ["READERS", "BACKBONES", "NECKS", "HEADS", "LOSSES", "DETECTORS", "SECOND_STAGE", "ROI_HEAD"]
    builders = ["build", "build_reader", "build_backbone", "build_neck", "build_head", "build_loss", "build_detector"]
    regs = {n: getattr(registry, n) for n in reg_names}
    blds = {n: getattr(registry, n) for n in builders}
    mod("det3d.models", **regs, **blds)
    mod("det3d.models.registry", **regs)
    mod("det3d.models.builder", **blds)
    mod("det3d.models.detectors", VoxelNet=detector.VoxelNet, SingleStageDetector=detector.SingleStageDetector,
        BaseDetector=detector.BaseDetector)
    mod("det3d.models.readers", VoxelFeatureExtractorV3=reader.VoxelFeatureExtractorV3)
    mod("det3d.models.backbones", SpMiddleResNetFHD=backbone.SpMiddleResNetFHD)
    mod("det3d.models.backbones.scn", SpMiddleResNetFHD=backbone.SpMiddleResNetFHD,
        SparseBasicBlock=backbone.SparseBasicBlock)
    mod("det3d.models.necks", RPN=neck.RPN)
    mod("det3d.models.bbox_heads", CenterHead=head.CenterHead)
    mod("det3d.models.bbox_heads.center_head", CenterHead=head.CenterHead, SepHead=head.SepHead)
    mod("det3d.datasets", PIPELINES=registry.PIPELINES, DATASETS=registry.DATASETS)
    mod("det3d.datasets.registry", PIPELINES=registry.PIPELINES, DATASETS=registry.DATASETS)
    mod("det3d.datasets.pipelines", Voxelization=pipelines.Voxelization)
    mod("det3d.core")
    mod("det3d.core.input")
    mod("det3d.core.input.voxel_generator", VoxelGenerator=voxel_generator.VoxelGenerator)
    # `import spconv` users of the reference backbone get the native classes
    if "spconv" not in sys.modules:
        sp = types.ModuleType("spconv")
        for n in ("SparseConvTensor", "SubMConv3d", "SparseConv3d", "SparseSequential", "SparseModule"):
            setattr(sp, n, getattr(sparse, n))
        sp.__futuredet_b200__ = True
        sys.modules["spconv"] = sp
    return root
