"""CPU: the C-ABI library builds, loads and exports every symbol include/futuredet_b200.h declares; the
det3d-compatible host layer loads the reference-style configs and builds the registry classes."""
import ctypes
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CFG_TEXT = '''
import itertools
import logging
from det3d.utils.config_tool import get_downsample_factor
timesteps = 7
TWO_STAGE = False
tasks = [dict(num_class=1, class_names=["car"])]
class_names = list(itertools.chain(*[t["class_names"] for t in tasks]))
model = dict(
    type="VoxelNet", pretrained=None,
    reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
    backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
    neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
              us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256, logger=logging.getLogger("RPN")),
    bbox_head=dict(type="CenterHead", in_channels=sum([256, 256]), tasks=tasks, dataset="nuscenes", weight=0.25,
                   code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
                   common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                   share_conv_channel=64, dcn_head=False, timesteps=timesteps, two_stage=TWO_STAGE, reverse=False,
                   sparse=False, dense=False, bev_map=False, forecast_feature=False, classify=False, wide_head=False))
assigner = dict(out_size_factor=get_downsample_factor(model), max_objs=1000)
train_cfg = dict(assigner=assigner)
test_cfg = dict(score_threshold=0.1, out_size_factor=get_downsample_factor(model))
voxel_generator = dict(range=[-54, -54, -5.0, 54, 54, 3.0], voxel_size=[0.075, 0.075, 0.2],
                       max_points_in_voxel=10, max_voxel_num=[120000, 160000], double_flip=False)
'''


def declared_symbols():
    text = open(os.path.join(REPO, "include", "futuredet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from futuredet_b200 import build, lib
    path = build.build_library()
    assert os.path.exists(path)
    dll = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(dll, s), "missing export %s" % s
    assert sorted(lib.SIGNATURES) == syms          # the ctypes table binds exactly the declared ABI
    assert lib.load().fd_version() == 4


def test_conv_desc_layout_matches_header():
    """sizeof(fd_conv_desc) as the C compiler sees it == the ctypes mirror."""
    import subprocess, tempfile
    from futuredet_b200 import lib
    src = '#include <stdio.h>\n#include "futuredet_b200.h"\nint main(){printf("%zu", sizeof(fd_conv_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        size = int(subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout)
    assert size == ctypes.sizeof(lib.ConvDesc)


def test_invalid_arguments_return_errors_not_aborts():
    from futuredet_b200 import lib
    L = lib.load()
    assert L.fd_conv_forward(None, None) < 0
    assert b"null descriptor" in L.fd_last_error()
    assert L.fd_voxelize_workspace_bytes(-1, 1, 1, 1) == 0
    assert L.fd_voxelize_workspace_bytes(300000, 1, 160000, 10) > 300000 * 8


def test_missing_library_fails_loudly(monkeypatch):
    from futuredet_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libfuturedet_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        lib.load()


def test_reference_style_config_loads_and_builds(tmp_path):
    import futuredet_b200 as fb
    p = tmp_path / "nusc_centerpoint_forecast_n3_detection.py"
    p.write_text(CFG_TEXT)
    cfg = fb.Config.fromfile(str(p))
    assert cfg.timesteps == 7 and cfg.TWO_STAGE is False
    assert cfg.assigner.out_size_factor == 8 and cfg["test_cfg"]["out_size_factor"] == 8
    assert cfg.get("missing", 3) == 3
    assert cfg.model.neck.ds_num_filters == [128, 256]
    model = fb.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
    sd = model.state_dict()
    # key layout / shapes of the reference checkpoint format
    assert sd["backbone.conv_input.0.weight"].shape == (3, 3, 3, 5, 16)
    assert "backbone.conv1.0.conv1.bias" in sd and "backbone.conv2.0.bias" not in sd
    assert sd["backbone.conv4.0.weight"].shape == (3, 3, 3, 64, 128)
    assert sd["backbone.extra_conv.0.weight"].shape == (3, 1, 1, 128, 128)
    assert sd["neck.blocks.0.1.weight"].shape == (128, 256, 3, 3) and "neck.blocks.0.0.weight" not in sd
    assert sd["neck.deblocks.1.0.weight"].shape == (256, 256, 2, 2)
    assert sd["bbox_head.tasks.0.vel.3.weight"].shape == (14, 64, 3, 3)       # 2 * timesteps velocity channels
    n_params = sum(p.numel() for p in model.parameters())
    assert n_params == 7_802_791 - 0 or n_params > 7_700_000
    assert sum(p.numel() for p in model.neck.parameters()) == 4_576_768
    assert sum(p.numel() for p in model.bbox_head.parameters()) == 530_711


def test_state_dict_keys_match_reference_golden(golden_dir):
    """Keys/shapes of our RPN/CenterHead == those saved from the reference classes."""
    import futuredet_b200 as fb
    g = torch.load(os.path.join(golden_dir, "neck_head.pt"), weights_only=False)
    neck = fb.build_neck(dict(g["neck_cfg"]))
    head = fb.build_head(dict(g["head_cfg"]))
    for mod, ref in ((neck, g["neck_state"]), (head, g["head_state"])):
        mine = mod.state_dict()
        assert sorted(mine.keys()) == sorted(ref.keys())
        for k in ref:
            assert mine[k].shape == ref[k].shape, k
        mod.load_state_dict(ref, strict=True)


def test_unsupported_variants_raise():
    import futuredet_b200 as fb
    with pytest.raises(NotImplementedError):          # deformable convolution is outside the hot-path scope
        fb.build_head(dict(type="CenterHead", in_channels=32, tasks=[dict(num_class=1, class_names=["car"])],
                           code_weights=[1.0] * 10, common_heads={"reg": (2, 2)}, dcn_head=True, classify=False))


REF_CFG_DIR = "/root/reference/configs/centerpoint"


@pytest.mark.skipif(not os.path.isdir(REF_CFG_DIR), reason="the reference tree exists only in the build container")
def test_every_reference_voxelnet_config_loads_and_builds():
    """The shipped configs load UNCHANGED through Config.fromfile and build through the registry: all 8 VoxelNet
    configs (n0 / n3 / n3dtf / n3dtfm, car and pedestrian).  The 2 PointPillars configs are out of scope."""
    import glob
    import futuredet_b200 as fb
    paths = sorted(glob.glob(os.path.join(REF_CFG_DIR, "*.py")))
    vox = [p for p in paths if "_pp_" not in os.path.basename(p)]
    assert len(vox) == 8
    for p in vox:
        cfg = fb.Config.fromfile(p)
        assert cfg.model.type == "VoxelNet"
        model = fb.build_detector(cfg.model, train_cfg=cfg.train_cfg, test_cfg=cfg.test_cfg)
        h = model.bbox_head
        name = os.path.basename(p)
        assert h.dense == ("n3dtf" in name) and h.forecast_feature == ("n3dtf" in name) and h.bev_map == ("n3dtfm" in name)
        assert len(h.tasks) == (7 if "n3dtf" in name else 1)
        assert model.precision is None and fb.default_precision() == "bf16x3"


def test_precision_api():
    import futuredet_b200 as fb
    from futuredet_b200 import precision as P
    assert fb.default_precision() == "bf16x3"                  # the tensor-core arm is the product default
    with fb.use_precision("fp32"):
        assert P.resolve(None) == "fp32" and P.act_fmt() == "fp32" and P.resolve("bf16x3") == "bf16x3"
    assert P.resolve(None) == "bf16x3" and P.act_fmt() == "split"
    with pytest.raises(ValueError):
        P.check("fp16")
    head = dict(type="CenterHead", in_channels=32, tasks=[dict(num_class=1, class_names=["car"])],
                code_weights=[1.0] * 10, common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2),
                                                       "vel": (2, 2)}, classify=False)
    cfg = dict(type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
               backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
               neck=dict(type="RPN", layer_nums=[1, 1], ds_layer_strides=[1, 2], ds_num_filters=[8, 16],
                         us_layer_strides=[1, 2], us_num_filters=[16, 16], num_input_features=256), bbox_head=head)
    m = fb.build_detector(cfg, test_cfg=dict(precision="fp32"))
    assert m.precision == "fp32" and m.neck.precision == "fp32" and m.bbox_head.tasks[0].precision == "fp32"
    assert all(c.precision == "fp32" for c in m.backbone.modules() if hasattr(c, "indice_key"))
    m.set_precision(None)
    assert m.neck.precision is None and m.backbone.conv_input[0].precision is None
