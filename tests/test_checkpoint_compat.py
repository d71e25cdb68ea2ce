"""CPU: reference checkpoints load unchanged.  The state_dict of every module on the path has exactly
the parameter / buffer names and shapes of the REFERENCE module: RPN and CenterHead as built by det3d.models, and
SpMiddleResNetFHD as built by the reference's own det3d/models/backbones/scn.py (tests/golden/state_keys.json, written
by oracle/gen_golden.py keys in the build container)."""
import json
import os

import futuredet_b200 as fb

HEAD = dict(type="CenterHead", in_channels=512, tasks=[dict(num_class=1, class_names=["car"])], dataset="nuscenes",
            weight=0.25, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
            common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
            share_conv_channel=64, dcn_head=False, classify=False)


def shapes(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def test_state_dict_layout_matches_reference_modules(golden_dir):
    want = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    bb = fb.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8))
    neck = fb.build_neck(dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                              us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256))
    assert shapes(bb) == want["backbone"]
    assert shapes(neck) == want["neck"]
    assert shapes(fb.build_head(dict(HEAD, timesteps=1))) == want["head_n0"]
    assert shapes(fb.build_head(dict(HEAD, timesteps=7))) == want["head_n3"]
    # head variants of the n3dtf / n3dtfm configs and the remaining flag combinations (center_head.py:99-126,320-372)
    for name, flags in (("head_n3dtf", dict(dense=True, forecast_feature=True)),
                        ("head_n3dtfm", dict(dense=True, forecast_feature=True, bev_map=True)),
                        ("head_two_stage", dict(two_stage=True)), ("head_wide", dict(wide_head=True)),
                        ("head_classify", dict(classify=True)), ("head_sparse", dict(sparse=True))):
        assert shapes(fb.build_head(dict(HEAD, timesteps=7, **{**dict(classify=False), **flags}))) == want[name], name
    # spconv-1.x weight layout [kD, kH, kW, Cin, Cout] and the 2*timesteps velocity channels of the n3 head
    assert want["backbone"]["conv2.0.weight"] == [3, 3, 3, 16, 32]
    assert want["head_n3"]["tasks.0.vel.3.weight"] == [14, 64, 3, 3]


def test_voxelnet_loads_a_reference_shaped_checkpoint(golden_dir):
    """A checkpoint with the reference's key layout (incl. the DDP `module.` prefix the released models carry) loads
    strictly into VoxelNet."""
    import torch
    want = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    ckpt = {}
    for part, prefix in (("backbone", "backbone."), ("neck", "neck."), ("head_n3", "bbox_head.")):
        for k, shp in want[part].items():
            ckpt["module." + prefix + k] = torch.zeros(shp, dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
    model = fb.build_detector(dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(HEAD, timesteps=7)))
    sd = {k[7:]: v for k, v in ckpt.items()}
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
