"""GPU parity of (a) the sparse backbone against the golden written by the REFERENCE's own scn.py source (run over the
CPU spconv shim, oracle/gen_golden.py backbone) and (b) the dense / forecast_feature / bev_map CenterHead variants of
the n3dtf / n3dtfm configs against the reference CenterHead class (forward incl. `feats`, dense loss, dense predict)."""
import os

import numpy as np
import pytest
import torch

import futuredet_b200 as fb
from oracle import spconv_ref as S

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_backbone_matches_reference_scn_golden(cuda, golden_dir, precision):
    g = torch.load(os.path.join(golden_dir, "backbone_scn.pt"), weights_only=False)
    bb = fb.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8)).eval()
    bb.load_state_dict(S.seeded_state(g["state_shapes"], g["state_seed"]), strict=True)
    bb.to(cuda)
    with fb.use_precision(precision):
        feats = g["features"]
        if precision != "fp32":                              # tensor-core stem wants rows padded to 8 channels
            feats = torch.cat([feats, torch.zeros(len(feats), 3)], 1)
        out, stages = bb(feats.to(cuda), g["coors"].to(cuda), g["batch_size"], g["grid"])
        assert out.shape == g["out"].shape
        torch.testing.assert_close(out.cpu(), g["out"], rtol=TOL, atol=TOL)
        for name in ("conv1", "conv2", "conv3", "conv4"):
            ref = g["stages"][name]
            st = stages[name]
            n = st.num_active()
            assert n == len(ref["indices"]) and list(st.spatial_shape) == list(ref["spatial_shape"])
            assert torch.equal(st.indices[:n].cpu(), ref["indices"].int()), name       # bit-exact active set and order
        f4 = stages["conv4"].features[:len(g["stages"]["conv4"]["features"])].cpu()
        torch.testing.assert_close(f4, g["stages"]["conv4"]["features"], rtol=TOL, atol=TOL)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["n3dtf", "n3dtfm"])
def test_head_variants_match_reference_class(cuda, golden_dir, name, precision):
    g = torch.load(os.path.join(golden_dir, "head_variants.pt"), weights_only=False)[name]
    head = fb.build_head(dict(g["cfg"])).eval()
    sd = S.seeded_state(g["state_shapes"], g["state_seed"])
    for k in sd:
        if k.endswith("hm.3.bias"):
            sd[k] = sd[k] - 2.19
    head.load_state_dict(sd, strict=True)
    head.to(cuda)
    with fb.use_precision(precision):
        bm = g["bev_map"].to(cuda) if g["bev_map"] is not None else None
        preds = head(g["x"].to(cuda), bm)
        assert len(preds) == len(g["preds"]) == 3
        for p, w in zip(preds, g["preds"]):
            assert set(p) == set(w)
            for k in w:
                assert p[k].shape == w[k].shape, k
                torch.testing.assert_close(p[k].cpu(), w[k], rtol=TOL, atol=TOL, msg=lambda s_, k=k: "%s: %s" % (k, s_))
        ex = {k: [[t.to(cuda) for t in per_t] for per_t in v] for k, v in g["example"].items()}
        loss = head.loss(ex, [{k: v.clone() for k, v in p.items()} for p in preds])
        for i in range(3):
            assert abs(float(loss["loss"][i]) - float(g["loss"]["loss"][i])) <= 2e-3
            assert abs(float(loss["hm_loss"][i]) - float(g["loss"]["hm_loss"][i])) <= 2e-3
            torch.testing.assert_close(loss["loc_loss_elem"][i].cpu(), g["loss"]["loc_loss_elem"][i], rtol=2e-3, atol=2e-3)


def test_dense_predict_matches_reference_golden(cuda, golden_dir):
    from futuredet_b200 import predict as P
    from test_gpu_predict import compare, to_head_views
    g = torch.load(os.path.join(golden_dir, "head_variants.pt"), weights_only=False)["dense_predict"]

    class H:
        dense, standard, timesteps, target_timesteps, num_classes = True, False, 3, 7, [1, 1, 1]
    ret = P.center_head_predict(H(), {}, [to_head_views(p, cuda) for p in g["preds"]], g["test_cfg"])
    compare(ret, g["ret"], exact_cells=False)


@pytest.mark.parametrize("bev", [False, True])
def test_detector_with_dense_forecast_head_end_to_end(cuda, bev):
    """VoxelNet built like the n3dtf / n3dtfm configs (dense + forecast_feature [+ bev_map], 7 chained SepHeads):
    `model(example, return_loss=True)` (eval-mode loss) and `return_loss=False` (dense predict) through the detector."""
    from oracle.gen_golden import make_targets
    from test_gpu_train import random_sites
    from test_gpu_predict import TEST_CFG
    torch.manual_seed(0)
    rng = np.random.default_rng(5)
    cfg = dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(type="CenterHead", in_channels=512, tasks=[dict(num_class=1, class_names=["car"])],
                       dataset="nuscenes", weight=0.25, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                       share_conv_channel=64, dcn_head=False, timesteps=7, dense=True, forecast_feature=True,
                       bev_map=bev, classify=False))
    model = fb.build_detector(cfg, test_cfg=fb.ConfigDict(TEST_CFG)).to(cuda).eval()
    assert len(model.bbox_head.tasks) == 7
    B, grid = 2, [64, 64, 40]
    c = random_sites(rng, B, [40, 64, 64], 4000)
    n = len(c)
    voxels = np.zeros((n, 10, 5), np.float32); voxels[:, 0] = rng.standard_normal((n, 5)).astype(np.float32)
    example = make_targets(B, 8, 8, 7, torch.Generator().manual_seed(3), max_objs=10)
    example = {k: [[t.to(cuda) for t in row] for row in v] for k, v in example.items()}
    example.update(voxels=torch.from_numpy(voxels).to(cuda), num_points=torch.ones(n, dtype=torch.int32, device=cuda),
                   coordinates=torch.from_numpy(c).to(cuda), num_voxels=torch.tensor([0] * B), shape=[np.array(grid)] * B,
                   bev_map=[torch.rand((B, 8, 8), device=cuda) for _ in range(6)])
    with torch.no_grad():
        losses = model(example, return_loss=True)
        dets = model(example, return_loss=False)
    assert len(losses["loss"]) == 7 and all(torch.isfinite(l).all() for l in losses["loss"])
    assert len(dets) == B and all(set(d) >= {"box3d_lidar", "scores", "label_preds"} for d in dets)
    assert all(d["box3d_lidar"].shape[1] == 9 for d in dets)
    assert all(len(d["label_preds"]) == 0 or int(d["label_preds"].max()) <= 6 for d in dets)
