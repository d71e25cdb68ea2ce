"""GPU parity of the det3d-facing modules: RPN + CenterHead vs golden tensors from the reference classes,
SpMiddleResNetFHD vs the spconv restatement, and the whole VoxelNet forward vs the chained oracle (<= 1e-3)."""
import os

import numpy as np
import pytest
import torch

import futuredet_b200 as fb
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene
from oracle import dense_ref as D
from oracle import spconv_ref as S
from oracle import voxelizer as V

pytestmark = pytest.mark.gpu
TOL = 1e-3      # north_star: heatmap / box tensors within 1e-3 fp32


def randomise_bn(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


@pytest.fixture(params=["fp32", "bf16x3"])
def precision(request):
    with fb.use_precision(request.param):
        yield request.param


def test_rpn_and_center_head_match_reference_golden(cuda, golden_dir, precision):
    g = torch.load(os.path.join(golden_dir, "neck_head.pt"), weights_only=False)
    neck = fb.build_neck(dict(g["neck_cfg"])).eval()
    head = fb.build_head(dict(g["head_cfg"])).eval()
    neck.load_state_dict(g["neck_state"]); head.load_state_dict(g["head_state"])
    neck.to(cuda); head.to(cuda)
    feat = neck(g["x"].to(cuda))
    assert feat.shape == g["neck_out"].shape
    torch.testing.assert_close(feat.cpu(), g["neck_out"], rtol=TOL, atol=TOL)
    preds = head(feat)
    assert len(preds) == 1 and set(preds[0]) == set(g["preds"][0])
    for k, v in g["preds"][0].items():
        assert preds[0][k].shape == v.shape
        torch.testing.assert_close(preds[0][k].cpu(), v, rtol=TOL, atol=TOL)


def build_model(timesteps, dev, seed=0):
    torch.manual_seed(seed)
    cfg = dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(type="CenterHead", in_channels=512, tasks=[dict(num_class=1, class_names=["car"])],
                       dataset="nuscenes", weight=0.25, code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                       share_conv_channel=64, dcn_head=False, timesteps=timesteps, classify=False))
    m = fb.build_detector(cfg).eval()
    randomise_bn(m, seed + 1)
    return m


def test_backbone_matches_oracle(cuda, precision):
    m = build_model(1, cuda)
    pts = synth_scene(40000, seed=0)
    vox = V.voxelize_batch_c([pts, synth_scene(30000, seed=1)], NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    sd = {k: v.clone() for k, v in m.backbone.state_dict().items()}
    feats = torch.from_numpy(vox["features"])
    want = S.backbone_forward(sd, feats, vox["coords"], 2, [1440, 1440, 40])
    m.to(cuda)
    got, stages = m.backbone(feats.to(cuda), torch.from_numpy(vox["coords"]).to(cuda), 2, [1440, 1440, 40])
    assert got.shape == (2, 256, 180, 180)
    torch.testing.assert_close(got.cpu(), want, rtol=TOL, atol=TOL)
    assert stages["conv4"].spatial_shape == [5, 180, 180]


@pytest.mark.parametrize("timesteps", [1, 7])
def test_voxelnet_forward_points_matches_chained_oracle(cuda, timesteps, precision):
    """forecast_n0 (timesteps=1) and forecast_n3 (7-timestep vel head) on a reduced scene, end to end."""
    m = build_model(timesteps, cuda, seed=timesteps)
    scenes = [synth_scene(50000, seed=3), synth_scene(36000, seed=4)]
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    vox = V.voxelize_batch_c(scenes, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    bev = S.backbone_forward({k[9:]: v for k, v in sd.items() if k.startswith("backbone.")},
                             torch.from_numpy(vox["features"]), vox["coords"], 2, [1440, 1440, 40])
    feat = D.rpn_forward({k[5:]: v for k, v in sd.items() if k.startswith("neck.")}, bev, [5, 5], [1, 2], [1, 2])
    names = [["reg", "height", "dim", "rot", "vel", "hm"]]
    want = D.center_head_forward({k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}, feat, names)
    m.to(cuda)
    m.configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                               max_voxel_num=[120000, 160000]))
    pts = torch.from_numpy(np.concatenate(scenes)).to(cuda)
    off = torch.tensor([0, len(scenes[0]), len(scenes[0]) + len(scenes[1])], dtype=torch.int32, device=cuda)
    preds, voxd = m.forward_points(pts, off, return_voxels=True)
    n = int(voxd["total"].item())
    assert np.array_equal(voxd["coords"][:n].cpu().numpy(), vox["coords"])
    assert preds[0]["vel"].shape == (2, 2 * timesteps, 180, 180)
    for k, v in want[0].items():
        torch.testing.assert_close(preds[0][k].cpu(), v, rtol=TOL, atol=TOL, msg=lambda s: "%s: %s" % (k, s))
    # reference-API entry (padded voxels through the `example` dict) agrees with the fused entry
    ex_pts = [V.points_to_voxel_c(s, NUSC_VOXEL, NUSC_RANGE, 10, 160000) for s in scenes]
    example = dict(voxels=torch.from_numpy(np.concatenate([e["voxels"] for e in ex_pts])).to(cuda),
                   coordinates=torch.from_numpy(vox["coords"]).to(cuda),
                   num_points=torch.from_numpy(vox["num_points"]).to(cuda),
                   num_voxels=torch.from_numpy(vox["num_voxels"].astype(np.int64)).to(cuda),
                   shape=np.array([[1440, 1440, 40]] * 2))
    data = dict(features=example["voxels"], num_voxels=example["num_points"], coors=example["coordinates"],
                batch_size=2, input_shape=example["shape"][0])
    x, _ = m.extract_feat(data)
    p2 = m.bbox_head(x)
    for k in want[0]:
        torch.testing.assert_close(p2[0][k], preds[0][k], rtol=1e-4, atol=1e-4)   # API path re-rounds at module boundaries


def test_center_head_loss_matches_reference_golden(cuda, golden_dir):
    """CenterHead.loss (focal + masked-L1, 3 forecast timesteps) vs the reference's own `head.loss` output."""
    g = torch.load(os.path.join(golden_dir, "neck_head.pt"), weights_only=False)
    head = fb.build_head(dict(g["head_cfg"])).eval()
    preds = [{k: v.clone().to(cuda) for k, v in p.items()} for p in g["preds"]]       # reference predictions (NCHW)
    ex = {k: [[t.to(cuda) for t in ts] for ts in v] for k, v in g["example"].items()}
    out = head.loss(ex, preds)
    ref = g["loss"]
    torch.testing.assert_close(out["loss"][0].cpu(), ref["loss"][0], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out["hm_loss"][0], ref["hm_loss"][0], rtol=1e-4, atol=1e-5)
    for a, b in zip(out["loc_loss"][0], ref["loc_loss"][0]):
        torch.testing.assert_close(a.cpu(), b, rtol=1e-4, atol=1e-5)
    for a, b in zip(out["loc_loss_elem"][0], ref["loc_loss_elem"][0]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    assert float(out["num_positive"][0]) == float(ref["num_positive"][0])
    # in-place _sigmoid side effect of the reference (center_head.py:392-394,402)
    want_hm = torch.clamp(torch.sigmoid(g["preds"][0]["hm"]), 1e-4, 1 - 1e-4)
    torch.testing.assert_close(preds[0]["hm"].cpu(), want_hm, rtol=1e-5, atol=1e-6)
    # and on the head's own channels-last outputs (strided views)
    neck = fb.build_neck(dict(g["neck_cfg"])).eval()
    neck.load_state_dict(g["neck_state"]); head.load_state_dict(g["head_state"])
    neck.to(cuda); head.to(cuda)
    preds2 = head(neck(g["x"].to(cuda)))
    out2 = head.loss(ex, preds2)
    torch.testing.assert_close(out2["loss"][0].cpu(), ref["loss"][0], rtol=1e-3, atol=1e-3)


def test_forward_host_matches_forward_points(cuda):
    """Host-buffer entry (pinned points in, pinned head tensors out) == device entry."""
    m = build_model(7, cuda).to(cuda)
    m.configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                               max_voxel_num=[120000, 160000]))
    scene = synth_scene(20000, seed=11)
    pts_h = torch.from_numpy(scene).pin_memory()
    off_h = torch.tensor([0, len(scene)], dtype=torch.int32).pin_memory()
    outs, layout = m.forward_host(pts_h, off_h)
    torch.cuda.synchronize()
    preds = m.forward_points(pts_h.to(cuda), off_h.to(cuda))
    assert len(outs) == 1 and outs[0].is_pinned() and outs[0].shape == (1, 180, 180, 23)
    col = 0
    for t_id, name, c in layout:
        want = preds[t_id][name].permute(0, 2, 3, 1).cpu()
        torch.testing.assert_close(outs[0][..., col:col + c], want, rtol=0, atol=0)
        col += c
    assert [n for _, n, _ in layout] == ["reg", "height", "dim", "rot", "vel", "hm"]
    # pipelined calls (upload / kernels / download on three streams): every result belongs to its own input
    scenes = [synth_scene(15000 + 2000 * i, seed=20 + i) for i in range(3)]
    hosts = [(torch.from_numpy(sc).pin_memory(), torch.tensor([0, len(sc)], dtype=torch.int32).pin_memory()) for sc in scenes]
    results = [m.forward_host(p, o)[0] for p, o in hosts]
    m.host_result_ready.synchronize()
    for (p, o), res in zip(hosts, results):
        want = m.forward_points(p.to(cuda), o.to(cuda))[0]["hm"].permute(0, 2, 3, 1).cpu()
        torch.testing.assert_close(res[0][..., -1:], want, rtol=0, atol=0)


def test_batched_inference_paths_are_bit_identical(cuda, monkeypatch):
    """From ops.SORT_MIN_BATCH scenes per forward the SubM / keyed strided convolutions run on pattern-sorted tiles and
    the stem gathers pre-split rows; both only move data differently -- every head tensor is bit-identical to the
    small-batch path (which the oracle tests above pin)."""
    from futuredet_b200 import ops
    monkeypatch.setattr(ops, "SORT_MIN_ROWS", 0)
    m = build_model(3, cuda).to(cuda)
    m.configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                               max_voxel_num=[120000, 160000]))
    scenes = [synth_scene(30000 + 5000 * i, seed=40 + i) for i in range(3)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(cuda)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=cuda)
    outs = []
    for min_batch in (100, 1):
        monkeypatch.setattr(ops, "SORT_MIN_BATCH", min_batch)
        outs.append(m.forward_points(pts, off))
    for a, b in zip(*outs):
        for k in a:
            assert torch.equal(a[k], b[k]), k


# ---------------------------------------------------------------------------------------------------------------
# The BENCHED configuration, end to end (BASELINE configs[1] / [2] / [4]): full 10-sweep scenes, the 160 k-voxel cap
# hit, the tensor-core arm that bench.py times, every head tensor <= 1e-3 from the fp32 oracle chain.
def _full_model(tasks, timesteps, dev, seed):
    torch.manual_seed(seed)
    cfg = dict(
        type="VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=dict(type="CenterHead", in_channels=512, tasks=tasks, dataset="nuscenes", weight=0.25,
                       code_weights=[1.0] * 6 + [0.2, 0.2, 1.0, 1.0],
                       common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                       share_conv_channel=64, dcn_head=False, timesteps=timesteps, classify=False))
    m = fb.build_detector(cfg).eval()
    randomise_bn(m, seed + 1)
    return m


def _oracle_chain(sd, vox, B, n_tasks):
    bev = S.backbone_forward({k[9:]: v for k, v in sd.items() if k.startswith("backbone.")},
                             torch.from_numpy(vox["features"]), vox["coords"], B, [1440, 1440, 40])
    feat = D.rpn_forward({k[5:]: v for k, v in sd.items() if k.startswith("neck.")}, bev, [5, 5], [1, 2], [1, 2])
    names = [["reg", "height", "dim", "rot", "vel", "hm"]] * n_tasks
    return D.center_head_forward({k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}, feat, names)


def _run_points(m, scenes, dev):
    m.to(dev).configure_voxelizer(dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                                       max_voxel_num=[120000, 160000]))
    pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
    with torch.no_grad():
        return m.forward_points(pts, off, return_voxels=True)


def test_benched_configuration_matches_oracle(cuda):
    """2 full 305 k-point scenes (160 k-voxel cap hit in both), product default precision (bf16x3 / tcgen05), the
    forecast_n0 head and the 7-timestep forecast_n3 head on the same backbone + neck weights."""
    assert fb.default_precision() == "bf16x3"
    scenes = [synth_scene(360000, seed=11), synth_scene(360000, seed=12)]
    assert all(len(s) > 300000 for s in scenes)
    vox = V.voxelize_batch_c(scenes, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    assert list(vox["num_voxels"]) == [160000, 160000]                     # the cap is hit
    m3 = _full_model([dict(num_class=1, class_names=["car"])], 7, cuda, seed=5)
    sd3 = {k: v.clone() for k, v in m3.state_dict().items()}
    m0 = _full_model([dict(num_class=1, class_names=["car"])], 1, cuda, seed=6)
    m0.backbone.load_state_dict(m3.backbone.state_dict()); m0.neck.load_state_dict(m3.neck.state_dict())
    m0.bbox_head.shared_conv.load_state_dict(m3.bbox_head.shared_conv.state_dict())
    sd0 = {k: v.clone() for k, v in m0.state_dict().items()}
    want3 = _oracle_chain(sd3, vox, 2, 1)
    # n0 differs from n3 only below the shared conv: reuse the oracle's features by recomputing just the head would need
    # the intermediate; the chain is cheap enough (a few seconds per scene) to run twice
    want0 = _oracle_chain(sd0, vox, 2, 1)
    for m, want, T in ((m3, want3, 7), (m0, want0, 1)):
        preds, voxd = _run_points(m, scenes, cuda)
        n = int(voxd["total"].item())
        assert n == 320000 and np.array_equal(voxd["coords"][:n].cpu().numpy(), vox["coords"])
        assert preds[0]["vel"].shape == (2, 2 * T, 180, 180)
        for k, v in want[0].items():
            err = float((preds[0][k].cpu() - v).abs().max())
            assert err <= TOL, "T=%d %s: max abs err %g" % (T, k, err)


def test_two_task_car_ped_500k_points_matches_oracle(cuda):
    """BASELINE configs[4]: mixed car + pedestrian heads (two SepHeads, center_head.py:351-372) on a 500 k-point
    dense scene."""
    scene = synth_scene(590000, seed=21)
    assert len(scene) >= 500000
    vox = V.voxelize_batch_c([scene], NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    tasks = [dict(num_class=1, class_names=["car"]), dict(num_class=1, class_names=["pedestrian"])]
    m = _full_model(tasks, 7, cuda, seed=8)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    want = _oracle_chain(sd, vox, 1, 2)
    preds, voxd = _run_points(m, [scene], cuda)
    n = int(voxd["total"].item())
    assert np.array_equal(voxd["coords"][:n].cpu().numpy(), vox["coords"])
    assert len(preds) == 2
    for t in range(2):
        for k, v in want[t].items():
            err = float((preds[t][k].cpu() - v).abs().max())
            assert err <= TOL, "task %d %s: max abs err %g" % (t, k, err)
