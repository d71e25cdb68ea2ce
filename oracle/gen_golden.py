"""ORACLE tooling (test infrastructure only): generate tests/golden/* by running the REFERENCE itself.

Runs only in the build container, where /root/reference exists; the fixtures it writes are committed so
that nothing on the GPU box needs the reference.  Usage:  python oracle/gen_golden.py

  voxel_<case>.npz   reference numba points_to_voxel (det3d/ops/point_cloud/point_cloud_ops.py:112-184)
                     + reference VoxelFeatureExtractorV3 (det3d/models/readers/voxel_encoder.py:17-24)
  neck_head.pt       reference RPN (det3d/models/necks/rpn.py) + CenterHead forward & loss
                     (det3d/models/bbox_heads/center_head.py:375-539) at reduced widths, random BN statistics
  predict.pt         reference `CenterHead.predict` (center_head.py:541-770: decode, masks, per-timestep merge) on
                     object-like head tensors, for a 1-timestep head (replicated 7x) and a 7-timestep head; only the
                     CUDA call inside rotate_nms_pcdet is replaced (oracle.predict_ref.rotate_nms_ref + float64 IoU)
                     (`python oracle/gen_golden.py predict` regenerates only this file)
  loader.npz         reference `LoadPointCloudFromFile.__call__` (det3d/datasets/pipelines/loading.py:102-147) on synthetic
                     nuScenes-format .bin sweeps written to a temp dir (`python oracle/gen_golden.py loader`)
  assign.npz         reference `AssignLabel.__call__` (det3d/datasets/pipelines/preprocess.py:336-909, standard sampler)
                     on synthetic car annotations, 3 timesteps (`python oracle/gen_golden.py assign`)
  state_keys.json    parameter / buffer names and shapes of the REFERENCE modules: RPN, CenterHead (n0 and n3 heads) built
                     from det3d.models, and SpMiddleResNetFHD built by executing the reference's own
                     det3d/models/backbones/scn.py with `spconv` bound to the spconv-1.x-shaped classes of this repo
                     (checkpoint compatibility; `python oracle/gen_golden.py keys`)
  backbone_scn.pt    the reference's OWN backbone source (det3d/models/backbones/scn.py:37-176, SparseBasicBlock +
                     SpMiddleResNetFHD, executed unmodified) run on the CPU with `spconv` bound to oracle/spconv_shim.py:
                     input voxels, state_dict, BEV output and per-stage features / indices on a reduced x/y grid
                     (`python oracle/gen_golden.py backbone`)
  head_variants.pt   reference CenterHead in the n3dtf (dense + forecast_feature) and n3dtfm (+ bev_map) modes
                     (center_head.py:99-124,268-390): forward incl. `feats`, dense loss (:413-415,485-507)
                     (`python oracle/gen_golden.py variants`)
  neck_head_train.pt the same reference classes in TRAINING mode: loss dict, every parameter gradient after
                     `sum(loss["loss"]).backward()` (trainer.py:85,317-344), the input gradient and the updated
                     BatchNorm running statistics  (`python oracle/gen_golden.py train` regenerates only this file)
"""
import importlib.util
import logging
import os
import sys
import types

import numpy as np
import torch
import torchvision  # noqa: F401  (must be imported before the reference path is appended)
import statistics   # noqa: F401  (real stdlib module first: /root/reference/statistics.py would shadow it)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, random_points, synth_scene  # noqa: E402


def load_ref_voxelizer():
    spec = importlib.util.spec_from_file_location("pco", REF + "/det3d/ops/point_cloud/point_cloud_ops.py")
    pco = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pco)
    return pco


def import_ref_models():
    class AttrDict(dict):
        def __init__(self, *a, **k):
            super().__init__()
            for key, v in dict(*a, **k).items():
                self[key] = AttrDict(v) if isinstance(v, dict) and not isinstance(v, AttrDict) else v
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    def shim(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m
    shim("addict", Dict=AttrDict)
    shim("terminaltables", AsciiTable=object)
    shim("pycocotools").mask = shim("pycocotools.mask")
    sys.path.append(REF)          # append, not insert
    import det3d.ops
    pkg = types.ModuleType("det3d.ops.iou3d_nms")
    pkg.__path__ = [REF + "/det3d/ops/iou3d_nms"]
    sys.modules["det3d.ops.iou3d_nms"] = pkg
    pkg.iou3d_nms_cuda = shim("det3d.ops.iou3d_nms.iou3d_nms_cuda")
    import det3d.models as M
    return M


VOXEL_CASES = {
    # name: (points, max_voxels)
    "random": lambda: (random_points(12000, seed=0), 160000),
    "boundary": lambda: (random_points(12000, seed=1, snap_frac=0.25), 160000),
    "pile": lambda: (random_points(12000, seed=2, pile=1500), 160000),
    "cap": lambda: (random_points(12000, seed=3), 1500),
    "scene": lambda: (synth_scene(16000, seed=4), 5000),
}

NECK_CFG = dict(type="RPN", layer_nums=[1, 2], ds_layer_strides=[1, 2], ds_num_filters=[8, 16], us_layer_strides=[1, 2],
                us_num_filters=[16, 16], num_input_features=8)
HEAD_CFG = dict(type="CenterHead", in_channels=32, tasks=[dict(num_class=1, class_names=["car"])], dataset="nuscenes",
                weight=0.25, code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0],
                common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2), "vel": (2, 2)},
                share_conv_channel=64, dcn_head=False, timesteps=3, two_stage=False, reverse=False, sparse=False,
                dense=False, bev_map=False, forecast_feature=False, classify=False, wide_head=False)


def randomise_bn(module, gen):
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)


def make_targets(B, H, W, timesteps, gen, max_objs=20):
    """example[...] targets in the collate layout: key -> [timestep][task] -> Tensor[B,...] (collate.py:208-232)."""
    ex = {k: [] for k in ("hm", "anno_box", "ind", "mask", "cat")}
    for _ in range(timesteps):
        hm = torch.rand((B, 1, H, W), generator=gen) ** 8
        ind = torch.randint(0, H * W, (B, max_objs), generator=gen)
        mask = (torch.rand((B, max_objs), generator=gen) < 0.6).to(torch.uint8)
        for b in range(B):
            hm[b, 0].view(-1)[ind[b][mask[b].bool()]] = 1.0
        ex["hm"].append([hm])
        ex["anno_box"].append([torch.randn((B, max_objs, 14), generator=gen)])
        ex["ind"].append([ind])
        ex["mask"].append([mask])
        ex["cat"].append([torch.zeros((B, max_objs), dtype=torch.int64)])
    return ex


def gen_train(M):
    """Reference RPN + CenterHead in training mode: forward, loss, backward."""
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(7)
    neck = M.build_neck(dict(NECK_CFG, logger=logging.getLogger("RPN")))
    head = M.build_head(dict(HEAD_CFG))
    randomise_bn(neck, gen)
    randomise_bn(head, gen)
    neck.train()
    head.train()
    init = dict(neck_state={k: v.clone() for k, v in neck.state_dict().items()},
                head_state={k: v.clone() for k, v in head.state_dict().items()})
    x = torch.randn((2, 8, 12, 12), generator=gen).requires_grad_(True)
    feat = neck(x)
    preds = head(feat)
    example = make_targets(2, feat.shape[2], feat.shape[3], 3, gen)
    loss = head.loss(example, preds)
    total = sum(loss["loss"])
    total.backward()
    grads = {"neck." + k: p.grad.clone() for k, p in neck.named_parameters()}
    grads.update({"head." + k: p.grad.clone() for k, p in head.named_parameters()})
    torch.save(dict(neck_cfg=NECK_CFG, head_cfg=HEAD_CFG, x=x.detach(), x_grad=x.grad.clone(), example=example,
                    total=total.detach(), grads=grads,
                    loss={k: [t.detach() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else v
                          for k, v in loss.items()},
                    neck_state_after={k: v.clone() for k, v in neck.state_dict().items()},
                    head_state_after={k: v.clone() for k, v in head.state_dict().items()}, **init),
               os.path.join(OUT, "neck_head_train.pt"))
    print("neck_head_train: total loss %.6f, %d parameter gradients" % (float(total), len(grads)))


TEST_CFG = dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_per_img=500,
                nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=83,
                         nms_iou_threshold=0.2),
                score_threshold=0.1, pc_range=[-54, -54], out_size_factor=8, voxel_size=[0.075, 0.075])


def gen_predict(M):
    """Reference CenterHead.predict with the CUDA NMS call swapped for the oracle's (same box convention / ordering)."""
    from det3d.core import box_torch_ops
    from oracle import predict_ref as PR
    box_torch_ops.rotate_nms_pcdet = lambda boxes, scores, thresh, pre_maxsize=None, post_max_size=None: \
        PR.rotate_nms_ref(boxes, scores, thresh, pre_maxsize, post_max_size, PR.nms_np)
    AttrDict = sys.modules["addict"].Dict
    cases = {}
    for name, T, H, W in (("t1", 1, 40, 40), ("t7", 7, 36, 44)):   # the reference merge (:693-713) only works for T in {1, 7}
        head = M.build_head(dict(HEAD_CFG, timesteps=T))
        preds = PR.synth_preds(2, H, W, T, seed=10 + T, n_obj=12)
        ret = head.predict({}, [{k: v.clone() for k, v in preds.items()}], AttrDict(TEST_CFG))
        cases[name] = dict(timesteps=T, preds=preds,
                           ret=[{k: v for k, v in r.items() if k != "metadata"} for r in ret])
        print("predict %s: %s boxes per sample" % (name, [len(r["scores"]) for r in ret]))
    torch.save(dict(test_cfg=TEST_CFG, cases=cases), os.path.join(OUT, "predict.pt"))


def gen_loader():
    """Run the reference loader on synthetic .bin files.  det3d.datasets pulls the nuscenes devkit at import time, so
    loading.py is executed inside stub packages (only its own code runs: read_file / remove_close / read_sweep /
    LoadPointCloudFromFile)."""
    import tempfile
    from oracle import loader_ref as LR
    import_ref_models()
    for name in ("det3d.datasets", "det3d.datasets.pipelines"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [REF + "/" + name.replace(".", "/")]
        sys.modules[name] = pkg
    reg = types.ModuleType("det3d.datasets.registry")

    class _Reg:
        def register_module(self, cls):
            return cls
    reg.PIPELINES = _Reg()
    sys.modules["det3d.datasets.registry"] = reg
    spec = importlib.util.spec_from_file_location("det3d.datasets.pipelines.loading",
                                                  REF + "/det3d/datasets/pipelines/loading.py")
    LD = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = LD
    spec.loader.exec_module(LD)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case, seed in (("a", 0), ("b", 1)):
            key, sweeps = LR.synth_sweeps(seed)
            kp = os.path.join(tmp, "key_%s.bin" % case)
            key.tofile(kp)
            infos = []
            for i, (rec, T, lag) in enumerate(sweeps):
                sp = os.path.join(tmp, "sweep_%s_%d.bin" % (case, i))
                rec.tofile(sp)
                infos.append(dict(lidar_path=sp, transform_matrix=T, time_lag=lag))
            res = dict(lidar=dict(nsweeps=len(sweeps) + 1), painted=False)
            LD.LoadPointCloudFromFile(dataset="NuScenesDataset")(res, dict(lidar_path=kp, sweeps=infos))
            # the reference visits the sweeps in rng.choice order (loading.py:121-122): record it so the oracle and
            # the native path can be fed the same order
            order = np.random.default_rng(0).choice(len(infos), len(infos), replace=False)
            out["combined_" + case] = res["lidar"]["combined"]
            out["order_" + case] = order
            out["seed_" + case] = np.int64(seed)
            print("loader %s: %d points" % (case, len(res["lidar"]["combined"])))
    np.savez_compressed(os.path.join(OUT, "loader.npz"), **out)


def _stub_dataset_packages():
    import_ref_models()
    for name in ("det3d.datasets", "det3d.datasets.pipelines"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [REF + "/" + name.replace(".", "/")]
        sys.modules[name] = pkg
    reg = types.ModuleType("det3d.datasets.registry")

    class _Reg:
        def register_module(self, cls):
            return cls
    reg.PIPELINES = _Reg()
    sys.modules["det3d.datasets.registry"] = reg
    b = types.ModuleType("det3d.builder")           # det3d.builder drags det3d.solver (py3.12-incompatible import)
    b.build_dbsampler = None
    sys.modules["det3d.builder"] = b


ASSIGN_CFG = dict(out_size_factor=8, gaussian_overlap=0.1, max_objs=500, min_radius=2, radius_mult=False,
                  sampler_type="standard")


def gen_assign():
    """Run the reference AssignLabel on synthetic annotations (2 samples x 3 timesteps, one car task)."""
    from oracle import assign_ref as AR
    _stub_dataset_packages()
    spec = importlib.util.spec_from_file_location("det3d.datasets.pipelines.preprocess",
                                                  REF + "/det3d/datasets/pipelines/preprocess.py")
    PP = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = PP
    spec.loader.exec_module(PP)
    AttrDict = sys.modules["addict"].Dict
    out = {}
    for radius_mult in (False, True):
        cfg = AttrDict(dict(ASSIGN_CFG, radius_mult=radius_mult))
        cfg.target_assigner = AttrDict(tasks=None)
        cfg.target_assigner.tasks = [AttrDict(num_class=1, class_names=["car"])]
        al = PP.AssignLabel(cfg=cfg)
        for sample, seed in enumerate((0, 1)):
            boxes = AR.synth_annotations(seed)
            n = len(boxes[0])
            res = dict(mode="train", type="NuScenesDataset", lidar=dict(
                voxels=dict(shape=np.array([1440, 1440, 40]), range=np.array(NUSC_RANGE, np.float32),
                            size=np.array(NUSC_VOXEL, np.float32)),
                annotations=dict(gt_boxes=[b.copy() for b in boxes], gt_names=[np.array(["car"] * n)] * 3,
                                 gt_classes=[np.ones(n, np.int32) for _ in range(3)],
                                 gt_trajectory=[np.array(["static"] * n)] * 3)))
            res, _ = al(res, {})
            tg = res["lidar"]["targets"]
            tag = "%d_%d" % (int(radius_mult), sample)
            for key in ("hm", "anno_box", "ind", "mask", "cat"):
                out[key + "_" + tag] = np.stack([tg[key][t][0] for t in range(3)])
            print("assign rm=%d sample %d: %d objects placed at t0" % (radius_mult, sample, int(tg["mask"][0][0].sum())))
    # sampler_type = "trajectory" (the n3dtf / n3dtfm configs, preprocess.py:573-897): the standard targets plus
    # `*_trajectory` (one 3-class task: static / linear / nonlinear) and `*_forecast` (one 7-class task: the boxes of all
    # timesteps, class = timestep) -- same per-task routine on regrouped classes
    for radius_mult in (False, True):
        cfg = AttrDict(dict(ASSIGN_CFG, radius_mult=radius_mult, sampler_type="trajectory"))
        cfg.target_assigner = AttrDict(tasks=None)
        cfg.target_assigner.tasks = [AttrDict(num_class=1, class_names=["car"])]
        al = PP.AssignLabel(cfg=cfg)
        boxes = AR.synth_annotations(2)
        n = len(boxes[0])
        traj = AR.synth_trajectories(2, n)
        res = dict(mode="train", type="NuScenesDataset", lidar=dict(
            voxels=dict(shape=np.array([1440, 1440, 40]), range=np.array(NUSC_RANGE, np.float32),
                        size=np.array(NUSC_VOXEL, np.float32)),
            annotations=dict(gt_boxes=[b.copy() for b in boxes], gt_names=[np.array(["car"] * n)] * 3,
                             gt_classes=[np.ones(n, np.int32) for _ in range(3)],
                             gt_trajectory=[traj.copy() for _ in range(3)])))
        res, _ = al(res, {})
        tg = res["lidar"]["targets"]
        tag = "traj%d" % int(radius_mult)
        for suffix in ("", "_trajectory", "_forecast"):
            for key in ("hm", "anno_box", "ind", "mask", "cat"):
                out[key + suffix + "_" + tag] = np.stack([tg[key + suffix][t][0] for t in range(3)])
        print("assign trajectory rm=%d: %d / %d / %d objects placed at t0 (standard / trajectory / forecast)" %
              (radius_mult, int(tg["mask"][0][0].sum()), int(tg["mask_trajectory"][0][0].sum()),
               int(tg["mask_forecast"][0][0].sum())))
    # heat maps are sparse: store them compressed
    np.savez_compressed(os.path.join(OUT, "assign.npz"), **out)


def load_ref_scn(spconv_module):
    """Execute the reference's det3d/models/backbones/scn.py with `spconv` bound to `spconv_module`."""
    import_ref_models()
    sp = types.ModuleType("spconv")
    for n in ("SparseConvTensor", "SubMConv3d", "SparseConv3d", "SparseSequential", "SparseModule"):
        setattr(sp, n, getattr(spconv_module, n))
    sys.modules["spconv"] = sp
    spec = importlib.util.spec_from_file_location("det3d.models.backbones.scn_ref", REF + "/det3d/models/backbones/scn.py")
    scn = importlib.util.module_from_spec(spec)
    scn.__package__ = "det3d.models.backbones"
    sys.modules[spec.name] = scn
    spec.loader.exec_module(scn)
    return scn


BACKBONE_GRID = [96, 80, 40]       # (x, y, z): z as in the configs (41 -> 21 -> 11 -> 5 -> 2), x / y reduced


def gen_backbone():
    """Reference scn.py forward (eval mode) over the CPU spconv shim."""
    from oracle import spconv_shim
    scn = load_ref_scn(spconv_shim)
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(5)
    from oracle.spconv_ref import seeded_state
    bb = scn.SpMiddleResNetFHD(num_input_features=5, ds_factor=8)
    bb.load_state_dict(seeded_state(bb, 33))        # the fixture stores the seed, not 11 MB of weights
    bb.eval()
    rng = np.random.default_rng(11)
    gx, gy, gz = BACKBONE_GRID
    B = 2
    coors = []
    for b in range(B):
        n = 2600 - 700 * b
        # clustered occupancy (objects + ground), unique cells, arbitrary (non-sorted) order as the voxelizer emits
        cz = np.clip(rng.normal(12, 7, n), 0, gz - 1).astype(np.int64)
        cy = np.clip(rng.normal(gy / 2, gy / 4, n), 0, gy - 1).astype(np.int64)
        cx = np.clip(rng.normal(gx / 2, gx / 4, n), 0, gx - 1).astype(np.int64)
        lin = np.unique((cz * gy + cy) * gx + cx)
        rng.shuffle(lin)
        c = np.stack([np.full(len(lin), b), lin // (gy * gx), (lin // gx) % gy, lin % gx], 1)
        coors.append(c)
    coors = np.concatenate(coors).astype(np.int32)
    feats = torch.from_numpy(rng.normal(0, 1, (len(coors), 5)).astype(np.float32))
    with torch.no_grad():
        out, stages = bb(feats, torch.from_numpy(coors), B, BACKBONE_GRID)
    torch.save(dict(grid=BACKBONE_GRID, batch_size=B, features=feats, coors=torch.from_numpy(coors), state_seed=33,
                    state_shapes={k: tuple(v.shape) for k, v in bb.state_dict().items()}, out=out,
                    stages={k: dict(features=v.features.clone() if k == "conv4" else None,
                                    feature_sum=v.features.double().sum(0).float(),
                                    indices=torch.as_tensor(np.asarray(v.indices)).clone(),
                                    spatial_shape=list(v.spatial_shape)) for k, v in stages.items()}),
               os.path.join(OUT, "backbone_scn.pt"))
    print("backbone_scn: %d voxels -> out %s, stage rows %s" % (len(coors), tuple(out.shape),
                                                              {k: len(v.features) for k, v in stages.items()}))


def gen_head_variants(M):
    """Reference CenterHead in the dense + forecast_feature (+ bev_map) modes of the n3dtf / n3dtfm configs."""
    out = {}
    for name, bev in (("n3dtf", False), ("n3dtfm", True)):
        torch.manual_seed(21)
        gen = torch.Generator().manual_seed(22)
        cfg = dict(HEAD_CFG, dense=True, forecast_feature=True, bev_map=bev)
        from oracle.spconv_ref import seeded_state
        head = M.build_head(dict(cfg))
        sd = seeded_state(head, 44)
        for k in sd:
            if k.endswith("hm.3.bias"):
                sd[k] = sd[k] - 2.19
        head.load_state_dict(sd)
        head.eval()
        x = torch.randn((2, 32, 12, 10), generator=gen)
        bm = torch.rand((2, 6, 12, 10), generator=gen) if bev else None
        with torch.no_grad():
            preds = head(x, bm)
            example = make_targets(2, 12, 10, 3, gen)
            loss = head.loss(example, [{k: v.clone() for k, v in p.items()} for p in preds])
        out[name] = dict(cfg=cfg, state_seed=44, state_shapes={k: tuple(v.shape) for k, v in head.state_dict().items()},
                         x=x, bev_map=bm,
                         preds=[{k: v.clone() for k, v in p.items()} for p in preds], example=example,
                         loss={k: [t.detach() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else v
                               for k, v in loss.items()})
        print("head %s: %d tasks, keys %s, loss %s" % (name, len(preds), sorted(preds[0]), [float(l) for l in loss["loss"]]))
    # dense-mode predict (center_head.py:606-607,693-713): every task decoded + NMS-ed on its own, labels offset per task
    from det3d.core import box_torch_ops
    from oracle import predict_ref as PR
    box_torch_ops.rotate_nms_pcdet = lambda boxes, scores, thresh, pre_maxsize=None, post_max_size=None: \
        PR.rotate_nms_ref(boxes, scores, thresh, pre_maxsize, post_max_size, PR.nms_np)
    AttrDict = sys.modules["addict"].Dict
    head = M.build_head(dict(HEAD_CFG, dense=True, forecast_feature=True))
    tasks = [PR.synth_preds(2, 36, 44, 1, seed=50 + t, n_obj=10) for t in range(3)]
    ret = head.predict({}, [{k: v.clone() for k, v in p.items()} for p in tasks], AttrDict(TEST_CFG))
    out["dense_predict"] = dict(test_cfg=TEST_CFG, preds=tasks,
                                ret=[{k: v for k, v in r.items() if k != "metadata"} for r in ret])
    print("dense predict: %s boxes per sample" % [len(r["scores"]) for r in ret])
    torch.save(out, os.path.join(OUT, "head_variants.pt"))


def gen_state_keys():
    import json
    M = import_ref_models()
    from futuredet_b200 import sparse as fsp
    sp = types.ModuleType("spconv")
    for n in ("SparseConvTensor", "SubMConv3d", "SparseConv3d", "SparseSequential", "SparseModule"):
        setattr(sp, n, getattr(fsp, n))
    sys.modules["spconv"] = sp
    spec = importlib.util.spec_from_file_location("det3d.models.backbones.scn_ref", REF + "/det3d/models/backbones/scn.py")
    scn = importlib.util.module_from_spec(spec)
    scn.__package__ = "det3d.models.backbones"
    sys.modules[spec.name] = scn
    spec.loader.exec_module(scn)
    out = {}
    bb = scn.SpMiddleResNetFHD(num_input_features=5, ds_factor=8)
    out["backbone"] = {k: list(v.shape) for k, v in bb.state_dict().items()}
    neck = M.build_neck(dict(type="RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                             us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256,
                             logger=logging.getLogger("RPN")))
    out["neck"] = {k: list(v.shape) for k, v in neck.state_dict().items()}
    for name, T in (("head_n0", 1), ("head_n3", 7)):
        head = M.build_head(dict(HEAD_CFG, in_channels=512, timesteps=T))
        out[name] = {k: list(v.shape) for k, v in head.state_dict().items()}
    for name, flags in (("head_n3dtf", dict(dense=True, forecast_feature=True)),
                        ("head_n3dtfm", dict(dense=True, forecast_feature=True, bev_map=True)),
                        ("head_two_stage", dict(two_stage=True)), ("head_wide", dict(wide_head=True)),
                        ("head_classify", dict(classify=True)), ("head_sparse", dict(sparse=True))):
        head = M.build_head(dict(HEAD_CFG, in_channels=512, timesteps=7, **flags))
        out[name] = {k: list(v.shape) for k, v in head.state_dict().items()}
    json.dump(out, open(os.path.join(OUT, "state_keys.json"), "w"), indent=0, sort_keys=True)
    print("state keys:", {k: len(v) for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    if "keys" in sys.argv[1:]:
        gen_state_keys()
        return
    if "assign" in sys.argv[1:]:
        gen_assign()
        return
    if "backbone" in sys.argv[1:]:
        gen_backbone()
        return
    if "variants" in sys.argv[1:]:
        gen_head_variants(import_ref_models())
        return
    if "loader" in sys.argv[1:]:
        gen_loader()
        return
    if "train" in sys.argv[1:]:
        gen_train(import_ref_models())
        return
    if "predict" in sys.argv[1:]:
        gen_predict(import_ref_models())
        return
    pco = load_ref_voxelizer()
    M = import_ref_models()
    from det3d.models.readers.voxel_encoder import VoxelFeatureExtractorV3
    vfe = VoxelFeatureExtractorV3(num_input_features=5)
    for name, make in VOXEL_CASES.items():
        pts, mv = make()
        voxels, coors, npts = pco.points_to_voxel(pts, np.array(NUSC_VOXEL, np.float32), np.array(NUSC_RANGE, np.float32),
                                                  10, True, mv)
        mean = vfe(torch.from_numpy(voxels), torch.from_numpy(npts)).numpy()
        np.savez_compressed(os.path.join(OUT, "voxel_%s.npz" % name), points=pts, max_voxels=np.int64(mv),
                            coors=coors, num_points=npts, mean=mean)
        print("voxel_%s: %d pts -> %d voxels" % (name, len(pts), len(coors)))

    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1)
    neck = M.build_neck(dict(NECK_CFG, logger=logging.getLogger("RPN")))
    head = M.build_head(dict(HEAD_CFG))
    randomise_bn(neck, gen)
    randomise_bn(head, gen)
    neck.eval()
    head.eval()
    x = torch.randn((2, 8, 12, 12), generator=gen)
    with torch.no_grad():
        feat = neck(x)
        preds = head(feat)
        example = make_targets(2, feat.shape[2], feat.shape[3], 3, gen)
        preds_for_loss = [{k: v.clone() for k, v in p.items()} for p in preds]
        loss = head.loss(example, preds_for_loss)
    torch.save(dict(neck_cfg=NECK_CFG, head_cfg=HEAD_CFG, neck_state=neck.state_dict(), head_state=head.state_dict(),
                    x=x, neck_out=feat, preds=[{k: v for k, v in p.items()} for p in preds], example=example,
                    loss={k: [t.detach() if torch.is_tensor(t) else t for t in v] if isinstance(v, list) else v
                          for k, v in loss.items()}),
               os.path.join(OUT, "neck_head.pt"))
    print("neck_head: neck_out", tuple(feat.shape), {k: tuple(v.shape) for k, v in preds[0].items()})
    print("loss keys", {k: (v[0] if isinstance(v, list) else v) for k, v in loss.items()})
    gen_train(M)
    gen_predict(M)
    gen_loader()
    gen_assign()
    gen_head_variants(M)
    gen_backbone()
    gen_state_keys()


if __name__ == "__main__":
    main()
