"""GPU parity: fused voxelize+VFE kernel (through the C ABI) vs reference goldens and the C oracle.
Indices, voxel order and counts are bit-exact; the fp32 mean is checked to 1e-6."""
import os

import numpy as np
import pytest
import torch

from futuredet_b200 import ops
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, random_points, synth_scene
from oracle import voxelizer as V

pytestmark = pytest.mark.gpu
CASES = ["random", "boundary", "pile", "cap", "scene"]


def run_gpu(scenes, max_voxels, dev, want_voxels=False):
    pts = torch.from_numpy(np.concatenate(scenes, 0)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
    r = ops.voxelize_vfe(pts, off, NUSC_VOXEL, NUSC_RANGE, 10, max_voxels, want_voxels=want_voxels)
    m = int(r["total"].item())
    out = dict(features=r["features"][:m].cpu().numpy(), coords=r["coords"][:m].cpu().numpy(),
               num_points=r["num_points"][:m].cpu().numpy(), num_voxels=r["num_voxels"].cpu().numpy())
    if want_voxels:
        out["voxels"] = r["voxels"][:m].cpu().numpy()
    return out


@pytest.mark.parametrize("case", CASES)
def test_matches_reference_golden(cuda, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "voxel_%s.npz" % case))
    r = run_gpu([g["points"]], int(g["max_voxels"]), cuda)
    assert np.array_equal(r["coords"][:, 1:], g["coors"])
    assert (r["coords"][:, 0] == 0).all()
    assert np.array_equal(r["num_points"], g["num_points"])
    np.testing.assert_allclose(r["features"], g["mean"], rtol=1e-6, atol=1e-6)


def test_config1_50k_bit_exact_vs_oracle(cuda):
    pts = random_points(50000, seed=0, snap_frac=0.1, pile=2000)
    o = V.points_to_voxel_c(pts, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    r = run_gpu([pts], 160000, cuda, want_voxels=True)
    assert np.array_equal(r["coords"][:, 1:], o["coors"])
    assert np.array_equal(r["num_points"], o["num_points"])
    assert np.array_equal(r["voxels"], o["voxels"])                  # padded point lists, bit-exact
    np.testing.assert_allclose(r["features"], o["mean"], rtol=1e-6, atol=1e-6)


def test_ragged_batch_with_empty_scene(cuda):
    scenes = [synth_scene(40000, seed=1), np.zeros((0, 5), np.float32), random_points(9000, seed=5),
              np.full((5, 5), 500.0, np.float32), synth_scene(20000, seed=2)]
    o = V.voxelize_batch_c(scenes, NUSC_VOXEL, NUSC_RANGE, 10, 6000)
    r = run_gpu(scenes, 6000, cuda)
    assert np.array_equal(r["num_voxels"], o["num_voxels"])
    assert r["num_voxels"][0] == 6000 and r["num_voxels"][1] == 0 and r["num_voxels"][3] == 0   # cap, empty, rejected
    assert np.array_equal(r["coords"], o["coords"])
    assert np.array_equal(r["num_points"], o["num_points"])
    np.testing.assert_allclose(r["features"], o["features"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n_target,cap", [(360000, 160000), (360000, 120000), (600000, 160000)])
def test_full_size_scene_vs_oracle(cuda, n_target, cap):
    """BASELINE configs 2-5 sizes (300k / 500k points; val and train caps are both hit)."""
    pts = synth_scene(n_target, seed=0)
    o = V.points_to_voxel_c(pts, NUSC_VOXEL, NUSC_RANGE, 10, cap, want_voxels=False)
    r = run_gpu([pts], cap, cuda)
    assert len(o["coors"]) == cap
    assert np.array_equal(r["coords"][:, 1:], o["coors"])
    assert np.array_equal(r["num_points"], o["num_points"])
    np.testing.assert_allclose(r["features"], o["mean"], rtol=1e-6, atol=1e-6)


def test_size_independent_properties(cuda):
    pts = synth_scene(360000, seed=3)
    r1 = run_gpu([pts], 160000, cuda)
    r2 = run_gpu([pts], 160000, cuda)
    for k in r1:                                                     # run-to-run determinism, bit for bit
        assert np.array_equal(r1[k], r2[k]), k
    c = r1["coords"].astype(np.int64)
    key = (c[:, 1] * 1440 + c[:, 2]) * 1440 + c[:, 3]
    assert len(np.unique(key)) == len(key)                           # no voxel emitted twice
    assert (r1["num_points"] >= 1).all() and (r1["num_points"] <= 10).all()
    # idempotence: voxelizing the voxel means lands every mean in its own voxel
    r3 = run_gpu([np.ascontiguousarray(r1["features"])], 160000, cuda)
    assert np.array_equal(r3["coords"], r1["coords"]) and (r3["num_points"] == 1).all()
    # duplicating the cloud as a second scene gives the same voxels with batch index 1
    rb = run_gpu([pts, pts], 160000, cuda)
    m = len(r1["coords"])
    assert np.array_equal(rb["coords"][m:, 1:], r1["coords"][:, 1:]) and (rb["coords"][m:, 0] == 1).all()
    assert np.array_equal(rb["features"][m:], r1["features"])


def test_voxel_generator_and_pipeline_api(cuda):
    """Legacy entry points keep the reference's return types (voxel_generator.py:19-30, preprocess.py:244-271)."""
    from futuredet_b200.pipelines import Voxelization
    pts = random_points(20000, seed=9)
    o = V.points_to_voxel_c(pts, NUSC_VOXEL, NUSC_RANGE, 10, 160000)
    stage = Voxelization(cfg=dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10,
                                  max_voxel_num=[120000, 160000], double_flip=False))
    res, _ = stage(dict(mode="val", lidar=dict(points=pts)), {})
    v = res["lidar"]["voxels"]
    assert np.array_equal(v["voxels"], o["voxels"]) and np.array_equal(v["coordinates"], o["coors"])
    assert np.array_equal(v["num_points"], o["num_points"])
    assert v["num_voxels"].dtype == np.int64 and v["num_voxels"][0] == len(o["coors"])
    assert list(v["shape"]) == [1440, 1440, 40]
    # reader on padded voxels == VFE oracle
    from futuredet_b200.reader import VoxelFeatureExtractorV3
    mean = VoxelFeatureExtractorV3(5)(torch.from_numpy(v["voxels"]).cuda(), torch.from_numpy(v["num_points"]).cuda())
    np.testing.assert_allclose(mean.cpu().numpy(), o["mean"], rtol=1e-6, atol=1e-6)
