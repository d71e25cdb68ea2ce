"""GPU: CUDA-graph replay of the forward pass and of the training step reproduces the eager path (bit-exact forward;
training within the fp32-atomics noise of the weight gradient), for inputs of different sizes fed to one captured graph."""
import numpy as np
import pytest
import torch

from futuredet_b200 import graphs, train
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene, synth_targets

pytestmark = pytest.mark.gpu
VOX = dict(range=NUSC_RANGE, voxel_size=NUSC_VOXEL, max_points_in_voxel=10, max_voxel_num=[120000, 160000])


def scene_tensors(seed, n, dev):
    sc = synth_scene(n, seed=seed)
    return torch.from_numpy(sc).to(dev), torch.tensor([0, len(sc)], dtype=torch.int32, device=dev)


def test_graphed_forward_matches_eager(cuda):
    from test_gpu_train import build_model
    if True:
        model = build_model(1, cuda).to(cuda).eval().set_precision("bf16x3")
        model.configure_voxelizer(VOX)
        gf = graphs.GraphedForward(model, max_points=40000, batch_size=1)
        for seed, n in ((0, 30000), (1, 36000), (2, 20000)):          # one graph, three different clouds (growing and shrinking)
            pts, off = scene_tensors(seed, n, cuda)
            with torch.no_grad():
                want = {k: v.clone() for k, v in model.forward_points(pts, off)[0].items()}
            got = gf(pts, off)[0]
            for k in want:
                assert torch.equal(got[k], want[k]), (seed, k)


def test_graphed_train_step_matches_eager(cuda):
    from test_gpu_train import build_model
    ex = synth_targets(1, 180, 180, 3, n_obj=20, max_objs=50, seed=0)
    ex = {k: [[t.to(cuda) for t in row] for row in v] for k, v in ex.items()}
    pts, off = scene_tensors(3, 30000, cuda)

    def run(graphed):
        model = build_model(3, cuda).to(cuda).train()
        model.configure_voxelizer(VOX, training=True)
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 0.0                                     # warm-up steps of the capture must not move the statistics
        tr = train.NativeTrainer(model, precision="bf16x3")
        opt = torch.optim.SGD(model.parameters(), lr=1e-3)
        step = graphs.GraphedTrainStep(tr, max_points=40000, batch_size=1) if graphed else None
        losses = []
        for _ in range(3):
            out = step(ex, pts, off) if graphed else tr.step(ex, points=pts, batch_offsets=off)
            losses.append(float(sum(out["loss"])))
            opt.step()
        return losses, {k: p.detach().clone() for k, p in model.named_parameters()}

    l_eager, p_eager = run(False)
    l_graph, p_graph = run(True)
    assert l_eager[2] < l_eager[0]
    np.testing.assert_allclose(l_graph, l_eager, rtol=2e-4)          # the replayed steps see the updated weights
    # weights after three updates: relative L2 (the weight gradient accumulates with fp32 atomics, and single entries of a
    # training-mode-BatchNorm model are ReLU-flip sensitive to that ordering noise)
    for k in p_eager:
        err, ref = float((p_graph[k] - p_eager[k]).norm()), float(p_eager[k].norm())
        assert err <= 2e-2 * ref + 1e-4, (k, err, ref)


def test_eval_after_training_sees_updated_weights_and_statistics(cuda):
    """Derived-weight caches (tensor-core packs, d-major first neck conv, folded BatchNorm, fused head convs) must follow
    the optimizer and the running statistics: inference before training, a few (graph-replayed) training steps, inference
    again == inference of a FRESH model loaded with the trained state_dict; and frozen parameters do not break the
    native backward."""
    from test_gpu_train import build_model
    from futuredet_b200.synth import synth_targets
    pts, off = scene_tensors(5, 30000, cuda)
    ex = synth_targets(1, 180, 180, 3, seed=1)
    ex = {k: [[t.to(cuda) for t in ts] for ts in v] for k, v in ex.items()}
    model = build_model(3, cuda).to(cuda).eval()
    model.configure_voxelizer(VOX)
    with torch.no_grad():
        before = {k: v.clone() for k, v in model.forward_points(pts, off)[0].items()}      # fills every cache
    model.train()
    for p in model.backbone.conv_input.parameters():
        p.requires_grad_(False)                                                            # frozen stem
    tr = train.NativeTrainer(model, precision="bf16x3")
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    step = graphs.GraphedTrainStep(tr, max_points=40000, batch_size=1)
    for _ in range(3):
        step(ex, pts, off)
        opt.step()
    model.eval()
    with torch.no_grad():
        after = {k: v.clone() for k, v in model.forward_points(pts, off)[0].items()}
    fresh = build_model(3, cuda).to(cuda).eval()
    fresh.load_state_dict(model.state_dict())
    fresh.configure_voxelizer(VOX)
    with torch.no_grad():
        want = fresh.forward_points(pts, off)[0]
    assert any(float((after[k] - before[k]).abs().max()) > 1e-4 for k in after)            # training changed the outputs
    for k in want:
        assert torch.equal(after[k], want[k]), k                                           # no stale derived weights
