"""CPU: host-side logic of the widened path that needs no kernel -- gradient buckets, the sweep-batch description handed
to fd_assemble_sweeps, head-tensor layout detection of predict, config-driven precision selection."""
import numpy as np
import torch
from torch import nn

from futuredet_b200 import loader, predict, train


def test_grad_buckets_layout_and_ready_callbacks():
    model = nn.Sequential(nn.Linear(4, 6), nn.BatchNorm1d(6), nn.Linear(6, 2))
    params = list(model.parameters())
    gb = train.GradBuckets(params, bucket_bytes=4 * 20)              # 24+6 | 6+6+12 | 2  -> buckets close at >= 20 floats
    # buckets follow REVERSE registration order (the order backward finalises gradients)
    order = [id(p) for _, plist in gb.buckets for p in plist]
    assert order == [id(p) for p in reversed(params)]
    assert sum(f.numel() for f, _ in gb.buckets) == sum(p.numel() for p in params)
    for p in params:                                                 # param.grad is a view of its bucket
        b, view = gb.where[id(p)]
        assert p.grad.data_ptr() == view.data_ptr() and view.shape == p.shape
        flat = gb.buckets[b][0]
        assert flat.data_ptr() <= view.data_ptr() < flat.data_ptr() + flat.numel() * 4
    fired = []
    gb.on_bucket_ready = lambda i, flat: fired.append(i)
    gb.zero()
    for p in reversed(params):
        gb.grad(p).fill_(1.0)
        gb.done(p)
    assert fired == list(range(len(gb.buckets)))                     # each bucket exactly once, in backward order
    gb.zero()
    assert all(float(f.abs().sum()) == 0.0 for f, _ in gb.buckets)
    fired.clear()
    gb.done(params[-1])                                              # only one gradient arrived: flush reports the rest
    gb.flush()
    assert sorted(fired) == list(range(len(gb.buckets)))
    # attach=False leaves param.grad to autograd (loss.backward() bridge)
    for p in params:
        p.grad = None
    train.GradBuckets(params, attach=False)
    assert all(p.grad is None for p in params)


def test_sweep_batch_description():
    sb = loader.SweepBatch()
    key = np.zeros((7, 5), np.float32)
    T = np.eye(4); T[0, 3] = 2.0
    sb.add_scene(key, [(np.ones((3, 5), np.float32), T, 0.05), (np.ones((0, 5), np.float32), None, 0.1)])
    sb.add_scene(key[:2], [])
    assert sb.offsets == [0, 7, 10, 10, 12] and sb.scene == [0, 0, 0, 1] and sb.n_scenes == 2
    assert sb.flags == [0, 3, 2, 0]                                  # key frames: nothing; sweeps: remove_close (+ transform)
    assert sb.lags == [0.0, 0.05, 0.1, 0.0]
    assert np.array_equal(sb.xforms[1], T) and np.array_equal(sb.xforms[2], np.eye(4))


def test_predict_layout_packs_separate_tensors():
    names = ["reg", "height", "dim", "rot", "vel", "hm"]
    chans = dict(reg=2, height=1, dim=3, rot=2, vel=14, hm=1)
    p = {n: torch.randn(2, c, 5, 6) for n, c in chans.items()}
    base, off = predict._channels_last_layout(p, names)
    assert base.shape == (2, 5, 6, 23) and off == dict(reg=0, height=2, dim=3, rot=6, vel=8, hm=22)
    for n in names:
        assert torch.equal(base[..., off[n]:off[n] + chans[n]], p[n].permute(0, 2, 3, 1))


def test_training_precision_selection():
    assert train._prec_for("fp32", 64, 27) == "fp32"
    assert train._prec_for("bf16x3", 5, 27) == "fp32"                # stem: no tensor-core tile for Cin = 5
    assert train._prec_for("bf16x3", 64, 27) == "bf16x3"
    t = torch.zeros(8, 23)[:, 2:3]                                   # channel slice of a head output: unaligned rows
    assert train._prec_for("bf16x3", 64, 9, t) == "fp32"
