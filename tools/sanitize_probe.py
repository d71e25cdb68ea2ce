"""Small instances of every hand-written kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(run on the GPU box; sizes are kept small because the sanitizers slow kernels down 10-100x).
  compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futuredet_b200 as fb  # noqa: E402
from futuredet_b200 import ops, predict as P, train_ops as T  # noqa: E402
from futuredet_b200.synth import NUSC_RANGE, NUSC_VOXEL, synth_scene  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
if os.environ.get("FD_NO_WATCHDOG"):          # the sanitizers slow kernels down enough to trip the 1 s mbarrier watchdog
    import ctypes as C
    from futuredet_b200 import lib as _lib
    _L = _lib.load()
    _L.fd_debug_set_tc.argtypes = [C.c_int, C.c_int]
    torch.zeros(1, device=dev)
    assert _L.fd_debug_set_tc(7, 0) == 0
which = sys.argv[1:] or ["voxelize", "conv", "sort", "dense", "tall", "wgrad", "bn", "predict", "model"]


def sites(B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = np.sort(rng.choice(cells, size=min(n, cells), replace=False))
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


if "voxelize" in which:
    sc = [synth_scene(12000, seed=1), synth_scene(9000, seed=2)]
    pts = torch.from_numpy(np.concatenate(sc)).to(dev)
    off = torch.tensor([0, len(sc[0]), len(sc[0]) + len(sc[1])], dtype=torch.int32, device=dev)
    v = ops.voxelize_vfe(pts, off, NUSC_VOXEL, NUSC_RANGE, 10, 4000, num_feat=5, feat_stride=8)
    print("voxelize: %d voxels" % int(v["total"].item()))

if "conv" in which or "wgrad" in which or "sort" in which:
    shape, B = [9, 24, 24], 2
    c = sites(B, shape, 3000)
    n = len(c)
    ct = torch.from_numpy(c).to(dev)
    nd = torch.tensor([n], dtype=torch.int32, device=dev)
    rb, _ = ops.rulebook_subm(ct, nd, n, shape, [3, 3, 3], batch_size=B)
    rb2, _ = ops.rulebook_conv(ct, nd, n, B, shape, [3, 3, 3], [2, 2, 2], [1, 1, 1])

if "conv" in which:
    for cin, cout in ((16, 16), (32, 32), (64, 64), (128, 128), (16, 32), (64, 128)):       # NT = 16 / 32 / 64 / 128
        x = ops.to_split(torch.randn((n, cin), device=dev))
        w = torch.randn((27, cin, cout), device=dev) / np.sqrt(27 * cin)
        y = ops.sparse_conv(x, w, rb, residual=x if cin == cout else None, relu=True, precision="bf16x3", out_fmt="split")
        y2 = ops.sparse_conv(x, w, rb2, relu=True, precision="bf16x3", out_fmt="fp32")
        ref = ops.sparse_conv(x.to_fp32(), w, rb2, relu=True, precision="fp32")
        no = int(rb2.n_out_dev.item())            # rows past the active count are undefined
        print("sparse conv %d->%d: max |tc - fp32| %.2e" % (cin, cout, float((y2[:no] - ref[:no]).abs().max())))

if "sort" in which:
    # pattern-sorted tiles: the counting sort (keys from the neighbour search and from the table), the row_perm epilogue
    ops.SORT_MIN_ROWS, ops.SORT_WINDOW = 0, 2048
    for r in (rb, rb2):
        r._sorted = None
    rb_k, _ = ops.rulebook_subm(ct, nd, n, shape, [3, 3, 3], batch_size=B)          # with row keys
    rb2.row_key = None                                                              # keys computed from the table
    for cin, cout, table in ((16, 16, rb_k), (64, 64, rb_k), (32, 64, rb2)):
        x = ops.to_split(torch.randn((n, cin), device=dev))
        w = torch.randn((27, cin, cout), device=dev) / np.sqrt(27 * cin)
        res = x if cin == cout else None
        ya = ops.sparse_conv(x, w, table, residual=res, relu=True, precision="bf16x3", out_fmt="split", sort_tiles=False)
        yb = ops.sparse_conv(x, w, table, residual=res, relu=True, precision="bf16x3", out_fmt="split", sort_tiles=True)
        no = int(table.n_out_dev.item())
        print("sorted conv %d->%d: identical %s" % (cin, cout, bool(torch.equal(ya.t[:no].view(torch.int32), yb.t[:no].view(torch.int32)))))

if "tall" in which:
    for cin, cout, H, W in ((64, 64, 21, 13), (128, 16, 16, 8), (128, 256, 18, 17)):   # tall TMA stages of the dense 3x3 layers
        x = torch.randn((2, H, W, cin), device=dev)
        w = torch.randn((9, cin, cout), device=dev) / np.sqrt(9 * cin)
        y = ops.conv2d_nhwc(ops.to_split(x), w, (3, 3), (1, 1), (1, 1), relu=True, precision="bf16x3", out_fmt="fp32")
        ref = ops.conv2d_nhwc(x, w, (3, 3), (1, 1), (1, 1), relu=True, precision="fp32")
        y = y.t if isinstance(y, ops.Feat) else y
        print("tall dense conv %d->%d %dx%d: max |tc - fp32| %.2e" % (cin, cout, H, W, float((y - ref).abs().max())))

if "dense" in which:
    for cin, cout, k, s in ((64, 128, 3, 1), (128, 64, 1, 1), (64, 64, 3, 2)):            # TMA tile loads (stride 1) / cp.async
        x = torch.randn((2, 20, 22, cin), device=dev)
        w = torch.randn((k * k, cin, cout), device=dev) / np.sqrt(k * k * cin)
        y = ops.conv2d_nhwc(ops.to_split(x), w, (k, k), (s, s), (k // 2, k // 2), relu=True, precision="bf16x3", out_fmt="fp32")
        ref = ops.conv2d_nhwc(x, w, (k, k), (s, s), (k // 2, k // 2), relu=True, precision="fp32")
        y = y.t if isinstance(y, ops.Feat) else y
        print("dense conv %d->%d k%d s%d: max |tc - fp32| %.2e" % (cin, cout, k, s, float((y - ref).abs().max())))

if "wgrad" in which:
    # bf16x3 = the output-stationary tcgen05 kernel (2 / 4 / 8 offsets per accumulator group, several passes, Cout tiles)
    for cin, cout in ((64, 64), (16, 32), (32, 32), (128, 128)):
        x = torch.randn((n, cin), device=dev)
        gy = torch.randn((n, cout), device=dev)
        dws = {}
        for prec in ("fp32", "bf16x3"):
            dws[prec] = torch.zeros((27, cin, cout), device=dev)
            T.sparse_conv_wgrad(x, gy, rb, dws[prec], precision=prec)
        print("wgrad %d->%d: max |tc - fp32| %.2e" % (cin, cout, float((dws["bf16x3"] - dws["fp32"]).abs().max())))
    for cin, cout, tr in ((128, 256, False), (256, 128, True)):       # dense 3x3 with two Cout tiles, ConvTranspose2d phases
        x = torch.randn((2, 10, 12, cin), device=dev)
        k = 2 if tr else 3
        gy = torch.randn((2, 20, 24, cout) if tr else (2, 10, 12, cout), device=dev)
        dws = {}
        for prec in ("fp32", "bf16x3"):
            dws[prec] = torch.zeros((k * k, cin, cout), device=dev)
            T.conv2d_wgrad(x, gy, dws[prec], (k, k), (2, 2) if tr else (1, 1), (0, 0) if tr else (1, 1), transposed=tr, precision=prec)
        print("dense wgrad %d->%d%s: max |tc - fp32| %.2e" % (cin, cout, " (convT)" if tr else "", float((dws["bf16x3"] - dws["fp32"]).abs().max())))

if "bn" in which:
    # train-mode BatchNorm kernels (128-bit paths) with the split-bf16 copies of the training step
    from torch import nn
    for Cc in (16, 128):
        bn = nn.BatchNorm1d(Cc).to(dev).train()
        x = torch.randn((5000, Cc), device=dev)
        saved = T.bn_train_stats(x, bn)
        ys = torch.empty_like(x)
        y = T.affine_act(x, saved.scale, saved.shift, x, True, split=(ys, 0))
        dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
        dx, dres, dxs = T.bn_backward(torch.randn_like(x), y, True, x, saved, bn.weight.detach(), dg, db, True, want_split=True)
        back = ops.Feat(ys, "split").to_fp32()
        print("bn C=%d: split copy max err %.2e" % (Cc, float((back - y).abs().max())))

if "predict" in which:
    from oracle import predict_ref as PR
    cfg = dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_per_img=500,
               nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=83,
                        nms_iou_threshold=0.2), score_threshold=0.1, pc_range=[-54, -54], out_size_factor=8,
               voxel_size=[0.075, 0.075])
    preds = PR.synth_preds(1, 48, 48, 3, seed=3, n_obj=12)

    class H:
        standard, dense, timesteps, target_timesteps = True, False, 3, 7
    ret = P.center_head_predict(H(), {}, [{k: v.to(dev) for k, v in preds.items()}], cfg)
    print("predict: %d boxes" % len(ret[0]["scores"]))

if "model" in which:
    import bench
    m = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
    sc = synth_scene(15000, seed=3)
    with torch.no_grad():
        p = m.forward_points(torch.from_numpy(sc).to(dev), torch.tensor([0, len(sc)], dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    print("model forward ok", float(p[0]["hm"].abs().max()))
print("probe done")
