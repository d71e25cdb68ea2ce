"""CPU: the training oracle (torch-autograd over the functional restatements in oracle/) reproduces the gradients,
losses and BatchNorm running statistics of the REFERENCE RPN + CenterHead classes in training mode
(tests/golden/neck_head_train.pt, written by oracle/gen_golden.py from /root/reference), and the sparse-backbone
restatement is differentiable (its gradients match finite differences of its own forward)."""
import os

import numpy as np
import torch

from oracle import dense_ref as D
from oracle import spconv_ref as S
from oracle.loss_ref import center_head_loss_ref

HEADS = ["reg", "height", "dim", "rot", "vel", "hm"]


def run_oracle_neck_head(g):
    nsd = {k: v.clone() for k, v in g["neck_state"].items()}
    hsd = {k: v.clone() for k, v in g["head_state"].items()}
    for sd in (nsd, hsd):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    nc, hc = g["neck_cfg"], g["head_cfg"]
    feat = D.rpn_forward(nsd, x, nc["layer_nums"], nc["ds_layer_strides"], nc["us_layer_strides"], train=0.01)
    preds = D.center_head_forward(hsd, feat, [HEADS], train=0.1)
    loss = center_head_loss_ref(preds, g["example"], hc["timesteps"], hc["code_weights"], hc["weight"])
    total = sum(loss["loss"])
    total.backward()
    return nsd, hsd, x, loss, total


def test_dense_training_oracle_matches_reference_gradients(golden_dir):
    g = torch.load(os.path.join(golden_dir, "neck_head_train.pt"), weights_only=False)
    nsd, hsd, x, loss, total = run_oracle_neck_head(g)
    torch.testing.assert_close(total.detach(), g["total"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(x.grad, g["x_grad"], rtol=1e-4, atol=1e-6)
    n = 0
    for name, want in g["grads"].items():
        part, key = name.split(".", 1)
        got = (nsd if part == "neck" else hsd)[key].grad
        torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-6, msg=lambda m, k=name: "%s: %s" % (k, m))
        n += 1
    assert n == len(g["grads"]) == 61
    for sd, after in ((nsd, g["neck_state_after"]), (hsd, g["head_state_after"])):
        for k, v in after.items():
            if "running" in k:
                torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-6)


def test_sparse_backbone_oracle_is_differentiable():
    """Gradient of the restated spconv backbone (train-mode BN) w.r.t. a weight matches a central finite difference."""
    import futuredet_b200 as fb
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    bb = fb.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8))
    sd = {k: v.clone().double() if v.is_floating_point() else v.clone() for k, v in bb.state_dict().items()}
    grid = [24, 24, 40]                       # (x, y, z): D 41 -> 21 -> 11 -> 5 -> 2
    cells = 41 * 24 * 24
    lin = np.sort(rng.choice(cells, 1500, replace=False))
    coors = np.stack([np.zeros_like(lin), lin // (24 * 24), (lin // 24) % 24, lin % 24], 1).astype(np.int32)
    feats = torch.from_numpy(rng.standard_normal((1500, 5))).double()
    proj = torch.from_numpy(rng.standard_normal((1, 256, 3, 3))).double()
    key = "conv2.3.conv1.weight"

    def f(w):
        s = dict(sd)
        s[key] = w
        s = {k: (v.clone() if "running" in k else v) for k, v in s.items()}
        dense = S.backbone_forward(s, feats, coors, 1, grid, bn_eval=S.bn_train(0.01))
        return (dense * proj).sum()

    w = sd[key].clone().requires_grad_(True)
    f(w).backward()
    idx = [(1, 1, 1, 3, 5), (0, 2, 1, 7, 0), (2, 0, 0, 31, 31)]
    for i in idx:
        e = torch.zeros_like(w)
        e[i] = 1e-7
        fd = (f(w.detach() + e) - f(w.detach() - e)) / 2e-7
        assert abs(float(fd) - float(w.grad[i])) <= 1e-4 * max(1.0, abs(float(fd))), (i, float(fd), float(w.grad[i]))
