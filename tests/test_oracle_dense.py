"""CPU: the functional RPN / CenterHead / loss oracle against golden tensors produced by the reference classes."""
import os

import pytest
import torch

from oracle import dense_ref as D


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "neck_head.pt"), weights_only=False)


def test_rpn_oracle_matches_reference(gold):
    c = gold["neck_cfg"]
    y = D.rpn_forward(gold["neck_state"], gold["x"], c["layer_nums"], c["ds_layer_strides"], c["us_layer_strides"])
    torch.testing.assert_close(y, gold["neck_out"], rtol=1e-5, atol=1e-5)


def test_center_head_oracle_matches_reference(gold):
    names = [list(p.keys()) for p in gold["preds"]]
    preds = D.center_head_forward(gold["head_state"], gold["neck_out"], names)
    for p, g in zip(preds, gold["preds"]):
        for k in g:
            torch.testing.assert_close(p[k], g[k], rtol=1e-5, atol=1e-5)
    assert gold["preds"][0]["vel"].shape[1] == 2 * gold["head_cfg"]["timesteps"]     # multi-timestep head


def test_loss_oracle_matches_reference(gold):
    from oracle.loss_ref import center_head_loss_ref
    cfg = gold["head_cfg"]
    out = center_head_loss_ref(gold["preds"], gold["example"], cfg["timesteps"], cfg["code_weights"], cfg["weight"])
    g = gold["loss"]
    torch.testing.assert_close(out["loss"][0], g["loss"][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out["hm_loss"][0], g["hm_loss"][0], rtol=1e-5, atol=1e-6)
    for a, b in zip(out["loc_loss_elem"][0], g["loc_loss_elem"][0]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-6)
