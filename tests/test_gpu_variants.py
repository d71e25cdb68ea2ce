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
