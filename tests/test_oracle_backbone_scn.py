"""CPU: the functional backbone restatement (oracle/spconv_ref.py::backbone_forward) reproduces the golden written by
the REFERENCE's own det3d/models/backbones/scn.py (SparseBasicBlock + SpMiddleResNetFHD executed unmodified over the CPU
spconv shim, oracle/gen_golden.py backbone): topology, indice_key sharing, bias / BN / residual / ReLU order, strides
and paddings, dense().view() -- all pinned to the reference source.  What stays unpinned is spconv's per-op arithmetic
(third-party, not installable here)."""
import os

import numpy as np
import torch

from oracle import spconv_ref as S


def test_restatement_matches_reference_scn_source(golden_dir):
    g = torch.load(os.path.join(golden_dir, "backbone_scn.pt"), weights_only=False)
    sd = S.seeded_state(g["state_shapes"], g["state_seed"])
    out, stages = S.backbone_forward(sd, g["features"], g["coors"].numpy(), g["batch_size"], g["grid"], return_stages=True)
    assert out.shape == g["out"].shape
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-5)
    for name in ("conv1", "conv2", "conv3", "conv4"):
        feats, coords, shape = stages[name]
        ref = g["stages"][name]
        assert list(shape) == list(ref["spatial_shape"])
        assert np.array_equal(np.asarray(coords), ref["indices"].numpy()), name      # same active set, same row order
        torch.testing.assert_close(feats.double().sum(0).float(), ref["feature_sum"], rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(stages["conv4"][0], g["stages"]["conv4"]["features"], rtol=1e-5, atol=1e-5)
