"""CPU: the predict oracle (oracle/predict_ref.py) against known rotated-IoU answers and against the output of the
REFERENCE `CenterHead.predict` (tests/golden/predict.pt, written by oracle/gen_golden.py from /root/reference)."""
import os

import numpy as np
import torch

from oracle import predict_ref as PR


def test_rotated_iou_known_answers():
    a = [0, 0, 0, 4, 2, 1, 0.0]
    assert abs(PR.iou_bev_np(a, a) - 1.0) < 1e-12
    assert abs(PR.iou_bev_np(a, [2, 0, 0, 4, 2, 1, 0.0]) - (4.0 / 12.0)) < 1e-12          # half overlap along x
    assert PR.iou_bev_np(a, [10, 0, 0, 4, 2, 1, 0.3]) == 0.0
    # the same rectangle turned by 90 degrees: the intersection is the 2 x 2 square
    assert abs(PR.iou_bev_np(a, [0, 0, 0, 4, 2, 1, np.pi / 2]) - 4.0 / (8 + 8 - 4)) < 1e-9
    # heading and swapped extents describe the same footprint
    assert abs(PR.iou_bev_np(a, [0, 0, 0, 2, 4, 1, np.pi / 2]) - 1.0) < 1e-9
    # square turned by 45 degrees inside a larger square: area ratio
    assert abs(PR.iou_bev_np([0, 0, 0, 4, 4, 1, 0], [0, 0, 0, 2, 2, 1, np.pi / 4]) - 4.0 / 16.0) < 1e-9


def test_greedy_nms_order_and_suppression():
    boxes = np.array([[0, 0, 0, 4, 2, 1, 0], [0.2, 0, 0, 4, 2, 1, 0], [5, 5, 0, 4, 2, 1, 1.0], [0.1, 0.1, 0, 4, 2, 1, 0.1],
                      [5.1, 5, 0, 4, 2, 1, 1.0]], np.float32)
    assert PR.nms_np(boxes, 0.2).tolist() == [0, 2]
    assert PR.nms_np(boxes, 0.99).tolist() == [0, 1, 2, 3, 4]
    assert PR.nms_np(boxes[:0], 0.2).tolist() == []


def test_predict_oracle_matches_reference_predict(golden_dir):
    g = torch.load(os.path.join(golden_dir, "predict.pt"), weights_only=False)
    for name, case in g["cases"].items():
        got = PR.predict_ref(case["preds"], case["timesteps"], g["test_cfg"])
        assert len(got) == len(case["ret"])
        for r, want in zip(got, case["ret"]):
            assert torch.equal(r["box3d_lidar"], want["box3d_lidar"]), name
            assert torch.equal(r["scores"], want["scores"]) and torch.equal(r["label_preds"], want["label_preds"]), name
            assert len(r["scores"]) % 7 == 0 and len(r["scores"]) > 0          # 7 forecast timesteps, NMS kept something
            assert r["label_preds"].tolist() == sorted(r["label_preds"].tolist())   # labels = timestep index blocks
