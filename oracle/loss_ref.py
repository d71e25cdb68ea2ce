"""ORACLE (test infrastructure only): CenterHead.loss, standard branch, restated on CPU torch.

Follows det3d/models/bbox_heads/center_head.py:396-539 (standard mode: lines 400-402, 420-421, 443-444,
457-469, 499-500, 512-516, 526-527) with FastFocalLoss / RegLoss from oracle/dense_ref.py.
Pinned by tests/golden/neck_head.pt (reference `head.loss` output)."""
import numpy as np
import torch

from .dense_ref import fast_focal_loss, reg_loss


def center_head_loss_ref(preds_dicts, example, timesteps, code_weights, weight):
    out = {k: [] for k in ("loss", "hm_loss", "loc_loss", "loc_loss_elem", "num_positive")}
    cw = torch.tensor(code_weights, dtype=torch.float32)
    cw_forecast = torch.tensor(list(np.array(code_weights) * np.array([0, 0, 0, 0, 0, 0, 1, 1, 0, 0])), dtype=torch.float32)
    for task_id, p in enumerate(preds_dicts):
        hm = torch.clamp(torch.sigmoid(p["hm"]), min=1e-4, max=1 - 1e-4)                      # _sigmoid, :392-394
        hm_loss = fast_focal_loss(hm, example["hm"][0][task_id], example["ind"][0][task_id],
                                  example["mask"][0][task_id], example["cat"][0][task_id])
        target_box = [example["anno_box"][i][task_id][..., [0, 1, 2, 3, 4, 5, 6, 7, -2, -1]] for i in range(timesteps)]
        anno = [torch.cat((p["reg"], p["height"], p["dim"], p["vel"][:, 2 * i:2 * i + 2], p["rot"]), dim=1)
                for i in range(timesteps)]
        box_loss = [reg_loss(anno[i], example["mask"][0][task_id], example["ind"][0][task_id], target_box[i])
                    for i in range(timesteps)]
        loc_loss = [(box_loss[i] * (cw if i == 0 else cw_forecast)).sum() for i in range(timesteps)]
        out["loss"].append(hm_loss + weight * sum(loc_loss))
        out["hm_loss"].append(hm_loss)
        out["loc_loss"].append(loc_loss)
        out["loc_loss_elem"].append(box_loss)
        out["num_positive"].append(sum(sum(sum(example["mask"][i][task_id].float() for i in range(timesteps)))))
    return out
