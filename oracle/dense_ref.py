"""ORACLE (test infrastructure only): RPN neck + CenterHead forward/loss restated in functional torch (CPU, fp32).

Follows, layer for layer and with the reference's state_dict key layout:
  det3d/models/necks/rpn.py:70-113,124-142,150-159     RPN (ZeroPad2d+Conv/BN/ReLU stacks, deblocks, cat)
  det3d/models/bbox_heads/center_head.py:129-152        SepHead (conv-BN-ReLU, conv; final_kernel=3, bn=True)
  det3d/models/bbox_heads/center_head.py:344-349,375-390 CenterHead.shared_conv / forward (standard mode)
  det3d/models/losses/centernet_loss.py:18-25,75-95     RegLoss / FastFocalLoss
  det3d/models/bbox_heads/center_head.py:396-539        CenterHead.loss (standard branch)
Pinned by tests/golden/neck_head_*.pt, produced by importing the *reference classes* in the build
container (oracle/gen_golden.py) at reduced channel widths.
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, p, eps, train=None):
    """train=None: eval mode; train=momentum: training mode (batch statistics, running stats of sd updated in place)."""
    if train is None:
        return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                            False, 0.0, eps)
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        True, train, eps)


def rpn_forward(sd, x, layer_nums, ds_layer_strides, us_layer_strides, prefix="", eps=1e-3, train=None):
    """RPN.forward (rpn.py:150-159), eval mode.  us stride > 1 -> ConvTranspose2d, else Conv2d (rpn.py:78-110)."""
    ups = []
    start = len(layer_nums) - len(us_layer_strides)
    for i, n in enumerate(layer_nums):
        b = "%sblocks.%d." % (prefix, i)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[b + "1.weight"], None, stride=ds_layer_strides[i])   # ZeroPad2d(1)+conv
        x = F.relu(_bn(x, sd, b + "2.", eps, train))
        for j in range(n):
            k = 4 + 3 * j
            x = F.conv2d(x, sd[b + "%d.weight" % k], None, padding=1)
            x = F.relu(_bn(x, sd, b + "%d." % (k + 1), eps, train))
        if i - start >= 0:
            d = "%sdeblocks.%d." % (prefix, i - start)
            s = us_layer_strides[i - start]
            if s > 1:
                y = F.conv_transpose2d(x, sd[d + "0.weight"], None, stride=int(s))
            else:
                s = int(round(1 / s))
                y = F.conv2d(x, sd[d + "0.weight"], None, stride=s)
            ups.append(F.relu(_bn(y, sd, d + "1.", eps, train)))
    return torch.cat(ups, dim=1) if ups else x


def center_head_forward(sd, x, head_names_per_task, prefix="", train=None):
    """CenterHead.forward in standard mode (center_head.py:375-390): list over tasks of {head: [B,c,H,W]}.
    BatchNorm2d in the head uses the torch default eps 1e-5 (center_head.py:347,136)."""
    p = prefix + "shared_conv."
    x = F.conv2d(x, sd[p + "0.weight"], sd[p + "0.bias"], padding=1)
    x = F.relu(_bn(x, sd, p + "1.", 1e-5, train))
    rets = []
    for t, names in enumerate(head_names_per_task):
        ret = {}
        for h in names:
            q = "%stasks.%d.%s." % (prefix, t, h)
            y = F.conv2d(x, sd[q + "0.weight"], sd[q + "0.bias"], padding=1)
            y = F.relu(_bn(y, sd, q + "1.", 1e-5, train))
            ret[h] = F.conv2d(y, sd[q + "3.weight"], sd[q + "3.bias"], padding=1)
        rets.append(ret)
    return rets


def _gather_feat(feat, ind):
    """_transpose_and_gather_feat (det3d/core/utils/center_utils.py:66-80)."""
    B, C, H, W = feat.shape
    f = feat.permute(0, 2, 3, 1).reshape(B, H * W, C)
    return f.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))


def fast_focal_loss(out, target, ind, mask, cat):
    """centernet_loss.py:75-95."""
    mask = mask.float()
    gt = torch.pow(1 - target, 4)
    neg_loss = (torch.log(1 - out) * torch.pow(out, 2) * gt).sum()
    pos_pred = _gather_feat(out, ind).gather(2, cat.unsqueeze(2))
    num_pos = mask.sum()
    pos_loss = (torch.log(pos_pred) * torch.pow(1 - pos_pred, 2) * mask.unsqueeze(2)).sum()
    if num_pos == 0:
        return -neg_loss
    return -(pos_loss + neg_loss) / num_pos


def reg_loss(output, mask, ind, target):
    """centernet_loss.py:18-25 (+ _reg_loss 5-16 semantics: masked L1 / (num + 1e-4), per channel)."""
    pred = _gather_feat(output, ind)
    m = mask.float().unsqueeze(2)
    loss = F.l1_loss(pred * m, target * m, reduction="none")
    loss = loss / (m.sum() + 1e-4)
    return loss.transpose(2, 0).sum(dim=2).sum(dim=1)
