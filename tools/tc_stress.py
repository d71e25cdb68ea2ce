"""Tensor-core conv kernel vs the fp32 CUDA-core arm on random sparse problems (debug / stress helper)."""
import sys
import numpy as np
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from futuredet_b200 import ops

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def sites(B, shape, n):
    cells = B * shape[0] * shape[1] * shape[2]
    lin = rng.choice(cells, size=min(n, cells), replace=False)
    lin.sort()
    c = np.empty((len(lin), 4), np.int32)
    c[:, 3] = lin % shape[2]; lin = lin // shape[2]
    c[:, 2] = lin % shape[1]; lin = lin // shape[1]
    c[:, 1] = lin % shape[0]; c[:, 0] = lin // shape[0]
    return c


CASES = [(16, 16, 3000, "fp32"), (16, 16, 3000, "split"), (32, 32, 5000, "split"), (64, 64, 40000, "split"),
                            (64, 128, 30000, "split"), (128, 128, 70000, "split"), (32, 64, 200000, "split"), (16, 32, 300000, "split")]
if len(sys.argv) > 1:
    CASES = [CASES[int(sys.argv[1])]]
for (cin, cout, n, fmt) in CASES:
    shape, B = [21, 200, 200], 2
    c = sites(B, shape, n)
    n = len(c)
    ct = torch.from_numpy(c).to(dev)
    nd = torch.tensor([n], dtype=torch.int32, device=dev)
    rb, _ = ops.rulebook_subm(ct, nd, n, shape, [3, 3, 3], batch_size=B)
    x = torch.randn((n, cin), device=dev)
    w = torch.randn((27, cin, cout), device=dev) / np.sqrt(27 * cin)
    ref = ops.sparse_conv(x, w, rb, precision="fp32")
    xin = ops.to_split(x) if fmt == "split" else x
    torch.cuda.synchronize()
    y = ops.sparse_conv(xin, w, rb, precision="bf16x3", out_fmt=fmt)
    torch.cuda.synchronize()
    y = y.to_fp32() if isinstance(y, ops.Feat) else y
    print(cin, cout, n, fmt, "max err", float((y - ref).abs().max()), flush=True)
print("stress ok")
