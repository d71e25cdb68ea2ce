"""Pipeline timeline of block 0 of one tensor-core conv launch (FD_TC_DEBUG=32). Perf triage helper."""
import ctypes as C, os, sys
import numpy as np, torch
os.environ["FD_TC_DEBUG"] = os.environ.get("FD_TC_DEBUG", "32")
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from futuredet_b200 import ops, lib
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
cin, cout, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
shape, B = [11, 360, 360], 1
cells = B * shape[0] * shape[1] * shape[2]
# clustered sites (like LiDAR surfaces): random walk-ish blobs
lin = np.unique((rng.integers(0, cells // 64, n // 4)[:, None] * 64 + rng.integers(0, 64, (n // 4, 8))).ravel())[:n]
c = np.empty((len(lin), 4), np.int32)
c[:, 3] = lin % shape[2]; l2 = lin // shape[2]
c[:, 2] = l2 % shape[1]; l2 = l2 // shape[1]
c[:, 1] = l2 % shape[0]; c[:, 0] = l2 // shape[0]
n = len(c)
ct = torch.from_numpy(c).to(dev); nd = torch.tensor([n], dtype=torch.int32, device=dev)
rb, _ = ops.rulebook_subm(ct, nd, n, shape, [3, 3, 3], batch_size=B)
print("rows", n, "pairs", int(rb.pair_num.sum()))
x = ops.to_split(torch.randn((n, cin), device=dev))
w = torch.randn((27, cin, cout), device=dev) / 30
for _ in range(3):
    y = ops.sparse_conv(x, w, rb, residual=x if cin == cout else None, relu=True, precision="bf16x3", out_fmt="split")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ops.sparse_conv(x, w, rb, residual=x if cin == cout else None, relu=True, precision="bf16x3", out_fmt="split"); e1.record()
torch.cuda.synchronize()
print("kernel ms", e0.elapsed_time(e1))
L = lib.load()
L.fd_debug_read_tc_trace.argtypes = [C.c_void_p, C.c_int]
def trace(role):
    buf = np.zeros(8192, np.int64)
    assert L.fd_debug_read_tc_trace(buf.ctypes.data_as(C.c_void_p), role) == 0
    return buf
m = trace(0); pr = trace(1); ep = trace(2)
k = int((m[0::2] > 0).sum())
ready, commit = m[0:2 * k:2], m[1:2 * k:2]
t0 = ready[0]
print("MMA stages traced:", k)
d = np.diff(ready)
print("MMA ready->ready cycles: mean %.0f median %.0f p10 %.0f p90 %.0f" % (d.mean(), np.median(d), np.percentile(d, 10), np.percentile(d, 90)))
print("MMA ready->commit issue: mean %.0f" % (commit - ready).mean())
print("MMA commit->next ready (waiting): mean %.0f median %.0f" % ((ready[1:] - commit[:-1]).mean(), np.median(ready[1:] - commit[:-1])))
f = trace(3)
bc, bw, af = f[0:4 * k:4], f[2:4 * k:4], f[3:4 * k:4]
print("MMA: before-commit->after-commit %.0f | after-commit->before-wait(next) %.0f | wait duration %.0f | fence %.0f | issue(fence->before commit) %.0f" % (
    (commit - bc).mean(), (bw[1:] - commit[:-1]).mean(), (ready - bw).mean(), (af - ready).mean(), (bc - af).mean()))
kp = int((pr[0::4] > 0).sum())
pw0, pw1, pi, pos = pr[0:4 * kp:4], pr[1:4 * kp:4], pr[2:4 * kp:4], pr[3:4 * kp:4]
print("producer(g0) stages:", kp, " wait-for-empty mean %.0f median %.0f ; issue mean %.0f ; loop period mean %.0f" % ((pw1 - pw0).mean(), np.median(pw1 - pw0), (pi - pw1).mean(), np.diff(pw0).mean()))
# fill latency: producer issue done -> MMA sees the stage ready
lat = []
for i in range(min(kp, 400)):
    g = int(pos[i])
    if g < k:
        lat.append(ready[g] - pi[i])
lat = np.array(lat)
print("issue-done -> MMA-ready latency: mean %.0f median %.0f p90 %.0f" % (lat.mean(), np.median(lat), np.percentile(lat, 90)))
X = []
for i in range(min(kp, 400)):
    g = int(pos[i])
    if 5 <= g < k + 5 and g - 5 < k:
        X.append(pw1[i] - commit[g - 5])
X = np.array(X)
print("commit(c-SA) -> producer sees slot empty: mean %.0f median %.0f p10 %.0f p90 %.0f" % (X.mean(), np.median(X), np.percentile(X, 10), np.percentile(X, 90)))
W = np.array([pw0[i] - commit[int(pos[i]) - 5] for i in range(min(kp, 400)) if 5 <= int(pos[i]) < k + 5])
print("producer starts waiting relative to that commit: mean %.0f median %.0f" % (W.mean(), np.median(W)))
ke = int((ep[0::2] > 0).sum())
print("epilogue units:", ke, "durations", (ep[1:2 * ke:2] - ep[0:2 * ke:2])[:6], "starts", (ep[0:2 * ke:2] - t0)[:6])
print("first 12 MMA ready times:", (ready[:12] - t0))
