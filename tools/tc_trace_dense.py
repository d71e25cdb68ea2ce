"""Pipeline timeline of block 0 of one dense tensor-core conv launch (FD_TC_DEBUG=32). Perf triage helper.
usage: python tools/tc_trace_dense.py cin cout [B H W]"""
import ctypes as C, os, sys
import numpy as np, torch
os.environ["FD_TC_DEBUG"] = os.environ.get("FD_TC_DEBUG", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from futuredet_b200 import ops, lib
dev = torch.device("cuda:0")
cin, cout = int(sys.argv[1]), int(sys.argv[2])
B, H, W = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (4, 180, 180)
x = ops.to_split(torch.randn((B, H, W, cin), device=dev))
w = torch.randn((9, cin, cout), device=dev) / 30
run = lambda: ops.conv2d_nhwc(x, w, (3, 3), (1, 1), (1, 1), relu=True, precision="bf16x3", out_fmt="split")
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("kernel ms %.4f  -> %.1f TFLOP/s algorithmic" % (ms, 2.0 * B * H * W * 9 * cin * cout / ms / 1e9))
L = lib.load()
L.fd_debug_read_tc_trace.argtypes = [C.c_void_p, C.c_int]
def trace(role):
    buf = np.zeros(8192, np.int64)
    assert L.fd_debug_read_tc_trace(buf.ctypes.data_as(C.c_void_p), role) == 0
    return buf
m = trace(0); ep = trace(2); f = trace(3)
k = int((m[0::2] > 0).sum())
ready, commit = m[0:2 * k:2], m[1:2 * k:2]
print("MMA stages traced:", k)
d = np.diff(ready)
print("MMA ready->ready cycles: mean %.0f median %.0f p10 %.0f p90 %.0f" % (d.mean(), np.median(d), np.percentile(d, 10), np.percentile(d, 90)))
print("MMA ready->commit (issue): mean %.0f median %.0f" % ((commit - ready).mean(), np.median(commit - ready)))
print("MMA commit->next ready: mean %.0f median %.0f p90 %.0f" % ((ready[1:] - commit[:-1]).mean(), np.median(ready[1:] - commit[:-1]), np.percentile(ready[1:] - commit[:-1], 90)))
bw, af = f[2:4 * k:4], f[3:4 * k:4]
print("MMA: wait duration (before-wait -> ready) mean %.0f median %.0f | commit -> before-wait(next) mean %.0f" % (
    (ready - bw).mean(), np.median(ready - bw), (bw[1:] - commit[:-1]).mean()))
ke = int((ep[0::2] > 0).sum())
print("epilogue units:", ke, "durations", (ep[1:2 * ke:2] - ep[0:2 * ke:2])[:8], "starts", (ep[0:2 * ke:2] - ready[0])[:8])
print("first 40 MMA ready deltas:", d[:40])
