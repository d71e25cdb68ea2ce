"""Pipeline timeline of block 0 of a level-1 SubM tensor-core conv over real (synthetic LiDAR) voxels, with and without
pattern-sorted tiles (FD_TC_DEBUG=32).  Perf triage helper:  python tools/tc_trace_sorted.py [cin cout scenes residual]"""
import ctypes as C, os, sys
import numpy as np, torch
os.environ["FD_TC_DEBUG"] = os.environ.get("FD_TC_DEBUG", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from futuredet_b200 import ops, lib
from futuredet_b200.synth import synth_scene
dev = torch.device("cuda:0")
cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16, 16)
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
use_res = (int(sys.argv[4]) if len(sys.argv) > 4 else 1) and cin == cout
model = bench.build_model().to(dev).configure_voxelizer(bench.VOXEL_CFG)
scenes = [synth_scene(bench.N_TARGET, seed=100 + b) for b in range(B)]
pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
vox = model.voxelize(pts, off)
coords, nd = vox["coords"], vox["total"]
n_cap = coords.shape[0]
n = int(nd.item())
rb, _ = ops.rulebook_subm(coords, nd, n_cap, [41, 1440, 1440], [3, 3, 3], batch_size=B)
x = ops.to_split(torch.randn((n_cap, cin), device=dev))
w = torch.randn((27, cin, cout), device=dev) / 30
L = lib.load()
L.fd_debug_read_tc_trace.argtypes = [C.c_void_p, C.c_int]
def trace(role):
    buf = np.zeros(8192, np.int64)
    assert L.fd_debug_read_tc_trace(buf.ctypes.data_as(C.c_void_p), role) == 0
    return buf
for srt in (False, True):
    run = lambda: ops.sparse_conv(x, w, rb, residual=x if use_res else None, relu=True, precision="bf16x3", out_fmt="split", sort_tiles=srt)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    print("==== sort_tiles=%s rows %d kernel ms %.3f" % (srt, n, e0.elapsed_time(e1)))
    m = trace(0); ep = trace(2)
    k = int((m[0::2] > 0).sum())
    ready, commit = m[0:2 * k:2], m[1:2 * k:2]
    d = np.diff(ready)
    print("MMA stages traced (block 0): %d; ready->ready mean %.0f median %.0f p90 %.0f; total span %.0f" % (k, d.mean(), np.median(d), np.percentile(d, 90), ready[-1] - ready[0]))
    ke = int((ep[0::2] > 0).sum())
    es, ee = ep[0:2 * ke:2], ep[1:2 * ke:2]
    print("epilogue units: %d; duration mean %.0f median %.0f; start->start mean %.0f; idle between units mean %.0f" % (
        ke, (ee - es).mean(), np.median(ee - es), np.diff(es).mean(), (es[1:] - ee[:-1]).mean()))
    big = np.sort(d)[-max(ke, 1):]
    print("largest MMA gaps (unit boundaries?): mean %.0f; stages/unit %.1f" % (big.mean(), k / max(ke, 1)))
