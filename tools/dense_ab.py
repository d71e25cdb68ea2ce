import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from futuredet_b200 import ops, lib
dev = torch.device('cuda:0')
model = bench.build_model().set_precision('bf16x3').to(dev).configure_voxelizer(bench.VOXEL_CFG)
L = lib.load(); L.fd_debug_set_tc.argtypes=[C.c_int, C.c_int]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = ops.to_split(torch.randn((B, 180, 180, 256), device=dev))
def run():
    return model.bbox_head(model.neck(x, out_fmt='split'))
with torch.no_grad():
    for tall in (0, 1, 0, 1):
        L.fd_debug_set_tc(8, tall)
        for _ in range(2): run()
        torch.cuda.synchronize()
        ops.PROFILE = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
        per = [round(r['start'].elapsed_time(r['end']), 3) for r in recs]
        print('tall=%d neck+head %.3f ms; per launch:' % (tall, e0.elapsed_time(e1)), per)
