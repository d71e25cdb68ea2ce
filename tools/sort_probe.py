"""Perf triage of the pattern-sorted tiles (fd_rulebook_sort_rows): per sparse-conv launch, CUDA-event time and the number
of (tile, K stage) items the kernel works through, with the sort off and on, plus the time of the sort calls themselves.
    python tools/sort_probe.py [batch]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench                                           # noqa: E402
from futuredet_b200 import ops                         # noqa: E402
from futuredet_b200.synth import synth_scene           # noqa: E402


def stage_items(mask, n_rows, K, cin, cout):
    """(tile, stage) items of conv_tc_kernel for this tile-mask array (same rule as unit_gmask)."""
    acc = 2 * cout if cout <= 64 else cout
    acc = max(acc, 32) if cout <= 64 else acc
    T = min(4, 256 // acc)
    tiles = (n_rows + 127) // 128
    while T > 1 and -(-tiles // T) * max(cout // 128, 1) < 148:
        T >>= 1
    m = mask[:tiles].astype(np.uint32)
    pad = (-tiles) % T
    m = np.concatenate([m, np.zeros(pad, np.uint32)]).reshape(-1, T)
    u = np.bitwise_or.reduce(m, 1)
    if cin >= 64:
        per = np.array([int(v).bit_count() for v in u]) * (cin // 64)
    else:
        opk = 64 // cin
        ns = -(-K // opk)
        per = np.zeros(len(u), np.int64)
        for s in range(ns):
            per += ((u >> np.uint32(s * opk)) & np.uint32((1 << opk) - 1)) != 0
    per = np.maximum(per, 1)
    live = np.full(len(u), T)
    if pad:
        live[-1] = T - pad
    return int((per * live).sum()), T


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda", 0)
    model = bench.build_model().set_precision("bf16x3")
    model.to(dev).configure_voxelizer(bench.VOXEL_CFG)
    scenes = [synth_scene(bench.N_TARGET, seed=100 + b) for b in range(B)]
    pts = torch.from_numpy(np.concatenate(scenes)).to(dev)
    off = torch.tensor(np.r_[0, np.cumsum([len(s) for s in scenes])], dtype=torch.int32, device=dev)
    calls = []
    real_conv, real_sorted = ops.sparse_conv, ops.Rulebook.sorted_tiles
    sort_ev = []

    def conv(x, w, rb, *a, **k):
        calls.append((rb, int(w.shape[1]), int(w.shape[2]), bool(k.get("sort_tiles")) and ops.SORT_TILES and rb.K == 27))
        return real_conv(x, w, rb, *a, **k)

    def sorted_tiles(self):
        if self._sorted is not None:
            return self._sorted
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = real_sorted(self)
        e1.record()
        sort_ev.append((self.n_out_cap, e0, e1, self))
        return r

    ops.sparse_conv, ops.Rulebook.sorted_tiles = conv, sorted_tiles
    import futuredet_b200.sparse as sp
    sp.ops = ops
    rows = {}
    with torch.no_grad():
        for mode in (False, True):
            ops.SORT_TILES = mode
            for _ in range(2):
                model.forward_points(pts, off)
            torch.cuda.synchronize()
            calls.clear(); sort_ev.clear()
            ops.PROFILE = []
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.forward_points(pts, off)
            e1.record()
            torch.cuda.synchronize()
            recs, ops.PROFILE = [r for r in ops.PROFILE if r["kind"].startswith("sparse3d")], None
            print("sort_tiles=%s: forward %.3f ms, %d sparse launches" % (mode, e0.elapsed_time(e1), len(recs)))
            assert len(recs) == len(calls)
            for i, (r, (rb, cin, cout, srt)) in enumerate(zip(recs, calls)):
                n = int(rb.n_out_dev.item())
                mask = (rb._sorted[2] if srt else rb.tile_mask).cpu().numpy().view(np.uint32)
                items, T = stage_items(mask, n, rb.K, max(cin, 8), cout)
                ms = r["start"].elapsed_time(r["end"])
                rows.setdefault(i, {})[mode] = (n, cin, cout, T, items, ms)
            if mode:
                tot = 0.0
                for cap, a, b, rb in sort_ev:
                    ms = a.elapsed_time(b)
                    tot += ms
                    print("   sort: n %8d cap %9d  %.3f ms" % (int(rb.n_out_dev.item()), cap, ms))
                print("   sort total %.3f ms" % tot)
    print("%3s %8s %4s %4s %2s | %9s %8s %7s | %9s %8s %7s | items x  time x" % ("#", "rows", "cin", "cout", "T", "items", "ms", "cyc/it", "items", "ms", "cyc/it"))
    ta = tb = 0.0
    for i in sorted(rows):
        a, b = rows[i][False], rows[i][True]
        cy = lambda r: r[5] * 1e-3 * 1.9e9 * 148 / max(r[4], 1)
        ta += a[5]; tb += b[5]
        print("%3d %8d %4d %4d %2d | %9d %8.3f %7.0f | %9d %8.3f %7.0f | %6.2f %6.2f" %
              (i, a[0], a[1], a[2], a[3], a[4], a[5], cy(a), b[4], b[5], cy(b), b[4] / max(a[4], 1), b[5] / a[5]))
    print("sparse conv total: %.3f -> %.3f ms" % (ta, tb))


if __name__ == "__main__":
    main()
