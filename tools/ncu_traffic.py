"""Turn an ncu --csv metrics log of a forward pass into per-family sums (time share, DRAM / L2 bytes).
usage: python tools/ncu_traffic.py <log.csv> [out.json]"""
import collections
import csv
import json
import sys

rows = collections.OrderedDict()
for l in csv.reader(open(sys.argv[1])):
    if len(l) < 15 or l[0] == "ID":
        continue
    r = rows.setdefault(int(l[0]), {"name": l[4]})
    try:
        r[l[12]] = float(l[14].replace(",", ""))
    except ValueError:                      # "n/a" (a metric the kernel does not report)
        pass
fam = collections.OrderedDict()
for r in rows.values():
    n = r["name"]
    key = ("conv_tc" if "conv_tc_kernel" in n else "conv_simt" if "conv_simt" in n else "voxelize" if "vox_" in n else
           "rowsort" if "rowsort" in n else
           "rulebook" if any(k in n for k in ("neighbors", "outset", "coord_index", "nbr_")) else "scan" if "scan" in n else
           "wgrad" if "wgrad" in n else "other")
    f = fam.setdefault(key, collections.Counter())
    f["launches"] += 1
    for k, v in r.items():
        if k != "name" and "pct" not in k:
            f[k] += v
tot = sum(f["gpu__time_duration.sum"] for f in fam.values())
out = {}
for k, f in fam.items():
    out[k] = dict(launches=int(f["launches"]), us=f["gpu__time_duration.sum"] / 1e3, share=f["gpu__time_duration.sum"] / tot,
                  dram_mb=(f.get("dram__bytes_read.sum", 0) + f.get("dram__bytes_write.sum", 0)) / 1e6,
                  l2_gb=f.get("lts__t_bytes.sum", 0) / 1e9)
    print("%-10s %3d launches %9.1f us  %5.1f %%  dram %8.1f MB  L2 %6.2f GB" %
          (k, out[k]["launches"], out[k]["us"], 100 * out[k]["share"], out[k]["dram_mb"], out[k]["l2_gb"]))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
