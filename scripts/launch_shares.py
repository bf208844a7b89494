"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step).
Usage: python scripts/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_list_shares.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
n = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = name.replace("decnet::", "")
    agg[name][0] += 1
    agg[name][1] += us
    total += us
    n += 1
ours = sum(v[1] for k, v in agg.items() if any(t in k for t in ("conv2d::", "conv2dtc::", "conv2dnhwc::", "conv2drows::", "conv3d::", "glue::", "sparse::", "detail::", "codec::", "featext::")))
print(f"# total {total:.1f} us over {n} launches; hand-written decnet kernels: {100 * ours / total:.1f} % of the device time")
print("share%  launches  total_us  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * v[1] / total:6.2f}  {v[0]:6d}  {v[1]:10.1f}  {k[:100]}")
