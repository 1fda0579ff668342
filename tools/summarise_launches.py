"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarise_launches.py gpurun_out/launches_cfg3.csv "command line that was profiled" > profiles/rNN_launches_cfg3.csv
"""
import csv
import re
import sys
from collections import OrderedDict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"# ncu launch list: {cmd}")
print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)")
print(f"# total {total / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
print("kernel,launches,total_ms,avg_us,share")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name},{n},{ns / 1e6:.4f},{ns / n / 1e3:.2f},{ns / total:.5f}")
