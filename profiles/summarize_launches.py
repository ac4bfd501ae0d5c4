"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
tot = 0.0
n = 0
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = r[ki].split("(")[0][:70]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    n += 1
print(f"Launch list: {n} launches, {tot/1e3:.2f} ms of kernel time (ncu per-launch times are cold-cache and")
print("serialised: compare SHARES, not absolutes).\n")
print("| kernel | launches | total us | avg us | share |")
print("|---|---:|---:|---:|---:|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"| `{k}` | {c} | {t:.1f} | {t/c:.1f} | {100*t/tot:.2f}% |")
