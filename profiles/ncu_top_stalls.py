"""Top stall-sample SASS lines of one kernel in an ncu report (source page).
usage: python profiles/ncu_top_stalls.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ci, si = h.index("# Samples"), h.index("Source")
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((float(r[ci]), idx, r[si][:120]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("samples", tot)
for d in sorted(data, reverse=True)[:n]:
    print(f"{d[0] / tot * 100:5.1f}%  #{d[1]:<5} {d[2]}")
