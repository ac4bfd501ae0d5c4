"""Per-instruction warp-stall samples of the chain loop from an `ncu --set full --import-source on` report.
usage: python profiles/chain_source.py report.ncu-rep [kernel-substring] [min-samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "psi_fwd_cl_kernel"
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kern.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for k in kern:
    if want not in k["name"]:
        continue
    h = k["hdr"]
    si, ws, ie = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    data = [(int(r[ws] or 0), int(r[ie] or 0), r[si].strip()) for r in k["rows"] if len(r) > ie]
    tot = sum(d[0] for d in data)
    # the chain loop = the execution-count bucket (2 % wide) that collects the most stall samples
    buckets = {}
    for d in data:
        if d[1] > 0:
            key = round(50 * __import__("math").log(d[1]))
            buckets[key] = buckets.get(key, 0) + d[0]
    best = max(buckets, key=buckets.get)
    sel = [(i, d) for i, d in enumerate(data) if d[1] > 0 and round(50 * __import__("math").log(d[1])) == best]
    loop = sum(d[0] for _, d in sel)
    print(f"{k['name'][:60]}: {tot} samples, chain loop {loop} ({100*loop/tot:.0f}%), {len(sel)} instructions")
    for i, d in sel:
        if d[0] >= thr:
            print(f"  {i:5d} {d[0]:7d} {100*d[0]/loop:5.1f}%  {d[2][:80]}")
