#!/usr/bin/env python
"""Top stall sites of one kernel of an ncu report: tools/ncu_stalls.py REPORT.ncu-rep LAUNCH_INDEX [N]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 24
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip), "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
print(rows[0][1][:100])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 1 and r[0] != "Address"]
idx = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: int(r[idx[k]] or 0)
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
keys = [k for k in hdr if k.startswith("stall_") and "(" not in k]
agg = {k: sum(num(r, k) for r in data) for k in keys}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    st = sorted([(k, num(r, k)) for k in keys if num(r, k) > 0], key=lambda kv: -kv[1])[:2]
    print(str(num(r, "# Samples")).rjust(6), r[idx["Address"]][-5:], r[idx["Source"]][:64].ljust(64), st)
