#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> per-kernel shares as a markdown table.
usage: tools/launch_summary.py LAUNCHES.csv OUT.md "title / command note" """
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
agg = {}
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("ocrb::", "").strip()
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[14].replace(",", "")) / 1e3)
tot = sum(t for _, t in agg.values())
tc = sum(t for k, (_, t) in agg.items() if k.startswith(("conv_halo", "conv_tc", "conv_lateral", "stem_tc")))
rec = sum(t for k, (_, t) in agg.items() if k.startswith(("rec_", "conv_fp32")))
lines = [f"# {note}", "",
         f"`ncu --metrics gpu__time_duration.sum --clock-control none --csv` — raw list: `{src.split('/')[-1]}` ({len(rows)} launches, serialised, cold cache)."
         f"  Detector tcgen05 kernels: {100 * tc / tot:.1f} % of the GPU time, glyph-net kernels {100 * rec / tot:.1f} % "
         "(bench.py's `roofline.share_of_step`, measured with CUDA events in the pipelined step, must agree).", "",
         "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot < 0.001:
        continue
    lines.append(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f} % |")
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
