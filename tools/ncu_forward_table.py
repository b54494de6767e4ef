#!/usr/bin/env python
"""REPORT.ncu-rep (ncu --set full of tools/forward_once.py N) -> markdown table + traffic JSON.
usage: tools/ncu_forward_table.py REPORT.ncu-rep N_IMAGES OUT_PREFIX [capture note]"""
import csv
import json
import subprocess
import sys

rep, n_img, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
# (layer, MACs per 800x800 image)
L = [("stem", 400 * 400 * 64 * 49)]
for b in ("0", "1"):
    L += [(f"l1.{b}.c1", 200 * 200 * 64 * 576), (f"l1.{b}.c2", 200 * 200 * 64 * 576)]
for li, (hw, c) in enumerate(((100, 128), (50, 256), (25, 512)), start=2):
    k = 9 * c
    L += [(f"l{li}.0.c1(s2)", hw * hw * c * (k // 2)), (f"l{li}.0.c2(+ds)", hw * hw * c * (k + c // 2)),
          (f"l{li}.1.c1", hw * hw * c * k), (f"l{li}.1.c2", hw * hw * c * k)]
L += [("in5", 25 * 25 * 256 * 512), ("in4(+sum)", 50 * 50 * 256 * 256)]
L += [("out5(x4)", 25 * 25 * 64 * 2304), ("out4(x2)", 50 * 50 * 64 * 2304)]
# FPN levels 3 and 2 without their 256-channel intermediates (detector.cu prep_fused_fpn_level): four 4-tap class
# convolutions of the level above's backbone feature + the composed 3x3 on the level's own; MACs actually executed
# parity classes run as pairs: classes (a,1) | (a,0) of neighbouring columns share one operand tile (N = 128, conv_halo pair mode)
L += [(f"out3.pair{a}", 50 * 50 * 128 * 1024) for a in (0, 1)] + [("out3.x(+res)", 100 * 100 * 64 * 1152)]
L += [(f"out2.pair{a}", 100 * 100 * 128 * 512) for a in (0, 1)] + [("out2.x(+res)", 200 * 200 * 64 * 576)]
# cat3 = [p5^4 | p4^2 | p3] reaches bin_conv1 through its class pairs (prep_fused_bin_p3); the main part reads p2
L += [(f"bin_conv1.pair{a}", 100 * 100 * 128 * 768) for a in (0, 1)] + [("bin_conv1.main(+res)", 200 * 200 * 64 * 576)]
L += [("head", 200 * 200 * 256 * 64 + 400 * 400 * 4 * 64)]

# REPORT may also be the `ncu -i REPORT.ncu-rep --page raw --csv` export (reports over 64 MiB do not travel back from the GPU box)
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def val(r, name, scaled=False):
    i = col[name]
    v = float(r[i].replace(",", "") or 0)
    return v * SC.get(units[i], 1.0) if scaled else v


if len(data) != len(L):
    sys.exit(f"{len(data)} launches in the report, {len(L)} layers expected")
lines = [f"# ncu --set full, one detector forward ({n_img} images 800x800, BF16 mode)", "", note, "",
         f"| layer | kernel | us / {n_img} images | TFLOP/s | tensor pipe active % | DRAM read MB | DRAM write MB | DRAM % | L2 % | L1TEX % | issue % |",
         "|---|---|---|---|---|---|---|---|---|---|---|"]
per_layer, tot_us, tot_bytes, tot_flop = [], 0.0, 0.0, 0.0
for (name, macs), r in zip(L, data):
    us = val(r, "gpu__time_duration.sum", True)
    rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
    flop = 2.0 * macs * n_img
    kern = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("ocrb::", "")
    g = lambda k: val(r, k) if k in col else float("nan")
    lines.append(f"| {name} | {kern} | {us:.1f} | {flop / us / 1e6:.0f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | "
                 f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                 f"{g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {g('sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} |")
    per_layer.append({"layer": name, "us": us, "dram_bytes": rd + wr})
    tot_us += us; tot_bytes += rd + wr; tot_flop += flop
lines += ["", f"Sum: {tot_us / 1e3:.2f} ms / {n_img} images = {tot_us / n_img:.1f} us/image, {tot_flop / tot_us / 1e6:.0f} TFLOP/s overall; "
          f"DRAM {tot_bytes / 1e9:.2f} GB = {tot_bytes / n_img / 1e6:.0f} MB/image."]
open(prefix + "_table.md", "w").write("\n".join(lines) + "\n")
json.dump({"capture": note, "launches": len(data), "images": n_img, "dram_bytes_total": tot_bytes, "dram_bytes_per_launch": tot_bytes / len(data),
           "dram_bytes_per_image": tot_bytes / n_img,
           "algorithmic_bytes_note": "unfused bf16 activation traffic 285.8 MB/image (SURVEY 8d)", "per_layer": per_layer},
          open(prefix + "_traffic.json", "w"), indent=1)
print("\n".join(lines[-8:]))
