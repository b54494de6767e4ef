#!/usr/bin/env python
"""profiles/r2_sass_extract.md: per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, UTMALDG / UTMASTG = TMA tensor loads / stores, LDTM = tcgen05.ld),
from `cuobjdump -sass ocr_rs_b200/libocrb.so`.  Runs in the build container (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "ocr_rs_b200", "libocrb.so")], capture_output=True, text=True).stdout
    rows, tot = [], collections.Counter()
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        c = collections.Counter()
        for m in re.finditer(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG|UTMASTG|LDTM|FFMA2)\b", f):
            c[m.group(1)] += 1
        if c.get("UTCHMMA", 0) + c.get("UTCHMMA.2CTA", 0) + c.get("LDTM", 0) > 0:
            rows.append((name, c))
        tot.update(c)
    names = subprocess.run(["c++filt"], input="\n".join(n for n, _ in rows), capture_output=True, text=True).stdout.splitlines()
    out = ["# SASS evidence of the tensor-core / TMA path (round 2)", "",
           "`python tools/sass_extract.py` = `cuobjdump -sass ocr_rs_b200/libocrb.so` (sm_100a), mnemonics counted per kernel.",
           "`UTCHMMA` = tcgen05.mma, `.2CTA` = cta_group::2, `UTMALDG` / `UTMASTG` = TMA tensor loads / stores, `LDTM` = tcgen05.ld "
           "(TMEM -> registers), `FFMA2` = packed fp32x2 FMA.", "",
           "| kernel | UTCHMMA | UTCHMMA.2CTA | UTMALDG | UTMASTG | LDTM | FFMA2 |", "|---|---|---|---|---|---|---|"]
    for (name, c), d in sorted(zip(rows, names), key=lambda r: r[1]):
        d = re.sub(r"\(.*", "", d)[:120]
        out.append(f"| `{d}` | {c.get('UTCHMMA', 0)} | {c.get('UTCHMMA.2CTA', 0)} | {c.get('UTMALDG', 0)} | {c.get('UTMASTG', 0)} | {c.get('LDTM', 0)} | {c.get('FFMA2', 0)} |")
    out += ["", f"Library totals: {tot.get('UTCHMMA', 0)} UTCHMMA + {tot.get('UTCHMMA.2CTA', 0)} UTCHMMA.2CTA, {tot.get('UTMALDG', 0)} UTMALDG, "
            f"{tot.get('UTMASTG', 0)} UTMASTG, {tot.get('LDTM', 0)} LDTM, {tot.get('FFMA2', 0)} FFMA2 in {len(rows)} tensor-core kernels.", ""]
    path = os.path.join(ROOT, "profiles", "r2_sass_extract.md")
    open(path, "w").write("\n".join(out))
    print(out[-2])


if __name__ == "__main__":
    sys.exit(main())
