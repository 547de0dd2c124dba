#!/usr/bin/env python
"""Aggregates the SASS-level source page of an ncu report by CUDA source line:
   python profiles/ncu_lines.py report.ncu-rep kernel_regex [min_pct]
prints, per source line, the share of executed warp instructions and of stall samples (needs -lineinfo and
`ncu --import-source on`)."""
import csv
import subprocess
import sys
from collections import OrderedDict

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, recs, cur_file, cur_line, cur_src = None, [], "", None, ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; continue
    if not hdr or len(r) < len(hdr) - 2:
        continue
    if r[0] != "":
        cur_line = (cur_file.split("/")[-1], r[0]); cur_src = r[1]
    if r[2].startswith("0x"):
        try:
            n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        recs.append((cur_line, cur_src, n, s))
tot = sum(x[2] for x in recs) or 1; tots = sum(x[3] for x in recs) or 1
print(f"# {kern}: {tot} warp instructions, {tots} stall samples")
agg = OrderedDict()
for line, src, n, s in recs:
    a = agg.setdefault(line, [src, 0, 0]); a[1] += n; a[2] += s
for k, v in agg.items():
    if 100 * v[1] / tot >= min_pct or 100 * v[2] / tots >= min_pct:
        print(f"{k[0][:14]:14s} {k[1]:>5s} {100 * v[1] / tot:5.1f}%i {100 * v[2] / tots:5.1f}%s  {v[0].strip()[:110]}")
