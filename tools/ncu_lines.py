#!/usr/bin/env python
"""Per-source-line summary of an ncu report: python tools/ncu_lines.py report.ncu-rep kernel_regex units [min_share]
(units = the number the instruction counts are divided by, e.g. 32-read tiles)."""
import csv
import subprocess
import sys

rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, fname, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        out.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), fname, int(r[0]), r[1].strip()[:110]))
    except Exception:
        pass
tot, tots = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"instructions per unit {tot / units:.1f}, samples {tots}")
for o in sorted(out, key=lambda o: (o[2], o[3])):
    if o[0] / tot >= min_share or o[1] / max(tots, 1) >= min_share:
        print(f"{o[0] / units:7.1f} i/u  samp {100.0 * o[1] / max(tots, 1):5.1f}%  {o[2]}:{o[3]}  {o[4]}")
