#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: runs of SASS instructions with the same execution count.
Usage: python tools/ncu_blocks.py file.csv [n_tiles] [--dump FROM TO]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tiles = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
tots = sum(int(r[ix["# Samples"]]) for r in data)
print("warp instructions", tot, "per tile", round(tot / tiles, 1), "samples", tots)
if "--dump" in sys.argv:
    a, b = sys.argv[sys.argv.index("--dump") + 1: sys.argv.index("--dump") + 3]
    on = False
    for r in data:
        t = r[0][-5:]
        if t == a:
            on = True
        if on:
            print(t, r[1][:80], r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]], r[ix["# Samples"]])
        if t == b:
            break
    sys.exit(0)
prev, start, acc, sacc, n = None, 0, 0, 0, 0
for i, r in enumerate(data + [None]):
    e = int(r[ix["Instructions Executed"]]) if r else -1
    s = int(r[ix["# Samples"]]) if r else 0
    if prev is None or abs(e - prev) > 0.02 * max(e, prev, 1):
        if prev is not None and acc > 0.004 * tot:
            print(f"{data[start][0][-5:]}..{data[i-1][0][-5:]} n={n} exec/instr={prev} share={acc/tot:.3f} samp={sacc/tots:.3f} thr={data[start][ix['Avg. Threads Executed']]}")
        start, acc, sacc, n = i, 0, 0, 0
    prev = e
    acc += max(e, 0)
    sacc += s
    n += 1
