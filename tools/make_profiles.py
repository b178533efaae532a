#!/usr/bin/env python
"""Copies the round's measurement files from gpurun_out/ into profiles/ and derives profiles/roofline_traffic.json (the DRAM
bytes of the scan stage per launch, what bench.py reports as roofline.traffic) and the per-kernel shares of one device-resident
step from the ncu launch list.  Usage: python tools/make_profiles.py r02"""
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
copied = []
for f in sorted(os.listdir(src)):
    if f.startswith(tag + "_") and (f.endswith(".txt") or f.endswith(".json") or f.endswith(".csv")) and os.path.getsize(os.path.join(src, f)) > 0:
        if f.endswith("_launches_bench.csv"):
            continue  # filtered below: the raw list also holds the torch kernels of the read generator
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        copied.append(f)

# ---- launch list of the bench command: our kernels only, and the shares of the first complete device-resident chain ----
rows = list(csv.reader(open(os.path.join(src, f"{tag}_launches_bench.csv"))))
hdr, ours = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and "idp::" in r[hdr.index("Kernel Name")]:
        ours.append((r[hdr.index("Kernel Name")], float(r[hdr.index("Metric Value")]) / 1000.0))
with open(os.path.join(dst, f"{tag}_launches_bench.csv"), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gapped-stress "
             "--no-workloads --no-whole-reads`: the library's kernels in launch order (the torch kernels of the read generator are left out), microseconds\n")
    for name, us in ours:
        fh.write(f"\"{name}\",{us:.3f}\n")
chain = []
for name, us in ours:
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("idp::", "")
    if chain and short.startswith("scan_fx_kernel"):
        break
    chain.append((short, us))
total = sum(us for _, us in chain)
shares = {}
for short, us in chain:
    key = short if short not in shares else short + " (2nd)"
    shares[key] = {"us": round(us, 1), "share": round(us / total, 3)}

# ---- DRAM traffic of the scan stage from the --set full summary ----
txt = open(os.path.join(src, f"{tag}_ncu_scan_c2_10Mpairs.txt")).read()
per = re.findall(r"== launch \d+: (.*?)\n(.*?)(?=\n== launch|\Z)", txt, flags=re.S)
scan = {}
for name, body in per:
    short = "scan_fx_kernel" if "scan_fx" in name else "scan_kernel"
    rd = float(re.search(r"DRAM read\s+([0-9.]+) (\w+)", body).group(1)) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[re.search(r"DRAM read\s+[0-9.]+ (\w+)", body).group(1)]
    wr = float(re.search(r"DRAM written\s+([0-9.]+) (\w+)", body).group(1)) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}[re.search(r"DRAM written\s+[0-9.]+ (\w+)", body).group(1)]
    dur = float(re.search(r"duration\s+([0-9.]+) us", body).group(1))
    scan[short] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "ncu_duration_us": dur}
bench = json.load(open(os.path.join(src, f"{tag}_bench_1gpu.json")))
out = {
    "workload": "c2", "pairs": 10000000, "layout": bench["config"]["layout"],
    "scan_kernel_dram_bytes_per_launch": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in scan.values()),
    "kernels": scan,
    "source": f"ncu --set full --clock-control none -k regex:scan_ -c 2 on `python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gapped-stress "
              f"--no-workloads --no-whole-reads` (B200, round 2); summary in profiles/{tag}_ncu_scan_c2_10Mpairs.txt, launch list profiles/{tag}_launches_bench.csv. "
              "The two compact_mask_kernel launches of the stage (5 MB of masks, 12 MB of list entries) are not in the figure.",
    "launch_shares_device_resident_step": shares,
    "note": "per-launch times under ncu are cold-cache and serialised; the CUDA-event share of the scan stage in the bench line is "
            f"{bench['stages_ms']['ms_scan']:.3f} of {bench['stages_ms']['ms_total']:.3f} ms",
}
json.dump(out, open(os.path.join(dst, "roofline_traffic.json"), "w"), indent=1)
print("copied", len(copied), "files;", "scan stage DRAM bytes per launch:", out["scan_kernel_dram_bytes_per_launch"])
print(json.dumps(shares, indent=1))
