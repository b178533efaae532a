#!/usr/bin/env python
"""Text summary of an `ncu --set full` report, one block per captured kernel launch (what goes under profiles/):
    python tools/ncu_summary.py report.ncu-rep [units_per_launch ...] > profiles/rNN_ncu_<kernel>.txt
units_per_launch (optional, one per launch): e.g. the number of 32-read tiles, to print instructions per unit."""
import csv
import subprocess
import sys

rep = sys.argv[1]
units = [float(x) for x in sys.argv[2:]]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, unit_row = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "blocks / SM by registers"), ("launch__occupancy_limit_shared_mem", "blocks / SM by shared memory"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads per instruction"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput (%)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "  of which bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "DRAM active (%)"),
    ("sass__inst_executed_local_loads", "local loads (spills)"), ("sass__inst_executed_local_stores", "local stores (spills)"),
]
STALL = "smsp__average_warps_issue_stalled_"
for n, v in enumerate(rows[2:]):
    name = v[ix["Kernel Name"]]
    print(f"== launch {n}: {name}")
    dur_ns = None
    for key, label in KEYS:
        if key not in ix:
            continue
        val, u = v[ix[key]], unit_row[ix[key]]
        print(f"   {label:34s} {val} {u}")
        if key == "gpu__time_duration.sum":
            try:
                dur_ns = float(val) * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "second": 1e9, "nsecond": 1.0}.get(u, 1.0)
            except ValueError:
                pass
    try:
        rd = float(v[ix["dram__bytes_read.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit_row[ix["dram__bytes_read.sum"]], 1.0)
        wr = float(v[ix["dram__bytes_write.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit_row[ix["dram__bytes_write.sum"]], 1.0)
        if dur_ns:
            print(f"   {'DRAM bytes per launch':34s} {rd + wr:.0f} ({(rd + wr) / dur_ns:.0f} GB/s over the launch)")
    except (KeyError, ValueError):
        pass
    if n < len(units) and units[n] > 0:
        print(f"   {'warp instructions per unit':34s} {float(v[ix['smsp__inst_executed.sum']]) / units[n]:.1f} ({units[n]:.0f} units)")
    stalls = []
    for h, i in ix.items():
        if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(v[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("   stalls (warps per issue slot):     " + ", ".join(f"{s} {x:.2f}" for x, s in stalls[:6]))
