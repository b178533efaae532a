#!/usr/bin/env python
"""Runs one BASELINE.md workload through the device-resident entry point a few times (profiling helper for ncu).
--layout window: the batch cut to idp_panel_window bases per read and 16-byte records (what the plugin ships by default
for 5'-only runs); --layout whole: 75 B rows; --records 16|32."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import synth, synth_gpu

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--layout", default="auto", choices=["auto", "window", "whole"])
ap.add_argument("--records", type=int, default=16, choices=[16, 32])
args = ap.parse_args()
dev = torch.device("cuda", 0)
w = synth.WORKLOADS[args.workload]
panel = synth.make_panel(w["pairs"], w["panel_seed"], degenerate=w["degenerate"], mapped=w["mapped"])
opts = dict(w["options"])
tool = idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, device=0, **opts)
do3 = opts.get("three_prime", -1) >= 0
window = args.layout == "window" or (args.layout == "auto" and not do3)
pt = synth_gpu.PanelTensors(panel, dev)
seqs, flags, locs = [], [], [[], [], []]
for ci, lo in enumerate(range(0, args.pairs, 1_000_000)):
    g = synth_gpu.make_reads(pt, min(1_000_000, args.pairs - lo), w["read_seed"] + ci, indel_frac=w["indel_frac"], mapped_frac=w["mapped_frac"])
    seqs.append(synth_gpu.pack(synth_gpu.window_codes(g["codes"], g["flags"], tool.window()) if window else g["codes"]))
    flags.append(g["flags"])
    if g["ref_idx"] is not None:
        for k, key in enumerate(("ref_idx", "ustart", "uend")):
            locs[k].append(g[key])
    del g
s4, fl = torch.cat(seqs), torch.cat(flags)
loc = [torch.cat(x) for x in locs] if locs[0] else None
n = fl.numel()
fixed_len = tool.window() if window else 150
f5 = torch.empty((n, args.records), dtype=torch.uint8, device=dev)
f3 = torch.empty((n, args.records), dtype=torch.uint8, device=dev) if do3 else None
tool.set_stream(torch.cuda.current_stream().cuda_stream)
print(f"{args.workload}: {args.pairs} pairs, fixed_len {fixed_len} ({s4.shape[1]} B rows), {args.records}-byte records")
for i in range(args.iters):
    tool.process_batch_device(n, s4.data_ptr(), fl.data_ptr(), f5.data_ptr(), f3.data_ptr() if f3 is not None else 0, fixed_len=fixed_len,
                              ref_idx_ptr=loc[0].data_ptr() if loc else 0, ustart_ptr=loc[1].data_ptr() if loc else 0, uend_ptr=loc[2].data_ptr() if loc else 0,
                              compact=args.records == 16)
    t = tool.timings()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()})
if t["sw_cells"] and t["ms_sw_score"]:
    print("GCUPS", t["sw_cells"] / t["ms_sw_score"] / 1e6)
tool.close()
