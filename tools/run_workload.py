#!/usr/bin/env python
"""Runs one BASELINE.md workload through the device-resident entry point a few times (profiling helper for ncu)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
panel, reads, opts = synth.workload(args.workload, args.pairs)
dev = torch.device("cuda", 0)
tool = idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, device=0, **opts)
s4 = torch.from_numpy(reads.seq4()).to(dev)
fl = torch.from_numpy(reads.flags.view(np.int16)).to(dev)
loc = [torch.from_numpy(a).to(dev) for a in (reads.ref_idx, reads.ustart, reads.uend)] if reads.ref_idx is not None else None
f5 = torch.empty((reads.n, 32), dtype=torch.uint8, device=dev)
f3 = torch.empty((reads.n, 32), dtype=torch.uint8, device=dev) if opts.get("three_prime", -1) >= 0 else None
tool.set_stream(torch.cuda.current_stream().cuda_stream)
for i in range(args.iters):
    tool.process_batch_device(reads.n, s4.data_ptr(), fl.data_ptr(), f5.data_ptr(), f3.data_ptr() if f3 is not None else 0, fixed_len=150,
                              ref_idx_ptr=loc[0].data_ptr() if loc else 0, ustart_ptr=loc[1].data_ptr() if loc else 0, uend_ptr=loc[2].data_ptr() if loc else 0)
    t = tool.timings()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()})
if t["sw_cells"] and t["ms_sw_score"]:
    print("GCUPS", t["sw_cells"] / t["ms_sw_score"] / 1e6)
tool.close()
