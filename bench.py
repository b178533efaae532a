#!/usr/bin/env python
"""bench.py -- read pairs/sec of the IdentifyPrimers matching path on B200 (BASELINE.json metric), weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--pairs P] [--impl reference]

A step is one pass of the matcher chain over one batch of synthetic read pairs (default: workload c2 = 10M read pairs
2x150 bp against a 96-pair IUPAC-degenerate panel, -k 6 -K 10, per GPU).  `value` is timed with the batch resident in
HBM (CUDA events on the launching stream); `e2e` goes through idp_match_batch with pinned HOST buffers, H2D and D2H
inside the timed region.  `--impl reference` times the reference's algorithm on the host cores: the reference is
Scala/JVM and no JVM exists here, so that arm is the CPU oracle port (oracle/), the only other place it is executed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read_pairs_per_sec_identifyprimers"
UNIT = "read pairs/s"
ALGO_BYTES_PER_PAIR = 182  # 2 reads x (75 B packed bases in + 16 B result out)


def workload_desc(name, pairs):
    from fg_idprimer_b200 import synth
    w = synth.WORKLOADS[name]
    o = w["options"]
    return (f"{name}: {pairs} read pairs 2x150 bp vs {w['pairs']}-pair {'IUPAC-degenerate ' if w['degenerate'] else ''}"
            f"{'mapped ' if w['mapped'] else ''}amplicon panel, -k {o.get('ungapped_kmer_length')} -K {o.get('gapped_kmer_length')}"
            f"{' -3 %d' % o['three_prime'] if 'three_prime' in o else ''}, full Location->Ungapped->Gapped chain")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_panel(panel, opts):
    from oracle import oracle_py as O
    ref = {n: i for i, n in enumerate(panel.ref_names)}

    class _P:
        pass
    prs = []
    for p in panel.primers:
        q = _P()
        q.__dict__.update(pair_id=p.pair_id, primer_id=p.primer_id, sequence=p.sequence, positive_strand=p.positive_strand,
                          ref_name=p.ref_name, start=p.start, end=p.end, ref_idx=ref.get(p.ref_name, -1))
        prs.append(q)
    params = O.make_params(k_ungapped=opts.get("ungapped_kmer_length"), k_gapped=opts.get("gapped_kmer_length"), three_prime=opts.get("three_prime", -1))
    return O.Panel(prs, params)


def time_oracle(name, target_seconds, cores):
    """Times the CPU oracle on a bounded sample of the workload; returns (pairs/s, sample pairs, seconds)."""
    from fg_idprimer_b200 import synth
    panel, reads, opts = synth.workload(name, 20000)
    op = oracle_panel(panel, opts)
    t0 = time.perf_counter()
    op.process_batch(reads.oracle_arrays(), n_threads=cores)
    rate = 20000 / max(time.perf_counter() - t0, 1e-6)
    n = int(min(max(rate * target_seconds, 20000), 3_000_000))
    panel, reads, opts = synth.workload(name, n)
    arrays = reads.oracle_arrays()
    passes, dt = 0, 0.0
    while dt < target_seconds and passes < 64:  # repeat the bounded sample until ~target_seconds of CPU work are timed
        t0 = time.perf_counter()
        op.process_batch(arrays, n_threads=cores)
        dt += time.perf_counter() - t0
        passes += 1
    return n * passes / dt, n * passes, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fg_idprimer_b200 import synth
    cores = os.cpu_count() or 1
    panel, reads, opts = synth.workload(args.workload, 20000)
    op = oracle_panel(panel, opts)
    t0 = time.perf_counter()
    op.process_batch(reads.oracle_arrays(), n_threads=cores)
    rate = 20000 / max(time.perf_counter() - t0, 1e-6)
    total = args.steps + args.warmup
    n = int(min(max(rate * min(8.0, 150.0 / total), 20000), 2_000_000))
    panel, reads, opts = synth.workload(args.workload, n)
    arrays = reads.oracle_arrays()
    times = []
    for i in range(total):
        t0 = time.perf_counter()
        op.process_batch(arrays, n_threads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    value = n * args.steps / dt
    sample = f"{n} read pairs of {args.workload} per step, CPU oracle port (pthreads x {cores}), matching only (no BAM I/O)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, args.pairs), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gapped-stress", action="store_true")
    ap.add_argument("--e2e-window", action="store_true",
                    help="also time idp_match_batch on the windowed batch (idp_panel_window bases per read; 5'-only workloads)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fg_idprimer_b200 as idp
    from fg_idprimer_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the IdentifyPrimers GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: the NCCL banner ("NCCL version ...") is written to fd 1 by the library when the
        # communicator comes up, so fd 1 points at stderr until the first collective has run
        if "IDP_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["IDP_NCCL_DEBUG"]
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = torch.device("cuda", local)

    # ---- synthetic batch, generated in chunks straight into pinned host memory ----
    w = synth.WORKLOADS[args.workload]
    panel = synth.make_panel(w["pairs"], w["panel_seed"], degenerate=w["degenerate"], mapped=w["mapped"])
    opts = dict(w["options"])
    n_pairs, n_reads, stride = args.pairs, 2 * args.pairs, 75
    h_seq4 = torch.empty((n_reads, stride), dtype=torch.uint8, pin_memory=True)
    h_flags = torch.empty(n_reads, dtype=torch.int16, pin_memory=True)
    h_loc = [torch.empty(n_reads, dtype=torch.int32, pin_memory=True) for _ in range(3)] if w["mapped"] else None
    h_five = torch.empty((n_reads, 32), dtype=torch.uint8, pin_memory=True)
    h_three = torch.empty((n_reads, 32), dtype=torch.uint8, pin_memory=True) if opts.get("three_prime", -1) >= 0 else None
    # opt-in: the same reads cut to the bases the 5' chain can reach (what idp_bam_to_batch_window ships), for `e2e_window`
    win = 0
    if args.e2e_window and opts.get("three_prime", -1) < 0:
        ks = [k for k in (opts.get("ungapped_kmer_length"), opts.get("gapped_kmer_length")) if k]
        win = max([max(p.length for p in panel.primers), max(len(p.sequence) for p in panel.primers) + opts.get("slop", 5)] + ks)
        win = min(win, 150)
    h_win = torch.empty((n_reads, (win + 1) // 2), dtype=torch.uint8, pin_memory=True) if win else None
    t_gen = time.perf_counter()
    chunk = 1_000_000
    sample_reads = None
    for ci, lo in enumerate(range(0, n_pairs, chunk)):
        n = min(chunk, n_pairs - lo)
        rd = synth.make_reads(panel, n, w["read_seed"] + 1000 * rank + ci, indel_frac=w["indel_frac"], mapped_frac=w["mapped_frac"])
        h_seq4.numpy()[2 * lo:2 * (lo + n)] = rd.seq4()
        h_flags.numpy().view(np.uint16)[2 * lo:2 * (lo + n)] = rd.flags
        if win:  # first `win` stored bases, or the last ones of a mapped reverse-strand read (PM:39-45)
            from_end = ((rd.flags & 0x4) == 0) & ((rd.flags & 0x10) != 0)
            cw = np.where(from_end[:, None], rd.codes[:, 150 - win:], rd.codes[:, :win])
            h_win.numpy()[2 * lo:2 * (lo + n)] = synth.pack_fixed(cw)
        if h_loc is not None:
            for t, a in zip(h_loc, (rd.ref_idx, rd.ustart, rd.uend)):
                t.numpy()[2 * lo:2 * (lo + n)] = a
        if ci == 0:
            sample_reads = rd
    t_gen = time.perf_counter() - t_gen

    tool = idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, device=local, **opts)
    d_seq4, d_flags = h_seq4.to(dev, non_blocking=True), h_flags.to(dev, non_blocking=True)
    d_loc = [t.to(dev, non_blocking=True) for t in h_loc] if h_loc is not None else None
    d_five = torch.empty((n_reads, 32), dtype=torch.uint8, device=dev)
    d_three = torch.empty((n_reads, 32), dtype=torch.uint8, device=dev) if h_three is not None else None
    metrics_dev = torch.zeros(160, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    tool.set_stream(stream.cuda_stream)

    def step_resident():
        m, _ = tool.process_batch_device(n_reads, d_seq4.data_ptr(), d_flags.data_ptr(), d_five.data_ptr(),
                                         d_three.data_ptr() if d_three is not None else 0, fixed_len=150,
                                         ref_idx_ptr=d_loc[0].data_ptr() if d_loc else 0, ustart_ptr=d_loc[1].data_ptr() if d_loc else 0,
                                         uend_ptr=d_loc[2].data_ptr() if d_loc else 0)
        if world > 1:  # the only cross-GPU traffic: the 160 metric counters
            metrics_dev.copy_(torch.from_numpy(m))
            dist.all_reduce(metrics_dev)
        return tool.timings()

    host_batch = idp.ReadBatch.__new__(idp.ReadBatch)
    host_batch.seq4, host_batch.flags = h_seq4.numpy().reshape(-1), h_flags.numpy().view(np.uint16)
    host_batch.seq_off = host_batch.len = None
    host_batch.fixed_len = 150
    host_batch.ref_idx, host_batch.unclipped_start, host_batch.unclipped_end = ([t.numpy() for t in h_loc] if h_loc else [None] * 3)

    from fg_idprimer_b200 import _lib as L
    import ctypes as C

    def step_e2e():
        keep = []
        rs = host_batch.as_struct(keep)
        metrics = np.zeros(160, dtype=np.int64)
        n_align = np.zeros(1, dtype=np.int64)
        L.check(L.lib().idp_match_batch(tool._ctx, C.byref(rs), n_reads, h_five.data_ptr(), h_three.data_ptr() if h_three is not None else None,
                                        metrics.ctypes.data, n_align.ctypes.data))
        if world > 1:
            metrics_dev.copy_(torch.from_numpy(metrics))
            dist.all_reduce(metrics_dev)
        return tool.timings()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, use_events):
        for _ in range(args.warmup):
            fn()
        barrier()
        acc = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            acc.append(fn())
        ev1.record(stream)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        ms = ev0.elapsed_time(ev1) if use_events else wall_ms
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, acc

    clk = ClockSampler(local)
    clk.__enter__()
    ms_res, tm_res = timed(step_resident, True)
    # parity spot check of what was just timed (the checker, not the thing measured): first 20k pairs against the oracle
    torch.cuda.synchronize()
    parity = None
    if rank == 0:
        try:
            op = oracle_panel(panel, opts)
            k = min(20000, sample_reads.n // 2)
            want5, _, _, _ = op.process_batch(sample_reads.oracle_arrays(0, 2 * k), n_threads=os.cpu_count() or 1)
            got = d_five[:2 * k].cpu().numpy().view(L.RESULT_DTYPE).reshape(-1)
            parity = all(np.array_equal(got[f], want5[f]) for f in ("primer", "type", "status", "multi", "mm", "next_mm", "raw_score", "raw_next", "t_start", "t_end"))
        except Exception as e:  # pragma: no cover
            parity = f"check failed to run: {e}"
    ms_e2e, tm_e2e = timed(step_e2e, False)
    clk.__exit__(None, None, None)
    clocks = clk.summary()  # sampled across both timed regions

    value = world * n_pairs * args.steps / (ms_res / 1e3)
    e2e_value = world * n_pairs * args.steps / (ms_e2e / 1e3)
    scan_ms = statistics.mean(t["ms_scan"] for t in tm_res)
    achieved = ALGO_BYTES_PER_PAIR * n_pairs / (scan_ms / 1e3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and tj.get("pairs") == n_pairs:
            traffic = tj["scan_kernel_dram_bytes_per_launch"]
    except Exception:
        pass

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, n_pairs) + " (per GPU)", "parallelism": f"reads sharded over {world} GPU(s), panel replicated, all-reduce of 160 counters",
                   "l2": f"inputs {n_reads * 77 / 1e9:.2f} GB per step exceed the 126 MB L2, no flush needed", "generator_seconds": round(t_gen, 1)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tm_e2e[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(tm_e2e[-1]["d2h_bytes"]),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(sum(t["kernel_launches"] for t in tm_res)),
        "roofline": {"bound": "hbm", "kernel": "scan_kernel (k-mer prefilter + location + ungapped IUPAC scan)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": scan_ms,
                     "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PAIR * n_pairs},
        "clocks": clocks,
        "stages_ms": {k: statistics.mean(t[k] for t in tm_res) for k in ("ms_total", "ms_scan", "ms_gapped", "ms_three_prime", "ms_sw_score")},
        "work": {k: int(tm_res[-1][k]) for k in ("n_fallthrough", "n_alignments_5p", "n_alignments_3p", "sw_cells")},
        "parity_spot_check_vs_oracle": parity,
    }
    sw_ms = out["stages_ms"]["ms_sw_score"]
    if sw_ms > 0 and out["work"]["sw_cells"] > 0:
        out["gapped"] = {"gcups": out["work"]["sw_cells"] / (sw_ms / 1e3) / 1e9, "cells": out["work"]["sw_cells"], "sw_kernel_ms": sw_ms, "workload": args.workload}

    if win:
        try:
            assert tool.window() == win, (tool.window(), win)
            h_five_full = h_five.numpy().copy()
            wb = idp.ReadBatch.__new__(idp.ReadBatch)
            wb.seq4, wb.flags = h_win.numpy().reshape(-1), h_flags.numpy().view(np.uint16)
            wb.seq_off = wb.len = None
            wb.fixed_len = win
            wb.ref_idx, wb.unclipped_start, wb.unclipped_end = ([t.numpy() for t in h_loc] if h_loc else [None] * 3)

            def step_e2e_window():
                keep = []
                rs = wb.as_struct(keep)
                metrics = np.zeros(160, dtype=np.int64)
                n_align = np.zeros(1, dtype=np.int64)
                L.check(L.lib().idp_match_batch(tool._ctx, C.byref(rs), n_reads, h_five.data_ptr(), None, metrics.ctypes.data, n_align.ctypes.data))
                if world > 1:
                    metrics_dev.copy_(torch.from_numpy(metrics))
                    dist.all_reduce(metrics_dev)
                return tool.timings()

            ms_w, tm_w = timed(step_e2e_window, False)
            out["e2e_window"] = {"value": world * n_pairs * args.steps / (ms_w / 1e3), "unit": UNIT, "window_bases": win,
                                 "h2d_bytes_per_step": int(tm_w[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(tm_w[-1]["d2h_bytes"]),
                                 "ms_per_step": ms_w / args.steps,
                                 "same_results_as_whole_reads": bool(np.array_equal(h_five.numpy(), h_five_full))}
        except Exception as e:  # pragma: no cover
            out["e2e_window"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_gapped_stress:
        # gapped-alignment stress: no -K filter, every fall-through read aligns against all 2000 primers
        try:
            p3, r3, o3 = synth.workload("c3b", 200_000)
            with idp.IdentifyPrimers(p3.primers, device=local, **o3) as t3:
                t3.set_stream(stream.cuda_stream)
                s4, fl = torch.from_numpy(r3.seq4()).to(dev), torch.from_numpy(r3.flags.view(np.int16)).to(dev)
                f5 = torch.empty((r3.n, 32), dtype=torch.uint8, device=dev)
                tms = []
                for i in range(4):
                    t3.process_batch_device(r3.n, s4.data_ptr(), fl.data_ptr(), f5.data_ptr(), fixed_len=150)
                    if i:
                        tms.append(t3.timings())
                cells = tms[-1]["sw_cells"]; ms = statistics.mean(t["ms_sw_score"] for t in tms)
                out["gapped_stress"] = {"workload": "c3b: 200000 read pairs vs 1000-pair panel, 5% indel reads, -k 6, no -K", "gcups": cells / (ms / 1e3) / 1e9,
                                        "cells": int(cells), "alignments": int(tms[-1]["n_alignments_5p"]), "sw_kernel_ms": ms}
                try:
                    # integer issue roofline of the packed aligner (two alignments per register): per PAIR of cells
                    # 3 full-rate instructions (SHF, LOP3, LOP3) and 5 half-rate ones (IMAD, VIADD.16x2, 2 x VIADDMNMX.S16x2,
                    # VIMNMX3.S16x2), at the issue rates measured by tools/int_peak.cu and tools/int_peak16.cu
                    with open(os.path.join(ROOT, "profiles", "int_peak_b200.json")) as fh:
                        ip = json.load(fh)
                    with open(os.path.join(ROOT, "profiles", "int_peak16_b200.json")) as fh:
                        ip16 = json.load(fh)
                    full = ip["lop_iadd_per_clk_per_sm"]
                    half = min(ip16["per_clk_per_sm"]["viaddmax_s16x2"], ip16["per_clk_per_sm"]["vimax3_s16x2"], ip16["per_clk_per_sm"]["imad"])
                    clk_per_pair = 3.0 / full + 5.0 / half
                    peak = 2.0 / clk_per_pair * ip["sms"] * ip["sm_clock_mhz_max"] * 1e6 / 1e9
                    out["gapped_stress"].update({"roofline_gcups": peak, "frac": out["gapped_stress"]["gcups"] / peak,
                                                 "peak_source": "integer issue roofline: 3 full-rate + 5 half-rate instructions per two cells at the measured "
                                                                "rates (profiles/int_peak_b200.json, profiles/int_peak16_b200.json)",
                                                 "roofline_gcups_32bit_9_instr_per_cell": ip["lop_iadd_mix_ops_per_s"] / 9 / 1e9})
                except Exception:
                    pass
        except Exception as e:  # pragma: no cover
            out["gapped_stress"] = {"error": str(e)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, n, dt = time_oracle(args.workload, 12.0, cores)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n} read pairs of {args.workload} in {dt:.1f} s, CPU oracle port (pthreads x {cores}), matching only"}
    tool.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
