#!/usr/bin/env python
"""bench.py -- read pairs/sec of the IdentifyPrimers matching path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--pairs P] [--impl reference]

A step is one pass of the matcher chain over one batch of synthetic read pairs.  The headline workload is c2 (10M read
pairs 2x150 bp per GPU against a 96-pair IUPAC-degenerate panel, -k 6 -K 10, full Location -> Ungapped -> Gapped chain).
What the plugin ships by default for a run without --three-prime is the WINDOWED batch (idp_bam_to_batch_window: the
leading bases the 5' chain can reach, 18 B instead of 75 B per read) and 16-byte result records (idp_match_batch16):

  value              device-resident windowed batch -> idp_match_batch16_device, CUDA events on the launching stream; the steps
                     are enqueued back to back in the context's asynchronous mode 2 (no per-stage events, counters collected once);
                     stages_ms / roofline come from a second loop over the same steps with the per-stage events recorded
  e2e                pinned HOST windowed batch -> idp_match_batch16, H2D + D2H inside the timed region (wall clock)
  resident_whole_reads / e2e_whole_reads   the same through the 75 B rows and 32-byte records (idp_match_batch*)
  workloads          c3 (gapped-heavy), c4 (5,000-pair panel, 50M pairs strong-split over the GPUs), c5 (mapped/unmapped,
                     --three-prime 5: whole reads), each timed the same two ways, with its own oracle spot check

The 160 metric counters are accumulated per rank across the timed steps and all-reduced ONCE after the loop (the
reference merges counters per batch on the host and writes them once, IP:425-426, 356-360).  `--impl reference` times the
reference's algorithm on the host cores: the reference is Scala/JVM and no JVM exists here, so that arm is the CPU oracle
port (oracle/), rebuilt -O3 -march=native on the box; it is the only other place the oracle is executed.
"""
import argparse
import ctypes as C
import json
import os
import shutil
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
READ_LEN = 150


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


def oracle_panel(panel, opts, fast=False):
    from oracle import oracle_py as O
    if fast:
        O.use_fast_build()   # -O3 -march=native, compiled on this box (the checker elsewhere keeps the -O2 -g build)
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


def jvm_probe():
    """the reference's own JVM tool is timed when a JVM and an fg-idprimer jar exist on the box."""
    java = shutil.which("java")
    jars = []
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        for d, _, files in os.walk(ref):
            jars += [os.path.join(d, f) for f in files if f.endswith(".jar")]
    return {"java": java, "jars": len(jars)}


def oracle_rate(name, cores, seconds):
    from fg_idprimer_b200 import synth
    panel, reads, opts = synth.workload(name, 20000)
    op = oracle_panel(panel, opts, fast=True)
    t0 = time.perf_counter()
    op.process_batch(reads.oracle_arrays(), n_threads=cores)
    return op, panel, opts, 20000 / max(time.perf_counter() - t0, 1e-6)


def oracle_arrays_chunked(panel, w, n_pairs, seed):
    """numpy synth-v1 reads in 1M-pair chunks -> one oracle-style dict (ASCII bases)."""
    from fg_idprimer_b200 import synth
    bases, flags, loc = [], [], [[], [], []]
    for ci, lo in enumerate(range(0, n_pairs, 1_000_000)):
        n = min(1_000_000, n_pairs - lo)
        rd = synth.make_reads(panel, n, seed + ci, indel_frac=w["indel_frac"], mapped_frac=w["mapped_frac"])
        a = rd.oracle_arrays()
        bases.append(a["bases"]); flags.append(a["flags"])
        for k, key in enumerate(("ref_idx", "ustart", "uend")):
            loc[k].append(a[key])
    n_reads = 2 * n_pairs
    return dict(bases=np.concatenate(bases), off=np.arange(n_reads + 1, dtype=np.int64) * READ_LEN, flags=np.concatenate(flags),
                ref_idx=np.concatenate(loc[0]), ustart=np.concatenate(loc[1]), uend=np.concatenate(loc[2]))


def time_oracle(name, target_seconds, cores):
    """cpu_baseline leg: the CPU oracle (-O3 -march=native build) on a bounded sample of the workload."""
    from fg_idprimer_b200 import synth
    op, panel, opts, rate = oracle_rate(name, cores, target_seconds)
    n = int(min(max(rate * target_seconds, 20000), 3_000_000))
    arrays = oracle_arrays_chunked(panel, synth.WORKLOADS[name], n, synth.WORKLOADS[name]["read_seed"])
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
    op, panel, opts, rate = oracle_rate(args.workload, cores, 0)
    total = args.steps + args.warmup
    # the arm's own config: the same pair count as the GPU arm when (warmup + steps) passes fit in ~150 s, else a bounded sample
    n = args.pairs if args.pairs * total / rate <= 150.0 else int(min(max(rate * 150.0 / total, 20000), args.pairs))
    w = synth.WORKLOADS[args.workload]
    arrays = oracle_arrays_chunked(panel, w, n, w["read_seed"])
    times = []
    for i in range(total):
        t0 = time.perf_counter()
        op.process_batch(arrays, n_threads=cores)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    value = n * args.steps / dt
    same = n == args.pairs
    sample = (f"{n} read pairs of {args.workload} per step ({'the full workload' if same else 'bounded sample of the ' + str(args.pairs) + '-pair workload'}), "
              f"CPU oracle port built -O3 -march=native (pthreads x {cores}), whole reads, matching only (no BAM I/O)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, args.pairs), "sample": sample, "jvm_probe": jvm_probe(),
                   "why_port": "the reference is Scala on the JVM; no JVM / jar exists on this box, so the arm is the C restatement of its algorithm"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
class Batch:
    """One synthetic workload batch: device-resident and pinned-host copies in the layouts the entry points take."""

    def __init__(self, torch, dev, name, n_pairs, seed_offset, want_whole, win, panel=None):
        from fg_idprimer_b200 import synth, synth_gpu
        w = synth.WORKLOADS[name]
        self.name, self.w, self.n_pairs, self.n_reads, self.win = name, w, n_pairs, 2 * n_pairs, win
        self.panel = panel or synth.make_panel(w["pairs"], w["panel_seed"], degenerate=w["degenerate"], mapped=w["mapped"])
        self.opts = dict(w["options"])
        pt = synth_gpu.PanelTensors(self.panel, dev)
        n_reads = self.n_reads
        wb = (win + 1) // 2
        self.d_whole = torch.empty((n_reads, READ_LEN // 2), dtype=torch.uint8, device=dev) if want_whole else None
        self.d_win = torch.empty((n_reads, wb), dtype=torch.uint8, device=dev) if win else None
        self.d_flags = torch.empty(n_reads, dtype=torch.int16, device=dev)
        self.d_loc = [torch.empty(n_reads, dtype=torch.int32, device=dev) for _ in range(3)] if w["mapped"] else None
        t0 = time.perf_counter()
        self.sample = None
        chunk = 1_000_000
        for ci, lo in enumerate(range(0, n_pairs, chunk)):
            n = min(chunk, n_pairs - lo)
            g = synth_gpu.make_reads(pt, n, w["read_seed"] + 7919 * seed_offset + ci, indel_frac=w["indel_frac"], mapped_frac=w["mapped_frac"])
            sl = slice(2 * lo, 2 * (lo + n))
            if want_whole:
                self.d_whole[sl] = synth_gpu.pack(g["codes"])
            if win:
                self.d_win[sl] = synth_gpu.pack(synth_gpu.window_codes(g["codes"], g["flags"], win))
            self.d_flags[sl] = g["flags"]
            if self.d_loc is not None:
                for t, key in zip(self.d_loc, ("ref_idx", "ustart", "uend")):
                    t[sl] = g[key]
            if ci == 0:  # what the oracle spot check sees: the first pairs, whole reads
                k = min(20000, n)
                self.sample = dict(codes=g["codes"][:2 * k].cpu().numpy(), flags=g["flags"][:2 * k].cpu().numpy().view(np.uint16),
                                   loc=[g[key][:2 * k].cpu().numpy() for key in ("ref_idx", "ustart", "uend")] if self.d_loc is not None else None)
            del g
        torch.cuda.synchronize()
        self.gen_seconds = time.perf_counter() - t0

    def oracle_arrays(self):
        from fg_idprimer_b200 import synth
        c = self.sample["codes"]
        n = c.shape[0]
        z = np.zeros(n, dtype=np.int32)
        loc = self.sample["loc"]
        return dict(bases=synth.codes_to_ascii(c).reshape(-1), off=np.arange(n + 1, dtype=np.int64) * READ_LEN, flags=self.sample["flags"],
                    ref_idx=loc[0] if loc else z - 1, ustart=loc[1] if loc else z, uend=loc[2] if loc else z)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--pairs", type=int, default=10_000_000, help="read pairs per GPU per step (headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gapped-stress", action="store_true")
    ap.add_argument("--no-workloads", action="store_true", help="skip the c3 / c4 / c5 lines")
    ap.add_argument("--no-whole-reads", action="store_true", help="skip the 75 B / 32 B comparison legs of the headline workload")
    ap.add_argument("--c4-pairs", type=int, default=50_000_000, help="c4 read pairs in total (strong-split over the GPUs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import fg_idprimer_b200 as idp
    from fg_idprimer_b200 import _lib as L

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
    # a stream of our own: torch's default current stream is the legacy NULL stream, and NULL means "the context's own stream" to
    # idp_ctx_set_stream -- the events below must sit on the stream the kernels are enqueued on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, peak_src = float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    int_peak = None
    try:
        with open(os.path.join(ROOT, "profiles", "int_peak_b200.json")) as fh:
            int_peak = json.load(fh)
    except Exception:
        pass

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, use_events, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        acc = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
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

    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    def host_link_ceiling():
        """Pinned H2D and D2H copies of 256 MB issued together on two streams by every rank at once: what the host's PCIe /
        memory path gives this rank while all ranks are pushing -- the ceiling of any end-to-end number on this box."""
        nbytes = 256 << 20
        h_in, h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True), torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        res = {}
        for label, do_in, do_out in (("h2d_alone", True, False), ("d2h_alone", False, True), ("both", True, True)):
            for rep in range(2):
                barrier()
                t0 = time.perf_counter()
                for _ in range(3):
                    if do_in:
                        with torch.cuda.stream(s1):
                            d_in.copy_(h_in, non_blocking=True)
                    if do_out:
                        with torch.cuda.stream(s2):
                            h_out.copy_(d_out, non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[label] = 3 * nbytes / float(t.item()) / 1e9
        return {"h2d_gbs_per_gpu": res["h2d_alone"], "d2h_gbs_per_gpu": res["d2h_alone"], "duplex_gbs_per_gpu_each_way": res["both"],
                "ranks_copying_at_once": world, "how": "3 x 256 MB pinned copies per direction, all ranks at once, slowest rank"}

    totals = {}  # workload -> int64[161] accumulated over the timed steps of this rank (160 bins + alignment count)

    def bench_workload(name, n_pairs, steps, warmup, whole_legs, seed_offset):
        """Times one workload resident and end to end; returns its JSON object."""
        from fg_idprimer_b200 import synth
        w = synth.WORKLOADS[name]
        o = dict(w["options"])
        do3 = o.get("three_prime", -1) >= 0
        panel = synth.make_panel(w["pairs"], w["panel_seed"], degenerate=w["degenerate"], mapped=w["mapped"])
        tool = idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, device=local, **o)
        win = 0 if do3 else min(tool.window(), READ_LEN)
        b = Batch(torch, dev, name, n_pairs, seed_offset, want_whole=do3 or whole_legs, win=win, panel=panel)
        n_reads = b.n_reads
        tool.set_stream(stream.cuda_stream)
        d_five = torch.empty((n_reads, 16), dtype=torch.uint8, device=dev)
        d_three = torch.empty((n_reads, 16), dtype=torch.uint8, device=dev) if do3 else None
        d_seq = b.d_whole if do3 else b.d_win
        fixed_len = READ_LEN if do3 else win
        loc = b.d_loc
        acc = np.zeros(161, dtype=np.int64)

        # device-resident legs run in the library's asynchronous mode: batches are enqueued back to back on torch's current stream
        # and the counters are collected once after the timed loop (idp_ctx_set_async / idp_ctx_sync)
        # The timed loop runs without the per-stage events and the per-batch counter copy (mode 2: each is a small bubble between two
        # kernels of the stream); the stage times and the roofline's kernel time come from the same steps run once more with them on
        tool.set_async(True, stage_timing=False)

        def step_resident(compact=True, seq=d_seq, flen=fixed_len, five=d_five, three=d_three):
            tool.process_batch_device(n_reads, seq.data_ptr(), b.d_flags.data_ptr(), five.data_ptr(), three.data_ptr() if three is not None else 0,
                                      fixed_len=flen, ref_idx_ptr=loc[0].data_ptr() if loc else 0, ustart_ptr=loc[1].data_ptr() if loc else 0,
                                      uend_ptr=loc[2].data_ptr() if loc else 0, compact=compact)

        def timed_resident(fn):
            for _ in range(warmup):
                fn()
            tool.sync()
            ms, _ = timed(fn, True, steps, 0)
            m, na = tool.sync()   # counters of exactly `steps` steps
            return ms, m, na, tool.timings()

        ms_res, m_res, na_res, _ = timed_resident(step_resident)
        acc[:160] = m_res; acc[160] = na_res
        tool.set_async(True, stage_timing=True)
        ms_staged, _, _, tm_last = timed_resident(step_resident)
        tm_res = [tm_last]
        totals[name] = acc.copy()
        # parity spot check of what was just timed (the checker, not the thing measured): first pairs against the oracle
        torch.cuda.synchronize()
        parity = None
        if rank == 0:
            try:
                op = oracle_panel(b.panel, o)
                arrays = b.oracle_arrays()
                k = len(arrays["flags"])
                want5, want3, _, _ = op.process_batch(arrays, n_threads=os.cpu_count() or 1)
                fields = ("primer", "type", "status", "multi", "mm", "next_mm", "raw_score", "raw_next", "q_len", "q_len_next", "t_start", "t_end")
                got = L.unpack16(d_five[:k].cpu().numpy().view(L.RESULT16_DTYPE).reshape(-1))
                parity = all(np.array_equal(got[f], want5[f]) for f in fields)
                if do3:
                    got3 = L.unpack16(d_three[:k].cpu().numpy().view(L.RESULT16_DTYPE).reshape(-1))
                    parity = parity and all(np.array_equal(got3[f], want3[f]) for f in fields)
            except Exception as e:  # pragma: no cover
                parity = f"check failed to run: {e}"

        tool.set_async(False)
        # ---- end to end: pinned host buffers through the C-ABI, copies inside the timed region ----
        h_seq, h_flags = pinned(d_seq), pinned(b.d_flags)
        h_loc = [pinned(t) for t in loc] if loc else None
        h_five = torch.empty((n_reads, 16), dtype=torch.uint8, pin_memory=True)
        h_three = torch.empty((n_reads, 16), dtype=torch.uint8, pin_memory=True) if do3 else None

        def make_e2e(fn, seq, flen, five, three):
            rs = L.IdpReads()
            rs.seq4, rs.flags, rs.fixed_len = seq.data_ptr(), h_flags.data_ptr(), flen
            rs.seq_off = rs.len = None
            rs.ref_idx, rs.unclipped_start, rs.unclipped_end = ([t.data_ptr() for t in h_loc] if h_loc else [None] * 3)
            metrics, n_align = np.zeros(160, dtype=np.int64), np.zeros(1, dtype=np.int64)

            def step():
                L.check(fn(tool._ctx, C.byref(rs), n_reads, five.data_ptr(), three.data_ptr() if three is not None else None,
                           metrics.ctypes.data, n_align.ctypes.data))
                return tool.timings()
            return step

        e2e_steps = max(2, min(steps, 5))
        ms_e2e, tm_e2e = timed(make_e2e(L.lib().idp_match_batch16, h_seq, fixed_len, h_five, h_three), False, e2e_steps, 2)
        same = bool(torch.equal(h_five.to(dev), d_five)) and (not do3 or bool(torch.equal(h_three.to(dev), d_three)))

        scan_ms = statistics.mean(t["ms_scan"] for t in tm_res)
        stages = {k: statistics.mean(t[k] for t in tm_res) for k in ("ms_total", "ms_scan", "ms_gapped", "ms_three_prime", "ms_sw_score")}
        # per step: the alignment / cell totals accumulate on the device over the asynchronous steps, the list lengths are the last batch's
        work = {k: int(tm_res[-1][k]) // steps for k in ("n_alignments_5p", "n_alignments_3p", "sw_cells")}
        work.update({k: int(tm_res[-1][k]) for k in ("n_fallthrough", "n_scan_parked", "n_scan_deferred")})
        shipped = 2 * ((fixed_len + 1) // 2 + 2 + 16 * (2 if do3 else 1) + 1 + (12 if loc else 0))
        out = {
            "workload": workload_desc(name, n_pairs) + " (per GPU)",
            "layout": (f"whole reads, {READ_LEN // 2} B rows" if do3 else f"windowed rows: {win} bases = {(win + 1) // 2} B per read") + ", 16-byte records",
            "value": world * n_pairs * steps / (ms_res / 1e3), "unit": UNIT, "ms_per_step": ms_res / steps, "steps": steps,
            "e2e": {"value": world * n_pairs * e2e_steps / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(tm_e2e[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(tm_e2e[-1]["d2h_bytes"]),
                    "same_bytes_as_resident_run": same},
            "stages_ms": stages, "work": work, "gpu_launches_per_step": int(tm_res[-1]["kernel_launches"]),
            "resident_mode": "asynchronous: the steps are enqueued back to back (idp_ctx_set_async), counters collected once by idp_ctx_sync after the timed loop; "
                             "stages_ms and the roofline's kernel time come from a second loop of the same steps with the per-stage CUDA events recorded",
            "ms_per_step_with_stage_events": ms_staged / steps,
            "parity_spot_check_vs_oracle": parity, "generator_seconds": round(b.gen_seconds, 1),
            "roofline": {"bound": "hbm", "kernel": "scan stage = scan_fx_kernel (exact-prefix table + mismatchAlign + k-mer rule) + compact_mask_kernel + scan_kernel over the parked reads (location, bit-sliced filter / seed index / exact walk) + compact_mask_kernel",
                         "achieved": ALGO_BYTES_PER_PAIR * n_pairs / (scan_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": ALGO_BYTES_PER_PAIR * n_pairs / (scan_ms / 1e3) / 1e9 / peak, "kernel_ms": scan_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PAIR * n_pairs,
                         "bytes_per_pair_this_layout": shipped, "achieved_this_layout": shipped * n_pairs / (scan_ms / 1e3) / 1e9,
                         "note": "achieved = SURVEY 8(d) algorithmic bytes (182 B per pair: whole 75 B reads in, 16 B records out) / scan kernel time; "
                                 "the kernel of a windowed batch touches only bytes_per_pair_this_layout"},
        }
        sw_ms = stages["ms_sw_score"]
        if sw_ms > 0 and work["sw_cells"] > 0:
            g = {"gcups": work["sw_cells"] / (sw_ms / 1e3) / 1e9, "cells": work["sw_cells"], "sw_kernel_ms": sw_ms,
                 "kernels": "sw_pairs_kernel (two pairs per thread in 16-bit half-words, traceback-equivalent)" +
                            (" + sw_all_kernel (packed 16x2, Local)" if do3 else "")}
            if int_peak:
                g["roofline_gcups_9_instr_per_cell"] = int_peak["lop_iadd_mix_ops_per_s"] / 9 / 1e9   # INT32 ops/s / 9
                g["frac"] = g["gcups"] / g["roofline_gcups_9_instr_per_cell"]
                g["note"] = ("the 9-instructions-per-cell figure is SURVEY 8(d)'s 32-bit estimate; the kernel issues 4 full-rate + 5 half-rate "
                             "instructions per TWO cells.  On c2 the stage is a few hundred thousand short alignments: launch and tail bound")
            out["gapped"] = g

        if whole_legs and not do3:
            # the round-1 layout for comparison: 75 B rows in, 32-byte records out
            d5w = torch.empty((n_reads, 32), dtype=torch.uint8, device=dev)
            tool.set_async(True)
            ms_w, _, _, tw_last = timed_resident(lambda: step_resident(False, b.d_whole, READ_LEN, d5w, None))
            tool.set_async(False)
            tm_w = [tw_last]
            ok = bool(np.array_equal(L.unpack16(d_five[:200000].cpu().numpy().view(L.RESULT16_DTYPE).reshape(-1)).view(np.uint8),
                                     d5w[:200000].cpu().numpy().reshape(-1)))
            sw = statistics.mean(t["ms_scan"] for t in tm_w)
            out["resident_whole_reads"] = {"value": world * n_pairs * steps / (ms_w / 1e3), "unit": UNIT, "ms_per_step": ms_w / steps, "scan_kernel_ms": sw,
                                           "scan_hbm_frac": ALGO_BYTES_PER_PAIR * n_pairs / (sw / 1e3) / 1e9 / peak, "layout": "75 B rows in, 32-byte records out",
                                           "same_results_as_windowed": ok}
            h_whole = pinned(b.d_whole)
            h5w = torch.empty((n_reads, 32), dtype=torch.uint8, pin_memory=True)
            ms_we, tm_we = timed(make_e2e(L.lib().idp_match_batch, h_whole, READ_LEN, h5w, None), False, e2e_steps, 2)
            out["e2e_whole_reads"] = {"value": world * n_pairs * e2e_steps / (ms_we / 1e3), "unit": UNIT, "ms_per_step": ms_we / e2e_steps,
                                      "h2d_bytes_per_step": int(tm_we[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(tm_we[-1]["d2h_bytes"])}
            del h_whole, h5w, d5w
        tool.close()
        return out, tm_res

    clk = ClockSampler(local)
    clk.__enter__()
    head, tm_head = bench_workload(args.workload, args.pairs, args.steps, args.warmup, not args.no_whole_reads, rank)
    workloads = {}
    if not args.no_workloads:
        sub_steps = max(2, min(args.steps, 5))
        plan = [("c3", 4_000_000, "weak"), ("c4", max(args.c4_pairs // world, 1), "strong"), ("c5", 2_000_000, "weak")]
        for name, n, scaling in plan:
            if name == args.workload:
                continue
            try:
                workloads[name], _ = bench_workload(name, n, sub_steps, 3, False, rank)
                workloads[name]["scaling"] = scaling
                if name == "c4":
                    workloads[name]["workload"] = workload_desc(name, args.c4_pairs) + f" in total, strong-split: {n} pairs on each of {world} GPU(s)"
            except Exception as e:  # pragma: no cover
                workloads[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
    clk.__exit__(None, None, None)
    clocks = clk.summary()  # sampled across all timed regions
    link = host_link_ceiling()

    # ---- the one collective of the path: metric counters summed over the ranks, once, after the timed loops ----
    names = sorted(totals)
    red = torch.from_numpy(np.stack([totals[k] for k in names])).to(dev)
    if world > 1:
        dist.all_reduce(red)
    red = red.cpu().numpy()
    counters = {}
    for k, row in zip(names, red):
        s17 = np.zeros(17, dtype=np.int64)
        L.lib().idp_summary_metrics(np.ascontiguousarray(row[:160]).ctypes.data, s17.ctypes.data)
        steps_k = head["steps"] if k == args.workload else workloads[k]["steps"]
        pairs_k = args.pairs if k == args.workload else {"c3": 4_000_000, "c4": max(args.c4_pairs // world, 1), "c5": 2_000_000}[k]
        counters[k] = {"templates": int(s17[0]), "expected_templates": int(world * steps_k * pairs_k), "gapped_alignments": int(row[160]),
                       "location": int(s17[13]), "ungapped": int(s17[14]), "gapped": int(s17[15]), "no_match": int(s17[16])}

    out = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": head["workload"], "layout": head["layout"],
                   "parallelism": f"reads sharded over {world} GPU(s), panel replicated, one all-reduce of the 160 counters after the timed loop",
                   "l2": f"inputs + outputs {head['roofline']['bytes_per_pair_this_layout'] * args.pairs / 1e9:.2f} GB per step exceed the 126 MB L2, no flush needed",
                   "generator": "synth-v1 distributions drawn on the GPU (fg_idprimer_b200/synth_gpu.py)", "generator_seconds": head["generator_seconds"]},
        "e2e": head["e2e"],
        "gpu_launches": int(tm_head[-1]["kernel_launches"]) * args.steps,
        "roofline": dict(head["roofline"], traffic=None),
        "clocks": clocks,
        "stages_ms": head["stages_ms"], "work": head["work"],
        "parity_spot_check_vs_oracle": head["parity_spot_check_vs_oracle"],
        "host_link_ceiling": link,
        "counters_reduced_once": counters,
    }
    try:  # DRAM bytes of the scan kernel from this round's ncu capture of the same workload, layout and size
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            tj = json.load(fh)
        if tj.get("workload") == args.workload and tj.get("pairs") == args.pairs and tj.get("layout") == head["layout"]:
            out["roofline"]["traffic"] = tj["scan_kernel_dram_bytes_per_launch"]
            out["roofline"]["traffic_source"] = tj.get("source")
    except Exception:
        pass
    for k in ("gapped", "resident_whole_reads", "e2e_whole_reads", "resident_mode", "ms_per_step_with_stage_events"):
        if k in head:
            out[k] = head[k]
    if workloads:
        out["workloads"] = workloads

    if rank == 0 and world == 1 and not args.no_gapped_stress:
        # gapped-alignment stress: no -K filter, every fall-through read aligns against all 2000 primers
        try:
            from fg_idprimer_b200 import synth, synth_gpu
            n3 = 1_000_000
            b3 = Batch(torch, dev, "c3b", n3, 0, want_whole=False, win=35)
            with idp.IdentifyPrimers(b3.panel.primers, device=local, **b3.opts) as t3:
                assert t3.window() == 35
                t3.set_stream(stream.cuda_stream)
                f5 = torch.empty((b3.n_reads, 16), dtype=torch.uint8, device=dev)
                tms = []
                for i in range(4):
                    t3.process_batch_device(b3.n_reads, b3.d_win.data_ptr(), b3.d_flags.data_ptr(), f5.data_ptr(), fixed_len=35, compact=True)
                    if i:
                        tms.append(t3.timings())
                cells = tms[-1]["sw_cells"]; ms = statistics.mean(t["ms_sw_score"] for t in tms)
                out["gapped_stress"] = {"workload": f"c3b: {n3} read pairs vs 1000-pair panel, 5% indel reads, -k 6, no -K", "gcups": cells / (ms / 1e3) / 1e9,
                                        "cells": int(cells), "alignments": int(tms[-1]["n_alignments_5p"]), "sw_kernel_ms": ms,
                                        "kernel": "sw_all_kernel (two alignments per register, S16x2 DPX)"}
                try:
                    # integer issue roofline of the packed aligner (two alignments per register): per PAIR of cells
                    # 1 full-rate instruction (LOP3) and 5 half-rate ones (PRMT, VIADD.16x2, 2 x VIADDMNMX.S16x2, VIMNMX3.S16x2),
                    # at the issue rates measured by tools/int_peak.cu and tools/int_peak16.cu.  (Before the mismatch masks came from
                    # one PRMT the mix was 3 full + 5 half: that older, lower roofline is kept beside it.)
                    with open(os.path.join(ROOT, "profiles", "int_peak16_b200.json")) as fh:
                        ip16 = json.load(fh)
                    ip = int_peak
                    full = ip["lop_iadd_per_clk_per_sm"]
                    r16 = ip16["per_clk_per_sm"]
                    half = min(r16["viaddmax_s16x2"], r16["vimax3_s16x2"], r16["vadd2"], r16["prmt"])
                    scale = ip["sms"] * ip["sm_clock_mhz_max"] * 1e6 / 1e9
                    pk = 2.0 / (1.0 / full + 5.0 / half) * scale
                    pk_old = 2.0 / (3.0 / full + 5.0 / half) * scale
                    out["gapped_stress"].update({"roofline_gcups": pk, "frac": out["gapped_stress"]["gcups"] / pk,
                                                 "peak_source": "integer issue roofline: 1 full-rate + 5 half-rate instructions per two cells at the measured "
                                                                "rates (profiles/int_peak_b200.json, profiles/int_peak16_b200.json)",
                                                 "roofline_gcups_round1_mix_3_full_5_half": pk_old, "frac_of_round1_mix": out["gapped_stress"]["gcups"] / pk_old,
                                                 "roofline_gcups_32bit_9_instr_per_cell": ip["lop_iadd_mix_ops_per_s"] / 9 / 1e9})
                except Exception:
                    pass
        except Exception as e:  # pragma: no cover
            out["gapped_stress"] = {"error": str(e)}

    if rank == 0 and world == 1 and not args.no_gapped_stress:
        # the AsyncAligner seam (AA:37-54) as a batch: idp_align_batch with HOST ASCII arrays, copies inside the timed region
        try:
            from fg_idprimer_b200 import synth
            p3 = synth.make_panel(1000, 1003)
            rd = synth.make_reads(p3, 500_000, 77, indel_frac=0.05)
            n_al = rd.n
            rng = np.random.default_rng(5)
            pick = np.where(rd.truth >= 0, 2 * np.maximum(rd.truth, 0) + (np.arange(n_al) & 1), rng.integers(0, 2000, n_al))
            qs = [p3.primers[i].sequence.encode() for i in pick]
            q_off = np.zeros(n_al + 1, dtype=np.int64); q_off[1:] = np.cumsum([len(q) for q in qs])
            q_all = b"".join(qs)
            out["align_batch"] = {}
            for mode_name, mode, tlen in (("glocal", L.MODE_GLOCAL, 35), ("local", L.MODE_LOCAL, READ_LEN)):
                t_all = synth.codes_to_ascii(rd.codes[:, :tlen]).tobytes()
                t_off = np.arange(n_al + 1, dtype=np.int64) * tlen
                sc, ts, te = (np.zeros(n_al, dtype=np.int32) for _ in range(3))
                al = idp.AsyncCudaAligner(mode=mode, device=local)
                times = []
                for it in range(4):
                    t0 = time.perf_counter()
                    L.check(L.lib().idp_align_batch(al._tool._ctx, mode, q_all, q_off.ctypes.data, t_all, t_off.ctypes.data, n_al,
                                                    sc.ctypes.data, ts.ctypes.data, te.ctypes.data, None, None))
                    if it:
                        times.append(time.perf_counter() - t0)
                al.close()
                dt = statistics.mean(times)
                cells = float(np.sum(np.diff(q_off)) * tlen)
                out["align_batch"][mode_name] = {"alignments_per_s": n_al / dt, "ms_per_call": 1e3 * dt, "alignments_per_call": n_al,
                                                 "target_bases": tlen, "gcups_end_to_end": cells / dt / 1e9, "mean_score": float(sc.mean()),
                                                 "what": "idp_align_batch, pageable host ASCII in, (score, targetStart, targetEnd) out, wall clock"}
        except Exception as e:  # pragma: no cover
            out["align_batch"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, n, dt = time_oracle(args.workload, 12.0, cores)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n} read pairs of {args.workload} processed in {dt:.1f} s (a bounded sample of at most 3M pairs, repeated), CPU oracle port built "
                                         f"-O3 -march=native (pthreads x {cores}), whole reads, matching only"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
