#!/usr/bin/env python
"""Rewrites the round-2 table of README.md (between the r02-table markers) from profiles/r02_bench_1gpu.json,
profiles/r02_bench_reference.json and profiles/r02a_bench_2gpu.json / r02_bench_8gpu.json when present:
    python tools/fill_readme.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            txt = fh.read().strip()
        return json.loads(txt.splitlines()[-1]) if not txt.startswith("{\n") else json.loads(txt)
    except Exception:
        return None


d = load("r02_bench_1gpu.json")
ref = load("r02_bench_reference.json")
G = 1e9
st = d["stages_ms"]
rf = d["roofline"]
w = d["workloads"]
gs = d["gapped_stress"]
ab = d["align_batch"]
rows = []
rows.append(("device-resident matcher chain, windowed batch + 16-byte records (the default the plugin ships)",
             f"{d['value'] / G:.1f} G read pairs/s ({d['ms_per_step']:.3f} ms per 10 M pairs: scan stage {st['ms_scan']:.3f}, gapped stage {st['ms_gapped']:.3f}, "
             f"metrics + counters {st['ms_total'] - st['ms_scan'] - st['ms_gapped'] - st['ms_three_prime']:.3f}; round 1: 8.9 G, 1.13 ms)"))
e = d["e2e"]
hl = d["host_link_ceiling"]
rows.append(("end to end through `idp_match_batch16` (pinned host buffers, H2D + D2H inside)",
             f"{e['value'] / G:.2f} G read pairs/s ({e['ms_per_step']:.1f} ms per step for {e['h2d_bytes_per_step'] / 1e9:.2f} GB in + {e['d2h_bytes_per_step'] / 1e9:.2f} GB out; "
             f"this box copies {hl['h2d_gbs_per_gpu']:.0f} / {hl['d2h_gbs_per_gpu']:.0f} GB/s one way and {hl['duplex_gbs_per_gpu_each_way']:.0f} GB/s each way at once: PCIe-bound; round 1: 0.33 G)"))
wr, we = d.get("resident_whole_reads"), d.get("e2e_whole_reads")
if wr and we:
    rows.append(("the same through whole 75 B reads and 32-byte records (round-1 layout)",
                 f"{wr['value'] / G:.1f} G resident, {we['value'] / G:.2f} G end to end (same results, checked in the run)"))
rows.append(("scan stage (`scan_fx_kernel` + compaction + `scan_kernel` over the parked reads + deferred k-mer walks + compaction)",
             f"{rf['kernel_ms']:.3f} ms = {rf['frac']:.2f} of the measured HBM copy peak ({rf['peak']:.0f} GB/s) on the 182 B/pair algorithmic figure "
             f"(round 1: 0.40); DRAM traffic {rf['traffic'] / 1e9:.2f} GB per launch pair (ncu), {rf['bytes_per_pair_this_layout']} B per pair touched in this layout"))
def rate(x):
    return f"{x / G:.2f} G" if x >= 0.1 * G else f"{x / 1e6:.0f} M"
def wl(k):
    v = w[k]
    return f"{k} {rate(v['value'])} pairs/s resident ({v['ms_per_step']:.2f} ms per step) / {rate(v['e2e']['value'])} end to end"
rows.append(("other workloads: c3 (4 M pairs, 1,000-pair panel, 5 % indels), c4 (50 M pairs, 5,000-pair panel), c5 (2 M pairs, mapped + `--three-prime 5`, whole reads)",
             "; ".join(wl(k) for k in ("c3", "c4", "c5")) +
             f" — c4: {w['c4']['stages_ms']['ms_sw_score']:.1f} ms of it Smith-Waterman ({w['c4']['work']['n_alignments_5p'] / 1e6:.0f} M alignments); "
             f"c5: {w['c5']['ms_per_step'] / 2:.1f} ms per 1 M pairs (round 1: 22), {w['c5']['stages_ms']['ms_sw_score']:.1f} ms of the step the 3′ Local alignments"))
rows.append(("gapped-alignment stress (c3b, no `-K`, 1 M pairs)",
             f"{gs['gcups']:.0f} GCUPS (round 1: 2,697) = {gs['frac']:.2f} of the integer issue roofline of the packed cell as it is now "
             f"({gs['roofline_gcups']:.0f} GCUPS: 1 full-rate + 5 half-rate instructions per two cells); {gs['frac_of_round1_mix']:.2f} of round 1's roofline (3 + 5)"))
rows.append(("`idp_align_batch` (the `AsyncAligner` seam, host ASCII in)",
             f"{ab['glocal']['alignments_per_s'] / 1e6:.0f} M Glocal alignments/s (35-base targets), {ab['local']['alignments_per_s'] / 1e6:.0f} M Local (150-base targets)"))
if ref:
    rows.append(("CPU oracle port, 16 host cores, `-O3 -march=native` (`--impl reference`, full 10 M pairs per step)",
                 f"{ref['value'] / G:.4f} G read pairs/s ({ref['ms_per_step'] / 1e3:.2f} s per step)"))
for name, n in (("r02_bench_2gpu.json", 2), ("r02_bench_4gpu.json", 4), ("r02_bench_8gpu.json", 8)):
    m = load(name)
    if m and not any(r[0].startswith(f"{n} GPUs") for r in rows):
        h = m.get("host_link_ceiling", {})
        rows.append((f"{n} GPUs (weak scaling, 10 M pairs per GPU; `profiles/{name}`)",
                     f"{m['value'] / G:.1f} G read pairs/s resident ({m['ms_per_step']:.3f} ms per step on the slowest rank" +
                     (f", measured with the per-stage events still in the loop: the single-GPU step takes {d['ms_per_step_with_stage_events']:.3f} ms that way, "
                      f"so {d['ms_per_step_with_stage_events'] / m['ms_per_step']:.2f} of linear" if "ms_per_step_with_stage_events" not in m and "ms_per_step_with_stage_events" in d
                      else f" = {m['value'] / n / d['value']:.2f} of the single-GPU line above") +
                     "; another box than the single-GPU line, whose runs of this build ranged 0.66-0.70 ms; the counters are reduced once after the run), "
                     f"{m['e2e']['value'] / G:.2f} G end to end (host link with all ranks copying at once: {h.get('duplex_gbs_per_gpu_each_way', 0):.0f} GB/s per GPU each way)"))
table = "| | value |\n|---|---|\n" + "\n".join(f"| {a} | {b} |" for a, b in rows) + "\n"
p = os.path.join(ROOT, "README.md")
s = open(p).read()
b, e_ = "<!-- r02-table-begin -->\n", "<!-- r02-table-end -->\n"
i, j = s.index(b) + len(b), s.index(e_)
open(p, "w").write(s[:i] + table + s[j:])
print(table)
