#!/usr/bin/env python
"""Generates tests/golden/classify_golden.json by calling the REFERENCE's own classifier,
/root/reference/scripts/post_process_bam.py::_classify_primer_pair, on random forward/reverse primer matches.
Runs in the build container only (the reference tree does not travel); pysam and defopt, which the script imports
at module level but the classifier does not use, are stubbed."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

for name in ("pysam", "defopt"):
    sys.modules.setdefault(name, types.ModuleType(name))
spec = importlib.util.spec_from_file_location("post_process_bam", "/root/reference/scripts/post_process_bam.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(42)
primers = []
for i in range(12):
    contig = f"chr{1 + i % 3}"
    start = 100 + 90 * i
    primers.append(dict(pair_id=f"p{i}", primer_id=f"p{i}_F", sequence="ACGTACGTACGTACGTAC", positive_strand=True, ref_name=contig, start=start, end=start + 17))
    primers.append(dict(pair_id=f"p{i}", primer_id=f"p{i}_R", sequence="TTGCATTGCATTGCATTG", positive_strand=False, ref_name=contig, start=start + 120, end=start + 137))
# a pair whose two primers sit on the same strand (multi-primer panels) and one on another contig
primers.append(dict(pair_id="q", primer_id="q_F1", sequence="GGGGACGTACGTACGTAC", positive_strand=True, ref_name="chr1", start=5000, end=5017))
primers.append(dict(pair_id="q", primer_id="q_F2", sequence="CCCCACGTACGTACGTAC", positive_strand=True, ref_name="chr2", start=5100, end=5117))


def to_match(i):
    if i < 0:
        return None
    p = primers[i]
    return ref.PrimerMatch(pair_id=p["pair_id"], primer_id=p["primer_id"], ref_name=p["ref_name"], start=p["start"], end=p["end"],
                           positive_strand=p["positive_strand"], read_num=1, match_type="Ungapped", match_type_info=["0", "na"])


cases = []
for _ in range(3000):
    f5 = int(rng.integers(-1, len(primers)))
    r5 = int(rng.integers(-1, len(primers)))
    if rng.random() < 0.3 and f5 >= 0:
        r5 = f5 ^ 1 if f5 < 24 else r5          # the canonical partner
    if rng.random() < 0.1 and f5 >= 0:
        r5 = f5                                   # self dimer
    lo, hi = (50, 250) if rng.random() < 0.7 else (int(rng.integers(0, 200)), int(rng.integers(150, 600)))
    cases.append(dict(f5=f5, r5=r5, min_insert=lo, max_insert=hi, expected=ref._classify_primer_pair(to_match(f5), to_match(r5), lo, hi)))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "classify_golden.json")
with open(out, "w") as fh:
    json.dump(dict(primers=primers, cases=cases, source="scripts/post_process_bam.py:56-78 (_classify_primer_pair), run in the build container"), fh)
print("wrote", out, {k: sum(1 for c in cases if c["expected"] == k) for k in ("Canonical", "SelfDimer", "CrossDimer", "NonCanonical", "Single", "NoMatch")})
