"""Run in a subprocess by tests/test_align_policy.py with IDP_LIB_PATH / IDP_ORACLE_LIB pointing at the builds whose AlignPolicy /
ALIGN_POLICY switches are all flipped: the CUDA path and the oracle must still agree bit for bit (gapped-heavy batches through
idp_match_batch, and idp_align_batch in its three modes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import _lib as L
from fg_idprimer_b200 import synth
from oracle import oracle_py as O

want_bits = int(sys.argv[1])
assert L.lib().idp_align_policy() == want_bits, L.lib().idp_align_policy()
assert O.lib().orc_align_policy_bits() == want_bits, O.lib().orc_align_policy_bits()
FIELDS = ["primer", "type", "status", "multi", "mm", "next_mm", "raw_score", "raw_next", "q_len", "q_len_next", "t_start", "t_end"]

for name, n_pairs, extra in (("c3", 6000, {}), ("c5", 1500, {}), ("c3", 3000, dict(gap_open=0, gap_extend=-2, mismatch_score=-3)),
                             ("c1", 3000, dict(gap_open=-1, gap_extend=-1, match_score=2, mismatch_score=-1, min_alignment_score_rate=0.0))):
    panel, reads, opts = synth.workload(name, n_pairs)
    opts.update(extra)
    ref = {n: i for i, n in enumerate(panel.ref_names)}

    class _P:
        pass
    prs = []
    for p in panel.primers:
        q = _P()
        q.__dict__.update(pair_id=p.pair_id, primer_id=p.primer_id, sequence=p.sequence, positive_strand=p.positive_strand,
                          ref_name=p.ref_name, start=p.start, end=p.end, ref_idx=ref.get(p.ref_name, -1))
        prs.append(q)
    params = O.make_params(k_ungapped=opts.get("ungapped_kmer_length"), k_gapped=opts.get("gapped_kmer_length"), three_prime=opts.get("three_prime", -1),
                           **{k: v for k, v in opts.items() if k in ("gap_open", "gap_extend", "match_score", "mismatch_score", "min_alignment_score_rate")})
    want5, want3, want_metrics, want_align = O.Panel(prs, params).process_batch(reads.oracle_arrays(), n_threads=4)
    with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
        batch = idp.ReadBatch(reads.seq4().reshape(-1), reads.flags, fixed_len=150, ref_idx=reads.ref_idx, unclipped_start=reads.ustart,
                              unclipped_end=reads.uend)
        got5, got3 = tool.process_batch(batch, compact=True)
        for f in FIELDS:
            assert np.array_equal(got5[f], want5[f]), (name, extra, f)
            if got3 is not None:
                assert np.array_equal(got3[f], want3[f]), (name, extra, "3'", f)
        assert np.array_equal(tool.metric_counter, want_metrics) and tool.num_alignments == want_align
        assert (got5["type"] == L.GAPPED).sum() > 20, (name, extra)

rng = np.random.default_rng(9)
queries, targets = [], []
for i in range(400):
    n, m = int(rng.integers(1, 40)), int(rng.integers(0, 80))
    q = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
    t = "".join("ACGT"[x] for x in rng.integers(0, 4, m))
    if i % 2 and m > n:
        pos = int(rng.integers(0, m - n + 1))
        qq = list(q)
        if n > 4 and i % 4 == 1:
            del qq[n // 2]
        if n > 6 and i % 8 == 3:
            qq.insert(n // 3, "ACGT"[i % 4]); qq.insert(n // 3, "ACGT"[(i + 1) % 4])
        t = t[:pos] + "".join(qq) + t[pos + len(qq):]
    queries.append(q.encode()); targets.append(t.encode())
differs = 0
for scores in [dict(match_score=1, mismatch_score=-4, gap_open=-6, gap_extend=-1), dict(match_score=2, mismatch_score=-1, gap_open=-1, gap_extend=-1),
               dict(match_score=1, mismatch_score=-3, gap_open=0, gap_extend=-2), dict(match_score=1, mismatch_score=-1, gap_open=0, gap_extend=-1)]:
    params = O.make_params(**scores)
    for mode in (L.MODE_LOCAL, L.MODE_GLOCAL, L.MODE_GLOBAL):
        al = idp.AsyncCudaAligner(mode=mode, **scores)
        s, ts, te, qs, qe = al.align(queries, targets, query_coordinates=True)
        al.close()
        for i, (q, t) in enumerate(zip(queries, targets)):
            a = O.align(params, mode, q, t)
            assert (int(s[i]), int(ts[i]), int(te[i]), int(qs[i]), int(qe[i])) == (a["score"], a["target_start"], a["target_end"], a["query_start"], a["query_end"]), (mode, scores, q, t, a)
print("POLICY-PARITY-OK", want_bits)
