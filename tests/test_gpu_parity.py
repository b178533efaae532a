"""GPU parity: the CUDA path (through the C-ABI, host buffers) against the CPU oracle on the same inputs, bit-exact on
every result field, the 160-bin metric counter and the alignment count; plus the reference's own known-answer numbers."""
import os

import numpy as np
import pytest

import helpers as H
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import _lib as L
from fg_idprimer_b200 import synth
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
K = H.kats()
FIELDS = ["primer", "type", "status", "multi", "mm", "next_mm", "raw_score", "raw_next", "q_len", "q_len_next", "t_start", "t_end"]


def _opts_to_oracle(opts):
    return O.make_params(slop=opts.get("slop", 5), three_prime=opts.get("three_prime", -1),
                         max_mismatch_rate=opts.get("max_mismatch_rate", 0.05),
                         min_alignment_score_rate=opts.get("min_alignment_score_rate", 0.01), match_score=opts.get("match_score", 1),
                         mismatch_score=opts.get("mismatch_score", -4), gap_open=opts.get("gap_open", -6), gap_extend=opts.get("gap_extend", -1),
                         k_ungapped=opts.get("ungapped_kmer_length"), k_gapped=opts.get("gapped_kmer_length"),
                         score_matrix=opts.get("scoring_matrix"))


def assert_same(got, want, what=""):
    for f in FIELDS:
        if not np.array_equal(got[f], want[f]):
            bad = np.nonzero(got[f] != want[f])[0]
            i = int(bad[0])
            raise AssertionError(f"{what}: field {f} differs at {len(bad)} reads, first read {i}: gpu={got[i]} oracle={want[i]}")


def run_both(primers, arrays, ref_names=H.DEFAULT_DICT, oracle_threads=8, **opts):
    """primers: helpers.Primer list; arrays: oracle-style dict (ASCII bases + offsets)."""
    gp = [idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end) for p in primers]
    name_to_idx = {n: i for i, n in enumerate(ref_names)}
    for p in primers:
        if p.ref_name and p.ref_idx in (-1, None):
            p.ref_idx = name_to_idx.get(p.ref_name, -2)
    panel = O.Panel(primers, _opts_to_oracle(opts))
    want5, want3, want_metrics, want_align = panel.process_batch(arrays, n_threads=oracle_threads)
    with idp.IdentifyPrimers(gp, ref_names=ref_names, **opts) as tool:
        batch = idp.ReadBatch.from_ascii(arrays["bases"], arrays["off"], arrays["flags"], arrays["ref_idx"], arrays["ustart"], arrays["uend"])
        got5, got3 = tool.process_batch(batch)
        assert_same(got5, want5, "5'")
        if opts.get("three_prime", -1) >= 0:
            assert_same(got3, want3, "3'")
        assert np.array_equal(tool.metric_counter, want_metrics)
        assert tool.num_alignments == want_align
        t = tool.timings()
        assert t["kernel_launches"] > 0
        summary = tool.summary()
        # the same batch through the 16-byte records (idp_match_batch16), with the scan kernel's exact-prefix shortcut switched
        # off: the rows unpack to the very same bytes, and the counters double
        os.environ["IDP_SCAN_NO_FX"] = "1"
        try:
            c5, c3 = tool.process_batch(batch, compact=True)
        finally:
            del os.environ["IDP_SCAN_NO_FX"]
        assert c5.tobytes() == got5.tobytes(), "16-byte records / filter path differ (5')"
        if got3 is not None:
            assert c3.tobytes() == got3.tobytes(), "16-byte records / filter path differ (3')"
        assert np.array_equal(tool.metric_counter, 2 * want_metrics) and tool.num_alignments == 2 * want_align
        if opts.get("three_prime", -1) < 0:
            # and cut to the bases the 5' chain can reach (what idp_bam_to_batch_window ships): same rows again
            cut = idp.ReadBatch.from_ascii(arrays["bases"], arrays["off"], arrays["flags"], arrays["ref_idx"], arrays["ustart"], arrays["uend"],
                                           window=tool.window())
            w5, _ = tool.process_batch(cut, compact=True)
            assert w5.tobytes() == got5.tobytes(), "windowed batch differs"
            assert np.array_equal(tool.metric_counter, 3 * want_metrics) and tool.num_alignments == 3 * want_align
        return got5, got3, summary, t


# ------------------------------------------------------------------------------------------------------------------
# the reference's own end-to-end known answers, through the GPU
# ------------------------------------------------------------------------------------------------------------------
def _ipt_opts():
    p = K["ipt_params"]
    return dict(slop=p["slop"], max_mismatch_rate=p["max_mismatches"] / p["primer_length"], min_alignment_score_rate=0.0,
                match_score=p["match_score"], mismatch_score=p["mismatch_score"], gap_open=p["gap_open"], gap_extend=p["gap_extend"])


@pytest.mark.parametrize("primers_mapped", [True, False])
@pytest.mark.parametrize("reads_mapped", [True, False])
def test_ipt_metrics_on_gpu(primers_mapped, reads_mapped):  # IdentifyPrimersTest.scala:299-403
    primers = H.primers_from(K["ipt_panel"], mapped=primers_mapped)
    recs = H.ipt_metrics_reads(H.primers_from(K["ipt_panel"], mapped=True))
    if not reads_mapped:
        recs = H.make_unmapped(recs)
    _, _, summary, _ = run_both(primers, H.recs_to_arrays(recs), **_ipt_opts())
    exp = dict(K["ipt_metrics"]["common"])
    for k, v in list(exp.items()):
        if isinstance(v, dict):
            exp[k] = v["mapped" if reads_mapped else "unmapped"]
    exp.update(K["ipt_metrics"]["mapped_primers_mapped_reads" if (primers_mapped and reads_mapped) else "otherwise"])
    assert summary == exp


def test_ipt_two_pairs_per_read_on_gpu():  # IdentifyPrimersTest.scala:238-297
    primers = H.primers_from(K["ipt_panel_two_pairs"])
    recs = H.pair_records(H.build_builder(primers))
    gp = [idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end) for p in primers]
    arrays = H.recs_to_arrays(recs)
    run_both(primers, arrays, **_ipt_opts())
    with idp.IdentifyPrimers(gp, ref_names=H.DEFAULT_DICT, **_ipt_opts()) as tool:
        five, _ = tool.process_batch(idp.ReadBatch.from_ascii(arrays["bases"], arrays["off"], arrays["flags"], arrays["ref_idx"], arrays["ustart"], arrays["uend"]))
        exp = K["ipt_per_read"]
        for i in range(0, len(recs), 2):
            tags = tool.pair_tags(recs[i].flags, recs[i + 1].flags, five[i], five[i + 1])
            for rec in (recs[i], recs[i + 1]):
                info = tags["f5"] if not rec.negative else tags["r5"]
                got = info.split(",")[5] if "," in info else info
                et = rec.tags.get("et")
                if et is None:
                    want = "none"
                elif et == "no_edits":
                    want = exp["no_edits"]
                else:
                    kind, n = et.split("_"); n = int(n)
                    want = {"mis": "Ungapped" if n < exp["mis_lt"] else "none", "ins": "Gapped" if n < exp["ins_lt"] else "none",
                            "del": "Gapped" if n <= exp["del_le"] else "none"}[kind]
                assert got == want, (et, rec.bases, info)


@pytest.mark.parametrize("three_prime", [-1, 5])
def test_bam_in_bam_out(three_prime):
    """BAM record bytes -> idp_bam_to_batch -> GPU -> idp_bam_annotate: the records written back carry the strings
    addTagsToPair / addTagsToFragment produce (IdentifyPrimers.scala:522-569), and the matches behind them are the oracle's."""
    from test_bam_ingest import recs_to_bam, _records
    primers = H.primers_from(K["ipt_panel"], mapped=True)
    recs = H.ipt_metrics_reads(primers)
    bam = recs_to_bam(recs)
    opts = dict(_ipt_opts(), three_prime=three_prime)
    gp = [idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end) for p in primers]
    with idp.IdentifyPrimers(gp, ref_names=H.DEFAULT_DICT, **opts) as tool:
        batch, _ = idp.ReadBatch.from_bam(bam)
        five, three = tool.process_batch(batch)
        out = _records(tool.annotate_bam(bam, five, three if three_prime >= 0 else None))
        names = ["f5", "r5"] + (["f3", "r3"] if three_prime >= 0 else [])
        i = k = 0
        while i < len(batch):
            if batch.flags[i] & H.F_MATE_NEXT:
                tags = tool.pair_tags(int(batch.flags[i]), int(batch.flags[i + 1]), five[i], five[i + 1],
                                      three[i] if three_prime >= 0 else None, three[i + 1] if three_prime >= 0 else None)
                want, n = [(t, tags[t]) for t in names], 2
            else:
                tags = tool.fragment_tags(int(batch.flags[i]), five[i], three[i] if three_prime >= 0 else None)
                want, n = [("f5", tags["f5"]), ("r5", "none")], 1
            for rec in out[k:k + n]:
                assert [(t, v) for t, ty, v in rec["aux"] if t in ("f5", "r5", "f3", "r3")] == want
            i += n
            k += n
        assert k == len(out) == len(recs)
        assert sum(1 for r in out for t, ty, v in r["aux"] if t == "f5" and v != "none") > 40
    run_both(primers, H.recs_to_arrays(recs), **opts)


def test_pmt_panel_location_and_ungapped_on_gpu():  # PrimerMatcherTest.scala:236-374
    primers = H.primers_from(K["pmt_panel"])
    recs = []
    for adjust in (0, 5, 6):
        for p in primers:
            start = p.start if p.positive_strand else p.end - 20 + 1
            genomic = p.sequence if p.positive_strand else H.revcomp(p.sequence)
            bases = genomic + "N" * 10 if p.positive_strand else "N" * 10 + genomic
            recs.append(H.Rec(bases, False, not p.positive_strand, p.ref_idx, start + adjust))
            recs.append(H.Rec(H.Rec(bases, False, not p.positive_strand).seq_order, True, False))
    for rate in (0.0, 0.1, 1.0):
        for k in (None, 5):
            run_both(primers, H.recs_to_arrays(recs), slop=5, max_mismatch_rate=rate, ungapped_kmer_length=k, gapped_kmer_length=k)


# ------------------------------------------------------------------------------------------------------------------
# synthetic workloads of BASELINE.md at oracle-sized N
# ------------------------------------------------------------------------------------------------------------------
def _synth_primers(panel):
    ref = {n: i for i, n in enumerate(panel.ref_names)}
    return [H.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end, ref.get(p.ref_name, -1)) for p in panel.primers]


@pytest.mark.parametrize("name,n_pairs", [("c1", 20000), ("c2", 20000), ("c3", 6000), ("c3b", 400), ("c4", 1500), ("c5", 3000)])
def test_synthetic_workloads(name, n_pairs):
    panel, reads, opts = synth.workload(name, n_pairs)
    got5, got3, summary, t = run_both(_synth_primers(panel), reads.oracle_arrays(), ref_names=panel.ref_names, **opts)
    assert summary["templates"] == n_pairs
    assert (got5["type"] != 0).mean() > 0.9
    if name == "c5":
        assert summary["location"] > 0 and (got3["type"] == L.GAPPED).any()
    if name == "c3b":
        assert t["n_alignments_5p"] == t["n_fallthrough"] * len(panel.primers)


def test_no_kmer_filters_and_unusual_options():
    panel, reads, _ = synth.workload("c3", 1500)
    prs = _synth_primers(panel)[:120]
    arrays = reads.oracle_arrays()
    run_both(prs, arrays, ref_names=panel.ref_names)                                            # all defaults (IP:170-183)
    run_both(prs, arrays, ref_names=panel.ref_names, slop=0, max_mismatch_rate=0.2, ungapped_kmer_length=4, gapped_kmer_length=5)
    run_both(prs, arrays, ref_names=panel.ref_names, slop=12, match_score=2, mismatch_score=-1, gap_open=-2, gap_extend=-2,
             min_alignment_score_rate=0.5, gapped_kmer_length=8)
    run_both(prs, arrays, ref_names=panel.ref_names, gap_open=0, gap_extend=0, mismatch_score=0, match_score=0, gapped_kmer_length=12)
    run_both(prs, arrays, ref_names=panel.ref_names, three_prime=0, gapped_kmer_length=10, ungapped_kmer_length=6)


@pytest.mark.parametrize("scores", [(1, -4, -6, -1), (7, -9, -11, -3), (40, -45, -30, -5), (60, -70, -64, -12), (300, -400, -250, -60),
                                    (5, 0, 0, 0), (0, -3, -2, 0)])
def test_score_ranges_all_primers_and_three_prime(scores):
    """Every primer is aligned (no -K) and 3' matching is on: the packed 16-bit aligner where the score range allows it,
    the 32-bit one above that (the last but two rows), odd primer counts and mixed primer lengths."""
    match, mismatch, gap_open, gap_extend = scores
    panel, reads, _ = synth.workload("c3", 700)
    prs = _synth_primers(panel)[:75]          # odd count: the last thread pairs a primer with itself
    arrays = reads.oracle_arrays()
    for slop, three in ((5, -1), (0, 4), (11, 0)):
        run_both(prs, arrays, ref_names=panel.ref_names, slop=slop, three_prime=three, gapped_kmer_length=None, ungapped_kmer_length=6,
                 match_score=match, mismatch_score=mismatch, gap_open=gap_open, gap_extend=gap_extend, min_alignment_score_rate=0.05)


@pytest.mark.parametrize("scores", [(1, -4, -6, -1), (1, -32, -30, -1), (1, -1, 0, 0), (1, 0, 0, -1), (1, -2, -1, 0), (2, -3, -4, -2), (1, -33, -6, -1)])
@pytest.mark.parametrize("read_len", [150, 254, 300])
def test_three_prime_two_pairs_per_thread(scores, read_len):
    """3' Local alignments with traceback, two (read, primer) pairs per thread in 16-bit half-words (sw_trace2_local16): at the
    edges of its value range, and past them -- rows * match above 31, a mismatch below -32, reads above 254 bases -- where the
    pairs take the 32-bit aligner one at a time.  Mapped and unmapped reads, both strands, primers of mixed lengths."""
    match, mismatch, gap_open, gap_extend = scores
    w = synth.WORKLOADS["c5"]
    panel = synth.make_panel(w["pairs"], w["panel_seed"], degenerate=False, mapped=True, read_len=read_len)
    reads = synth.make_reads(panel, 1501, w["read_seed"], indel_frac=0.1, mapped_frac=0.6, read_len=read_len)
    opts = dict(w["options"])
    opts.update(match_score=match, mismatch_score=mismatch, gap_open=gap_open, gap_extend=gap_extend, min_alignment_score_rate=0.2)
    got5, got3, _, _ = run_both(_synth_primers(panel), reads.oracle_arrays(), ref_names=panel.ref_names, **opts)
    assert (got3["type"] == L.GAPPED).any()


@pytest.mark.parametrize("scores,slop", [((1, -4, -6, -1), 5), ((1, -4, -6, -1), 40), ((1, -4, -6, -1), 0), ((2, -3, -4, 0), 5), ((1, -12, -10, -1), 7),
                                         ((3, -2, -1, -1), 5), ((1, -20, -10, -1), 5), ((0, -1, -1, -1), 3), ((1, 0, 0, 0), 9)])
def test_five_prime_two_pairs_per_thread(scores, slop):
    """5' Glocal alignments with traceback, two (read, primer) pairs per thread in 16-bit half-words (sw_trace2_glocal16): at the
    edges of its value range and past them -- targets above 63 bases (slop 40), scores whose shifted values leave eight bits --
    where single pairs (or the whole panel) take the 32-bit aligner.  Indel reads of c3 behind the -K candidate filter."""
    match, mismatch, gap_open, gap_extend = scores
    panel, reads, opts = synth.workload("c3", 2501)
    opts.update(match_score=match, mismatch_score=mismatch, gap_open=gap_open, gap_extend=gap_extend, slop=slop, min_alignment_score_rate=0.1)
    got5, _, _, t = run_both(_synth_primers(panel)[:400], reads.oracle_arrays(), ref_names=panel.ref_names, **opts)
    assert t["n_alignments_5p"] > 0


@pytest.mark.parametrize("scores", [(1, -4, -6, -1), (1, -1, -1, 0), (2, -3, -4, -2), (1, -8, -1, -1), (3, -1, -9, 0)])
@pytest.mark.parametrize("n_panel_pairs", [96, 7, 300])
def test_three_prime_all_primers(scores, n_panel_pairs):
    """3' matching of templates without a 5' match aligns every primer (IP:497): half of the templates are off-target here, panels
    of 14, 192 and 600 primers (one, two and several rounds of the all-primers kernel, odd thread counts)."""
    match, mismatch, gap_open, gap_extend = scores
    w = synth.WORKLOADS["c5"]
    panel = synth.make_panel(n_panel_pairs, w["panel_seed"] + n_panel_pairs, degenerate=n_panel_pairs == 7, mapped=True)
    reads = synth.make_reads(panel, 400, w["read_seed"], indel_frac=0.05, mapped_frac=0.3, offtarget=0.5)
    opts = dict(w["options"])
    opts.update(match_score=match, mismatch_score=mismatch, gap_open=gap_open, gap_extend=gap_extend, min_alignment_score_rate=0.0)
    got5, got3, _, t = run_both(_synth_primers(panel), reads.oracle_arrays(), ref_names=panel.ref_names, **opts)
    assert (got5["type"] == 0).sum() > 300 and (got3["type"] == L.GAPPED).sum() > 300


def test_ties_near_duplicate_primers():
    """Candidate order decides ties (PM:241, PM:309): near-duplicate primers, with and without the k-mer filters."""
    rng = np.random.default_rng(11)
    base = "ACGTTGCAAGGCTTACGATCGGA"
    primers = []
    for i in range(24):
        s = list(base)
        for j in rng.choice(len(base), size=i % 3, replace=False):
            s[j] = "ACGT"[(("ACGT".index(s[j])) + 1) % 4]
        seq = "".join(s)[: 18 + i % 6]
        primers.append(H.Primer(f"p{i}", f"p{i}+", seq, True))
        primers.append(H.Primer(f"p{i}", f"p{i}-", H.revcomp(seq)[: 18 + (i * 5) % 6], False))
    recs = []
    for i in range(600):
        p = primers[int(rng.integers(0, len(primers)))]
        s = list(p.sequence + "".join("ACGT"[x] for x in rng.integers(0, 4, 30)))
        for j in rng.choice(20, size=int(rng.integers(0, 4)), replace=False):
            s[j] = "ACGTN"[int(rng.integers(0, 5))]
        if rng.random() < 0.3:
            del s[int(rng.integers(2, 15))]
        r = H.Rec("".join(s), True, False, paired=False)
        recs.append(r)
    arrays = H.recs_to_arrays(recs)
    for k, kk in [(None, None), (4, 4), (6, 8), (3, 3)]:
        run_both(primers, arrays, max_mismatch_rate=0.15, ungapped_kmer_length=k, gapped_kmer_length=kk, min_alignment_score_rate=0.0)
        run_both(primers, arrays, max_mismatch_rate=0.15, ungapped_kmer_length=k, gapped_kmer_length=kk, three_prime=3, mismatch_score=-2)


def test_ragged_short_empty_and_iupac_reads():
    primers = H.primers_from(K["pmt_panel"]) + [H.Primer("7", "7+", "ACGTRYACGTNACGTMKACG", True), H.Primer("7", "7-", "TTGACCAGTWSGATTACA", False)]
    rng = np.random.default_rng(5)
    recs = [H.Rec("", True, False), H.Rec("A", True, False), H.Rec("ACGT", True, False), H.Rec("N" * 30, True, False),
            H.Rec("=" * 12, True, False), H.Rec("ACGTAYACGTGACGTAGACG" + "RYKM=NBDHV", True, False)]
    for _ in range(400):
        n = int(rng.integers(0, 60))
        alphabet = "ACGT" * 6 + "NRYKMSWBDHV="
        s = "".join(alphabet[int(x)] for x in rng.integers(0, len(alphabet), n))
        if rng.random() < 0.5 and n > 12:
            p = primers[int(rng.integers(0, len(primers)))]
            s = (p.sequence + s)[:n]
        mapped = rng.random() < 0.4
        neg = bool(rng.random() < 0.5)
        stored = H.revcomp(s) if (mapped and neg) else s
        recs.append(H.Rec(stored, not mapped, neg, ref_idx=int(rng.integers(0, 4)), start=int(rng.integers(-3, 1100))))
    # pair some of them up, leave the rest as fragments
    paired = H.pair_records(recs[:200]) + recs[200:]
    arrays = H.recs_to_arrays(paired)
    for opts in [dict(), dict(ungapped_kmer_length=5, gapped_kmer_length=5), dict(max_mismatch_rate=0.3, slop=3, three_prime=2),
                 dict(max_mismatch_rate=1.0, min_alignment_score_rate=0.0, ungapped_kmer_length=2)]:
        run_both(primers, arrays, **opts)


def test_long_primers_use_wider_kernels():
    rng = np.random.default_rng(9)
    for plen in (40, 64, 100, 200):
        primers = []
        for i in range(6):
            s = "".join("ACGT"[x] for x in rng.integers(0, 4, plen - i))
            primers.append(H.Primer(f"p{i}", f"p{i}+", s, True))
            primers.append(H.Primer(f"p{i}", f"p{i}-", "".join("ACGT"[x] for x in rng.integers(0, 4, plen // 2 + i)), False))
        recs = []
        for i in range(120):
            p = primers[int(rng.integers(0, len(primers)))]
            s = list(p.sequence + "".join("ACGT"[x] for x in rng.integers(0, 4, 60)))
            if i % 3 == 0:
                del s[plen // 3]
            if i % 3 == 1:
                s.insert(plen // 4, "G")
            if i % 5 == 0:
                s[3] = "N"
            recs.append(H.Rec("".join(s)[: 250], True, False))
        arrays = H.recs_to_arrays(recs)
        run_both(primers, arrays, max_mismatch_rate=0.05, gapped_kmer_length=10, ungapped_kmer_length=8)
        run_both(primers, arrays, max_mismatch_rate=0.05, three_prime=0)


def test_scoring_matrix():
    sm = np.full((128, 128), np.iinfo(np.int32).min, dtype=np.int32)
    for a in "ACGT":
        for b in "ACGTN":
            sm[ord(a), ord(b)] = 2 if a == b else (-1 if b == "N" else -3)
    for a, members in {"R": "AG", "Y": "CT", "N": "ACGT"}.items():
        for b in "ACGT":
            sm[ord(a), ord(b)] = 1 if b in members else -3
    primers = [H.Primer("1", "1+", "ACGTRYACGTNACGTACGAT", True), H.Primer("1", "1-", "TTGACCAGTAGGATTACA", False),
               H.Primer("2", "2+", "GGATCCRYTTAAGCCGTA", True), H.Primer("2", "2-", "CCATGGTTAACCGGTTAA", False)]
    rng = np.random.default_rng(2)
    recs = []
    for i in range(300):
        p = primers[i % 4]
        s = list(p.sequence.replace("R", "A").replace("Y", "C").replace("N", "G") + "".join("ACGT"[x] for x in rng.integers(0, 4, 40)))
        if i % 2:
            del s[5 + i % 7]
        if i % 7 == 0:
            s[8] = "N"
        if i % 50 == 0:
            s[9] = "K"   # a read base the matrix has no entry for -> the reference would throw (status 4)
        recs.append(H.Rec("".join(s), True, False))
    got5, _, _, _ = run_both(primers, H.recs_to_arrays(recs), scoring_matrix=sm, gap_open=-4, gap_extend=-1, min_alignment_score_rate=0.0)
    assert (got5["status"] == 4).any() and (got5["type"] == L.GAPPED).any()


def test_align_batch_matches_oracle_aligner():  # AsyncAlignerTest.scala:33-57, all three modes
    rng = np.random.default_rng(4)
    queries, targets = [], []
    for i in range(500):
        n, m = int(rng.integers(1, 40)), int(rng.integers(0, 80))
        q = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
        t = "".join("ACGT"[x] for x in rng.integers(0, 4, m))
        if i % 2 and m > n:
            pos = int(rng.integers(0, m - n + 1))
            qq = list(q)
            if n > 4 and i % 4 == 1:
                del qq[n // 2]
            t = t[:pos] + "".join(qq) + t[pos + len(qq):]
        queries.append(q.encode()); targets.append(t.encode())
    queries += [b"ACGTG", b"AATCGATA", b"GATTACA"]; targets += [b"ACGTG", b"AATGATA", b"TTTGATTACATTT"]
    for scores in [dict(match_score=1, mismatch_score=-4, gap_open=-6, gap_extend=-1), dict(match_score=2, mismatch_score=-1, gap_open=-1, gap_extend=-1),
                   dict(match_score=1, mismatch_score=-3, gap_open=0, gap_extend=-2)]:
        params = O.make_params(**scores)
        for mode in (L.MODE_LOCAL, L.MODE_GLOCAL, L.MODE_GLOBAL):
            al = idp.AsyncCudaAligner(mode=mode, **scores)
            score, ts, te = al.align(queries, targets)                                   # register-resident aligners (Local, Glocal)
            s2, ts2, te2, qs, qe = al.align(queries, targets, query_coordinates=True)     # Local: the generic one, with query coordinates
            short = [i for i, q in enumerate(queries) if len(q) <= 32]                   # a batch that fits the 32-row variant
            s3, ts3, te3 = al.align([queries[i] for i in short], [targets[i] for i in short])
            al.close()
            for i, (q, t) in enumerate(zip(queries, targets)):
                a = O.align(params, mode, q, t)
                assert (int(score[i]), int(ts[i]), int(te[i])) == (a["score"], a["target_start"], a["target_end"]), (mode, scores, q, t, a)
                assert (int(s2[i]), int(ts2[i]), int(te2[i]), int(qs[i]), int(qe[i])) == (a["score"], a["target_start"], a["target_end"], a["query_start"], a["query_end"]), (mode, scores, q, t, a)
            assert np.array_equal(s3, score[short]) and np.array_equal(ts3, ts[short]) and np.array_equal(te3, te[short])


def test_align_batch_long_queries_and_bad_letters():
    rng = np.random.default_rng(6)
    queries, targets = [], []
    for i in range(120):   # queries beyond the 64 register rows: the generic aligner, all modes
        n, m = int(rng.integers(1, 200)), int(rng.integers(0, 300))
        q = "".join("ACGTN"[x] for x in rng.integers(0, 5, n))
        t = "".join("ACGTN"[x] for x in rng.integers(0, 5, m))
        if i % 2 and m > n:
            pos = int(rng.integers(0, m - n + 1))
            t = t[:pos] + q + t[pos + n:]
        queries.append(q.encode()); targets.append(t.encode())
    scores = dict(match_score=1, mismatch_score=-4, gap_open=-6, gap_extend=-1)
    params = O.make_params(**scores)
    for mode in (L.MODE_LOCAL, L.MODE_GLOCAL, L.MODE_GLOBAL):
        al = idp.AsyncCudaAligner(mode=mode, **scores)
        s, ts, te, qs, qe = al.align(queries, targets, query_coordinates=True)
        for i, (q, t) in enumerate(zip(queries, targets)):
            a = O.align(params, mode, q, t)
            assert (int(s[i]), int(ts[i]), int(te[i]), int(qs[i]), int(qe[i])) == (a["score"], a["target_start"], a["target_end"], a["query_start"], a["query_end"]), (mode, q, t, a)
        with pytest.raises(L.IdpError) as err:   # lower case is not accepted (checked on the GPU while packing)
            al.align([b"ACGT", b"ACgT"], [b"ACGTACGT", b"ACGTACGT"])
        assert err.value.code == L.ERR_UNSUPPORTED
        s4, _, _ = al.align([b"ACGT"], [b"TTACGTTT"])   # the context is still usable
        assert int(s4[0]) == (4 if mode != L.MODE_GLOBAL else O.align(params, mode, b"ACGT", b"TTACGTTT")["score"])
        al.close()


# ------------------------------------------------------------------------------------------------------------------
# size-independent properties at a size the oracle could not finish quickly
# ------------------------------------------------------------------------------------------------------------------
def test_large_batch_properties():
    n_pairs = 1_000_000
    panel, reads, opts = synth.workload("c2", n_pairs)
    gp = panel.primers
    seq4 = reads.seq4()
    with idp.IdentifyPrimers(gp, **opts) as tool:
        batch = idp.ReadBatch(seq4.reshape(-1), reads.flags, fixed_len=150)
        five, _ = tool.process_batch(batch)
        m1 = tool.metric_counter.copy()
        assert tool.summary()["templates"] == n_pairs and tool.summary()["match_attempts"] == 2 * n_pairs
        # idempotence and chunk-size independence
        import os
        os.environ["IDP_CHUNK_READS"] = "77777"
        five_b, _ = tool.process_batch(batch)
        del os.environ["IDP_CHUNK_READS"]
        assert np.array_equal(five.view(np.uint8), five_b.view(np.uint8))
        assert np.array_equal(tool.metric_counter, 2 * m1)
        # permutation of templates permutes results
        perm = np.random.default_rng(1).permutation(n_pairs)
        ridx = np.stack([2 * perm, 2 * perm + 1], axis=1).reshape(-1)
        five_p, _ = tool.process_batch(idp.ReadBatch(seq4[ridx].reshape(-1), reads.flags[ridx], fixed_len=150))
        assert np.array_equal(five_p.view(np.uint8), five[ridx].view(np.uint8))
        # matched primers belong to the amplicon the read was drawn from, a bounded prefix agrees with the oracle
        hit = five["type"] != 0
        assert (five["primer"][hit] // 2 == reads.truth[hit]).mean() > 0.999
    lo, hi = 0, 40000
    panel_o = O.Panel(_synth_primers(panel), _opts_to_oracle(opts))
    want5, _, _, _ = panel_o.process_batch(reads.oracle_arrays(lo, hi), n_threads=8)
    assert_same(five[lo:hi], want5, "prefix")


# ------------------------------------------------------------------------------------------------------------------
# the bit-sliced fast path of the scan kernel (all accept thresholds <= 1) against the oracle and against the exact walk
# ------------------------------------------------------------------------------------------------------------------
def test_fast_filter_ties_and_iupac_dense_primers():
    rng = np.random.default_rng(21)
    base = "ACGTTGCAAGGCTTACGATCGGATTC"
    primers = []
    for i in range(20):   # near-duplicates: several primers accept the same read -> candidate order decides
        s = list(base[: 20 + i % 6])
        if i % 4:
            j = int(rng.integers(0, len(s)))
            s[j] = "ACGT"[("ACGT".index(s[j]) + 1 + i % 3) % 4]
        primers.append(H.Primer(f"d{i}", f"d{i}+", "".join(s), True))
        primers.append(H.Primer(f"d{i}", f"d{i}-", H.revcomp("".join(s)), False))
    for i in range(10):   # IUPAC letter every few bases: accepted by mismatch count, but no literal k-mer on the diagonal
        s = list("".join("ACGT"[x] for x in rng.integers(0, 4, 22)))
        for j in range(2 + i % 3, 22, 4 + i % 3):
            s[j] = {"A": "R", "C": "Y", "G": "R", "T": "Y"}[s[j]]
        primers.append(H.Primer(f"u{i}", f"u{i}+", "".join(s), True))
        primers.append(H.Primer(f"u{i}", f"u{i}-", "".join("ACGT"[x] for x in rng.integers(0, 4, 21)), False))
    inst = {"R": "AG", "Y": "CT"}
    recs = []
    for i in range(1500):
        p = primers[int(rng.integers(0, len(primers)))]
        s = [inst[ch][int(rng.integers(0, 2))] if ch in inst else ch for ch in p.sequence]
        s += ["ACGT"[x] for x in rng.integers(0, 4, 30)]
        if rng.random() < 0.5:
            s[int(rng.integers(0, 20))] = "ACGTN"[int(rng.integers(0, 5))]
        if rng.random() < 0.1:
            s = s[: int(rng.integers(3, 25))]
        recs.append(H.Rec("".join(s), True, False))
    arrays = H.recs_to_arrays(recs)
    for k in (None, 4, 6, 8):
        for rate in (0.0, 0.05):
            got5, _, _, _ = run_both(primers, arrays, max_mismatch_rate=rate, ungapped_kmer_length=k, gapped_kmer_length=10)
            assert (got5["type"] == L.UNGAPPED).any()


def test_kmer_rule_depends_on_mismatch_position():
    """PM:108-116 with degenerate primers: the primer is a candidate only through a run of k literally equal letters, so
    whether an accepted primer (at most one mismatch) is reported depends on where the mismatch sits.  Primers with a single
    plain run, with runs of exactly k, with no plain run at all; reads with the mismatch at every position, reads that
    spell the primer's ambiguity letter literally, reads carrying N."""
    rng = np.random.default_rng(77)
    deg = {"A": "R", "C": "Y", "G": "R", "T": "Y"}
    inst = {"R": "AG", "Y": "CT"}
    primers = []

    def add(name, seq):
        primers.append(H.Primer(name, name + "+", seq, True))
        primers.append(H.Primer(name, name + "-", "".join("ACGT"[x] for x in rng.integers(0, 4, 20)), False))

    for i in range(6):   # one plain run of exactly k = 6 at a different place each time, everything else degenerate
        s = ["ACGT"[x] for x in rng.integers(0, 4, 20 + i)]
        lo = 2 * i + 1
        add(f"one{i}", "".join(ch if lo <= j < lo + 6 else deg[ch] for j, ch in enumerate(s)))
    for d in (5, 6, 7, 8):   # a degenerate letter every d bases: plain runs of d-1
        s = ["ACGT"[x] for x in rng.integers(0, 4, 26)]
        add(f"every{d}", "".join(deg[ch] if j % d == d - 1 else ch for j, ch in enumerate(s)))
    add("plain", "".join("ACGT"[x] for x in rng.integers(0, 4, 24)))
    recs = []
    for p in primers[::2]:
        n = len(p.sequence)
        for pos in range(-1, n):   # -1: no mismatch
            for literal in (False, True):
                s = [(ch if literal else inst[ch][int(rng.integers(0, 2))]) if ch in inst else ch for ch in p.sequence]
                if pos >= 0:
                    allowed = inst.get(p.sequence[pos], p.sequence[pos])
                    s[pos] = rng.choice([b for b in "ACGTN" if b not in allowed])
                s += ["ACGT"[x] for x in rng.integers(0, 4, 40)]
                recs.append(H.Rec("".join(s), True, False))
    arrays = H.recs_to_arrays(recs)
    for k in (5, 6, 7):
        got5, _, _, _ = run_both(primers, arrays, max_mismatch_rate=0.05, ungapped_kmer_length=k, gapped_kmer_length=8)
        types = got5["type"]
        assert (types == L.UNGAPPED).any() and (types != L.UNGAPPED).any()


@pytest.mark.parametrize("name,n_pairs", [("c2", 300_000), ("c3", 100_000), ("c5", 100_000)])
def test_fast_filter_equals_exact_walk(name, n_pairs):
    import os
    panel, reads, opts = synth.workload(name, n_pairs)
    batch = idp.ReadBatch(reads.seq4().reshape(-1), reads.flags, fixed_len=150, ref_idx=reads.ref_idx, unclipped_start=reads.ustart,
                          unclipped_end=reads.uend)
    out = []
    for exact in (False, True):
        if exact:
            os.environ["IDP_SCAN_EXACT_WALK"] = "1"
        try:
            with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
                five, three = tool.process_batch(batch)
                out.append((five.copy(), None if three is None else three.copy(), tool.metric_counter.copy(), tool.num_alignments))
        finally:
            os.environ.pop("IDP_SCAN_EXACT_WALK", None)
    assert np.array_equal(out[0][0].view(np.uint8), out[1][0].view(np.uint8))
    if out[0][1] is not None:
        assert np.array_equal(out[0][1].view(np.uint8), out[1][1].view(np.uint8))
    assert np.array_equal(out[0][2], out[1][2]) and out[0][3] == out[1][3]


# ------------------------------------------------------------------------------------------------------------------
# input layouts: fixed-length batches of odd sizes (TMA tile tails), chunked host path on ragged reads, device pointers
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("read_len,n_reads", [(150, 1), (150, 31), (150, 33), (149, 1001), (151, 777), (36, 513), (21, 64)])
def test_fixed_length_layouts(read_len, n_reads):
    panel, reads, opts = synth.workload("c1", 1200)
    codes = reads.codes[:n_reads, :read_len] if read_len <= 150 else np.concatenate(
        [reads.codes[:n_reads], reads.codes[n_reads:2 * n_reads, :read_len - 150]], axis=1)
    flags = np.full(n_reads, H.F_UNMAPPED, dtype=np.uint16)      # fragments
    arrays = dict(bases=synth.codes_to_ascii(codes).reshape(-1), off=np.arange(n_reads + 1, dtype=np.int64) * read_len,
                  ref_idx=np.full(n_reads, -1, np.int32), ustart=np.zeros(n_reads, np.int32), uend=np.zeros(n_reads, np.int32), flags=flags)
    want5, _, want_metrics, want_align = O.Panel(_synth_primers(panel), _opts_to_oracle(opts)).process_batch(arrays, n_threads=4)
    with idp.IdentifyPrimers(panel.primers, **opts) as tool:
        got5, _ = tool.process_batch(idp.ReadBatch(synth.pack_fixed(codes).reshape(-1), flags, fixed_len=read_len))
        assert_same(got5, want5, f"fixed_len={read_len}")
        assert np.array_equal(tool.metric_counter, want_metrics) and tool.num_alignments == want_align


def test_chunked_host_path_on_ragged_mapped_reads():
    import os
    panel, reads, opts = synth.workload("c5", 700)
    rng = np.random.default_rng(3)
    arr = reads.oracle_arrays()
    # ragged: trim every read to a random length (mapped minus-strand reads keep their 3' end = sequencing 5' end)
    lens = rng.integers(12, 151, reads.n)
    bases, off = [], [0]
    neg = (reads.flags & H.F_REVERSE) != 0
    for i in range(reads.n):
        row = arr["bases"][i * 150:(i + 1) * 150]
        bases.append(row[150 - lens[i]:] if neg[i] else row[:lens[i]])
        off.append(off[-1] + int(lens[i]))
    ustart, uend = arr["ustart"].copy(), arr["uend"].copy()
    ustart[neg] = uend[neg] - lens[neg] + 1
    uend[~neg] = ustart[~neg] + lens[~neg] - 1
    arrays = dict(bases=np.concatenate(bases), off=np.array(off, dtype=np.int64), ref_idx=arr["ref_idx"], ustart=ustart, uend=uend, flags=arr["flags"])
    want5, want3, want_metrics, want_align = O.Panel(_synth_primers(panel), _opts_to_oracle(opts)).process_batch(arrays, n_threads=4)
    for chunk in ("37", "256", "100000"):
        os.environ["IDP_CHUNK_READS"] = chunk
        try:
            with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
                got5, got3 = tool.process_batch(idp.ReadBatch.from_ascii(arrays["bases"], arrays["off"], arrays["flags"], arrays["ref_idx"], ustart, uend))
                assert_same(got5, want5, f"chunk {chunk} 5'")
                assert_same(got3, want3, f"chunk {chunk} 3'")
                assert np.array_equal(tool.metric_counter, want_metrics) and tool.num_alignments == want_align
        finally:
            os.environ.pop("IDP_CHUNK_READS", None)


def test_device_pointer_entry_with_offsets():
    import torch
    panel, reads, opts = synth.workload("c5", 1200)
    arr = reads.oracle_arrays()
    want5, want3, want_metrics, want_align = O.Panel(_synth_primers(panel), _opts_to_oracle(opts)).process_batch(arr, n_threads=4)
    batch = idp.ReadBatch.from_ascii(arr["bases"], arr["off"], arr["flags"], arr["ref_idx"], arr["ustart"], arr["uend"])
    dev = torch.device("cuda", 0)
    t = {k: torch.from_numpy(v.view(np.int64) if v.dtype == np.uint64 else (v.view(np.int16) if v.dtype == np.uint16 else v)).to(dev)
         for k, v in dict(seq4=batch.seq4, seq_off=batch.seq_off, len=batch.len, ref=batch.ref_idx, us=batch.unclipped_start, ue=batch.unclipped_end, fl=batch.flags).items()}
    f5 = torch.zeros((reads.n, 32), dtype=torch.uint8, device=dev)
    f3 = torch.zeros((reads.n, 32), dtype=torch.uint8, device=dev)
    with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
        tool.set_stream(torch.cuda.current_stream().cuda_stream)
        metrics, n_align = tool.process_batch_device(reads.n, t["seq4"].data_ptr(), t["fl"].data_ptr(), f5.data_ptr(), f3.data_ptr(),
                                                     seq_off_ptr=t["seq_off"].data_ptr(), len_ptr=t["len"].data_ptr(), ref_idx_ptr=t["ref"].data_ptr(),
                                                     ustart_ptr=t["us"].data_ptr(), uend_ptr=t["ue"].data_ptr())
        torch.cuda.synchronize()
        assert_same(f5.cpu().numpy().view(L.RESULT_DTYPE).reshape(-1), want5, "device 5'")
        assert_same(f3.cpu().numpy().view(L.RESULT_DTYPE).reshape(-1), want3, "device 3'")
        assert np.array_equal(metrics, want_metrics) and n_align == want_align


def test_maximum_sizes_and_empty_batch():
    rng = np.random.default_rng(17)
    seq = lambda n: "".join("ACGT"[x] for x in rng.integers(0, 4, n))
    # longest supported primer (256 bases) and reads up to the 4095-base limit, 5' and 3' matching
    primers = [H.Primer("L", "L+", seq(256), True), H.Primer("L", "L-", seq(255), False), H.Primer("S", "S+", seq(20), True), H.Primer("S", "S-", seq(33), False)]
    recs = []
    for i, n in enumerate([4095, 4000, 1000, 300, 256, 255, 40, 20]):
        p = primers[i % 4]
        body = list(p.sequence + seq(4095))[:n]
        if i % 2 and n > 30:
            del body[11]
            body.append("A")
        recs.append(H.Rec("".join(body), True, False))
        recs.append(H.Rec(H.revcomp(p.sequence)[: max(1, n // 2)] + seq(10), True, False))   # primer at the 3' end of a short read
    arrays = H.recs_to_arrays(recs)
    run_both(primers, arrays, max_mismatch_rate=0.02, gapped_kmer_length=12, ungapped_kmer_length=8)
    run_both(primers, arrays, max_mismatch_rate=0.02, three_prime=3, min_alignment_score_rate=0.0)
    # an empty batch is legal
    with idp.IdentifyPrimers([idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand) for p in primers], three_prime=0) as tool:
        five, three = tool.process_batch(idp.ReadBatch(np.zeros(0, np.uint8), np.zeros(0, np.uint16), seq_off=np.zeros(0, np.uint64), length=np.zeros(0, np.int32)))
        assert len(five) == 0 and len(three) == 0 and tool.metric_counter.sum() == 0
    # beyond the limits the library refuses instead of approximating
    with pytest.raises(idp.IdpError):
        idp.IdentifyPrimers([idp.Primer("x", "x+", seq(257), True), idp.Primer("x", "x-", seq(20), False)], device=None)


# ------------------------------------------------------------------------------------------------------------------
# downstream classification as a device epilogue (scripts/post_process_bam.py:56-78): against the host function, which
# tests/test_host_abi.py pins to the outputs of the reference's own Python function (tests/golden/classify_golden.json)
# ------------------------------------------------------------------------------------------------------------------
def _golden_classify_panel():
    import json
    g = json.load(open(os.path.join(H.ROOT, "tests", "golden", "classify_golden.json")))
    return g, [idp.Primer(p["pair_id"], p["primer_id"], p["sequence"], p["positive_strand"], p["ref_name"], p["start"], p["end"]) for p in g["primers"]]


@pytest.mark.parametrize("compact", [False, True])
def test_device_classification_of_crafted_records(compact):
    """Every (R1 match, R2 match) combination over the golden panel, pairs and fragments, through idp_classify_batch_device."""
    import torch
    g, primers = _golden_classify_panel()
    n_p = len(primers)
    rng = np.random.default_rng(12)
    m1, m2 = np.meshgrid(np.arange(-1, n_p), np.arange(-1, n_p), indexing="ij")
    m1, m2 = m1.reshape(-1), m2.reshape(-1)
    frag = rng.integers(-1, n_p, 500)
    prim = np.concatenate([np.stack([m1, m2], axis=1).reshape(-1), frag])
    flags = np.concatenate([np.tile([H.F_PAIRED | H.F_FIRST | H.F_MATE_NEXT | H.F_UNMAPPED, H.F_PAIRED | H.F_SECOND | H.F_UNMAPPED], len(m1)),
                            np.where(rng.random(500) < 0.5, H.F_UNMAPPED, H.F_UNMAPPED | H.F_PAIRED | H.F_FIRST)]).astype(np.uint16)
    five = np.zeros(len(prim), dtype=L.RESULT_DTYPE)
    five["primer"] = prim
    five["type"] = np.where(prim >= 0, rng.integers(1, 4, len(prim)), 0)
    five["next_mm"] = np.where(five["type"] == L.UNGAPPED, L.NEXT_NA, 0)
    with idp.IdentifyPrimers(primers, ref_names=["chr1", "chr2", "chr3", "chrX"]) as tool:
        for lo_hi in ((50, 250), (102, 207), (0, 10)):
            want_cls, want_counts = tool.classify_templates(flags, five, *lo_hi)
            dev = torch.device("cuda", 0)
            if compact:
                rec = np.zeros(len(prim), dtype=L.RESULT16_DTYPE)
                rec["head"] = (prim + 1).astype(np.uint32) | (five["type"].astype(np.uint32) << 24)
                rec["b"] = np.where(five["type"] == L.UNGAPPED, L.NEXT16_NA, 0)
                assert L.unpack16(rec).tobytes() == five.tobytes()
                d5 = torch.from_numpy(rec.view(np.uint8).reshape(-1, 16)).to(dev)
            else:
                d5 = torch.from_numpy(five.view(np.uint8).reshape(-1, 32)).to(dev)
            dfl = torch.from_numpy(flags.view(np.int16)).to(dev)
            dcls = torch.full((len(prim),), 255, dtype=torch.uint8, device=dev)
            counts = tool.classify_batch_device(dfl.data_ptr(), d5.data_ptr(), len(prim), compact, dcls.data_ptr(), *lo_hi)
            assert counts == want_counts
            assert np.array_equal(dcls.cpu().numpy(), want_cls)
            assert len(set(want_cls.tolist())) == 6   # every class occurs


@pytest.mark.parametrize("name,compact", [("c5", True), ("c1", False)])
def test_classification_epilogue_inside_process_batch(name, compact):
    panel, reads, opts = synth.workload(name, 3000)
    with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
        batch = idp.ReadBatch(reads.seq4().reshape(-1), reads.flags, fixed_len=150, ref_idx=reads.ref_idx, unclipped_start=reads.ustart,
                              unclipped_end=reads.uend)
        five, _ = tool.process_batch(batch)
        want_cls, want_counts = tool.classify_templates(reads.flags, five, 60, 240)
        tool.set_classify(60, 240)
        os.environ["IDP_CHUNK_READS"] = "1111"   # several chunks: the class bytes come back chunk by chunk
        try:
            five2, _ = tool.process_batch(batch, compact=compact)
        finally:
            del os.environ["IDP_CHUNK_READS"]
        assert five2.tobytes() == five.tobytes()
        assert np.array_equal(tool.last_classes, want_cls)
        assert dict(zip(tool.CLASS_NAMES, tool.class_counts.tolist())) == want_counts
        assert want_counts["Canonical"] > 1000


@pytest.mark.parametrize("stage_timing", [True, False])
def test_asynchronous_device_mode_accumulates_counters(stage_timing):
    """idp_ctx_set_async / idp_ctx_sync: batches enqueued back to back give the same records, and the counters of all of them at the
    sync -- with the per-stage events (mode 1) and without them and the per-batch counter copy (mode 2)."""
    import torch
    panel, reads, opts = synth.workload("c5", 2000)
    arr = reads.oracle_arrays()
    want5, want3, want_metrics, want_align = O.Panel(_synth_primers(panel), _opts_to_oracle(opts)).process_batch(arr, n_threads=4)
    dev = torch.device("cuda", 0)
    s4 = torch.from_numpy(reads.seq4()).to(dev)
    fl = torch.from_numpy(reads.flags.view(np.int16)).to(dev)
    loc = [torch.from_numpy(a).to(dev) for a in (reads.ref_idx, reads.ustart, reads.uend)]
    outs = [(torch.zeros((reads.n, 16), dtype=torch.uint8, device=dev), torch.zeros((reads.n, 16), dtype=torch.uint8, device=dev)) for _ in range(3)]
    with idp.IdentifyPrimers(panel.primers, ref_names=panel.ref_names, **opts) as tool:
        tool.set_stream(torch.cuda.current_stream().cuda_stream)
        tool.set_async(True, stage_timing=stage_timing)
        for f5, f3 in outs:
            m, na = tool.process_batch_device(reads.n, s4.data_ptr(), fl.data_ptr(), f5.data_ptr(), f3.data_ptr(), fixed_len=150, ref_idx_ptr=loc[0].data_ptr(),
                                              ustart_ptr=loc[1].data_ptr(), uend_ptr=loc[2].data_ptr(), compact=True)
            assert not m.any() and na == 0   # nothing is reported before the sync
        metrics, n_align = tool.sync()
        assert np.array_equal(metrics, 3 * want_metrics) and n_align == 3 * want_align
        assert np.array_equal(tool.metric_counter, 3 * want_metrics)
        t = tool.timings()
        assert t["ms_total"] > 0 and (t["ms_scan"] > 0) == stage_timing and t["n_alignments_5p"] == 3 * want_align
        for f5, f3 in outs:
            assert_same(L.unpack16(f5.cpu().numpy().view(L.RESULT16_DTYPE).reshape(-1)), want5, "async 5'")
            assert_same(L.unpack16(f3.cpu().numpy().view(L.RESULT16_DTYPE).reshape(-1)), want3, "async 3'")
        assert not tool.sync()[0].any()   # nothing pending: a no-op
        tool.set_async(False)
        m, na = tool.process_batch_device(reads.n, s4.data_ptr(), fl.data_ptr(), outs[0][0].data_ptr(), outs[0][1].data_ptr(), fixed_len=150,
                                          ref_idx_ptr=loc[0].data_ptr(), ustart_ptr=loc[1].data_ptr(), uend_ptr=loc[2].data_ptr(), compact=True)
        assert np.array_equal(m, want_metrics) and na == want_align
