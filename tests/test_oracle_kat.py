"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for the hot path
(tests/golden/reference_kats.json, lifted from PrimerMatcherTest.scala / IdentifyPrimersTest.scala)."""
import numpy as np
import pytest

import helpers as H
from oracle import oracle_py as O

K = H.kats()


def test_num_mismatches_truth_table():  # PrimerMatcherTest.scala:161-185
    for row in K["num_mismatches"]:
        got = O.num_mismatches(row["read"].encode(), row["primer"].encode(), row["start"], row["max_mismatches"])
        assert got == row["expected"], row


def test_mismatch_align():  # PrimerMatcherTest.scala:187-218
    spec = K["mismatch_align"]
    names = list(spec["primers"])
    primers = H.primers_from([spec["primers"][n] for n in names])
    panel = O.Panel(primers, O.make_params(max_mismatch_rate=spec["max_mismatch_rate"]))
    for stored, unmapped, negative in [(spec["read_seq_order"], True, False), (spec["read_seq_order"], False, False),
                                       (spec["minus_read_stored"], False, True)]:
        so = H.Rec(stored, unmapped, negative).seq_order
        for i, n in enumerate(names):
            assert panel.mismatch_align(so.encode(), i) == spec["expected"][n], (n, stored)


def test_kmer_candidates():  # PrimerMatcherTest.scala:89-145
    primers = H.primers_from(K["pmt_panel"])
    k = K["kmer"]["k"]
    panel = O.Panel(primers, O.make_params(max_mismatch_rate=0.1, k_ungapped=k))
    for case in K["kmer"]["cases"]:
        got = panel.candidates(0, case["bases"].encode(), k)
        assert len(got) == case["n"], case
        assert len(set(got)) == len(got)
        for i in got:
            assert case["must_contain"] in primers[i].sequence
    for t in K["kmer"]["tasks"]:
        so = H.Rec(t["stored"], t["unmapped"], t["negative"]).seq_order
        got = panel.candidates(0, so.encode(), k)
        assert sorted(primers[i].primer_id for i in got) == t["ids"], t


def _frag_from_primer(p, read_length=20, start_adjust=0):  # PrimerMatcherTest.fragFromPrimer (:50-62)
    start = p.start if p.positive_strand else p.end - read_length + 1
    genomic = p.sequence if p.positive_strand else H.revcomp(p.sequence)
    rem = read_length - p.length
    bases = genomic + "N" * rem if p.positive_strand else "N" * rem + genomic
    return H.Rec(bases, False, not p.positive_strand, p.ref_idx, start + start_adjust)


def test_location_overlaps_and_find():  # PrimerMatcherTest.scala:236-300
    primers = H.primers_from(K["pmt_panel"])
    for slop, adjust, found in [(0, 0, True), (5, 5, True), (5, 6, False)]:
        panel = O.Panel(primers, O.make_params(slop=slop, max_mismatch_rate=1.0))
        for i, p in enumerate(primers):
            r = _frag_from_primer(p, start_adjust=adjust)
            assert panel.location_overlaps(r.ref_idx, r.negative, r.start, r.end) == ([i] if found else []), (slop, adjust, p)
    # find with rate 0 and an all-N read -> empty (:270-277); with rate 1 -> the primer with 0 mismatches (:279-285, 287-295)
    for rate, bases_fn, expect in [(0.0, lambda r: "N" * 20, None), (1.0, lambda r: r.bases, 0)]:
        panel = O.Panel(primers, O.make_params(slop=0, max_mismatch_rate=rate, match_score=1, mismatch_score=-1000, gap_open=-1000, gap_extend=-1000))
        recs = []
        for p in primers:
            r = _frag_from_primer(p)
            r.bases = bases_fn(r)
            recs.append(r)
        five, _, _, _ = panel.process_batch(H.recs_to_arrays(recs))
        for i, res in enumerate(five):
            if expect is None:
                assert res["type"] != O.LOCATION
            else:
                assert (res["type"], res["primer"], res["mm"]) == (O.LOCATION, i, 0)
    # two same-strand overlapping primers -> None (:297-309)
    shifted = [H.Primer(p.pair_id + "s", p.primer_id + "s", p.sequence, p.positive_strand, p.ref_name, p.start + 1, p.end + 1, p.ref_idx) for p in primers[:2]]
    panel = O.Panel(primers[:2] + shifted, O.make_params(slop=5, max_mismatch_rate=1.0))
    for p in primers[:2]:
        r = _frag_from_primer(p)
        assert len(panel.location_overlaps(r.ref_idx, r.negative, r.start, r.end)) == 2
        five, _, _, _ = panel.process_batch(H.recs_to_arrays([r]))
        assert five[0]["type"] != O.LOCATION


def test_best_ungapped():  # PrimerMatcherTest.scala:322-340
    panel = O.Panel([H.Primer("1", "1+", "GATTACA", True)], O.make_params())
    for case in K["best_ungapped"]:
        res = panel.best_ungapped([0] * len(case["mms"]), case["mms"])
        if case["expected"] is None:
            assert res["type"] == O.NONE
        else:
            assert (res["type"], res["mm"], res["next_mm"], res["status"]) == (O.UNGAPPED, case["expected"]["mm"], case["expected"]["next"], 0)


def test_ungapped_find_all_primers():  # PrimerMatcherTest.scala:342-374
    primers = H.primers_from(K["pmt_panel"])
    panel = O.Panel(primers, O.make_params(slop=0, max_mismatch_rate=1.0))
    for i, p in enumerate(primers):
        r = _frag_from_primer(p)
        so = r.seq_order.encode()
        all_mm = {j: panel.mismatch_align(so, j) for j in range(len(primers))}
        all_mm = {j: m for j, m in all_mm.items() if m is not None}
        # force the ungapped matcher: an unmapped copy of the read in sequencing order
        five, _, _, _ = panel.process_batch(H.recs_to_arrays([H.Rec(r.seq_order, True, False)]))
        best = five[0]
        assert best["type"] == O.UNGAPPED and best["mm"] == 0
        assert all_mm[i] == 0
        others = [m for j, m in all_mm.items() if j != best["primer"]]
        assert best["next_mm"] == (min(others) if others else O.NEXT_NA)


def test_best_gapped():  # PrimerMatcherTest.scala:396-422
    for case in K["best_gapped"]:
        n = len(case["scores"])
        res = O.best_gapped([0] * n, case["scores"], [7] * n, [0] * n, [6] * n)
        if case["expected"] is None:
            assert res["type"] == O.NONE
            continue
        e = case["expected"]
        assert res["type"] == O.GAPPED and res["status"] == 0
        assert O.result_score(res) == e["score"] and O.result_second(res) == e["second"]
        assert (res["t_start"], res["t_end"]) == (e["start"], e["end"])


def _run_ipt(primers, recs):
    p = K["ipt_params"]
    params = O.make_params(slop=p["slop"], max_mismatch_rate=p["max_mismatches"] / p["primer_length"], min_alignment_score_rate=0.0,
                           match_score=p["match_score"], mismatch_score=p["mismatch_score"], gap_open=p["gap_open"], gap_extend=p["gap_extend"])
    panel = O.Panel(primers, params)
    return panel, panel.process_batch(H.recs_to_arrays(recs), n_threads=2)


def test_ipt_two_primer_pairs_per_read():  # IdentifyPrimersTest.scala:238-297
    primers = H.primers_from(K["ipt_panel_two_pairs"])
    recs = H.pair_records(H.build_builder(primers))
    panel, (five, _, _, _) = _run_ipt(primers, recs)
    names = {O.NONE: "none", O.LOCATION: "Location", O.UNGAPPED: "Ungapped", O.GAPPED: "Gapped"}
    exp = K["ipt_per_read"]
    for i in range(0, len(recs), 2):
        r1, r2 = recs[i], recs[i + 1]
        m1, m2 = five[i], five[i + 1]
        pos = lambda m: None if m["type"] == O.NONE else primers[m["primer"]].positive_strand
        fwd_is_r1 = H.pair_tag_slots(pos(m1), pos(m2))
        fwd, rev = (m1, m2) if fwd_is_r1 else (m2, m1)
        for rec in (r1, r2):
            m = fwd if not rec.negative else rev       # `if (rec.positiveStrand) rec("f5") else rec("r5")`
            got = names[int(m["type"])]
            et = rec.tags.get("et")
            if et is None:
                want = "none"
            elif et == "no_edits":
                want = exp["no_edits"]
            else:
                kind, n = et.split("_"); n = int(n)
                want = {"mis": "Ungapped" if n < exp["mis_lt"] else "none", "ins": "Gapped" if n < exp["ins_lt"] else "none",
                        "del": "Gapped" if n <= exp["del_le"] else "none"}[kind]
            assert got == want, (et, rec.bases, got, want)


@pytest.mark.parametrize("primers_mapped", [True, False])
@pytest.mark.parametrize("reads_mapped", [True, False])
def test_ipt_end_to_end_metrics(primers_mapped, reads_mapped):  # IdentifyPrimersTest.scala:299-403
    primers = H.primers_from(K["ipt_panel"], mapped=primers_mapped)
    recs = H.ipt_metrics_reads(H.primers_from(K["ipt_panel"], mapped=True))
    assert len(recs) == 126
    if not reads_mapped:
        recs = H.make_unmapped(recs)
    panel, (five, _, metrics, n_align) = _run_ipt(primers, recs)
    got = O.summary_metrics(metrics)
    exp = dict(K["ipt_metrics"]["common"])
    for k, v in list(exp.items()):
        if isinstance(v, dict):
            exp[k] = v["mapped" if reads_mapped else "unmapped"]
    exp.update(K["ipt_metrics"]["mapped_primers_mapped_reads" if (primers_mapped and reads_mapped) else "otherwise"])
    assert got == exp
    assert got["match_attempts"] == got["location"] + got["ungapped"] + got["gapped"] + got["no_match"]
    assert not five["status"].any()


def test_java_format_2f():  # PrimerMatch.scala:115 / PrimerMatchTest
    assert O.java_format_2f(0.125) == "0.13"      # HALF_UP on the shortest repr, not C's banker/binary rounding
    assert O.java_format_2f(3 / 7.0) == "0.43"
    assert O.java_format_2f(0.0) == "0.00"
    assert O.java_format_2f(-1.005) == "-1.01"
    assert O.java_format_2f(12.0) == "12.00"
    assert O.java_format_2f(0.995) == "1.00"
    assert O.java_format_2f(99.995) == "100.00"
    assert O.java_format_2f(0.001) == "0.00"


def test_kmer_rule_follows_the_mismatch_position():
    """PM:108-116 from first principles (no GPU): the ungapped matcher only sees primers that share a k-mer STRING with the
    read window, so a degenerate primer (IUPAC letters match A/C/G/T reads by PM:143 but are different letters) is reported
    only when a run of k plain letters survives the read's mismatch.  One plain run of exactly k at positions 7..12."""
    k = 6
    seq = "RYRYRYR" + "ACGTCA" + "YRYRYRY"            # 20 letters: degenerate, plain run of 6, degenerate
    primers = [H.Primer("p", "p+", seq, True), H.Primer("p", "p-", "TTGACCATGACCATGACCAT", False)]
    panel = O.Panel(primers, O.make_params(max_mismatch_rate=0.05, k_ungapped=k, min_alignment_score_rate=1e9))
    inst = {"R": "A", "Y": "C"}
    recs, expect = [], []
    for pos in range(-1, len(seq)):
        s = [inst.get(ch, ch) for ch in seq]
        if pos >= 0:
            s[pos] = "T" if s[pos] != "T" and (seq[pos] not in "Y") else "G"   # a base the primer letter does not allow
            assert O.num_mismatches("".join(s).encode(), seq.encode(), 0, 5.0) == 1
        recs.append(H.Rec("".join(s) + "ACGT" * 10, True, False))
        expect.append(not (7 <= pos <= 12))               # the only literal 6-mer is gone when the mismatch sits inside it
    # a read that spells the ambiguity letters themselves is literally equal everywhere
    recs.append(H.Rec(seq + "ACGT" * 10, True, False))
    expect.append(True)
    five, _, _, _ = panel.process_batch(H.recs_to_arrays(recs), n_threads=1)
    got = [int(t) == 2 and int(p) == 0 for t, p in zip(five["type"], five["primer"])]   # 2 = ungapped match to p+
    assert got == expect


def test_three_prime_candidate_rules_from_first_principles():
    """matchThreePrime (IdentifyPrimers.scala:483-516), no GPU: which primers are aligned against the reverse complement of
    the WHOLE read (the target-length quirk, :501-505) in Local mode -- the mate's 5' primer if it has one, else the opposite
    strand primers of the read's own pair, else every primer -- on read-through amplicons where the answer is evident."""
    rng = np.random.default_rng(3)

    def rnd(n):
        return "".join("ACGT"[x] for x in rng.integers(0, 4, n))

    fwd = [rnd(22), rnd(24)]
    rev = [rnd(23), rnd(21)]
    primers = []
    for i in range(2):
        primers.append(H.Primer(f"a{i}", f"a{i}F", fwd[i], True))
        primers.append(H.Primer(f"a{i}", f"a{i}R", rev[i], False))
    panel = O.Panel(primers, O.make_params(three_prime=5))
    insert = rnd(30)
    r1 = fwd[0] + insert + H.revcomp(rev[0])                 # reads through the whole amplicon 0
    r2 = rev[0] + H.revcomp(insert) + H.revcomp(fwd[0])
    junk = rnd(len(r1))
    tail_only = rnd(40) + H.revcomp(rev[1])                  # no 5' primer, the reverse primer of amplicon 1 at its 3' end
    recs = [
        H.Rec(r1, True, paired=True, first=True, mate_next=True), H.Rec(r2, True, paired=True, second=True),      # both 5' ends match
        H.Rec(r1, True, paired=True, first=True, mate_next=True), H.Rec(junk, True, paired=True, second=True),    # only R1 matches
        H.Rec(tail_only, True, paired=True, first=True, mate_next=True), H.Rec(junk, True, paired=True, second=True),  # neither
        H.Rec(tail_only, True),                                                                                   # a fragment
    ]
    five, three, _, _ = panel.process_batch(H.recs_to_arrays(recs), n_threads=1)
    F0, R0, F1, R1 = 0, 1, 2, 3
    assert [int(p) for p in five["primer"]] == [F0, R0, F0, -1, -1, -1, -1]
    # template 1: each read is checked against its MATE's 5' primer only, found at the start of the reversed read
    assert (int(three["primer"][0]), int(three["raw_score"][0]), int(three["t_start"][0]), int(three["t_end"][0])) == (R0, len(rev[0]), 1, len(rev[0]))
    assert (int(three["primer"][1]), int(three["raw_score"][1]), int(three["t_start"][1]), int(three["t_end"][1])) == (F0, len(fwd[0]), 1, len(fwd[0]))
    assert int(three["multi"][0]) == 0 and int(three["multi"][1]) == 0
    # template 2: R1 has no mate match -> the opposite-strand primer of its own pair; R2 -> its mate's primer (F0), which a
    # random read only matches by chance: a short local hit, still reported (min rate 0.01)
    assert int(three["primer"][2]) == R0 and int(three["raw_score"][2]) == len(rev[0])
    assert int(three["primer"][3]) in (F0, -1) and int(three["raw_score"][3]) < 12
    # template 3 and the fragment: nothing matched at 5' -> every primer is aligned; R1 of amplicon 1 wins with a full score
    for i in (4, 6):
        assert (int(three["primer"][i]), int(three["raw_score"][i]), int(three["t_start"][i]), int(three["t_end"][i])) == (R1, len(rev[1]), 1, len(rev[1]))
        assert int(three["multi"][i]) == 1 and int(three["q_len"][i]) == len(rev[1])
