#!/usr/bin/env python
"""Extracts the reference's known-answer vectors from its ScalaTest sources into reference_kats.json.

Run in the build container only (``python tests/golden/extract_kats.py``): it reads /root/reference, which does
not exist on the GPU box.  The JSON it writes is committed and is what the tests read.  The reference's tests
cannot be executed here (no JVM, fgbio/htsjdk un-vendored), so the vectors are lifted textually, each with the
file:line it came from.
"""
import json
import os
import re

REF = "/root/reference/src/test/scala/com/fulcrumgenomics/identifyprimers"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")

PRIMER_RE = re.compile(r'Primer\("([^"]*)",\s*"([^"]*)",\s*"([^"]*)",\s*(true|false)(?:,\s*"([^"]*)",\s*(\d+),\s*(\d+))?\)')


def numbered(path):
    with open(path) as fh:
        return list(enumerate(fh.read().split("\n"), start=1))


def primers_between(lines, lo, hi, src):
    out = []
    for no, text in lines:
        if lo <= no <= hi:
            for m in PRIMER_RE.finditer(text):
                out.append(dict(pair_id=m.group(1), primer_id=m.group(2), sequence=m.group(3), positive_strand=m.group(4) == "true",
                                ref_name=m.group(5) or "", start=int(m.group(6) or 0), end=int(m.group(7) or 0), src=f"{src}:{no}"))
    return out


def main():
    kats = {}
    pmt = numbered(os.path.join(REF, "PrimerMatcherTest.scala"))
    ipt = numbered(os.path.join(REF, "IdentifyPrimersTest.scala"))

    # --- PrimerMatcherTest: shared 12-primer panel
    lo = next(no for no, t in pmt if "val primers = IndexedSeq(" in t)
    hi = next(no for no, t in pmt if no > lo and t.strip() == ")")
    kats["pmt_panel"] = primers_between(pmt, lo, hi, "PrimerMatcherTest.scala")
    assert len(kats["pmt_panel"]) == 12

    # --- numMismatches truth table
    rows = []
    for no, t in pmt:
        m = re.search(r'TestCase\("([^"]*)",\s*"([^"]*)",\s*(\d+),\s*([0-9.]+)d?\)', t)
        if m and "val testCase" not in t:
            rows.append(dict(read=m.group(1), primer=m.group(2), expected=int(m.group(3)), max_mismatches=float(m.group(4)),
                             start=0, src=f"PrimerMatcherTest.scala:{no}"))
    for no, t in pmt:
        m = re.search(r'tool\.numMismatches\(testCase\.bases, testCase\.primer, (\d+), ([0-9.]+)d?\) shouldBe (\d+)', t)
        if m:
            rows.append(dict(read="TTAAA", primer="AAAAA", expected=int(m.group(3)), max_mismatches=float(m.group(2)),
                             start=int(m.group(1)), src=f"PrimerMatcherTest.scala:{no}"))
    assert len(rows) == 18, len(rows)
    kats["num_mismatches"] = rows

    # --- mismatchAlign primers (maxMismatchRate 0.25)
    lo = next(no for no, t in pmt if 'val primer      = Primer(' in t)
    named = {}
    for no, t in pmt:
        if lo <= no <= lo + 6:
            m = re.search(r'val (\w+)\s*= (Primer\(.*\))', t)
            if m:
                named[m.group(1)] = primers_between([(no, m.group(2))], no, no, "PrimerMatcherTest.scala")[0]
    assert set(named) == {"primer", "mmPrimer", "shortPrimer", "longPrimer", "mmShort", "mmLong"}
    kats["mismatch_align"] = dict(max_mismatch_rate=0.25, primers=named,
                                  # PrimerMatcherTest.scala:200-217: read bases in sequencing order are AAAAAAAAAA in all three cases
                                  read_seq_order="AAAAAAAAAA", minus_read_stored="TTTTTTTTTT",
                                  expected=dict(primer=0, mmPrimer=1, shortPrimer=0, longPrimer=0, mmShort=1, mmLong=None),
                                  src="PrimerMatcherTest.scala:187-218")

    # --- IdentifyPrimersTest: 12-primer panel, 4-primer panel, expected metrics
    lo = next(no for no, t in ipt if "private val primers = IndexedSeq(" in t)
    hi = next(no for no, t in ipt if no > lo and t.strip() == ")")
    kats["ipt_panel"] = primers_between(ipt, lo, hi, "IdentifyPrimersTest.scala")
    assert len(kats["ipt_panel"]) == 12
    lo = next(no for no, t in ipt if "val inputPrimers = Seq(" in t)
    kats["ipt_panel_two_pairs"] = primers_between(ipt, lo, lo + 5, "IdentifyPrimersTest.scala")
    assert len(kats["ipt_panel_two_pairs"]) == 4

    def const(name):
        for no, t in ipt:
            m = re.search(rf'private val {name}\s*=\s*(-?\d+)', t)
            if m:
                return int(m.group(1))
        raise KeyError(name)
    kats["ipt_params"] = dict(primer_length=const("primerLength"), max_mismatches=const("maxMismatches"), slop=const("slop"),
                              match_score=const("matchScore"), mismatch_score=const("mismatchScore"), gap_open=const("gapOpen"),
                              gap_extend=const("gapExtend"), min_alignment_score_rate=0.0, src="IdentifyPrimersTest.scala:39-53,264-276")
    common = {}
    for no, t in ipt:
        m = re.match(r'\s*(\w+)\s*=\s*(\d+),?\s*$', t)
        if m and 345 <= no <= 378:
            common[m.group(1)] = int(m.group(2))
        m = re.match(r'\s*(\w+)\s*=\s*if \(readsAreMapped\) (\d+) else (\d+),', t)
        if m:
            common[m.group(1)] = dict(mapped=int(m.group(2)), unmapped=int(m.group(3)))
    both, other = {}, {}
    block = None
    for no, t in ipt:
        if "if (primersAreMapped && readsAreMapped)" in t:
            block = both
        elif block is both and t.strip() == "else {":
            block = other
        m = re.match(r'\s*(location|ungapped|gapped)\s*=\s*(\d+),?\s*$', t)
        if m and block is not None:
            block[m.group(1)] = int(m.group(2))
    assert both == dict(location=14, ungapped=78, gapped=4) and other == dict(location=0, ungapped=92, gapped=4), (both, other)
    kats["ipt_metrics"] = dict(common=common, mapped_primers_mapped_reads=both, otherwise=other, src="IdentifyPrimersTest.scala:353-397")
    # per-read expectations of the two-pair run (IdentifyPrimersTest.scala:282-296)
    kats["ipt_per_read"] = dict(no_edits="Location", mis_lt=3, ins_lt=2, del_le=2, src="IdentifyPrimersTest.scala:282-296")

    # --- getBestUngappedAlignment / getBestGappedAlignment (PrimerMatcherTest.scala:322-340, 396-422)
    kats["best_ungapped"] = [
        dict(mms=[], expected=None, src="PrimerMatcherTest.scala:322-325"),
        dict(mms=[0], expected=dict(mm=0, next=2147483647), src="PrimerMatcherTest.scala:327-330"),
        dict(mms=[0, 0], expected=dict(mm=0, next=0), src="PrimerMatcherTest.scala:332-335"),
        dict(mms=[0, 1, 3, 1], expected=dict(mm=0, next=1), src="PrimerMatcherTest.scala:337-340"),
    ]
    kats["best_gapped"] = [  # Alignment("GATTACA","GATTACA",0,0,"7M",score): targetStart 0, targetEnd 6
        dict(scores=[], expected=None, src="PrimerMatcherTest.scala:405-407"),
        dict(scores=[0], expected=dict(score=0.0, second=0.0, start=0, end=6), src="PrimerMatcherTest.scala:409-411"),
        dict(scores=[0, 0], expected=dict(score=0.0, second=0.0, start=0, end=6), src="PrimerMatcherTest.scala:413-415"),
        dict(scores=[0, 1, 3, 1], expected=dict(score=3 / 7.0, second=1 / 7.0, start=0, end=6), src="PrimerMatcherTest.scala:417-420"),
    ]
    # --- k-mer lookups (PrimerMatcherTest.scala:89-145), kmerLength 5 over the pmt panel
    kats["kmer"] = dict(k=5, cases=[
        dict(bases="", n=0, src="PrimerMatcherTest.scala:93"),
        dict(bases="AAAA", n=0, src="PrimerMatcherTest.scala:94"),
        dict(bases="AAAAA", n=3, must_contain="AAAAA", src="PrimerMatcherTest.scala:97-101"),
        dict(bases="AAAAAAAAAA", n=3, must_contain="AAAAA", src="PrimerMatcherTest.scala:104-108"),
        dict(bases="GATTA", n=2, must_contain="GATTA", src="PrimerMatcherTest.scala:111-115"),
    ], tasks=[
        dict(stored="AAAAAAAAAAGGGGGGGGGG", unmapped=True, negative=False, ids=["1+", "2-", "3-"], src="PrimerMatcherTest.scala:119-124"),
        dict(stored="A" * 20, unmapped=False, negative=False, ids=["1+", "2-", "3-"], src="PrimerMatcherTest.scala:133-136"),
        dict(stored="A" * 20, unmapped=False, negative=True, ids=["1-", "2+", "3+"], src="PrimerMatcherTest.scala:140-143"),
    ])
    with open(OUT, "w") as fh:
        json.dump(kats, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
