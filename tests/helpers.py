"""Shared test helpers: a tiny stand-in for fgbio's SamBuilder and the read factory of the reference's
IdentifyPrimersTest (IdentifyPrimersTest.scala:100-235), so its known-answer numbers can be checked without a JVM."""
from __future__ import annotations

import json
import os
import sys
from dataclasses import dataclass, field, replace
from typing import List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_PAIRED, F_UNMAPPED, F_REVERSE, F_FIRST, F_SECOND, F_FR_PAIR, F_MATE_NEXT = 0x1, 0x4, 0x10, 0x40, 0x80, 0x1000, 0x2000

# fgbio SamBuilder default dictionary: chr1..chr22, chrX, chrY, chrM
DEFAULT_DICT = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY", "chrM"]


def kats() -> dict:
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as fh:
        return json.load(fh)


@dataclass
class Primer:
    pair_id: str
    primer_id: str
    sequence: str
    positive_strand: bool
    ref_name: str = ""
    start: int = 0
    end: int = 0
    ref_idx: int = -1

    @property
    def length(self):  # Primer.scala:206
        return self.end - self.start + 1 if self.ref_name else len(self.sequence)


def primers_from(dicts, dict_names=DEFAULT_DICT, mapped=True) -> List[Primer]:
    out = []
    for d in dicts:
        ref = d["ref_name"] if mapped else ""
        idx = dict_names.index(ref) if ref in dict_names else (-1 if not ref else -2)
        out.append(Primer(d["pair_id"], d["primer_id"], d["sequence"].upper(), bool(d["positive_strand"]), ref, d["start"], d["end"], idx))
    return out


_COMP = {ord(a): ord(b) for a, b in zip("ACGTacgt", "TGCAtgca")}


def revcomp(s: str) -> str:
    """htsjdk 2.16 SequenceUtil.reverseComplement: only ACGTacgt are complemented."""
    return s[::-1].translate(_COMP)


@dataclass
class Rec:
    bases: str                      # as stored in the BAM record (genomic order when mapped)
    unmapped: bool = False
    negative: bool = False
    ref_idx: int = -1
    start: int = 0                  # alignment start; cigar is always <len>M here
    paired: bool = False
    first: bool = False
    second: bool = False
    mate_next: bool = False
    fr: bool = False
    tags: dict = field(default_factory=dict)

    @property
    def end(self):
        return self.start + len(self.bases) - 1

    @property
    def flags(self):
        f = 0
        if self.paired: f |= F_PAIRED
        if self.unmapped: f |= F_UNMAPPED
        if self.negative: f |= F_REVERSE
        if self.first: f |= F_FIRST
        if self.second: f |= F_SECOND
        if self.fr: f |= F_FR_PAIR
        if self.mate_next: f |= F_MATE_NEXT
        return f

    @property
    def seq_order(self):  # PrimerMatcher.scala:39-45
        return self.bases if (self.unmapped or not self.negative) else revcomp(self.bases)


def recs_to_arrays(recs: List[Rec]) -> dict:
    lens = np.array([len(r.bases) for r in recs], dtype=np.int64)
    off = np.zeros(len(recs) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    bases = np.frombuffer("".join(r.bases for r in recs).encode(), dtype=np.uint8).copy() if recs else np.zeros(0, np.uint8)
    return dict(bases=bases, off=off,
                ref_idx=np.array([(-1 if r.unmapped else r.ref_idx) for r in recs], dtype=np.int32),
                ustart=np.array([r.start for r in recs], dtype=np.int32),
                uend=np.array([r.end for r in recs], dtype=np.int32),
                flags=np.array([r.flags for r in recs], dtype=np.uint16))


# ---------------------------------------------------------------------------------------------------------------
# IdentifyPrimersTest read factory (IdentifyPrimersTest.scala:100-235)
# ---------------------------------------------------------------------------------------------------------------
def _add_ns_in_the_middle(seq, n, fwd):  # IdentifyPrimersTest.scala:178-188
    rem = len(seq) - n
    if rem <= 0:
        s = seq
    elif rem == 1:
        s = "N" + seq[1:] if fwd else seq[1:] + "N"
    else:
        s = seq[:(rem + 1) // 2] + "N" * n + (seq[-(rem // 2):] if rem // 2 else "")
    assert s.count("N") == n and len(s) == len(seq), (seq, n, s)
    return s


def _delete_in_the_middle(seq, n):  # :191-197
    rem = len(seq) - n
    assert rem > 0
    return seq[:(rem + 1) // 2] + (seq[-(rem // 2):] if rem // 2 else "")


def _insert_in_the_middle(seq, n):  # :200-204
    return seq[:(len(seq) + 1) // 2] + "N" * n + (seq[-(len(seq) // 2):] if len(seq) // 2 else "")


def from_primer(primer: Primer, edit: Optional[tuple] = None, dict_names=DEFAULT_DICT) -> Rec:  # :207-235
    new_ref = dict_names[-1]
    if edit is None:
        p, et = primer, "no_edits"
    elif edit[0] == "mis":
        p = replace(primer, ref_name=new_ref, sequence=_add_ns_in_the_middle(primer.sequence, edit[1], primer.positive_strand))
        et = f"mis_{edit[1]}"
    elif edit[0] == "ins":
        p = replace(primer, ref_name=new_ref, sequence=_insert_in_the_middle(primer.sequence, edit[1]), end=primer.end + edit[1])
        et = f"ins_{edit[1]}"
    else:
        p = replace(primer, ref_name=new_ref, sequence=_delete_in_the_middle(primer.sequence, edit[1]), start=primer.start + edit[1])
        et = f"del_{edit[1]}"
    assert p.length == len(p.sequence)  # the factory writes cigar <newPrimer.length>M over these bases
    bases = p.sequence if p.positive_strand else revcomp(p.sequence)  # basesInGenomicOrder (Primer.scala:220)
    return Rec(bases=bases, unmapped=False, negative=not p.positive_strand, ref_idx=dict_names.index(p.ref_name), start=p.start,
               tags={"et": et})


def build_builder(primers: List[Primer], read_length=10) -> List[Rec]:  # :100-135
    recs = [from_primer(p) for p in primers]
    for n in range(0, 3 + 2):  # Range.inclusive(0, maxMismatches+1)
        recs += [from_primer(p, ("mis", n)) for p in primers]
    for L in (1, 2):
        recs += [from_primer(p, ("del", L)) for p in primers]
        recs += [from_primer(p, ("ins", L)) for p in primers]
    recs.append(Rec("G" * read_length, False, False, 0, 100000))
    recs.append(Rec("C" * read_length, False, True, 0, 100000))
    recs.append(Rec("G" * read_length, True, False, 0, 100000))
    recs.append(Rec("C" * read_length, True, True, 0, 100000))
    return recs


def _insert_size(a: Rec, b: Rec) -> int:  # htsjdk SamPairUtil.computeInsertSize
    if a.unmapped or b.unmapped or a.ref_idx != b.ref_idx:
        return 0
    p1 = a.end if a.negative else a.start
    p2 = b.end if b.negative else b.start
    return p2 - p1 + (1 if p2 >= p1 else -1)


def _is_fr_pair(r: Rec, mate: Rec) -> bool:  # fgbio SamRecord.isFrPair + htsjdk SamPairUtil.getPairOrientation
    if not r.paired or r.unmapped or mate.unmapped or r.ref_idx != mate.ref_idx:
        return False
    if r.negative == mate.negative:
        return False  # TANDEM
    isize = _insert_size(r, mate)
    pos5 = mate.start if r.negative else r.start
    neg5 = r.end if r.negative else r.start + isize
    return pos5 < neg5


def pair_records(frags: List[Rec]) -> List[Rec]:  # :137-151
    assert len(frags) % 2 == 0
    out = []
    for i in range(0, len(frags), 2):
        r1, r2 = replace(frags[i]), replace(frags[i + 1])
        r1.paired = r2.paired = True
        r1.first, r2.second = True, True
        r1.mate_next = True
        r1.fr, r2.fr = _is_fr_pair(r1, r2), _is_fr_pair(r2, r1)
        out += [r1, r2]
    return out


def make_unmapped(recs: List[Rec]) -> List[Rec]:  # htsjdk SAMUtils.makeReadUnmapped (IdentifyPrimersTest.scala:311-315)
    out = []
    for r in recs:
        q = replace(r)
        if q.negative and not q.unmapped:
            q.bases = revcomp(q.bases)
        elif q.negative and q.unmapped:
            q.bases = revcomp(q.bases)  # makeReadUnmapped reverse-complements whenever the negative flag is set
        q.negative = False
        q.unmapped = True
        q.ref_idx, q.start, q.fr = -1, 0, False
        out.append(q)
    # mate information is not consulted by the path, but isFrPair needs a mapped read
    return out


def ipt_metrics_reads(primers: List[Primer]) -> List[Rec]:  # :153-169
    frags = build_builder(primers)
    frags += [from_primer(primers[0]), from_primer(primers[3])]
    return pair_records(frags)


# ---------------------------------------------------------------------------------------------------------------
# f5/r5 slot logic of addTagsToPair (IdentifyPrimers.scala:522-554), host-side formatting
# ---------------------------------------------------------------------------------------------------------------
def pair_tag_slots(r1_five_primer_positive: Optional[bool], r2_five_primer_positive: Optional[bool]):
    """Returns (forward_is_r1). Arguments are None when the read has no 5' match, else primer.positive_strand."""
    if r1_five_primer_positive is not None:
        return bool(r1_five_primer_positive)
    if r2_five_primer_positive is not None:
        return not r2_five_primer_positive  # m.primer.negativeStrand => (r1, r2)
    return True
