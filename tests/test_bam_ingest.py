"""Host ingest: uncompressed BAM alignment records -> the SoA batch, no GPU needed.
The records are built here byte by byte (SAMv1 section 4.2) from the same read factory that reproduces the reference's
IdentifyPrimersTest, so the packer's flags / unclipped ends / isFrPair can be checked against that known answer
(fr_pairs = 61 of 62 mapped pairs, IdentifyPrimersTest.scala:367)."""
import struct

import numpy as np
import pytest

import helpers as H
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import _lib as L

K = H.kats()
_CIGAR_OPS = "MIDNSHP=X"


def bam_record(name, flag, ref_id, pos1, cigar, bases, next_ref=-1, next_pos1=0, tlen=0, mapq=60):
    """One BAM alignment record (block_size-prefixed).  cigar: [(op_char, len)], pos1 / next_pos1 are 1-based."""
    seq = bytearray((len(bases) + 1) // 2)
    lib = L.lib()
    arr = np.zeros(len(seq), dtype=np.uint8)
    lib.idp_pack_bases(bases.encode(), len(bases), arr.ctypes.data)
    qname = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | _CIGAR_OPS.index(op)) for op, n in cigar)
    body = struct.pack("<iiBBHHHiiii", ref_id, pos1 - 1, len(qname), mapq, 0, len(cigar), flag, len(bases), next_ref, next_pos1 - 1, tlen)
    body += qname + cig + arr.tobytes() + bytes([30] * len(bases)) + b"XYZ\x05"   # one tag, to be skipped
    return struct.pack("<i", len(body)) + body


def recs_to_bam(recs):
    """helpers.Rec pairs/fragments (R1 carries mate_next) -> BAM bytes with htsjdk-style mate fields."""
    out = []
    i = 0
    t = 0
    while i < len(recs):
        r1 = recs[i]
        r2 = recs[i + 1] if r1.mate_next else None
        group = [(r1, r2)] + ([(r2, r1)] if r2 is not None else [])
        for r, mate in group:
            flag = (0x1 if r.paired else 0) | (0x4 if r.unmapped else 0) | (0x10 if r.negative else 0) | (0x40 if r.first else 0) | (0x80 if r.second else 0)
            nref, npos, tlen = -1, 0, 0
            if mate is not None:
                flag |= (0x8 if mate.unmapped else 0) | (0x20 if mate.negative else 0)
                if not mate.unmapped:
                    nref, npos = mate.ref_idx, mate.start
                if not r.unmapped and not mate.unmapped and r.ref_idx == mate.ref_idx:   # SamPairUtil.computeInsertSize
                    p1 = r.end if r.negative else r.start
                    p2 = mate.end if mate.negative else mate.start
                    tlen = p2 - p1 + (1 if p2 >= p1 else -1)
            out.append(bam_record(f"t{t}", flag, -1 if r.unmapped else r.ref_idx, 0 if r.unmapped else r.start,
                                  [] if r.unmapped else [("M", len(r.bases))], r.bases, nref, npos, tlen))
        i += 2 if r2 is not None else 1
        t += 1
    return b"".join(out)


def test_bam_packer_reproduces_the_factory_batch():
    primers = H.primers_from(K["ipt_panel"], mapped=True)
    recs = H.ipt_metrics_reads(primers)
    bam = recs_to_bam(recs)
    batch, rec_index = idp.ReadBatch.from_bam(bam)
    want = H.recs_to_arrays(recs)
    want_batch = idp.ReadBatch.from_ascii(want["bases"], want["off"], want["flags"], want["ref_idx"], want["ustart"], want["uend"])
    assert rec_index.tolist() == list(range(len(recs)))
    assert np.array_equal(batch.seq4, want_batch.seq4) and np.array_equal(batch.seq_off, want_batch.seq_off) and np.array_equal(batch.len, want_batch.len)
    assert np.array_equal(batch.ref_idx, want["ref_idx"])
    mapped = want["ref_idx"] >= 0
    assert np.array_equal(batch.unclipped_start[mapped], want["ustart"][mapped]) and np.array_equal(batch.unclipped_end[mapped], want["uend"][mapped])
    assert np.array_equal(batch.flags, want["flags"])
    # the reference's own number: 61 FR pairs among the 62 mapped pairs (IdentifyPrimersTest.scala:364-367)
    r1 = (batch.flags & H.F_MATE_NEXT) != 0
    assert int(((batch.flags & H.F_FR_PAIR) != 0)[r1].sum()) == K["ipt_metrics"]["common"]["fr_pairs"]["mapped"]


def test_bam_packer_clips_order_and_skipped_records():
    recs = b"".join([
        # R2 arrives before R1, soft/hard clips on both sides, a supplementary and a secondary record in between
        bam_record("q1", 0x1 | 0x80 | 0x10 | 0x20 ^ 0x20, 2, 500, [("H", 3), ("S", 2), ("M", 10), ("I", 1), ("D", 4), ("M", 5), ("S", 6)], "ACGTACGTACGTACGTACGTACGT", 2, 300, -250),
        bam_record("q1", 0x1 | 0x40 | 0x800, 2, 900, [("M", 5)], "ACGTA"),
        bam_record("q1", 0x1 | 0x40 | 0x20, 2, 300, [("S", 4), ("M", 20)], "TTTTACGTACGTACGTACGTACGT", 2, 500, 250),
        bam_record("q1", 0x1 | 0x40 | 0x100, 3, 100, [("M", 5)], "ACGTA"),
        # an unmapped fragment and an unpaired mapped fragment
        bam_record("q2", 0x4, -1, 0, [], "NNNNACGT"),
        bam_record("q3", 0x10, 0, 77, [("M", 9)], "ACGTNACGT"),
    ])
    batch, rec_index = idp.ReadBatch.from_bam(recs)
    assert rec_index.tolist() == [2, 0, 4, 5]                      # R1 first, secondary / supplementary skipped
    assert batch.len.tolist() == [24, 24, 8, 9]
    assert batch.ref_idx.tolist() == [2, 2, -1, 0]
    assert (batch.unclipped_start.tolist()[0], batch.unclipped_end.tolist()[0]) == (296, 319)       # 300-4 .. 300+20-1
    assert (batch.unclipped_start.tolist()[1], batch.unclipped_end.tolist()[1]) == (495, 524)       # 500-5 .. 500+19-1+6
    assert (batch.unclipped_start.tolist()[3], batch.unclipped_end.tolist()[3]) == (77, 85)
    f = batch.flags.tolist()
    assert f[0] == (H.F_PAIRED | H.F_FIRST | H.F_MATE_NEXT | H.F_FR_PAIR)
    assert f[1] == (H.F_PAIRED | H.F_SECOND | H.F_REVERSE | H.F_FR_PAIR)
    assert f[2] == H.F_UNMAPPED and f[3] == H.F_REVERSE
    out = np.zeros(24, dtype=np.uint8)
    import ctypes as C
    buf = C.create_string_buffer(24)
    L.lib().idp_unpack_bases(batch.seq4.ctypes.data + int(batch.seq_off[0]), 24, buf)
    assert buf.raw == b"TTTTACGTACGTACGTACGTACGT"


def test_bam_packer_rejects_malformed_input():
    good = bam_record("q", 0x4, -1, 0, [], "ACGT")
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(good[:-3])
    two_r1 = bam_record("q", 0x1 | 0x40 | 0x4, -1, 0, [], "ACGT") * 2
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(two_r1)
    only_r2 = bam_record("q", 0x1 | 0x80 | 0x4, -1, 0, [], "ACGT")
    with pytest.raises(idp.IdpError) as e:
        idp.ReadBatch.from_bam(only_r2)
    assert "did not have an R1" in str(e.value)            # IdentifyPrimers.scala:411
