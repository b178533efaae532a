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


def bam_record(name, flag, ref_id, pos1, cigar, bases, next_ref=-1, next_pos1=0, tlen=0, mapq=60, aux=b"XYZq\0"):
    """One BAM alignment record (block_size-prefixed).  cigar: [(op_char, len)], pos1 / next_pos1 are 1-based."""
    seq = bytearray((len(bases) + 1) // 2)
    lib = L.lib()
    arr = np.zeros(len(seq), dtype=np.uint8)
    lib.idp_pack_bases(bases.encode(), len(bases), arr.ctypes.data)
    qname = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (n << 4) | _CIGAR_OPS.index(op)) for op, n in cigar)
    body = struct.pack("<iiBBHHHiiii", ref_id, pos1 - 1, len(qname), mapq, 0, len(cigar), flag, len(bases), next_ref, next_pos1 - 1, tlen)
    body += qname + cig + arr.tobytes() + bytes([30] * len(bases)) + aux   # one tag by default, skipped by the ingest
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
    import struct
    body = bytearray(good[4:])
    # a record whose l_seq says more bases + qualities than block_size holds (the packed bases alone would still fit), and one
    # whose name + CIGAR already overrun it: both are rejected before anything is read past the record
    long_seq = bytearray(body); struct.pack_into("<i", long_seq, 16, 10)
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(struct.pack("<i", len(long_seq)) + bytes(long_seq))
    many_ops = bytearray(body); struct.pack_into("<H", many_ops, 12, 4000)
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(struct.pack("<i", len(many_ops)) + bytes(many_ops))
    neg = bytearray(body); struct.pack_into("<i", neg, 16, -5)
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(struct.pack("<i", len(neg)) + bytes(neg))
    two_r1 = bam_record("q", 0x1 | 0x40 | 0x4, -1, 0, [], "ACGT") * 2
    with pytest.raises(idp.IdpError):
        idp.ReadBatch.from_bam(two_r1)
    only_r2 = bam_record("q", 0x1 | 0x80 | 0x4, -1, 0, [], "ACGT")
    with pytest.raises(idp.IdpError) as e:
        idp.ReadBatch.from_bam(only_r2)
    assert "did not have an R1" in str(e.value)            # IdentifyPrimers.scala:411


# ---- host egress: the tags written back into the records (IdentifyPrimers.scala:522-569) ----

def _aux_fields(rec_body, l_seq, n_cigar, l_name):
    """(tag, type, value) of a record's aux area."""
    a = 32 + l_name + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
    out = []
    while a < len(rec_body):
        tag, ty = rec_body[a:a + 2].decode(), chr(rec_body[a + 2])
        a += 3
        if ty == "Z":
            e = rec_body.index(b"\0", a)
            out.append((tag, ty, rec_body[a:e].decode()))
            a = e + 1
        else:
            n = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[ty]
            out.append((tag, ty, rec_body[a:a + n]))
            a += n
    return out


def _records(bam):
    off, out = 0, []
    while off < len(bam):
        (bs,) = struct.unpack_from("<i", bam, off)
        body = bam[off + 4: off + 4 + bs]
        l_name, n_cigar, flag, l_seq = body[8], struct.unpack_from("<H", body, 12)[0], struct.unpack_from("<H", body, 14)[0], struct.unpack_from("<i", body, 16)[0]
        out.append(dict(name=body[32:32 + l_name - 1].decode(), flag=flag, body=body, aux=_aux_fields(body, l_seq, n_cigar, l_name)))
        off += 4 + bs
    assert off == len(bam)
    return out


def _rec_with_aux(name, flag, bases, aux=b"XYZhello\0NMC\x05"):
    return bam_record(name, flag, -1, 0, [], bases, aux=aux)


def _row(**kw):
    r = np.zeros(1, dtype=L.RESULT_DTYPE)
    r["primer"] = -1
    for k, v in kw.items():
        r[k] = v
    return r[0]


@pytest.mark.parametrize("three_prime", [-1, 3])
def test_bam_annotate_sets_the_tags_like_addTagsToPair_and_addTagsToFragment(three_prime):
    primers = H.primers_from(K["ipt_panel"], mapped=True)
    tool = idp.IdentifyPrimers(primers, ref_names=H.DEFAULT_DICT, device=None, three_prime=three_prime)
    bam = b"".join([
        _rec_with_aux("a", 0x1 | 0x80 | 0x4, "ACGTAC"),                                # R2 arrives first
        _rec_with_aux("a", 0x1 | 0x40 | 0x4 | 0x800, "AC"),                            # supplementary: passes through untouched
        _rec_with_aux("a", 0x1 | 0x40 | 0x4, "ACGTACGT", aux=b"f5Zstale\0XYZhello\0"),  # R1, already carrying an f5 tag
        _rec_with_aux("b", 0x4, "ACGTAC"),                                             # fragment
        _rec_with_aux("c", 0x1 | 0x40 | 0x4, "ACGT"), _rec_with_aux("c", 0x1 | 0x80 | 0x4, "ACGT"),   # nothing matched
    ])
    batch, rec_index = idp.ReadBatch.from_bam(bam)
    assert rec_index.tolist() == [2, 0, 3, 4, 5]
    neg = next(i for i, p in enumerate(primers) if not p.positive_strand)
    pos = next(i for i, p in enumerate(primers) if p.positive_strand)
    five = np.array([_row(primer=neg, type=L.UNGAPPED, mm=1, next_mm=3),           # R1 matched a NEGATIVE primer: R2 is "forward"
                     _row(primer=pos, type=L.LOCATION, mm=0),
                     _row(primer=pos, type=L.GAPPED, raw_score=7, raw_next=4, q_len=10, q_len_next=10, t_start=1, t_end=10, multi=1),
                     _row(), _row()], dtype=L.RESULT_DTYPE)
    three = np.array([_row(), _row(primer=neg, type=L.GAPPED, raw_score=5, raw_next=0, q_len=10, q_len_next=0, t_start=2, t_end=9),
                      _row(primer=neg, type=L.GAPPED, raw_score=6, raw_next=0, q_len=10, q_len_next=0, t_start=3, t_end=8), _row(), _row()],
                     dtype=L.RESULT_DTYPE)
    out = tool.annotate_bam(bam, five, three if three_prime >= 0 else None)
    recs = _records(out)
    assert [(r["name"], r["flag"] & 0x8C0) for r in recs] == [("a", 0x40), ("a", 0x80), ("a", 0x840), ("b", 0), ("c", 0x40), ("c", 0x80)]  # allReads order
    want_pair = tool.pair_tags(int(batch.flags[0]), int(batch.flags[1]), five[0], five[1], three[0], three[1])
    names = ["f5", "r5"] + (["f3", "r3"] if three_prime >= 0 else [])
    for r in recs[:2]:
        ours = [(t, v) for t, ty, v in r["aux"] if t in ("f5", "r5", "f3", "r3")]
        assert ours == [(t, want_pair[t]) for t in names]
        assert ("XY", "Z", "hello") in r["aux"] and not any(v == "stale" for _, _, v in r["aux"])
    assert want_pair["f5"].split(",")[1] == primers[pos].primer_id and want_pair["f5"].split(",")[4] == "2"   # the forward slot is R2's match
    assert recs[2]["aux"] == [("XY", "Z", "hello"), ("NM", "C", b"\x05")]
    want_frag = tool.fragment_tags(int(batch.flags[2]), five[2], three[2])
    assert [(t, v) for t, ty, v in recs[3]["aux"] if t[0] in "fr"] == [("f5", want_frag["f5"]), ("r5", "none")]
    assert ("Gapped" in want_frag["f5"]) and (",3,8" in want_frag["f5"]) == (three_prime >= 0)    # the 3' match overwrites f5 (IP:565-568)
    for r in recs[4:]:
        assert [(t, v) for t, ty, v in r["aux"] if t[0] in "fr"] == [(t, "none") for t in names]
    tool.close()


def test_bam_annotate_rejects_mismatched_inputs_and_header_comments():
    primers = H.primers_from(K["ipt_panel"], mapped=True)
    tool = idp.IdentifyPrimers(primers, ref_names=H.DEFAULT_DICT, device=None)
    bam = _rec_with_aux("b", 0x4, "ACGTAC")
    with pytest.raises(idp.IdpError):
        tool.annotate_bam(bam, np.zeros(2, dtype=L.RESULT_DTYPE))       # one read, two result rows
    with pytest.raises(idp.IdpError):
        tool.annotate_bam(bam[:-2], np.zeros(1, dtype=L.RESULT_DTYPE))
    co = tool.header_comments()
    assert len(co) == 3 and co[0].startswith("The f5/r5/f3/r3 tags store") and "<pair_id>,<primer_id>,<ref_name>:<start>-<end>,<strand>,<read-num>" in co[1]
    tool.close()


# ------------------------------------------------------------------------------------------------------------------
# windowed ingest (idp_bam_to_batch_window): 5'-only runs ship only the bases the chain can reach
# ------------------------------------------------------------------------------------------------------------------
def _unpack(batch, i):
    import ctypes as C
    n = int(batch.len[i])
    buf = C.create_string_buffer(max(n, 1))
    L.lib().idp_unpack_bases(batch.seq4.ctypes.data + int(batch.seq_off[i]), n, buf)
    return buf.raw[:n].decode()


def _window_panel(rng, n_pairs, min_len, max_len):
    primers = []
    for i in range(n_pairs):
        for positive in (True, False):
            n = int(rng.integers(min_len, max_len + 1))
            seq = "".join("ACGT"[x] for x in rng.integers(0, 4, n))
            start = 1000 * (i + 1) + (0 if positive else 300)
            primers.append(H.Primer(f"p{i}", f"p{i}{'F' if positive else 'R'}", seq, positive, f"chr{1 + i % 5}", start, start + n - 1, i % 5))
    return primers


def _window_reads(rng, primers, n_reads, read_len_choices):
    """Reads that start (in sequencing order) with a primer carrying 0-3 edits or an indel, forward / reverse-strand mapped
    near the primer or unmapped, of several lengths (odd and even, shorter and longer than the window)."""
    recs = []
    for j in range(n_reads):
        p = primers[int(rng.integers(0, len(primers)))]
        s = list(p.sequence)
        for _ in range(int(rng.integers(0, 4))):
            s[int(rng.integers(0, len(s)))] = "ACGTN"[int(rng.integers(0, 5))]
        if rng.random() < 0.25:
            k = int(rng.integers(2, len(s) - 2))
            s = s[:k] + (["ACGT"[int(rng.integers(0, 4))]] if rng.random() < 0.5 else []) + s[k + (1 if rng.random() < 0.5 else 0):]
        n = int(rng.choice(read_len_choices))
        so = ("".join(s) + "".join("ACGT"[x] for x in rng.integers(0, 4, 200)))[:n]   # sequencing order
        kind = j % 4
        if kind == 0:
            recs.append(H.Rec(so, unmapped=True))
        elif kind == 1:
            recs.append(H.Rec(so, unmapped=True, negative=True))                 # the strand bit of an unmapped read is ignored (PM:41)
        elif p.positive_strand:
            recs.append(H.Rec(so, False, False, p.ref_idx, p.start + int(rng.integers(-6, 7))))
        else:
            end = p.end + int(rng.integers(-6, 7))
            recs.append(H.Rec(H.revcomp(so), False, True, p.ref_idx, end - n + 1))
    return recs


@pytest.mark.parametrize("max_len", [25, 26])   # windows of 30 and 31 bases: even and odd
def test_windowed_ingest_keeps_the_reachable_bases(max_len):
    rng = np.random.default_rng(5)
    primers = _window_panel(rng, 6, 18, max_len)
    tool = idp.IdentifyPrimers(primers, ref_names=H.DEFAULT_DICT, device=None, ungapped_kmer_length=6, gapped_kmer_length=10)
    w = tool.window()
    assert w == max(len(p.sequence) for p in primers) + 5                        # sequence + slop dominates here
    recs = _window_reads(rng, primers, 400, [12, 29, 30, 31, 32, 33, 70, 71, 150])
    bam = recs_to_bam(recs)
    full, rec_full = idp.ReadBatch.from_bam(bam)
    cut, rec_cut = idp.ReadBatch.from_bam(bam, window_for=tool)
    assert np.array_equal(rec_full, rec_cut)
    for a in ("ref_idx", "unclipped_start", "unclipped_end", "flags"):
        assert np.array_equal(getattr(full, a), getattr(cut, a)), a
    assert np.array_equal(cut.len, np.minimum(full.len, w))
    assert cut.seq4.nbytes < full.seq4.nbytes
    for i, r in enumerate(recs):
        stored = r.bases
        want = stored if len(stored) <= w else (stored[-w:] if (r.negative and not r.unmapped) else stored[:w])
        assert _unpack(cut, i) == want, (i, r)
        if len(want) % 2:                                                        # the spare nibble is zero, as in a BAM record
            assert cut.seq4[int(cut.seq_off[i]) + len(want) // 2] & 0x0F == 0
    tool.close()
    with idp.IdentifyPrimers(primers, ref_names=H.DEFAULT_DICT, device=None, three_prime=5) as t3:   # 3' matching reads the whole read
        assert t3.window() == -1
        with pytest.raises(idp.IdpError):
            idp.ReadBatch.from_bam(bam, window_for=t3)


@pytest.mark.parametrize("opts", [dict(), dict(ungapped_kmer_length=6, gapped_kmer_length=10), dict(ungapped_kmer_length=4, slop=2, max_mismatch_rate=0.1),
                                  dict(gapped_kmer_length=8, slop=9)])
def test_windowed_batch_gives_the_same_results(opts):
    """The oracle (= the reference's chain) on the whole reads and on the windowed batch: every 5' result field, the metric
    counter and the number of alignments are identical, so shipping the window is exact."""
    from oracle import oracle_py as O
    rng = np.random.default_rng(11)
    primers = _window_panel(rng, 12, 18, 27)
    primers[3] = H.Primer(primers[3].pair_id, primers[3].primer_id, primers[3].sequence, primers[3].positive_strand, primers[3].ref_name,
                          primers[3].start, primers[3].start + 33, primers[3].ref_idx)   # Primer.length 34 > its sequence (PR:206)
    recs = _window_reads(rng, primers, 1200, [8, 20, 33, 34, 35, 39, 40, 41, 75, 150])
    for i in range(0, len(recs) - 1, 3):                                        # some templates are pairs
        recs[i].paired = recs[i + 1].paired = True
        recs[i].first, recs[i + 1].second, recs[i].mate_next = True, True, True
    tool = idp.IdentifyPrimers(primers, ref_names=H.DEFAULT_DICT, device=None, **opts)
    bam = recs_to_bam(recs)
    full, _ = idp.ReadBatch.from_bam(bam)
    cut, _ = idp.ReadBatch.from_bam(bam, window_for=tool)
    w = tool.window()
    assert w >= 34 + 0 and int(cut.len.max()) == w
    params = O.make_params(slop=opts.get("slop", 5), max_mismatch_rate=opts.get("max_mismatch_rate", 0.05),
                           k_ungapped=opts.get("ungapped_kmer_length"), k_gapped=opts.get("gapped_kmer_length"))
    panel = O.Panel(primers, params)

    def arrays(batch):
        bases = "".join(_unpack(batch, i) for i in range(len(batch.len)))
        off = np.zeros(len(batch.len) + 1, dtype=np.int64)
        off[1:] = np.cumsum(batch.len)
        return dict(bases=np.frombuffer(bases.encode(), dtype=np.uint8).copy(), off=off, ref_idx=batch.ref_idx, ustart=batch.unclipped_start,
                    uend=batch.unclipped_end, flags=batch.flags)

    want5, _, want_metrics, want_align = panel.process_batch(arrays(full), n_threads=4)
    got5, _, got_metrics, got_align = panel.process_batch(arrays(cut), n_threads=4)
    for f in want5.dtype.names:
        assert np.array_equal(got5[f], want5[f]), f
    assert np.array_equal(got_metrics, want_metrics) and got_align == want_align
    types = set(int(t) for t in want5["type"])
    assert {0, 1, 2, 3} <= types                                                 # none, location, ungapped and gapped all occur
    tool.close()
