"""CPU-side checks of the product library: it loads, exports every symbol include/idprimer.h declares, and its
host-only logic (panel build, Scala iteration order, tag formatting, packing, metric roll-up) agrees with the oracle.
No compute entry point is called here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
import fg_idprimer_b200 as idp
from fg_idprimer_b200 import _lib as L
from oracle import oracle_py as O

K = H.kats()


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    header = open(os.path.join(H.ROOT, "include", "idprimer.h")).read()
    declared = set(re.findall(r"\b(idp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.idp_abi_version() == 2


def test_no_gpu_means_loud_failure_not_fallback():
    if L.lib().idp_device_count() > 0:
        pytest.skip("a GPU is present")
    primers = [idp.Primer("1", "1+", "ACGTACGTAC", True), idp.Primer("1", "1-", "TTGCATTGCA", False)]
    with pytest.raises(idp.IdpError) as e:
        idp.IdentifyPrimers(primers, device=0)
    assert e.value.code == L.ERR_CUDA and "no CPU fallback" in str(e.value)


def _tool(primers, **kw):
    ps = [idp.Primer(p.pair_id, p.primer_id, p.sequence, p.positive_strand, p.ref_name, p.start, p.end) for p in primers]
    return idp.IdentifyPrimers(ps, ref_names=H.DEFAULT_DICT, device=None, **kw)


def test_primer_length_and_scala_hash_match_oracle():
    primers = H.primers_from(K["pmt_panel"]) + [H.Primer("u", "u+", "ACGTNRY", True), H.Primer("u", "u-", "GGGTTTAA", False)]
    tool = _tool(primers)
    panel = O.Panel(primers, O.make_params())
    for i in range(len(primers)):
        assert tool.primer_length(i) == panel.primer_length(i) == primers[i].length
        assert tool.primer_hash(i) == panel.scala_hash(i)


def test_kmer_posting_order_matches_oracle():
    rng = np.random.default_rng(7)
    # many primers sharing k-mers so that posting lists cross the 4 -> 5 element (Set4 -> HashSet) boundary
    core = "ACGTTGCAAGGCTTAC"
    primers = []
    for i in range(40):
        tail = "".join("ACGT"[x] for x in rng.integers(0, 4, 8))
        primers.append(H.Primer(f"p{i}", f"p{i}+", core[: 8 + i % 8] + tail, True))
        primers.append(H.Primer(f"p{i}", f"p{i}-", tail + core[i % 5:], False))
    for k in (4, 6, 10):
        tool = _tool(primers, ungapped_kmer_length=k, gapped_kmer_length=k)
        panel = O.Panel(primers, O.make_params(k_ungapped=k, k_gapped=k))
        kmers = {p.sequence[i:i + k] for p in primers for i in range(len(p.sequence) - k + 1)}
        sizes = set()
        for km in sorted(kmers):
            got = tool.kmer_postings(0, km)
            # the oracle lists a single k-mer's primers when the "read" is exactly that k-mer
            want = panel.candidates(0, km.encode(), k)
            assert got == want, (k, km)
            assert got == tool.kmer_postings(1, km)
            sizes.add(len(got))
        if k == 4:
            assert max(sizes) >= 5 and 2 in sizes


def test_format_info_and_java_rounding_match_oracle():
    primers = H.primers_from(K["pmt_panel"])
    tool = _tool(primers)
    panel = O.Panel(primers, O.make_params())
    rows = np.zeros(6, dtype=L.RESULT_DTYPE)
    rows["primer"] = [-1, 0, 3, 3, 5, 7]
    rows["type"] = [0, 1, 2, 2, 3, 3]
    rows["mm"] = [0, 2, 1, 0, 0, 0]
    rows["next_mm"] = [0, 0, 3, L.NEXT_NA, 0, 0]
    rows["multi"] = [0, 0, 0, 0, 1, 0]
    rows["raw_score"] = [0, 0, 0, 0, 1, 7]
    rows["raw_next"] = [0, 0, 0, 0, -3, 0]
    rows["q_len"] = [0, 0, 0, 0, 8, 10]
    rows["q_len_next"] = [0, 0, 0, 0, 24, 0]
    rows["t_start"] = [0, 0, 0, 0, 1, 2]
    rows["t_end"] = [0, 0, 0, 0, 9, 11]
    for row in rows:
        for fl in (0, H.F_PAIRED | H.F_FIRST, H.F_PAIRED | H.F_SECOND):
            assert tool.info(row, fl) == panel.format_info(row.astype(O.RESULT_DTYPE), fl)
    assert tool.info(rows[4], 0) == "3,3-,chr2:1001-1010,-,1,Gapped,0.13,-0.13,1,9"   # 1/8 = 0.125 -> HALF_UP
    assert tool.info(rows[3], H.F_PAIRED | H.F_SECOND) == "2,2-,chr2:101-110,-,2,Ungapped,0,na"


def test_pack_unpack_roundtrip_and_numpy_packer():
    lib = L.lib()
    s = b"ACGTNMRWSYKVHDB=ACGTA"
    buf = np.zeros((len(s) + 1) // 2, dtype=np.uint8)
    lib.idp_pack_bases(s, len(s), buf.ctypes.data)
    out = C.create_string_buffer(len(s))
    lib.idp_unpack_bases(buf.ctypes.data, len(s), out)
    assert out.raw == s
    recs = [H.Rec("ACGTA"), H.Rec("NNRYACG"), H.Rec(""), H.Rec("T")]
    arr = H.recs_to_arrays(recs)
    seq4, boff, lens = idp.ReadBatch.pack(arr["bases"], arr["off"])
    assert lens.tolist() == [5, 7, 0, 1] and boff.tolist() == [0, 3, 7, 7]
    for r, o, n in zip(recs, boff, lens):
        out = C.create_string_buffer(int(n) + 1)
        lib.idp_unpack_bases(seq4.ctypes.data + int(o), int(n), out)
        assert out.raw[:n].decode() == r.bases


def test_summary_metrics_match_oracle():
    rng = np.random.default_rng(3)
    m = rng.integers(0, 50, 160).astype(np.int64)
    out = np.zeros(17, dtype=np.int64)
    L.lib().idp_summary_metrics(m.ctypes.data, out.ctypes.data)
    assert out.tolist() == list(O.summary_metrics(m).values())
    for tt in range(5):
        for fr in range(2):
            for a in range(4):
                for b in range(4):
                    assert L.lib().idp_metric_bin(tt, fr, a, b) == O.lib().orc_metric_bin(tt, fr, a, b)


def test_panel_validation_errors():
    good = [idp.Primer("1", "1+", "ACGTACGTAC", True), idp.Primer("1", "1-", "TTGCATTGCA", False)]
    with pytest.raises(idp.IdpError):
        idp.IdentifyPrimers(good, max_mismatch_rate=-0.1, device=None)          # IdentifyPrimers.scala:216
    with pytest.raises(idp.IdpError):
        idp.IdentifyPrimers(good, gap_open=1, device=None)                      # :219
    with pytest.raises(idp.IdpError):
        idp.IdentifyPrimers(good, ungapped_kmer_length=17, device=None)         # GPU path limit
    with pytest.raises(ValueError):
        idp.Primer("1,2", "x", "ACGT", True)                                    # Primer.scala:190
    with pytest.raises(ValueError):
        idp.Primer("1", "x", "ACGU", True)                                      # Primer.scala:194-200
    with pytest.raises(ValueError):
        idp.Primer.validate_primers([good[0], good[0]], False)                  # Primer.scala:146-150


def test_downstream_classification_matches_reference_script():
    """scripts/post_process_bam.py:56-78, golden vectors produced by the reference's own function (tests/golden/gen_classify_golden.py)."""
    import json
    g = json.load(open(os.path.join(H.ROOT, "tests", "golden", "classify_golden.json")))
    primers = [idp.Primer(p["pair_id"], p["primer_id"], p["sequence"], p["positive_strand"], p["ref_name"], p["start"], p["end"]) for p in g["primers"]]
    tool = idp.IdentifyPrimers(primers, ref_names=["chr1", "chr2", "chr3"], device=None)
    for c in g["cases"]:
        assert tool.classify_primer_pair(c["f5"], c["r5"], c["min_insert"], c["max_insert"]) == c["expected"], c
    # batch walk: slots of addTagsToPair (IdentifyPrimers.scala:530-540), fragments, counting only read-1 records
    five = np.zeros(7, dtype=L.RESULT_DTYPE)
    five["primer"] = [1, 0, -1, 3, 5, -1, 2]          # pair (p0_R, p0_F), pair (none, p1_R), fragment p2_R, pair (none, p1_F)
    five["type"] = [2, 2, 0, 2, 2, 0, 2]
    flags = np.array([H.F_PAIRED | H.F_FIRST | H.F_MATE_NEXT, H.F_PAIRED | H.F_SECOND, H.F_PAIRED | H.F_FIRST | H.F_MATE_NEXT, H.F_PAIRED | H.F_SECOND,
                      0, H.F_PAIRED | H.F_FIRST | H.F_MATE_NEXT, H.F_PAIRED | H.F_SECOND], dtype=np.uint16)
    cls, counts = tool.classify_templates(flags, five)
    names = [tool.CLASS_NAMES[c] for c in cls]
    assert names == ["Canonical", "Canonical", "Single", "Single", "Single", "Single", "Single"]
    assert counts == {"Canonical": 1, "SelfDimer": 0, "CrossDimer": 0, "NonCanonical": 0, "Single": 2, "NoMatch": 0}
    tool.close()
