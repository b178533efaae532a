"""ctypes binding of the CPU oracle (oracle/idp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package (fg_idprimer_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("IDP_ORACLE_LIB") or os.path.join(_HERE, "libidp_oracle.so")  # override: the alternative-policy build
ALT_POLICY_FLAGS = ["-DIDP_POLICY_LEFT_FROM_UP=1", "-DIDP_POLICY_UP_FROM_LEFT=1", "-DIDP_POLICY_EXTEND_OVER_OPEN=1", "-DIDP_POLICY_TIE_UP_BEFORE_LEFT=1"]
ALT_LIB_PATH = os.path.join(_HERE, "libidp_oracle_altpolicy.so")

NEXT_NA = 2147483647
NONE, LOCATION, UNGAPPED, GAPPED = 0, 1, 2, 3
MODE_LOCAL, MODE_GLOCAL, MODE_GLOBAL = 0, 1, 3
F_PAIRED, F_UNMAPPED, F_REVERSE, F_FIRST, F_SECOND, F_FR_PAIR, F_MATE_NEXT = 0x1, 0x4, 0x10, 0x40, 0x80, 0x1000, 0x2000


class OrcPrimer(C.Structure):
    _fields_ = [("pair_id", C.c_char_p), ("primer_id", C.c_char_p), ("sequence", C.c_char_p),
                ("positive_strand", C.c_int), ("ref_name", C.c_char_p), ("start", C.c_int), ("end", C.c_int),
                ("ref_idx", C.c_int)]


class OrcParams(C.Structure):
    _fields_ = [("slop", C.c_int), ("three_prime", C.c_int), ("match_score", C.c_int), ("mismatch_score", C.c_int),
                ("gap_open", C.c_int), ("gap_extend", C.c_int), ("k_ungapped", C.c_int), ("k_gapped", C.c_int),
                ("max_mismatch_rate", C.c_double), ("min_alignment_score_rate", C.c_double),
                ("score_matrix", C.POINTER(C.c_int))]


class OrcReads(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("off", C.c_void_p), ("ref_idx", C.c_void_p),
                ("unclipped_start", C.c_void_p), ("unclipped_end", C.c_void_p), ("flags", C.c_void_p)]


class OrcAlignment(C.Structure):
    _fields_ = [("score", C.c_int32), ("query_start", C.c_int32), ("target_start", C.c_int32),
                ("query_end", C.c_int32), ("target_end", C.c_int32)]


RESULT_DTYPE = np.dtype([("primer", "<i4"), ("type", "u1"), ("status", "u1"), ("multi", "u1"), ("pad", "u1"),
                         ("mm", "<i4"), ("next_mm", "<i4"), ("raw_score", "<i4"), ("raw_next", "<i4"),
                         ("q_len", "<i2"), ("q_len_next", "<i2"), ("t_start", "<i2"), ("t_end", "<i2")])
assert RESULT_DTYPE.itemsize == 32


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (checker only; see oracle/Makefile)."""
    src = os.path.join(_HERE, "idp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "idp_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libidp_oracle.so"])
    return _LIB_PATH


def build_alt_policy(force: bool = False) -> str:
    """The oracle with every unpinned alignment choice flipped (idp_oracle.c, ALIGN_POLICY): tests/test_align_policy.py."""
    src = os.path.join(_HERE, "idp_oracle.c")
    if force or not os.path.exists(ALT_LIB_PATH) or os.path.getmtime(ALT_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "idp_oracle.h"))):
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-g", "-fPIC", "-std=gnu11", "-pthread", *ALT_POLICY_FLAGS, "-shared", "-o", ALT_LIB_PATH, src,
                               "-lm", "-lpthread"])
    return ALT_LIB_PATH


def use_fast_build() -> str:
    """bench.py's reference arm / cpu_baseline only: switch this process to a -O3 -march=native build of the same source
    (BASELINE.md section 2), compiled on the machine that runs it (the file name carries a hash of the CPU's flags, so a
    copy built elsewhere is never loaded).  The checker keeps the -O2 -g build."""
    global _LIB_PATH, _lib
    import hashlib
    try:
        with open("/proc/cpuinfo") as fh:
            flags = next((ln for ln in fh if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    tag = hashlib.sha1(flags.encode()).hexdigest()[:10]
    path = os.path.join(_HERE, f"libidp_oracle_fast_{tag}.so")
    src = os.path.join(_HERE, "idp_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "idp_oracle.h"))):
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O3", "-march=native", "-fPIC", "-std=gnu11", "-pthread", "-shared", "-o", path, src, "-lm", "-lpthread"])
    if path != _LIB_PATH:
        _LIB_PATH, _lib = path, None
    return path


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if os.path.basename(_LIB_PATH) == "libidp_oracle.so":
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_align_policy_bits.restype = C.c_int
        L.orc_panel_create.restype = C.c_void_p
        L.orc_panel_create.argtypes = [C.POINTER(OrcPrimer), C.c_int, C.POINTER(OrcParams)]
        L.orc_panel_destroy.argtypes = [C.c_void_p]
        L.orc_primer_length.argtypes = [C.c_void_p, C.c_int]
        L.orc_primer_scala_hash.argtypes = [C.c_void_p, C.c_int]
        L.orc_primer_scala_hash.restype = C.c_uint32
        L.orc_num_mismatches.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_double]
        L.orc_candidates_with_common_kmer.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_mismatch_align.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.orc_location_overlaps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_best_ungapped.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_best_gapped.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_align.argtypes = [C.POINTER(OrcParams), C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(OrcAlignment)]
        L.orc_process_batch.argtypes = [C.c_void_p, C.POINTER(OrcReads), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_result_score.argtypes = [C.c_void_p]
        L.orc_result_score.restype = C.c_double
        L.orc_result_second.argtypes = [C.c_void_p]
        L.orc_result_second.restype = C.c_double
        L.orc_metric_bin.argtypes = [C.c_int] * 4
        L.orc_summary_metrics.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_format_info.argtypes = [C.c_void_p, C.c_void_p, C.c_uint16, C.c_char_p, C.c_int]
        L.orc_java_format_2f.argtypes = [C.c_double, C.c_char_p, C.c_int]
        _lib = L
    return _lib


def make_params(slop=5, three_prime=-1, max_mismatch_rate=0.05, min_alignment_score_rate=0.01, match_score=1,
                mismatch_score=-4, gap_open=-6, gap_extend=-1, k_ungapped=None, k_gapped=None, score_matrix=None) -> OrcParams:
    """Defaults are IdentifyPrimers.scala:170-183."""
    p = OrcParams()
    p.slop, p.three_prime = slop, three_prime
    p.match_score, p.mismatch_score, p.gap_open, p.gap_extend = match_score, mismatch_score, gap_open, gap_extend
    p.k_ungapped = k_ungapped if k_ungapped else -1
    p.k_gapped = k_gapped if k_gapped else -1
    p.max_mismatch_rate, p.min_alignment_score_rate = max_mismatch_rate, min_alignment_score_rate
    if score_matrix is not None:
        sm = np.ascontiguousarray(score_matrix, dtype=np.int32)
        assert sm.shape == (128, 128)
        p._keep = sm
        p.score_matrix = sm.ctypes.data_as(C.POINTER(C.c_int))
    return p


class Panel:
    """Owns an orc_panel (the three matchers' immutable state, IdentifyPrimers.scala:223-234)."""

    def __init__(self, primers: Sequence, params: OrcParams):
        self.primers = list(primers)
        self.params = params
        arr = (OrcPrimer * max(1, len(self.primers)))()
        self._keep = []
        for i, p in enumerate(self.primers):
            strs = [str(p.pair_id).encode(), str(p.primer_id).encode(), str(p.sequence).encode(), str(p.ref_name).encode()]
            self._keep.append(strs)
            arr[i].pair_id, arr[i].primer_id, arr[i].sequence, arr[i].ref_name = strs
            arr[i].positive_strand = 1 if p.positive_strand else 0
            arr[i].start, arr[i].end = int(p.start), int(p.end)
            arr[i].ref_idx = int(getattr(p, "ref_idx", -1))
        self._arr = arr
        self.h = lib().orc_panel_create(arr, len(self.primers), C.byref(params))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_panel_destroy(self.h)
            self.h = None

    def primer_length(self, i): return lib().orc_primer_length(self.h, i)
    def scala_hash(self, i): return lib().orc_primer_scala_hash(self.h, i)

    def candidates(self, which: int, bases: bytes, k: int):
        out = np.zeros(len(self.primers) + 1, dtype=np.int32)
        n = lib().orc_candidates_with_common_kmer(self.h, which, bases, len(bases), k, out.ctypes.data, len(out))
        return out[:n].tolist()

    def mismatch_align(self, so_bases: bytes, primer: int):
        mm = C.c_int(0)
        ok = lib().orc_mismatch_align(self.h, so_bases, len(so_bases), primer, C.byref(mm))
        return (mm.value if ok else None)

    def location_overlaps(self, ref_idx, negative, ustart, uend):
        out = np.zeros(64, dtype=np.int32)
        n = lib().orc_location_overlaps(self.h, ref_idx, 1 if negative else 0, ustart, uend, out.ctypes.data, len(out))
        return out[:min(n, 64)].tolist()

    def best_ungapped(self, primers, mms):
        pr = np.asarray(primers, dtype=np.int32); mm = np.asarray(mms, dtype=np.int32)
        out = np.zeros(1, dtype=RESULT_DTYPE)
        lib().orc_best_ungapped(self.h, pr.ctypes.data, mm.ctypes.data, len(pr), out.ctypes.data)
        return out[0]

    def process_batch(self, reads: dict, n_threads: int = 1):
        """reads: dict(bases=uint8[], off=int64[n+1], ref_idx, ustart, uend int32[n], flags uint16[n])."""
        n = len(reads["flags"])
        rd = OrcReads()
        keep = {k: np.ascontiguousarray(reads[k], dtype=dt) for k, dt in
                [("bases", np.uint8), ("off", np.int64), ("ref_idx", np.int32), ("ustart", np.int32), ("uend", np.int32), ("flags", np.uint16)]}
        rd.bases, rd.off, rd.ref_idx = keep["bases"].ctypes.data, keep["off"].ctypes.data, keep["ref_idx"].ctypes.data
        rd.unclipped_start, rd.unclipped_end, rd.flags = keep["ustart"].ctypes.data, keep["uend"].ctypes.data, keep["flags"].ctypes.data
        five = np.zeros(n, dtype=RESULT_DTYPE)
        three = np.zeros(n, dtype=RESULT_DTYPE)
        three["primer"] = -1
        metrics = np.zeros(160, dtype=np.int64)
        n_align = np.zeros(1, dtype=np.int64)
        lib().orc_process_batch(self.h, C.byref(rd), n, five.ctypes.data, three.ctypes.data, metrics.ctypes.data, n_align.ctypes.data, n_threads)
        return five, three, metrics, int(n_align[0])

    def format_info(self, result, flags: int) -> str:
        r = np.array([result], dtype=RESULT_DTYPE)
        buf = C.create_string_buffer(1024)
        lib().orc_format_info(self.h, r.ctypes.data, flags, buf, 1024)
        return buf.value.decode()


def num_mismatches(bases: bytes, primer: bytes, start: int, max_mismatches: float) -> int:
    return lib().orc_num_mismatches(bases, len(bases), primer, len(primer), start, max_mismatches)


def best_gapped(primers, scores, qlens, tstarts, tends):
    a = [np.asarray(x, dtype=np.int32) for x in (primers, scores, qlens, tstarts, tends)]
    out = np.zeros(1, dtype=RESULT_DTYPE)
    lib().orc_best_gapped(*[x.ctypes.data for x in a], len(a[0]), out.ctypes.data)
    return out[0]


def result_score(result) -> float:
    r = np.array([result], dtype=RESULT_DTYPE)
    return lib().orc_result_score(r.ctypes.data)


def result_second(result) -> float:
    r = np.array([result], dtype=RESULT_DTYPE)
    return lib().orc_result_second(r.ctypes.data)


def align(params: OrcParams, mode: int, query: bytes, target: bytes):
    a = OrcAlignment()
    lib().orc_align(C.byref(params), mode, query, len(query), target, len(target), C.byref(a))
    return dict(score=a.score, query_start=a.query_start, target_start=a.target_start, query_end=a.query_end, target_end=a.target_end)


def summary_metrics(metrics160) -> dict:
    m = np.ascontiguousarray(metrics160, dtype=np.int64)
    out = np.zeros(17, dtype=np.int64)
    lib().orc_summary_metrics(m.ctypes.data, out.ctypes.data)
    names = ["templates", "pairs", "fragments", "mapped_pairs", "unpaired", "unmapped_pairs", "fr_pairs", "mapped_fragments",
             "unmapped_fragments", "paired_matches", "unpaired_matches", "no_paired_matches", "match_attempts", "location",
             "ungapped", "gapped", "no_match"]
    return dict(zip(names, out.tolist()))


def java_format_2f(v: float) -> str:
    buf = C.create_string_buffer(128)
    lib().orc_java_format_2f(v, buf, 128)
    return buf.value.decode()
