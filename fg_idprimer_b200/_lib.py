"""ctypes view of include/idprimer.h.  Loading fails loudly when the in-tree CUDA library is missing: there is no
Python or CPU fallback for the matching path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDP_LIB_PATH") or os.path.join(HERE, "libidprimer_b200.so")  # override: A/B builds of the same ABI

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, 1, 2, 3, 4
F_PAIRED, F_UNMAPPED, F_REVERSE, F_FIRST, F_SECOND, F_FR_PAIR, F_MATE_NEXT = 0x1, 0x4, 0x10, 0x40, 0x80, 0x1000, 0x2000
NONE, LOCATION, UNGAPPED, GAPPED = 0, 1, 2, 3
NEXT_NA = 2147483647
MODE_LOCAL, MODE_GLOCAL, MODE_GLOBAL = 0, 1, 3

# every symbol include/idprimer.h declares
EXPORTS = ["idp_abi_version", "idp_last_error", "idp_device_count", "idp_align_policy", "idp_panel_create", "idp_panel_destroy", "idp_panel_size",
           "idp_panel_primer_length", "idp_panel_primer_hash", "idp_panel_kmer_postings", "idp_ctx_create", "idp_ctx_destroy",
           "idp_ctx_set_stream", "idp_ctx_timings", "idp_ctx_set_async", "idp_ctx_sync", "idp_host_alloc", "idp_host_free", "idp_match_batch", "idp_match_batch_device",
           "idp_match_batch16", "idp_match_batch16_device", "idp_result16_unpack", "idp_align_batch", "idp_result_score", "idp_result_second", "idp_metric_bin", "idp_summary_metrics", "idp_format_info",
           "idp_pack_bases", "idp_unpack_bases", "idp_classify_primer_pair", "idp_classify_templates", "idp_classify_batch_device", "idp_ctx_set_classify", "idp_bam_scan", "idp_bam_to_batch", "idp_bam_to_batch_window", "idp_panel_window",
           "idp_bam_annotate", "idp_header_comments"]


class IdpParams(C.Structure):
    _fields_ = [("slop", C.c_int32), ("three_prime", C.c_int32), ("match_score", C.c_int32), ("mismatch_score", C.c_int32),
                ("gap_open", C.c_int32), ("gap_extend", C.c_int32), ("k_ungapped", C.c_int32), ("k_gapped", C.c_int32),
                ("max_mismatch_rate", C.c_double), ("min_alignment_score_rate", C.c_double), ("score_matrix", C.c_void_p)]


class IdpPrimer(C.Structure):
    _fields_ = [("pair_id", C.c_char_p), ("primer_id", C.c_char_p), ("sequence", C.c_char_p), ("ref_name", C.c_char_p),
                ("ref_idx", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("positive_strand", C.c_uint8)]


class IdpReads(C.Structure):
    _fields_ = [("seq4", C.c_void_p), ("seq_off", C.c_void_p), ("len", C.c_void_p), ("ref_idx", C.c_void_p),
                ("unclipped_start", C.c_void_p), ("unclipped_end", C.c_void_p), ("flags", C.c_void_p), ("fixed_len", C.c_int32)]


class IdpTimings(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_scan", C.c_float), ("ms_gapped", C.c_float), ("ms_three_prime", C.c_float),
                ("ms_sw_score", C.c_float), ("n_reads", C.c_int64), ("n_fallthrough", C.c_int64), ("n_alignments_5p", C.c_int64),
                ("n_alignments_3p", C.c_int64), ("sw_cells", C.c_int64), ("scan_base_compares", C.c_int64),
                ("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("n_scan_parked", C.c_int64), ("n_scan_deferred", C.c_int64)]


RESULT_DTYPE = np.dtype([("primer", "<i4"), ("type", "u1"), ("status", "u1"), ("multi", "u1"), ("reserved", "u1"),
                         ("mm", "<i4"), ("next_mm", "<i4"), ("raw_score", "<i4"), ("raw_next", "<i4"),
                         ("q_len", "<i2"), ("q_len_next", "<i2"), ("t_start", "<i2"), ("t_end", "<i2")])
assert RESULT_DTYPE.itemsize == 32
# idp_result16: head = primer + 1 (bits 0..23) | type << 24 | multi << 26 | status << 27
RESULT16_DTYPE = np.dtype([("head", "<u4"), ("a", "<i2"), ("b", "<i2"), ("q_len", "<u2"), ("q_len_next", "<u2"),
                           ("t_start", "<u2"), ("t_end", "<u2")])
assert RESULT16_DTYPE.itemsize == 16
NEXT16_NA = 32767


def unpack16(rows: np.ndarray) -> np.ndarray:
    """idp_result16 rows -> idp_result rows (idp_result16_unpack)."""
    rows = np.ascontiguousarray(rows, dtype=RESULT16_DTYPE)
    out = np.zeros(len(rows), dtype=RESULT_DTYPE)
    lib().idp_result16_unpack(rows.ctypes.data, len(rows), out.ctypes.data)
    return out

_lib = None


class IdpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"idprimer error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m fg_idprimer_b200.build` "
                          "(the IdentifyPrimers GPU path has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.idp_abi_version.restype = C.c_int
    L.idp_last_error.restype = C.c_char_p
    L.idp_device_count.restype = C.c_int
    L.idp_panel_create.argtypes = [C.POINTER(IdpPrimer), C.c_int32, C.POINTER(IdpParams), C.POINTER(C.c_void_p)]
    L.idp_panel_destroy.argtypes = [C.c_void_p]
    L.idp_panel_size.argtypes = [C.c_void_p]
    L.idp_panel_primer_length.argtypes = [C.c_void_p, C.c_int32]
    L.idp_panel_primer_hash.argtypes = [C.c_void_p, C.c_int32]
    L.idp_panel_primer_hash.restype = C.c_uint32
    L.idp_panel_kmer_postings.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int32]
    L.idp_ctx_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    L.idp_ctx_destroy.argtypes = [C.c_void_p]
    L.idp_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.idp_ctx_timings.argtypes = [C.c_void_p, C.POINTER(IdpTimings)]
    L.idp_ctx_set_async.argtypes = [C.c_void_p, C.c_int]
    L.idp_ctx_sync.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.idp_host_alloc.argtypes = [C.c_size_t]
    L.idp_host_alloc.restype = C.c_void_p
    L.idp_host_free.argtypes = [C.c_void_p]
    for fn in (L.idp_match_batch, L.idp_match_batch_device, L.idp_match_batch16, L.idp_match_batch16_device):
        fn.argtypes = [C.c_void_p, C.POINTER(IdpReads), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.idp_result16_unpack.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.idp_result16_unpack.restype = None
    L.idp_align_batch.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.idp_result_score.argtypes = [C.c_void_p]
    L.idp_result_score.restype = C.c_double
    L.idp_result_second.argtypes = [C.c_void_p]
    L.idp_result_second.restype = C.c_double
    L.idp_metric_bin.argtypes = [C.c_int] * 4
    L.idp_summary_metrics.argtypes = [C.c_void_p, C.c_void_p]
    L.idp_format_info.argtypes = [C.c_void_p, C.c_void_p, C.c_uint16, C.c_char_p, C.c_int32]
    L.idp_classify_primer_pair.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.idp_classify_templates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.idp_classify_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.idp_ctx_set_classify.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.idp_bam_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.idp_bam_to_batch.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 8
    L.idp_bam_to_batch_window.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 8
    L.idp_panel_window.argtypes = [C.c_void_p]
    L.idp_panel_window.restype = C.c_int32
    L.idp_bam_annotate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    L.idp_header_comments.argtypes = [C.c_char_p, C.c_int32]
    L.idp_pack_bases.argtypes = [C.c_char_p, C.c_int32, C.c_void_p]
    L.idp_unpack_bases.argtypes = [C.c_void_p, C.c_int32, C.c_char_p]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        raise IdpError(rc, lib().idp_last_error().decode(errors="replace"))
