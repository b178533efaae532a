"""Host-side mirror of the reference's interface for the matching path, over the C-ABI in include/idprimer.h.

Names follow the Scala they stand in for (no JVM exists in this image, see INTEGRATION.md for the JNI binding):
  Primer                               -> identifyprimers/Primer.scala:182-221 (+ read/validate, :54-165)
  LocationBasedPrimerMatch, ...        -> identifyprimers/PrimerMatch.scala:96-117
  IdentifyPrimers(...).process_batch   -> identifyprimers/IdentifyPrimers.scala:165-186 (options), :381-436 (processBatch)
  AsyncCudaAligner.align               -> alignment/AsyncAligner.scala:37-54
All matching runs in the CUDA library; this module only packs inputs and decodes results.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib as L

_VALID_BASES = set("ACGTNacgtn")                      # htsjdk SequenceUtil.isValidBase
_IUPAC = set(".aAbBcCdDgGhHkKmMnNrRsStTvVwWyY")       # htsjdk SequenceUtil.isIUPAC
_NIBBLE_LUT = np.full(256, 15, dtype=np.uint8)        # htsjdk bytesToCompressedBases: anything unknown is N
for _i, _c in enumerate("=ACMGRSVTWYHKDBN"):
    _NIBBLE_LUT[ord(_c)] = _i
    _NIBBLE_LUT[ord(_c.lower())] = _i


@dataclass(frozen=True)
class Primer:
    """case class Primer (Primer.scala:182-221)."""
    pair_id: str
    primer_id: str
    sequence: str
    positive_strand: bool
    ref_name: str = ""
    start: int = 0
    end: int = 0

    def __post_init__(self):
        if "," in self.pair_id:
            raise ValueError(f"pair_id contained a comma: '{self.pair_id}'")
        if "," in self.primer_id:
            raise ValueError(f"primer_id contained a comma: '{self.pair_id}'")
        for i, b in enumerate(self.sequence):
            if b not in _VALID_BASES and b not in _IUPAC:
                raise ValueError(f"Found an invalid base for primer {self.pair_id}/{self.primer_id}: "
                                 f"{self.sequence[:i]}[{b}]{self.sequence[i + 1:]}")

    @property
    def length(self) -> int:          # Primer.scala:206
        return self.end - self.start + 1 if self.ref_name else len(self.sequence)

    @property
    def mapped(self) -> bool:
        return bool(self.ref_name)

    @property
    def negative_strand(self) -> bool:
        return not self.positive_strand

    @staticmethod
    def is_positive_strand(strand: str) -> bool:  # Primer.scala:115-119
        s = strand.lower()
        if s in ("+", "f", "fwd", "for", "forward", "positive", "true"):
            return True
        if s in ("-", "r", "rev", "reverse", "negative", "false"):
            return False
        raise ValueError(f"Could not parse strand: {strand}")

    @staticmethod
    def read(path, multi_primer_pairs: bool = False) -> List["Primer"]:  # Primer.scala:54-112
        with open(path) as fh:
            lines = fh.read().split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        headers = lines[0].split("\t") if lines else []
        required = ["pair_id", "primer_id", "sequence", "strand", "ref_name", "start", "end"]
        out = []
        for idx, line in enumerate(lines[1:]):
            fields = line.split("\t")
            row = dict(zip(headers, fields))

            def need(cols):
                for h in cols:
                    if h not in row and not (h == "strand" and "positive_strand" in row):
                        raise ValueError(f"Missing value for {h} on line {idx + 1}")
            strand = lambda: Primer.is_positive_strand(row["strand"] if "strand" in row else row["positive_strand"])
            if len(fields) == 4:
                need(required[:4])
                out.append(Primer(row["pair_id"], row["primer_id"], row["sequence"].upper(), strand()))
            elif len(fields) == len(headers):
                need(required)
                to_int = lambda s: 0 if s == "" else int(s)
                out.append(Primer(row["pair_id"], row["primer_id"], row["sequence"].upper(), strand(), row["ref_name"],
                                  to_int(row["start"]), to_int(row["end"])))
            else:
                raise ValueError(f"Incorrect number of values in line: {idx + 1}.")
        Primer.validate_primers(out, multi_primer_pairs)
        return out

    @staticmethod
    def write(path, primers: Iterable["Primer"]) -> None:
        with open(path, "w") as fh:
            fh.write("pair_id\tprimer_id\tsequence\tpositive_strand\tref_name\tstart\tend\n")
            for p in primers:
                fh.write(f"{p.pair_id}\t{p.primer_id}\t{p.sequence}\t{'true' if p.positive_strand else 'false'}\t{p.ref_name}\t{p.start}\t{p.end}\n")

    @staticmethod
    def validate_primers(primers: Sequence["Primer"], multi_primer_pairs: bool) -> None:  # Primer.scala:128-165
        groups = {}
        for p in primers:
            groups.setdefault(p.pair_id, []).append(p)
        for pair_id, ps in groups.items():
            ids = ", ".join(p.primer_id for p in ps)
            if multi_primer_pairs:
                if len(ps) < 2:
                    raise ValueError(f"Found {len(ps)} primers for pair with id '{pair_id}': {ids}")
                if not any(p.positive_strand for p in ps) or all(p.positive_strand for p in ps):
                    tpe = "forward" if not any(p.positive_strand for p in ps) else "reverse"
                    raise ValueError(f"Found two {tpe} primers for pair with id '{pair_id}': {ids}")
            else:
                if len(ps) != 2:
                    raise ValueError(f"Found {len(ps)} primers for pair with id '{pair_id}': {ids}")
                if ps[0].positive_strand == ps[1].positive_strand:
                    tpe = "forward" if ps[0].positive_strand else "reverse"
                    raise ValueError(f"Found two {tpe} primers for pair with id '{pair_id}': {ids}")
        mapped = sorted((p for p in primers if p.mapped), key=lambda p: (p.positive_strand, p.ref_name, p.start, p.end))
        for a, b in zip(mapped, mapped[1:]):
            if (a.ref_name, a.start, a.end, a.positive_strand) == (b.ref_name, b.start, b.end, b.positive_strand):
                raise ValueError(f"Found multiple primers with the same coordinates: {a} and {b}")


# ---- PrimerMatch.scala:96-117 ----
@dataclass(frozen=True)
class LocationBasedPrimerMatch:
    primer: Primer
    num_mismatches: int
    to_name = "Location"


@dataclass(frozen=True)
class UngappedAlignmentPrimerMatch:
    primer: Primer
    num_mismatches: int
    next_num_mismatches: int
    to_name = "Ungapped"

    def __post_init__(self):
        if not self.num_mismatches <= self.next_num_mismatches:
            raise ValueError("requirement failed")  # PrimerMatch.scala:102


@dataclass(frozen=True)
class GappedAlignmentPrimerMatch:
    primer: Primer
    score: float
    second_best_score: float
    start: int
    end: int
    to_name = "Gapped"

    def __post_init__(self):
        if not self.score >= self.second_best_score:
            raise ValueError(f"requirement failed: second best score ({self.second_best_score}) must be less than or equal to the score ({self.score})")
        if not self.start <= self.end:
            raise ValueError("requirement failed")


class ReadBatch:
    """Structure-of-arrays batch in the layout idp_reads expects (BAM-native 4-bit bases)."""

    def __init__(self, seq4, flags, seq_off=None, length=None, ref_idx=None, unclipped_start=None, unclipped_end=None, fixed_len=0):
        self.seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
        self.flags = np.ascontiguousarray(flags, dtype=np.uint16)
        self.fixed_len = int(fixed_len)
        self.seq_off = None if seq_off is None else np.ascontiguousarray(seq_off, dtype=np.uint64)
        self.len = None if length is None else np.ascontiguousarray(length, dtype=np.int32)
        self.ref_idx = None if ref_idx is None else np.ascontiguousarray(ref_idx, dtype=np.int32)
        self.unclipped_start = None if unclipped_start is None else np.ascontiguousarray(unclipped_start, dtype=np.int32)
        self.unclipped_end = None if unclipped_end is None else np.ascontiguousarray(unclipped_end, dtype=np.int32)

    def __len__(self):
        return len(self.flags)

    @staticmethod
    def pack(bases: np.ndarray, off: np.ndarray):
        """ASCII bases (uint8, concatenated) + n+1 offsets -> (seq4, byte offsets, lengths); every read starts on a byte."""
        off = np.asarray(off, dtype=np.int64)
        lens = np.diff(off)
        nbytes = (lens + 1) // 2
        boff = np.zeros(len(lens) + 1, dtype=np.int64)
        boff[1:] = np.cumsum(nbytes)
        nib = _NIBBLE_LUT[np.asarray(bases, dtype=np.uint8)]
        padded = np.zeros(int(boff[-1]) * 2, dtype=np.uint8)
        if len(nib):
            read_of = np.repeat(np.arange(len(lens)), lens)
            dest = 2 * boff[read_of] + (np.arange(len(nib)) - off[read_of])
            padded[dest] = nib
        seq4 = (padded[0::2] << 4) | padded[1::2]
        return seq4.astype(np.uint8), boff[:-1].astype(np.uint64), lens.astype(np.int32)

    @classmethod
    def from_ascii(cls, bases, off, flags, ref_idx=None, unclipped_start=None, unclipped_end=None, window: int = 0) -> "ReadBatch":
        """window > 0 (IdentifyPrimers.window() of a tool without --three-prime): keep only that many bases of every read --
        the first ones, or the last ones of a mapped reverse-strand read, whose 5' end is the end of the stored bases
        (PM:39-45) -- as idp_bam_to_batch_window does."""
        if window > 0:
            bases = np.asarray(bases, dtype=np.uint8)
            off = np.asarray(off, dtype=np.int64)
            fl = np.asarray(flags, dtype=np.uint16)
            lens = np.diff(off)
            keep = np.minimum(lens, window)
            from_end = ((fl & L.F_UNMAPPED) == 0) & ((fl & L.F_REVERSE) != 0)
            start = off[:-1] + np.where(from_end, lens - keep, 0)
            new_off = np.zeros(len(lens) + 1, dtype=np.int64)
            new_off[1:] = np.cumsum(keep)
            read_of = np.repeat(np.arange(len(lens)), keep)
            src = start[read_of] + (np.arange(int(new_off[-1])) - new_off[read_of])
            bases, off = bases[src], new_off
        seq4, boff, lens = cls.pack(bases, off)
        return cls(seq4, flags, boff, lens, ref_idx, unclipped_start, unclipped_end)

    @classmethod
    def from_bam(cls, bam: bytes, window_for: "IdentifyPrimers" = None):
        """Uncompressed, queryname-grouped BAM alignment records -> (batch, rec_index); see idp_bam_to_batch.
        window_for: a tool without --three-prime; every read is then cut to the bases its 5' chain can reach
        (idp_bam_to_batch_window), which gives the same results from a much smaller batch."""
        lib = L.lib()
        buf = np.frombuffer(bam, dtype=np.uint8)
        n = C.c_int32(0)
        nbytes = C.c_uint64(0)
        L.check(lib.idp_bam_scan(buf.ctypes.data, len(buf), C.byref(n), C.byref(nbytes)))
        n = n.value
        seq4 = np.zeros(nbytes.value, dtype=np.uint8)
        seq_off = np.zeros(n, dtype=np.uint64)
        length, ref_idx, us, ue, rec = (np.zeros(n, dtype=np.int32) for _ in range(5))
        flags = np.zeros(n, dtype=np.uint16)
        if window_for is not None:
            L.check(lib.idp_bam_to_batch_window(window_for._panel, buf.ctypes.data, len(buf), seq4.ctypes.data, seq_off.ctypes.data, length.ctypes.data,
                                                ref_idx.ctypes.data, us.ctypes.data, ue.ctypes.data, flags.ctypes.data, rec.ctypes.data))
            used = int(seq_off[-1] + (length[-1] + 1) // 2) if n else 0
            seq4 = seq4[:used]
        else:
            L.check(lib.idp_bam_to_batch(buf.ctypes.data, len(buf), seq4.ctypes.data, seq_off.ctypes.data, length.ctypes.data, ref_idx.ctypes.data,
                                         us.ctypes.data, ue.ctypes.data, flags.ctypes.data, rec.ctypes.data))
        return cls(seq4, flags, seq_off, length, ref_idx, us, ue), rec

    def as_struct(self, keep: list) -> L.IdpReads:
        r = L.IdpReads()
        ptr = lambda a: None if a is None else a.ctypes.data
        r.seq4, r.seq_off, r.len = ptr(self.seq4), ptr(self.seq_off), ptr(self.len)
        r.ref_idx, r.unclipped_start, r.unclipped_end = ptr(self.ref_idx), ptr(self.unclipped_start), ptr(self.unclipped_end)
        r.flags, r.fixed_len = ptr(self.flags), self.fixed_len
        keep.append(self)
        return r


class IdentifyPrimers:
    """The option surface of the tool that reaches the hot path (IdentifyPrimers.scala:170-185) and processBatch."""

    def __init__(self, primers: Sequence[Primer], slop: int = 5, three_prime: int = -1, max_mismatch_rate: float = 0.05,
                 min_alignment_score_rate: float = 0.01, match_score: int = 1, mismatch_score: int = -4, gap_open: int = -6,
                 gap_extend: int = -1, ungapped_kmer_length: Optional[int] = None, gapped_kmer_length: Optional[int] = None,
                 scoring_matrix: Optional[np.ndarray] = None, ref_names: Optional[Sequence[str]] = None, device: Optional[int] = 0):
        lib = L.lib()
        self.primers = list(primers)
        self._keep = []
        arr = (L.IdpPrimer * max(1, len(self.primers)))()
        ref_names = list(ref_names) if ref_names is not None else []
        for i, p in enumerate(self.primers):
            strs = [p.pair_id.encode(), p.primer_id.encode(), p.sequence.encode(), p.ref_name.encode()]
            self._keep.append(strs)
            arr[i].pair_id, arr[i].primer_id, arr[i].sequence, arr[i].ref_name = strs
            explicit = getattr(p, "ref_idx", None)
            arr[i].ref_idx = explicit if explicit is not None else (ref_names.index(p.ref_name) if p.ref_name in ref_names else -1)
            arr[i].start, arr[i].end, arr[i].positive_strand = p.start, p.end, 1 if p.positive_strand else 0
        prm = L.IdpParams()
        prm.slop, prm.three_prime = slop, three_prime
        prm.match_score, prm.mismatch_score, prm.gap_open, prm.gap_extend = match_score, mismatch_score, gap_open, gap_extend
        prm.k_ungapped = ungapped_kmer_length if ungapped_kmer_length else -1
        prm.k_gapped = gapped_kmer_length if gapped_kmer_length else -1
        prm.max_mismatch_rate, prm.min_alignment_score_rate = max_mismatch_rate, min_alignment_score_rate
        if scoring_matrix is not None:
            sm = np.ascontiguousarray(scoring_matrix, dtype=np.int32)
            assert sm.shape == (128, 128)
            self._keep.append(sm)
            prm.score_matrix = sm.ctypes.data
        self.params = prm
        self.three_prime = three_prime
        self._panel = C.c_void_p()
        self._ctx = C.c_void_p()
        L.check(lib.idp_panel_create(arr, len(self.primers), C.byref(prm), C.byref(self._panel)))
        if device is not None:
            L.check(lib.idp_ctx_create(self._panel, device, C.byref(self._ctx)))
        self.metric_counter = np.zeros(160, dtype=np.int64)   # SimpleCounter[TemplateTypeMetric] (IP:285)
        self.num_alignments = 0                               # IP:236

    # ---- lifecycle ----
    def close(self):
        lib = L.lib()
        if getattr(self, "_ctx", None) and self._ctx.value:
            lib.idp_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()
        if getattr(self, "_panel", None) and self._panel.value:
            lib.idp_panel_destroy(self._panel)
            self._panel = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- host-side panel queries (no GPU needed) ----
    def primer_length(self, i: int) -> int:
        return L.lib().idp_panel_primer_length(self._panel, i)

    def window(self) -> int:
        """Leading bases (sequencing order) the 5' chain can reach; -1 with --three-prime (idp_panel_window)."""
        return L.lib().idp_panel_window(self._panel)

    def primer_hash(self, i: int) -> int:
        return L.lib().idp_panel_primer_hash(self._panel, i)

    def kmer_postings(self, which: int, kmer: str) -> List[int]:
        out = np.zeros(len(self.primers) + 1, dtype=np.int32)
        n = L.lib().idp_panel_kmer_postings(self._panel, which, kmer.encode(), out.ctypes.data, len(out))
        return out[:max(n, 0)].tolist()

    # ---- the hot path ----
    def _require_ctx(self):
        if not self._ctx.value:
            raise RuntimeError("this IdentifyPrimers was created with device=None (host-only panel); matching needs a GPU context")

    def process_batch(self, reads: ReadBatch, compact: bool = False):
        """processBatch (IP:381-436): returns (five, three) result arrays; adds to metric_counter / num_alignments.
        compact=True goes through idp_match_batch16 (16-byte records over the bus) and unpacks them on the host: the rows
        returned are the same idp_result rows either way."""
        self._require_ctx()
        keep = []
        rs = reads.as_struct(keep)
        n = len(reads)
        dt = L.RESULT16_DTYPE if compact else L.RESULT_DTYPE
        five = np.zeros(n, dtype=dt)
        three = np.zeros(n, dtype=dt) if self.three_prime >= 0 else None
        metrics = np.zeros(160, dtype=np.int64)
        n_align = np.zeros(1, dtype=np.int64)
        fn = L.lib().idp_match_batch16 if compact else L.lib().idp_match_batch
        self._arm_classify(n, device=False)
        L.check(fn(self._ctx, C.byref(rs), n, five.ctypes.data, None if three is None else three.ctypes.data,
                   metrics.ctypes.data, n_align.ctypes.data))
        self.metric_counter += metrics
        self.num_alignments += int(n_align[0])
        if compact:
            five, three = L.unpack16(five), (None if three is None else L.unpack16(three))
        return five, three

    def process_batch_device(self, n: int, seq4_ptr: int, flags_ptr: int, five_ptr: int, three_ptr: int = 0, fixed_len: int = 0,
                             seq_off_ptr: int = 0, len_ptr: int = 0, ref_idx_ptr: int = 0, ustart_ptr: int = 0, uend_ptr: int = 0,
                             compact: bool = False):
        """The same with device-resident reads and results (raw device pointers, e.g. torch tensors' data_ptr());
        compact=True: five_ptr / three_ptr receive idp_result16 records."""
        self._require_ctx()
        rs = L.IdpReads()
        rs.seq4, rs.seq_off, rs.len = seq4_ptr, seq_off_ptr or None, len_ptr or None
        rs.ref_idx, rs.unclipped_start, rs.unclipped_end = ref_idx_ptr or None, ustart_ptr or None, uend_ptr or None
        rs.flags, rs.fixed_len = flags_ptr, fixed_len
        metrics = np.zeros(160, dtype=np.int64)
        n_align = np.zeros(1, dtype=np.int64)
        fn = L.lib().idp_match_batch16_device if compact else L.lib().idp_match_batch_device
        self._arm_classify(n, device=True)
        L.check(fn(self._ctx, C.byref(rs), n, five_ptr, three_ptr or None, metrics.ctypes.data, n_align.ctypes.data))
        self.metric_counter += metrics
        self.num_alignments += int(n_align[0])
        return metrics, int(n_align[0])

    def set_async(self, enable=True, stage_timing: bool = True):
        """Asynchronous device-resident mode (idp_ctx_set_async): process_batch_device only enqueues; sync() collects the counters.
        stage_timing=False drops the per-stage events and the per-batch counter copy (mode 2)."""
        self._require_ctx()
        L.check(L.lib().idp_ctx_set_async(self._ctx, (1 if stage_timing else 2) if enable else 0))

    def sync(self):
        """idp_ctx_sync: waits for the enqueued batches, adds their counters to metric_counter / num_alignments, returns them."""
        self._require_ctx()
        metrics = np.zeros(160, dtype=np.int64)
        n_align = np.zeros(1, dtype=np.int64)
        L.check(L.lib().idp_ctx_sync(self._ctx, metrics.ctypes.data, n_align.ctypes.data))
        self.metric_counter += metrics
        self.num_alignments += int(n_align[0])
        return metrics, int(n_align[0])

    def set_stream(self, cuda_stream_ptr: int):
        L.check(L.lib().idp_ctx_set_stream(self._ctx, cuda_stream_ptr or None))

    def timings(self) -> dict:
        t = L.IdpTimings()
        L.check(L.lib().idp_ctx_timings(self._ctx, C.byref(t)))
        return {k: getattr(t, k) for k, _ in L.IdpTimings._fields_}

    # ---- decoding ----
    def to_match(self, row):
        """idp_result -> the PrimerMatch the reference would hold (None for no match)."""
        t = int(row["type"])
        if t == L.NONE:
            return None
        p = self.primers[int(row["primer"])]
        if t == L.LOCATION:
            return LocationBasedPrimerMatch(p, int(row["mm"]))
        if t == L.UNGAPPED:
            return UngappedAlignmentPrimerMatch(p, int(row["mm"]), int(row["next_mm"]))
        r = np.array([row], dtype=L.RESULT_DTYPE)
        return GappedAlignmentPrimerMatch(p, L.lib().idp_result_score(r.ctypes.data), L.lib().idp_result_second(r.ctypes.data),
                                          int(row["t_start"]), int(row["t_end"]))

    def info(self, row, read_flags: int) -> str:
        """PrimerMatch.info(rec) (PrimerMatch.scala:78-88), or "none" (IP:242)."""
        r = np.array([row], dtype=L.RESULT_DTYPE)
        buf = C.create_string_buffer(1024)
        L.lib().idp_format_info(self._panel, r.ctypes.data, read_flags, buf, 1024)
        return buf.value.decode()

    def pair_tags(self, r1_flags, r2_flags, r1_five, r2_five, r1_three=None, r2_three=None) -> dict:
        """addTagsToPair (IP:522-554): the f5/r5 (and f3/r3) strings both reads of a pair receive."""
        def positive(row):
            return None if int(row["type"]) == L.NONE else self.primers[int(row["primer"])].positive_strand
        p1, p2 = positive(r1_five), positive(r2_five)
        if p1 is not None:
            fwd_is_r1 = p1
        elif p2 is not None:
            fwd_is_r1 = not p2
        else:
            fwd_is_r1 = True
        fwd = (r1_flags, r1_five, r1_three) if fwd_is_r1 else (r2_flags, r2_five, r2_three)
        rev = (r2_flags, r2_five, r2_three) if fwd_is_r1 else (r1_flags, r1_five, r1_three)
        tags = {"f5": self.info(fwd[1], fwd[0]), "r5": self.info(rev[1], rev[0])}
        if self.three_prime >= 0:
            tags["f3"] = self.info(fwd[2], fwd[0])
            tags["r3"] = self.info(rev[2], rev[0])
        return tags

    def fragment_tags(self, flags, five, three=None) -> dict:
        """addTagsToFragment (IP:557-569), including its overwrite of f5 by the 3' match when --three-prime >= 0."""
        tags = {"f5": self.info(five, flags), "r5": "none"}
        if self.three_prime >= 0:
            tags["f5"] = self.info(three, flags)
        return tags

    def annotate_bam(self, bam: bytes, five, three=None) -> bytes:
        """The records of `bam` (the bytes ReadBatch.from_bam was given) with the f5/r5(/f3/r3) tags set as addTagsToPair /
        addTagsToFragment do (IP:522-569); `five` / `three` are the result rows of process_batch on that batch."""
        lib = L.lib()
        buf = np.frombuffer(bam, dtype=np.uint8)
        five = np.ascontiguousarray(five, dtype=L.RESULT_DTYPE)
        three_ptr = None
        if three is not None:
            three = np.ascontiguousarray(three, dtype=L.RESULT_DTYPE)
            three_ptr = three.ctypes.data
        need = C.c_size_t(0)
        L.check(lib.idp_bam_annotate(self._panel, buf.ctypes.data, len(buf), five.ctypes.data, three_ptr, len(five), None, 0, C.byref(need)))
        out = np.zeros(need.value, dtype=np.uint8)
        L.check(lib.idp_bam_annotate(self._panel, buf.ctypes.data, len(buf), five.ctypes.data, three_ptr, len(five), out.ctypes.data, len(out), C.byref(need)))
        return out.tobytes()

    @staticmethod
    def header_comments() -> List[str]:
        """The @CO lines the tool adds to the output header (IP:243-250)."""
        buf = C.create_string_buffer(2048)
        L.lib().idp_header_comments(buf, 2048)
        return buf.value.decode().split("\n")

    # ---- downstream classification (scripts/post_process_bam.py:56-78) ----
    CLASS_NAMES = ["Canonical", "SelfDimer", "CrossDimer", "NonCanonical", "Single", "NoMatch"]

    def classify_primer_pair(self, f5_primer: int, r5_primer: int, min_insert_length: int = 50, max_insert_length: int = 250) -> str:
        return self.CLASS_NAMES[L.lib().idp_classify_primer_pair(self._panel, f5_primer, r5_primer, min_insert_length, max_insert_length)]

    def classify_templates(self, flags, five, min_insert_length: int = 50, max_insert_length: int = 250):
        """Per-read class codes and the six match-type counts (one per template whose first record is read 1)."""
        flags = np.ascontiguousarray(flags, dtype=np.uint16)
        five = np.ascontiguousarray(five, dtype=L.RESULT_DTYPE)
        cls = np.zeros(len(flags), dtype=np.uint8)
        counts = np.zeros(6, dtype=np.int64)
        L.check(L.lib().idp_classify_templates(self._panel, flags.ctypes.data, five.ctypes.data, len(flags), min_insert_length, max_insert_length,
                                               cls.ctypes.data, counts.ctypes.data))
        return cls, dict(zip(self.CLASS_NAMES, counts.tolist()))

    def set_classify(self, min_insert_length: int = 50, max_insert_length: int = 250, enable: bool = True, cls_ptr: int = 0):
        """Device epilogue (idp_ctx_set_classify): every later process_batch* call also classifies the templates on the GPU.
        `class_counts` accumulates the six counts; process_batch() returns the per-read classes in `last_classes`;
        process_batch_device() writes them to the device array at cls_ptr (0 = counts only)."""
        self._require_ctx()
        self.class_counts = np.zeros(6, dtype=np.int64)
        self._cls_on, self._cls_dev_ptr, self._cls_args = enable, cls_ptr, (min_insert_length, max_insert_length)
        self.last_classes = None
        if not enable:
            L.check(L.lib().idp_ctx_set_classify(self._ctx, -1, 0, None, None))

    def _arm_classify(self, n: int, device: bool):
        if not getattr(self, "_cls_on", False):
            return
        if device:
            ptr = self._cls_dev_ptr or None
        else:
            self.last_classes = np.zeros(n, dtype=np.uint8)
            ptr = self.last_classes.ctypes.data
        L.check(L.lib().idp_ctx_set_classify(self._ctx, self._cls_args[0], self._cls_args[1], ptr, self.class_counts.ctypes.data))

    def classify_batch_device(self, flags_ptr: int, five_ptr: int, n: int, compact: bool, cls_ptr: int = 0, min_insert_length: int = 50,
                              max_insert_length: int = 250):
        """idp_classify_batch_device on device-resident flags and 5' records -> dict of the six counts."""
        self._require_ctx()
        counts = np.zeros(6, dtype=np.int64)
        L.check(L.lib().idp_classify_batch_device(self._ctx, flags_ptr, five_ptr, 1 if compact else 0, n, min_insert_length, max_insert_length,
                                                  cls_ptr or None, counts.ctypes.data))
        return dict(zip(self.CLASS_NAMES, counts.tolist()))

    # ---- metrics (TemplateTypeMetric.scala:34-56, IdentifyPrimersMetric.scala:34-89) ----
    def summary(self) -> dict:
        out = np.zeros(17, dtype=np.int64)
        L.lib().idp_summary_metrics(self.metric_counter.ctypes.data, out.ctypes.data)
        names = ["templates", "pairs", "fragments", "mapped_pairs", "unpaired", "unmapped_pairs", "fr_pairs", "mapped_fragments",
                 "unmapped_fragments", "paired_matches", "unpaired_matches", "no_paired_matches", "match_attempts", "location",
                 "ungapped", "gapped", "no_match"]
        return dict(zip(names, out.tolist()))

    def template_type_metrics(self) -> List[dict]:
        tt_names = ["MappedPair", "PartiallyMappedPair", "UnmappedPair", "MappedFragment", "UnmappedFragment"]
        m_names = ["NoPrimerMatch", "Location", "Ungapped", "Gapped"]
        total = float(self.metric_counter.sum())
        rows = []
        for b in np.nonzero(self.metric_counter)[0]:
            c = int(self.metric_counter[b])
            rows.append(dict(template_type=tt_names[b >> 5], is_fr=bool((b >> 4) & 1), r1_primer_match_type=m_names[(b >> 2) & 3],
                             r2_primer_match_type=m_names[b & 3], count=c, frac=c / total))
        rows.sort(key=lambda r: -r["count"])  # stable, like sortBy(-_.count)
        return rows


class AsyncCudaAligner:
    """A batched AsyncAligner (AsyncAligner.scala:37-54): align(query, target) for many pairs in one GPU call."""

    def __init__(self, match_score=1, mismatch_score=-4, gap_open=-6, gap_extend=-1, mode=L.MODE_GLOCAL, device=0):
        self._tool = IdentifyPrimers([Primer("p", "p+", "A", True), Primer("p", "p-", "T", False)], match_score=match_score,
                                     mismatch_score=mismatch_score, gap_open=gap_open, gap_extend=gap_extend, device=device)
        self.mode = mode
        self.num_aligned = 0

    def align(self, queries: Sequence[bytes], targets: Sequence[bytes], query_coordinates: bool = False):
        """-> (score, target_start, target_end) arrays, plus (query_start, query_end) when query_coordinates is set."""
        n = len(queries)
        q_off = np.zeros(n + 1, dtype=np.int64); q_off[1:] = np.cumsum([len(q) for q in queries])
        t_off = np.zeros(n + 1, dtype=np.int64); t_off[1:] = np.cumsum([len(t) for t in targets])
        score = np.zeros(n, dtype=np.int32); ts = np.zeros(n, dtype=np.int32); te = np.zeros(n, dtype=np.int32)
        qs = np.zeros(n, dtype=np.int32) if query_coordinates else None
        qe = np.zeros(n, dtype=np.int32) if query_coordinates else None
        L.check(L.lib().idp_align_batch(self._tool._ctx, self.mode, b"".join(queries), q_off.ctypes.data, b"".join(targets),
                                        t_off.ctypes.data, n, score.ctypes.data, ts.ctypes.data, te.ctypes.data,
                                        qs.ctypes.data if query_coordinates else None, qe.ctypes.data if query_coordinates else None))
        self.num_aligned += n
        return (score, ts, te, qs, qe) if query_coordinates else (score, ts, te)

    def close(self):
        self._tool.close()
