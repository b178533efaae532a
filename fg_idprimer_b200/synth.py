"""Deterministic synthetic workloads ("synth-v1", SURVEY.md section 8d / BASELINE.md section 3).

Panels of random amplicons (forward primer + insert + reverse-complemented reverse primer) and 2x150 bp read pairs
drawn from them with substitutions, Ns, off-target and cross-pair templates, optional indels in the primer site,
optional IUPAC-degenerate primers, optional mapping coordinates.  Everything is produced as BAM 4-bit codes
(`=ACMGRSVTWYHKDBN` -> 0..15); `codes_to_ascii` gives the ASCII the CPU oracle consumes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .api import Primer

NIBBLE_CHARS = "=ACMGRSVTWYHKDBN"
_ASCII_LUT = np.frombuffer(NIBBLE_CHARS.encode(), dtype=np.uint8)
ACGT = np.array([1, 2, 4, 8], dtype=np.uint8)
_COMP = np.arange(16, dtype=np.uint8)
_COMP[[1, 2, 4, 8]] = [8, 4, 2, 1]                       # htsjdk 2.16: only A/C/G/T are complemented
_FULL_COMP = np.array([int(f"{i:04b}"[::-1], 2) for i in range(16)], dtype=np.uint8)  # true IUPAC complement (bit reversal)
_POPC = np.array([bin(i).count("1") for i in range(16)], dtype=np.uint8)
_MEMBER = np.zeros((16, 12), dtype=np.uint8)             # _MEMBER[code][r] = a uniformly chosen member base for r in 0..11
for _c in range(1, 16):
    _bits = [b for b in (1, 2, 4, 8) if _c & b]
    for _r in range(12):
        _MEMBER[_c, _r] = _bits[_r % len(_bits)]
_SUPERSETS = {b: [c for c in range(1, 16) if (c & b) and _POPC[c] >= 2] for b in (1, 2, 4, 8)}

F_PAIRED, F_UNMAPPED, F_REVERSE, F_FIRST, F_SECOND, F_FR_PAIR, F_MATE_NEXT = 0x1, 0x4, 0x10, 0x40, 0x80, 0x1000, 0x2000
CONTIGS = [f"chr{i}" for i in range(1, 23)]


def codes_to_ascii(codes: np.ndarray) -> np.ndarray:
    return _ASCII_LUT[codes]


def codes_to_str(codes) -> str:
    return codes_to_ascii(np.asarray(codes, dtype=np.uint8)).tobytes().decode()


def pack_fixed(codes: np.ndarray) -> np.ndarray:
    """[n, L] codes -> [n, (L+1)//2] BAM bytes (first base in the high nibble)."""
    n, L = codes.shape
    if L % 2:
        codes = np.concatenate([codes, np.zeros((n, 1), dtype=np.uint8)], axis=1)
    return ((codes[:, 0::2] << 4) | codes[:, 1::2]).astype(np.uint8)


@dataclass
class Panel:
    primers: List[Primer]
    amp_fwd: np.ndarray      # [P, read_len] codes of the amplicon from the forward primer's 5' end; 0 = beyond the amplicon
    amp_rev: np.ndarray      # [P, read_len] codes of its reverse complement (starts with the reverse primer)
    fwd_len: np.ndarray
    rev_len: np.ndarray
    amp_len: np.ndarray
    contig: np.ndarray       # [P] contig index (mapped panels)
    start: np.ndarray        # [P] 1-based start of the amplicon (mapped panels)
    mapped: bool
    ref_names: List[str]


def make_panel(n_pairs: int, seed: int, degenerate: bool = False, mapped: bool = False, read_len: int = 150,
               min_len: int = 18, max_len: int = 30) -> Panel:
    rng = np.random.default_rng(seed)
    lf = rng.integers(min_len, max_len + 1, n_pairs)
    lr = rng.integers(min_len, max_len + 1, n_pairs)
    amp_len = rng.integers(120, 251, n_pairs)
    max_amp = 250
    amp = ACGT[rng.integers(0, 4, (n_pairs, max_amp))]
    seen = set()
    primers: List[Primer] = []
    amp_fwd = np.zeros((n_pairs, read_len), dtype=np.uint8)
    amp_rev = np.zeros((n_pairs, read_len), dtype=np.uint8)
    contig = (np.arange(n_pairs) % len(CONTIGS)).astype(np.int32)
    start = (1000 * (np.arange(n_pairs) + 1)).astype(np.int32)
    for i in range(n_pairs):
        a = amp[i, :amp_len[i]].copy()
        while True:  # reject a primer whose sequence duplicates an earlier one
            f = a[:lf[i]]
            r = _COMP[a[amp_len[i] - lr[i]:][::-1]]
            kf, kr = f.tobytes(), r.tobytes()
            if kf not in seen and kr not in seen and kf != kr:
                seen.add(kf); seen.add(kr)
                break
            a[:] = ACGT[rng.integers(0, 4, amp_len[i])]
        if degenerate:  # each primer base becomes, with prob 0.05, an IUPAC code containing it; at least one per primer
            for lo, hi in ((0, lf[i]), (amp_len[i] - lr[i], amp_len[i])):
                n = hi - lo
                pick = rng.random(n) < 0.05
                if not pick.any():
                    pick[rng.integers(0, n)] = True
                for j in np.nonzero(pick)[0]:
                    base = int(a[lo + j])
                    a[lo + j] = _SUPERSETS[base][rng.integers(0, len(_SUPERSETS[base]))]
        fseq = codes_to_str(a[:lf[i]])
        rseq = codes_to_str(_FULL_COMP[a[amp_len[i] - lr[i]:][::-1]])  # the reverse primer as ordered (its own 5'->3')
        n = min(read_len, amp_len[i])
        amp_fwd[i, :n] = a[:n]
        amp_rev[i, :n] = _FULL_COMP[a[::-1]][:n]
        if mapped:
            s, e = int(start[i]), int(start[i] + amp_len[i] - 1)
            primers.append(Primer(f"p{i}", f"p{i}_F", fseq, True, CONTIGS[contig[i]], s, s + int(lf[i]) - 1))
            primers.append(Primer(f"p{i}", f"p{i}_R", rseq, False, CONTIGS[contig[i]], e - int(lr[i]) + 1, e))
        else:
            primers.append(Primer(f"p{i}", f"p{i}_F", fseq, True))
            primers.append(Primer(f"p{i}", f"p{i}_R", rseq, False))
    return Panel(primers, amp_fwd, amp_rev, lf, lr, amp_len, contig, start, mapped, list(CONTIGS))


@dataclass
class Reads:
    codes: np.ndarray            # [2n, read_len] BAM codes as STORED in the record (R1, R2 interleaved)
    flags: np.ndarray            # uint16
    ref_idx: Optional[np.ndarray]
    ustart: Optional[np.ndarray]
    uend: Optional[np.ndarray]
    truth: np.ndarray            # [2n] amplicon index the read was drawn from, -1 off-target

    @property
    def n(self):
        return len(self.flags)

    def seq4(self) -> np.ndarray:
        return pack_fixed(self.codes)

    def oracle_arrays(self, lo: int = 0, hi: Optional[int] = None) -> dict:
        hi = self.n if hi is None else hi
        c = self.codes[lo:hi]
        n, L = c.shape
        z = np.zeros(n, dtype=np.int32)
        return dict(bases=codes_to_ascii(c).reshape(-1), off=np.arange(n + 1, dtype=np.int64) * L,
                    ref_idx=(self.ref_idx[lo:hi] if self.ref_idx is not None else z - 1),
                    ustart=(self.ustart[lo:hi] if self.ustart is not None else z),
                    uend=(self.uend[lo:hi] if self.uend is not None else z), flags=self.flags[lo:hi])


def make_reads(panel: Panel, n_pairs: int, seed: int, indel_frac: float = 0.0, mapped_frac: float = 0.0, sub_rate: float = 0.005,
               n_rate: float = 0.001, offtarget: float = 0.02, crosspair: float = 0.01, read_len: int = 150) -> Reads:
    rng = np.random.default_rng(seed)
    P = len(panel.fwd_len)
    a1 = rng.integers(0, P, n_pairs)
    a2 = a1.copy()
    u = rng.random(n_pairs)
    off = u < offtarget
    cross = (u >= offtarget) & (u < offtarget + crosspair)
    a2[cross] = (a1[cross] + 1 + rng.integers(0, max(P - 1, 1), int(cross.sum()))) % P
    codes = np.empty((2 * n_pairs, read_len), dtype=np.uint8)
    codes[0::2] = panel.amp_fwd[a1]
    codes[1::2] = panel.amp_rev[a2]
    truth = np.empty(2 * n_pairs, dtype=np.int32)
    truth[0::2] = a1; truth[1::2] = a2
    truth[0::2][off] = -1; truth[1::2][off] = -1
    codes[np.repeat(off, 2)] = 0
    rnd = np.left_shift(np.uint8(1), rng.integers(0, 4, codes.shape, dtype=np.uint8))   # iid A/C/G/T codes
    np.copyto(codes, rnd, where=(codes == 0))                # beyond the amplicon / off-target: iid bases
    del rnd
    deg = (codes & (codes - 1)) != 0                         # more than one bit set: instantiate degenerate positions
    if deg.any():
        idx = np.nonzero(deg)
        codes[idx] = _MEMBER[codes[idx], rng.integers(0, 12, len(idx[0]))]
    del deg
    total = codes.size
    k = rng.binomial(total, sub_rate)                        # substitutions
    pos = rng.integers(0, total, k)
    flat = codes.reshape(-1)
    cur = flat[pos]
    order = np.array([0, 0, 1, 0, 2, 0, 0, 0, 3], dtype=np.uint8)   # code -> 0..3
    flat[pos] = ACGT[(order[cur] + rng.integers(1, 4, k)) % 4]
    if indel_frac > 0:                                       # one indel of length 1-2 inside the primer site
        sel = np.nonzero((rng.random(2 * n_pairs) < indel_frac) & (truth >= 0))[0]
        plen = np.where(sel % 2 == 0, panel.fwd_len[truth[sel]], panel.rev_len[truth[sel]])
        d = rng.integers(1, 3, len(sel))
        is_ins = rng.random(len(sel)) < 0.5
        at = 1 + (rng.random(len(sel)) * (plen - 2)).astype(np.int64)   # interior position
        j = np.arange(read_len)[None, :]
        shift = np.where(j >= at[:, None], np.where(is_ins, -d, d)[:, None], 0)
        src = np.clip(j + shift, 0, read_len - 1)
        rows = codes[sel]
        new = np.take_along_axis(rows, src, axis=1)
        fresh = ACGT[rng.integers(0, 4, new.shape)]
        ins_mask = is_ins[:, None] & (j >= at[:, None]) & (j < (at + d)[:, None])
        tail_mask = (~is_ins)[:, None] & (j >= (read_len - d)[:, None])
        new[ins_mask | tail_mask] = fresh[ins_mask | tail_mask]
        codes[sel] = new
    k = rng.binomial(total, n_rate)                          # Ns
    flat = codes.reshape(-1)
    flat[rng.integers(0, total, k)] = 15
    flags = np.empty(2 * n_pairs, dtype=np.uint16)
    flags[0::2] = F_PAIRED | F_FIRST | F_MATE_NEXT | F_UNMAPPED
    flags[1::2] = F_PAIRED | F_SECOND | F_UNMAPPED
    ref_idx = ustart = uend = None
    if panel.mapped and mapped_frac > 0:
        is_mapped = (rng.random(n_pairs) < mapped_frac) & ~off & ~cross
        ref_idx = np.full(2 * n_pairs, -1, dtype=np.int32)
        ustart = np.zeros(2 * n_pairs, dtype=np.int32)
        uend = np.zeros(2 * n_pairs, dtype=np.int32)
        m = np.nonzero(is_mapped)[0]
        jit1, jit2 = rng.integers(-2, 3, len(m)), rng.integers(-2, 3, len(m))
        amp = a1[m]
        s = panel.start[amp]
        e = s + panel.amp_len[amp] - 1
        ref_idx[2 * m] = panel.contig[amp]; ref_idx[2 * m + 1] = panel.contig[amp]
        ustart[2 * m] = s + jit1; uend[2 * m] = s + jit1 + read_len - 1
        uend[2 * m + 1] = e + jit2; ustart[2 * m + 1] = e + jit2 - read_len + 1
        flags[2 * m] = F_PAIRED | F_FIRST | F_MATE_NEXT | F_FR_PAIR
        flags[2 * m + 1] = F_PAIRED | F_SECOND | F_REVERSE | F_FR_PAIR
        # a negative-strand record stores the reverse complement of what was sequenced
        codes[2 * m + 1] = _COMP[codes[2 * m + 1][:, ::-1]]
    return Reads(codes, flags, ref_idx, ustart, uend, truth)


# the BASELINE.md workloads: (panel pairs, options) -- reads are generated per call with the sizes the caller asks for
WORKLOADS = {
    "c1": dict(pairs=96, panel_seed=1001, read_seed=2001, degenerate=False, mapped=False, indel_frac=0.0, mapped_frac=0.0,
               options=dict(ungapped_kmer_length=6, gapped_kmer_length=10)),
    "c2": dict(pairs=96, panel_seed=1002, read_seed=2002, degenerate=True, mapped=False, indel_frac=0.0, mapped_frac=0.0,
               options=dict(ungapped_kmer_length=6, gapped_kmer_length=10)),
    "c3": dict(pairs=1000, panel_seed=1003, read_seed=2003, degenerate=False, mapped=False, indel_frac=0.05, mapped_frac=0.0,
               options=dict(ungapped_kmer_length=6, gapped_kmer_length=10)),
    "c3b": dict(pairs=1000, panel_seed=1003, read_seed=2003, degenerate=False, mapped=False, indel_frac=0.05, mapped_frac=0.0,
                options=dict(ungapped_kmer_length=6, gapped_kmer_length=None)),
    "c4": dict(pairs=5000, panel_seed=1004, read_seed=2004, degenerate=False, mapped=False, indel_frac=0.05, mapped_frac=0.0,
               options=dict(ungapped_kmer_length=6, gapped_kmer_length=10)),
    "c5": dict(pairs=96, panel_seed=1005, read_seed=2005, degenerate=False, mapped=True, indel_frac=0.05, mapped_frac=0.7,
               options=dict(ungapped_kmer_length=6, gapped_kmer_length=10, three_prime=5)),
}


def workload(name: str, n_pairs: int, read_seed_offset: int = 0):
    w = WORKLOADS[name]
    panel = make_panel(w["pairs"], w["panel_seed"], degenerate=w["degenerate"], mapped=w["mapped"])
    reads = make_reads(panel, n_pairs, w["read_seed"] + read_seed_offset, indel_frac=w["indel_frac"], mapped_frac=w["mapped_frac"])
    return panel, reads, dict(w["options"])
