"""Randomised differential test: random panels, options and reads through the GPU path and the CPU oracle.
Every seed draws its own option surface (slop, rates, -k/-K, 3' matching, scores), panel shape (lengths, IUPAC letters,
near-duplicates, mapped / unmapped primers) and read mix (edits, strands, ragged lengths, pairs and fragments)."""
import numpy as np
import pytest

import helpers as H
from test_gpu_parity import run_both

pytestmark = pytest.mark.gpu

IUPAC = {"A": "RWMN", "C": "YSMN", "G": "RSKN", "T": "YWKN"}
REFS = H.DEFAULT_DICT


def _seq(rng, n):
    return "".join("ACGT"[x] for x in rng.integers(0, 4, n))


def _instantiate(rng, primer_seq):
    members = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC", "N": "ACGT"}
    return "".join(members[c][int(rng.integers(0, len(members[c])))] if c in members else c for c in primer_seq)


def _panel(rng):
    n_pairs = int(rng.integers(2, 30))
    mapped = rng.random() < 0.5
    degenerate = rng.random() < 0.4
    lo, hi = (12, 24) if rng.random() < 0.3 else (18, 46)
    primers = []
    base = _seq(rng, 60)
    for i in range(n_pairs):
        for strand in (True, False):
            n = int(rng.integers(lo, hi))
            if rng.random() < 0.25:   # near-duplicate of a shared core
                s = list(base[:n])
                for j in rng.choice(n, size=int(rng.integers(0, 3)), replace=False):
                    s[j] = "ACGT"[int(rng.integers(0, 4))]
                seq = "".join(s)
            else:
                seq = _seq(rng, n)
            if degenerate:
                s = list(seq)
                for j in range(n):
                    if rng.random() < 0.06:
                        s[j] = IUPAC[s[j]][int(rng.integers(0, 4))] if s[j] in IUPAC else s[j]
                seq = "".join(s)
            if mapped:
                ref = int(rng.integers(0, 4))
                start = int(rng.integers(1, 3000))
                span = n if rng.random() < 0.8 else n + int(rng.integers(-3, 4))   # Primer.length may differ from the sequence length
                primers.append(H.Primer(f"p{i}", f"p{i}{'+' if strand else '-'}", seq, strand, REFS[ref], start, start + max(span, 1) - 1, ref))
            else:
                primers.append(H.Primer(f"p{i}", f"p{i}{'+' if strand else '-'}", seq, strand))
    # the tool refuses duplicate coordinates (Primer.scala:156-164); drop collisions
    seen, out = set(), []
    for p in primers:
        key = (p.ref_name, p.start, p.end, p.positive_strand) if p.ref_name else (p.primer_id,)
        if key in seen:
            continue
        seen.add(key)
        out.append(p)
    ids = {}
    for p in out:
        ids.setdefault(p.pair_id, []).append(p)
    return [p for p in out if len(ids[p.pair_id]) == 2]


def _reads(rng, primers, n):
    recs = []
    for _ in range(n):
        p = primers[int(rng.integers(0, len(primers)))]
        kind = rng.random()
        s = list(_instantiate(rng, p.sequence)) + list(_seq(rng, int(rng.integers(0, 160))))
        if kind < 0.15:
            s = list(_seq(rng, int(rng.integers(0, 120))))                       # unrelated
        elif kind < 0.45:
            for j in rng.choice(max(len(p.sequence), 1), size=int(rng.integers(1, 4)), replace=True):
                if j < len(s):
                    s[j] = "ACGTN"[int(rng.integers(0, 5))]                      # substitutions / Ns
        elif kind < 0.6 and len(s) > 8:
            j = int(rng.integers(1, min(len(p.sequence), len(s)) ))
            if rng.random() < 0.5:
                del s[j:j + int(rng.integers(1, 3))]
            else:
                s[j:j] = list(_seq(rng, int(rng.integers(1, 3))))
        if rng.random() < 0.1:
            s = s[: int(rng.integers(0, 25))]
        if rng.random() < 0.15:
            q = primers[int(rng.integers(0, len(primers)))]
            s = s + list(H.revcomp(_instantiate(rng, q.sequence))) + list(_seq(rng, int(rng.integers(0, 6))))   # a primer at the 3' end
        seq = "".join(s)[:400]
        is_mapped = bool(p.ref_name) and rng.random() < 0.6
        neg = bool(rng.random() < 0.5)
        if is_mapped:
            stored = H.revcomp(seq) if neg else seq
            jitter = int(rng.integers(-7, 8))
            if neg:  # unclippedEnd near primer.end
                start = p.end + jitter - len(stored) + 1
            else:
                start = p.start + jitter
            recs.append(H.Rec(stored, False, neg, p.ref_idx if rng.random() < 0.9 else int(rng.integers(0, 4)), start))
        else:
            recs.append(H.Rec(seq, True, neg))
    n_pairs = int(rng.integers(0, n // 2))
    return H.pair_records(recs[: 2 * n_pairs]) + recs[2 * n_pairs:]


@pytest.mark.parametrize("seed", range(0, 96, 2))
def test_random_configurations_seed_index(seed, monkeypatch):
    """The same draws with the seed index forced on (by default only panels of more than 256 primers use it)."""
    monkeypatch.setenv("IDP_SCAN_SEED", "1")
    test_random_configurations(seed)


@pytest.mark.parametrize("seed", range(0, 96, 3))
def test_random_configurations_bit_sliced(seed, monkeypatch):
    monkeypatch.setenv("IDP_SCAN_SEED", "0")
    test_random_configurations(seed)


def test_large_panel_uses_seed_index():
    """1,200 primers (38 groups): the seed index path by default; near-duplicate primers share seeds."""
    rng = np.random.default_rng(77)
    primers = []
    for i in range(600):
        for strand in (True, False):
            seq = _seq(rng, int(rng.integers(18, 34)))
            if i % 50 == 1 and primers:  # shares both seeds with the previous pair's primer, differs further in
                seq = primers[-2].sequence[:16] + _seq(rng, 8)
            primers.append(H.Primer(f"p{i}", f"p{i}{'+' if strand else '-'}", seq, strand))
    recs = _reads(rng, primers, 6000)
    run_both(primers, H.recs_to_arrays(recs), max_mismatch_rate=0.05, ungapped_kmer_length=6, gapped_kmer_length=10)
    run_both(primers, H.recs_to_arrays(recs[:2000]), max_mismatch_rate=0.05, ungapped_kmer_length=None, gapped_kmer_length=8)


@pytest.mark.parametrize("seed", range(96))
def test_random_configurations(seed):
    rng = np.random.default_rng(1000 + seed)
    primers = _panel(rng)
    if len(primers) < 2:
        pytest.skip("degenerate panel")
    opts = dict(slop=int(rng.integers(0, 9)), max_mismatch_rate=float(rng.choice([0.0, 0.05, 0.05, 0.1, 0.2, 0.5])),
                min_alignment_score_rate=float(rng.choice([0.0, 0.01, 0.3])), match_score=int(rng.integers(0, 4)),
                mismatch_score=-int(rng.integers(0, 6)), gap_open=-int(rng.integers(0, 8)), gap_extend=-int(rng.integers(0, 4)),
                ungapped_kmer_length=rng.choice([None, 3, 5, 6, 8, 11]), gapped_kmer_length=rng.choice([None, 4, 6, 10, 16]),
                three_prime=int(rng.choice([-1, -1, 0, 4])))
    for k in ("ungapped_kmer_length", "gapped_kmer_length"):
        opts[k] = None if opts[k] is None else int(opts[k])
    recs = _reads(rng, primers, int(rng.integers(200, 1500)))
    run_both(primers, H.recs_to_arrays(recs), **opts)
