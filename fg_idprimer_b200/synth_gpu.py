"""The "synth-v1" read generator of synth.py on the GPU (torch), for bench.py: the same distributions -- amplicon choice,
off-target and cross-pair templates, instantiation of degenerate positions, substitutions, primer-site indels, Ns, mapping
coordinates -- drawn from torch's generator instead of numpy's, so that tens of millions of pairs take seconds rather than
minutes.  Tests and the oracle keep using synth.make_reads; bench.py copies a prefix of what this produces back to the host
for its spot check against the oracle."""
from __future__ import annotations

import numpy as np
import torch

from . import synth

_ORDER = [0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0]  # code -> 0..3 for A/C/G/T


class PanelTensors:
    def __init__(self, panel: synth.Panel, device):
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        self.amp_fwd, self.amp_rev = t(panel.amp_fwd, torch.uint8), t(panel.amp_rev, torch.uint8)
        self.fwd_len, self.rev_len, self.amp_len = t(panel.fwd_len, torch.int64), t(panel.rev_len, torch.int64), t(panel.amp_len, torch.int64)
        self.contig, self.start = t(panel.contig, torch.int32), t(panel.start, torch.int32)
        self.mapped = panel.mapped
        self.P = len(panel.fwd_len)
        self.member = t(synth._MEMBER, torch.uint8)
        self.comp = t(synth._COMP, torch.uint8)
        self.acgt = t(synth.ACGT, torch.uint8)
        self.order = torch.tensor(_ORDER, dtype=torch.int64, device=device)


def make_reads(pt: PanelTensors, n_pairs: int, seed: int, indel_frac: float = 0.0, mapped_frac: float = 0.0, sub_rate: float = 0.005,
               n_rate: float = 0.001, offtarget: float = 0.02, crosspair: float = 0.01, read_len: int = 150):
    """-> dict(codes uint8 [2n, read_len] as stored in the record, flags int16, ref_idx / ustart / uend int32 or None)."""
    dev = pt.amp_fwd.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ri = lambda hi, shape, lo=0: torch.randint(lo, hi, shape, generator=g, device=dev)
    rf = lambda shape: torch.rand(shape, generator=g, device=dev)
    P = pt.P
    a1 = ri(P, (n_pairs,))
    u = rf((n_pairs,))
    off = u < offtarget
    cross = (u >= offtarget) & (u < offtarget + crosspair)
    a2 = torch.where(cross, (a1 + 1 + ri(max(P - 1, 1), (n_pairs,))) % P, a1)
    codes = torch.empty((2 * n_pairs, read_len), dtype=torch.uint8, device=dev)
    codes[0::2] = pt.amp_fwd[a1]
    codes[1::2] = pt.amp_rev[a2]
    truth = torch.stack([torch.where(off, -1, a1), torch.where(off, -1, a2)], dim=1).reshape(-1)
    codes[truth < 0] = 0
    rnd = pt.acgt[ri(4, codes.shape)]
    codes = torch.where(codes == 0, rnd, codes)              # beyond the amplicon / off-target: iid bases
    del rnd
    deg = (codes & (codes - 1)) != 0                          # degenerate positions: a uniformly chosen member base
    if bool(deg.any()):
        idx = deg.nonzero(as_tuple=True)
        codes[idx] = pt.member[codes[idx].long(), ri(12, (len(idx[0]),))]
    del deg
    total = codes.numel()
    flat = codes.view(-1)
    k = int(np.random.default_rng(seed).binomial(total, sub_rate))   # substitutions
    pos = ri(total, (k,))
    flat[pos] = pt.acgt[(pt.order[flat[pos].long()] + ri(4, (k,), lo=1)) % 4]
    if indel_frac > 0:                                        # one indel of length 1-2 inside the primer site
        sel = ((rf((2 * n_pairs,)) < indel_frac) & (truth >= 0)).nonzero(as_tuple=True)[0]
        tr = truth[sel]
        plen = torch.where(sel % 2 == 0, pt.fwd_len[tr], pt.rev_len[tr])
        d = ri(3, (len(sel),), lo=1)
        is_ins = rf((len(sel),)) < 0.5
        at = 1 + (rf((len(sel),)) * (plen - 2).float()).long()
        j = torch.arange(read_len, device=dev)[None, :]
        shift = torch.where(j >= at[:, None], torch.where(is_ins, -d, d)[:, None], torch.zeros_like(j))
        src = (j + shift).clamp(0, read_len - 1)
        new = torch.gather(codes[sel], 1, src)
        fresh = pt.acgt[ri(4, new.shape)]
        ins_mask = is_ins[:, None] & (j >= at[:, None]) & (j < (at + d)[:, None])
        tail_mask = (~is_ins)[:, None] & (j >= (read_len - d)[:, None])
        m = ins_mask | tail_mask
        new[m] = fresh[m]
        codes[sel] = new
    k = int(np.random.default_rng(seed + 1).binomial(total, n_rate))  # Ns
    codes.view(-1)[ri(total, (k,))] = 15
    S = synth
    flags = torch.empty(2 * n_pairs, dtype=torch.int32, device=dev)
    flags[0::2] = S.F_PAIRED | S.F_FIRST | S.F_MATE_NEXT | S.F_UNMAPPED
    flags[1::2] = S.F_PAIRED | S.F_SECOND | S.F_UNMAPPED
    ref_idx = ustart = uend = None
    if pt.mapped and mapped_frac > 0:
        is_mapped = (rf((n_pairs,)) < mapped_frac) & ~off & ~cross
        ref_idx = torch.full((2 * n_pairs,), -1, dtype=torch.int32, device=dev)
        ustart = torch.zeros(2 * n_pairs, dtype=torch.int32, device=dev)
        uend = torch.zeros(2 * n_pairs, dtype=torch.int32, device=dev)
        m = is_mapped.nonzero(as_tuple=True)[0]
        jit1, jit2 = ri(3, (len(m),), lo=-2).int(), ri(3, (len(m),), lo=-2).int()
        amp = a1[m]
        s = pt.start[amp]
        e = s + pt.amp_len[amp].int() - 1
        ref_idx[2 * m] = pt.contig[amp]; ref_idx[2 * m + 1] = pt.contig[amp]
        ustart[2 * m] = s + jit1; uend[2 * m] = s + jit1 + read_len - 1
        uend[2 * m + 1] = e + jit2; ustart[2 * m + 1] = e + jit2 - read_len + 1
        flags[2 * m] = S.F_PAIRED | S.F_FIRST | S.F_MATE_NEXT | S.F_FR_PAIR
        flags[2 * m + 1] = S.F_PAIRED | S.F_SECOND | S.F_REVERSE | S.F_FR_PAIR
        # a negative-strand record stores the reverse complement of what was sequenced
        codes[2 * m + 1] = pt.comp[codes[2 * m + 1].flip(1).long()]
    return dict(codes=codes, flags=flags.to(torch.int16), ref_idx=ref_idx, ustart=ustart, uend=uend)


def pack(codes: torch.Tensor) -> torch.Tensor:
    """[n, L] codes -> [n, (L+1)//2] BAM bytes (first base in the high nibble)."""
    n, L = codes.shape
    if L % 2:
        codes = torch.cat([codes, torch.zeros((n, 1), dtype=torch.uint8, device=codes.device)], dim=1)
    return (codes[:, 0::2] << 4) | codes[:, 1::2]


def window_codes(codes: torch.Tensor, flags: torch.Tensor, win: int) -> torch.Tensor:
    """The `win` bases of every read the 5' chain can reach: the first ones, or the last ones of a mapped reverse-strand read
    (what idp_bam_to_batch_window keeps)."""
    f = flags.to(torch.int32)
    from_end = ((f & synth.F_UNMAPPED) == 0) & ((f & synth.F_REVERSE) != 0)
    L = codes.shape[1]
    win = min(win, L)
    return torch.where(from_end[:, None], codes[:, L - win:], codes[:, :win])
