"""Multi-GPU host logic: the path shards by template (reads of one template stay together, IdentifyPrimers.scala:318-328
batches are independent), the panel is replicated, and the only exchange is a sum of the 160 TemplateTypeMetric
counters plus the alignment counter (IdentifyPrimers.scala:424-426).  Works with any torch.distributed backend."""
from __future__ import annotations

from typing import Tuple

import numpy as np

F_MATE_NEXT = 0x2000


def template_starts(flags: np.ndarray) -> np.ndarray:
    """Index of the first read of every template (R1, or a fragment)."""
    flags = np.asarray(flags)
    is_r2 = np.zeros(len(flags), dtype=bool)
    is_r2[1:] = (flags[:-1] & F_MATE_NEXT) != 0
    return np.nonzero(~is_r2)[0]


def shard_reads(flags: np.ndarray, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) read range of `rank`: templates are dealt in equal contiguous blocks, never split."""
    starts = template_starts(flags)
    n_t = len(starts)
    t_lo, t_hi = (n_t * rank) // world, (n_t * (rank + 1)) // world
    lo = int(starts[t_lo]) if t_lo < n_t else len(flags)
    hi = int(starts[t_hi]) if t_hi < n_t else len(flags)
    return lo, hi


def reduce_counters(metrics160: np.ndarray, n_alignments: int, device=None, group=None):
    """Sum of the metric histogram and alignment count over all ranks (the only cross-GPU traffic of the path)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(161, dtype=torch.int64, device=device)
    buf[:160] = torch.from_numpy(np.ascontiguousarray(metrics160, dtype=np.int64)).to(buf.device)
    buf[160] = int(n_alignments)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = buf.cpu().numpy()
    return out[:160].copy(), int(out[160])
