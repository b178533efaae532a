"""world_size-2 gloo test of the multi-GPU host logic on CPU: template-preserving shards + counter all-reduce.
Each rank runs the checker (oracle) on its shard; the reduced counters must equal a single-process run."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H  # noqa: F401  (sets sys.path)
from fg_idprimer_b200 import sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _oracle_run(lo, hi):
    import bench
    panel, reads, opts = synth.workload("c1", 3000)
    # make some fragments so that shard boundaries are not trivially even
    reads.flags[-5:] &= ~np.uint16(0x2000 | 0x1 | 0x40 | 0x80)
    op = bench.oracle_panel(panel, opts)
    arr = reads.oracle_arrays()
    sub = dict(bases=arr["bases"][lo * 150:hi * 150], off=arr["off"][:hi - lo + 1], ref_idx=arr["ref_idx"][lo:hi], ustart=arr["ustart"][lo:hi],
               uend=arr["uend"][lo:hi], flags=arr["flags"][lo:hi])
    five, _, metrics, n_align = op.process_batch(sub, n_threads=2)
    return reads, five, metrics, n_align


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    panel, reads, opts = synth.workload("c1", 3000)
    reads.flags[-5:] &= ~np.uint16(0x2000 | 0x1 | 0x40 | 0x80)
    lo, hi = sharding.shard_reads(reads.flags, world, rank)
    _, five, metrics, n_align = _oracle_run(lo, hi)
    total, total_align = sharding.reduce_counters(metrics, n_align)
    q.put((rank, lo, hi, total.tolist(), total_align, five["primer"].tolist()))
    dist.destroy_process_group()


def test_template_starts_and_shards_never_split_a_template():
    flags = np.array([0x2000, 0, 0, 0x2000, 0, 0x2000, 0, 0], dtype=np.uint16)  # pair, frag, pair, pair, frag
    assert sharding.template_starts(flags).tolist() == [0, 2, 3, 5, 7]
    for world in (1, 2, 3, 4, 5, 8):
        edges = [sharding.shard_reads(flags, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == len(flags)
        for (a, b), (c, d) in zip(edges, edges[1:]):
            assert b == c
        for lo, hi in edges:
            assert lo == len(flags) or lo in sharding.template_starts(flags).tolist()


def test_two_rank_gloo_reduction_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    reads, five, metrics, n_align = _oracle_run(0, 6000)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 6000
    for r in res:
        assert r[3] == metrics.tolist() and r[4] == n_align
    assert res[0][5] + res[1][5] == five["primer"].tolist()
