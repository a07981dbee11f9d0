"""world_size-2 test of the N>1 host logic on CPU (gloo): block sharding + merge of per-rank dense
partials by one all-reduce must equal the single-process result (CombineResults, aggregate.go:414-467).
The per-rank scan is the CPU oracle here; on GPUs it is the CUDA kernel and the all-reduce is NCCL
(sg_query_allreduce), exercised by bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sybil_b200.sharding import dense_layout, shard_range
from tests.util import Q, random_spec, run_oracle


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def dense(o, layout, width, aggs, nvalues):
    m = np.zeros((len(layout), width), np.int64)
    for k, r in o.Results.items():
        row = m[layout[k]]
        row[0] = r.Count
        for ai, a in enumerate(aggs):
            h = r.Hists.get(a)
            if h is None:
                continue
            base = 1 + ai * (2 + nvalues)
            row[base], row[base + 1] = h.Count, h.ExactSum
            row[base + 2:base + 2 + len(h.Values)] = h.Values
    return m


def worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = random_spec(77, nrows=6000, block_rows=500)  # 12 blocks
        q = Q(spec, int_filters=[("age", "gt", 12)], groups=["host", "state"], aggs=["lat"], op="hist")
        full_blocks = list(spec.blocks)
        first, count = shard_range(len(full_blocks), rank, world)
        spec.blocks = full_blocks[first:first + count]
        o = run_oracle(spec, q)
        keys = [None] * world
        dist.all_gather_object(keys, sorted(o.Results))
        layout, width = dense_layout([k for ks in keys for k in ks], 1, 1002)
        m = torch.from_numpy(dense(o, layout, width, q.aggs, 1002))
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
        matched = torch.tensor([o.MatchedCount])
        dist.all_reduce(matched)
        if rank == 0:
            spec.blocks = full_blocks
            ref = run_oracle(spec, q)
            layout_ref, _ = dense_layout(sorted(ref.Results), 1, 1002)
            ok = layout_ref == layout and np.array_equal(m.numpy(), dense(ref, layout, width, q.aggs, 1002)) \
                and int(matched.item()) == ref.MatchedCount
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_all_blocks_once():
    for n in (0, 1, 7, 1526, 15259):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                f, c = shard_range(n, r, w)
                seen += list(range(f, f + c))
            assert seen == list(range(n))


def test_two_rank_merge_equals_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
