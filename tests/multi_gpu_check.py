"""Multi-GPU parity check (run under torchrun on N GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank stages its shard of the blocks (sybil_b200.sharding.shard_range) — once with identically
seeded dictionaries, once with per-rank dictionaries (exchanged inside sg_query_allreduce) — scans on its GPU, merges with sg_query_allreduce (NCCL) and rank 0 compares the merged
result with the CPU oracle run over ALL blocks: the same bit-exact / tolerance contract as the
single-GPU parity tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from sybil_b200 import _ffi as F
from sybil_b200 import engine as E
from sybil_b200.sharding import shard_range
from tests.util import Q, compare, random_spec, run_oracle


def seed(table, spec):
    """Every rank interns the same strings / int values in the same order before staging."""
    strs = {"host": ["h%d" % i for i in range(5)], "state": ["s%d" % i for i in range(12)]}
    for name, vals in strs.items():
        b = [v.encode() for v in vals]
        offs = np.zeros(len(b) + 1, np.uint32)
        offs[1:] = np.cumsum([len(x) for x in b])
        blob = np.frombuffer(b"".join(b), np.uint8)
        table.ctx.check(table.lib.sg_table_dict_seed_str(table.h, spec.KeyTable[name], blob.ctypes.data, offs.ctypes.data, len(b)))
    ages = np.arange(10, 30, dtype=np.int64)
    table.ctx.check(table.lib.sg_table_dict_seed_int(table.h, spec.KeyTable["age"], ages.ctypes.data, len(ages)))


def run_mode(ctx, spec, queries, first, count, rank, world, seeded, quiet=False):
    """seeded=False: every rank interns strings / int keys in the order ITS shard shows them, so the
    ranks' slot spaces (and time axes) differ and sg_query_allreduce exchanges the dictionaries."""
    table = E.Table("mg", spec.key_table, ctx)
    table.IntInfo = dict(spec.IntInfo)
    if seeded:
        seed(table, spec)
    for b in spec.blocks[first:first + count]:
        table.add_block(b)
    ok = True
    for qi, q in enumerate(queries):
        qs = q.query_spec()
        ls = table.NewLoadSpec()
        for c in ("age", "lat", "big", "time"):
            ls.Int(c)
        for c in ("host", "state"):
            ls.Str(c)
        table.LoadAndQueryRecords(ls, qs, allreduce=True)
        if rank == 0:
            o = run_oracle(spec, q)
            try:
                compare(qs, o, q)
                if not quiet:
                    print("query %d (%s dictionaries): merged result over %d GPUs == oracle over all blocks (%d groups, %d matched)" % (
                        qi, "seeded" if seeded else "per-rank", world, len(qs.Results), qs.MatchedCount))
            except AssertionError as e:
                ok = False
                print("query %d (%s dictionaries) MISMATCH: %r" % (qi, "seeded" if seeded else "per-rank", e), file=sys.stderr)
    table.close()
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = E.get_context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    spec = random_spec(123, nrows=40000, block_rows=2500)  # 16 blocks
    queries = [
        Q(spec, groups=["host"], aggs=["age", "lat", "big"], op="avg"),
        Q(spec, int_filters=[("age", "gt", 12)], str_filters=[("state", "neq", "s3")], groups=["host", "age"], aggs=["lat"], op="hist"),
        Q(spec, groups=["state"], aggs=["big"], op="hist", loghist=True),
        Q(spec, groups=["host"], aggs=["lat"], op="hist", time_col="time", time_bucket=600),
    ]
    first, count = shard_range(len(spec.blocks), rank, world)
    ok = True
    for seeded in (True, False):
        ok &= run_mode(ctx, spec, queries, first, count, rank, world, seeded)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
