#!/usr/bin/env python
"""bench.py — rows/sec scanned and HBM GB/s of the group-by scan (BASELINE.json metric).

A "step" is one pass of the hot path (decode -> filter -> group-by -> aggregate ->
CombineResults) over one synthetic table.  At N=1 the workload is BASELINE.json
configs[1]: 100M rows, group-by 1 str column, sum+avg on 3 int columns ("c2").
`--workload c3|c4|c5` selects the other configs (they are parity-test cases and
scaling runs, not the default bench line).  With N>1 every rank scans its own
100M-row shard of an N x 100M-row table (weak scaling) and the per-group partials
are merged by one NCCL all-reduce inside the timed step.

value   rows/s with the encoded blocks already resident in HBM when the timed
        region starts (inputs 2.8 GB per GPU >> 126 MB of L2: no flush needed).
e2e     the same metric through the C ABI from pinned HOST buffers: every step
        re-stages all blocks (H2D inside the timed region), scans, and reads the
        result back.
roofline  algorithmic bytes (8 B per int column, 4 B per str column referenced,
        SURVEY.md §8d) / scan-kernel time measured with CUDA events by the library.
cpu_baseline  the CPU oracle (restatement of the reference's goroutine-per-block
        path; the Go reference cannot be built here) on a bounded sample, all host threads.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--rows", type=int, default=0, help="rows per GPU (default: the config's size, capped at 1e9/N for c3-c5)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def rows_per_gpu(args, spec_rows):
    if args.rows:
        return args.rows
    if args.workload == "c2":
        return 100_000_000  # weak scaling: 100M rows per GPU
    return spec_rows // max(args.gpus, 1) if args.gpus > 1 else spec_rows


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def seed_dicts(table, spec, F):
    """Identical key numbering on every rank (sg_table_dict_seed_*): the synthetic columns'
    value sets are known, so every rank seeds them in numeric order."""
    import numpy as np
    for c in spec.cols:
        if c.kind == F.SBG_STRKEY:
            strs = [(c.prefix + str(c.lo + v)).encode() for v in range(c.span)]
            offs = np.zeros(len(strs) + 1, np.uint32)
            offs[1:] = np.cumsum([len(s) for s in strs])
            blob = np.frombuffer(b"".join(strs), np.uint8)
            table.ctx.check(table.lib.sg_table_dict_seed_str(table.h, c.col_slot, blob.ctypes.data, offs.ctypes.data, len(strs)))
        elif c.kind == F.SBG_UNIFORM and c.span <= 5000:
            vals = np.arange(c.lo, c.lo + c.span, dtype=np.int64)
            table.ctx.check(table.lib.sg_table_dict_seed_int(table.h, c.col_slot, vals.ctypes.data, len(vals)))


def make_query(spec, synth, E):
    from tests.util import Q, Spec
    s = Spec(spec.key_table)
    s.IntInfo = dict(spec.IntInfo)
    return Q(s, **synth.query_for(spec))


def run_ours(args):
    import torch
    from sybil_b200 import _ffi as F
    from sybil_b200 import engine as E
    from sybil_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus
    if world != N and world > 1:
        N = world
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = E.get_context(local)
    lib = ctx.lib
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    base = synth.config(args.workload)
    per_gpu = rows_per_gpu(args, base.total_rows)
    spec = synth.config(args.workload, total_rows=per_gpu * world)
    nb_total = spec.num_blocks()
    from sybil_b200.sharding import shard_range
    first, nblocks = shard_range(nb_total, rank, world)
    bytes_per_row = synth.algorithmic_bytes_per_row(spec)

    # ---- inputs: generated on the host cores into pinned memory (untimed) ----------------
    # Tables whose encoded form exceeds CHUNK_LIMIT are generated and staged chunk by chunk
    # through one reusable pinned arena (the e2e leg, which needs every block in host memory,
    # is then skipped and reported as null).
    t0 = time.time()
    per_block_bytes = spec.block_rows * (bytes_per_row + 8) + (1 << 21)
    CHUNK_LIMIT = 12 << 30
    chunked = nblocks * per_block_bytes > CHUNK_LIMIT
    chunk_blocks = max(1, min(nblocks, CHUNK_LIMIT // per_block_bytes)) if chunked else nblocks
    arena_bytes = chunk_blocks * per_block_bytes + (1 << 20)
    arena = lib.sg_pinned_alloc(ctx.h, arena_bytes)
    if not arena:
        raise RuntimeError("pinned arena: " + ctx.err())
    table = E.Table(args.workload, spec.key_table, ctx)
    table.IntInfo = dict(spec.IntInfo)
    if world > 1:
        seed_dicts(table, spec, F)
    gen_s = stage_s = 0.0
    my_rows = 0
    store = None
    for c0 in range(0, nblocks, chunk_blocks):
        if store is not None:
            store.close()
        tg = time.time()
        nb = min(chunk_blocks, nblocks - c0)
        store = synth.generate(spec, first + c0, nb, arena_ptr=arena, arena_bytes=arena_bytes)
        gen_s += time.time() - tg
        ts = time.time()
        for i in range(nb):
            my_rows += store.block(i).contents.num_records
        ptrs, np_ = store.block_ptrs()
        table.add_blocks(ptrs, np_)
        table.sync()
        stage_s += time.time() - ts
    if chunked:
        args.no_e2e = True
    q = make_query(spec, synth, E)
    q.set_flags()

    def one_step(tbl, materialize=False):
        qs = q.query_spec()
        qs.materialize = materialize  # the C library always builds the full sorted result; the per-group
                                      # Python objects are only built for the checked step
        ls = tbl.NewLoadSpec()
        for c in spec.cols:
            (ls.Int if c.col_type == F.SG_COL_INT else ls.Str)(c.name)
        tbl.LoadAndQueryRecords(ls, qs, allreduce=world > 1)
        return qs

    # ---- value: resident inputs -----------------------------------------------------------
    for _ in range(args.warmup):
        qs = one_step(table)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        qs = one_step(table)
        kernel_ms += qs.stats.kernel_ms
        launches += qs.stats.kernel_launches
    barrier()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.finish() if sampler else None
    total_rows = sum_over_ranks(float(my_rows))
    ms_per_step = elapsed / args.steps * 1e3
    value = total_rows / (elapsed / args.steps)
    kernel_ms_avg = max_over_ranks(kernel_ms / args.steps)
    matched, ngroups = qs.MatchedCount, qs.NumGroups

    # ---- e2e: host buffers -> H2D -> scan -> result, every step ---------------------------
    e2e = None
    if not args.no_e2e:
        t2 = E.Table(args.workload + "_e2e", spec.key_table, ctx)
        t2.IntInfo = dict(spec.IntInfo)
        if world > 1:
            seed_dicts(t2, spec, F)

        e2e_ptrs, e2e_n = store.block_ptrs()

        def e2e_step():
            ctx.check(lib.sg_table_clear(t2.h))
            t2.add_blocks(e2e_ptrs, e2e_n)  # host buffers -> HBM inside the timed step
            r = one_step(t2)
            return r, lib.sg_table_h2d_bytes(t2.h)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r2, h2d = e2e_step()
        barrier()
        e_el = max_over_ranks(time.perf_counter() - t0)
        assert r2.MatchedCount == matched
        e2e = {"value": total_rows / (e_el / args.e2e_steps), "unit": "rows/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(r2.stats.d2h_bytes), "steps": args.e2e_steps,
               "ms_per_step": e_el / args.e2e_steps * 1e3}
        t2.close()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline ---------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = my_rows * bytes_per_row / (kernel_ms_avg * 1e-3) / 1e9 if kernel_ms_avg > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "sg::scan_kernel",
                "kernel_ms_per_launch": kernel_ms_avg, "algorithmic_bytes_per_row": bytes_per_row,
                "rows_per_launch": my_rows, "encoded_bytes_resident": int(lib.sg_table_encoded_bytes(table.h))}

    # ---- cpu baseline (rank 0, N = 1) -------------------------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_baseline(spec, store, q, store.num_blocks())

    out = {
        "metric": "rows/sec scanned (group-by sum+hist scan)", "value": value, "unit": "rows/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, WORKLOADS.get(args.workload, "")),
                   "rows_per_gpu": int(my_rows), "total_rows": int(total_rows), "blocks_per_gpu": nblocks,
                   "block_rows": spec.block_rows, "parallelism": "block-sharded x%d + 1 NCCL all-reduce" % world,
                   "l2": "inputs (%.1f GB per GPU) larger than L2; no flush" % (lib.sg_table_encoded_bytes(table.h) / 1e9),
                   "groups": ngroups, "matched_rows": int(matched)},
        "hbm_gbps": achieved * world, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "setup": {"generate_s": gen_s, "stage_s": stage_s},
    }
    print(json.dumps(out))
    table.close()
    store.close()
    if dist:
        dist.destroy_process_group()


WORKLOADS = {
    "c2": "100M rows/GPU, group-by 1 str col (64 keys), sum+avg on 3 int cols",
    "c3": "1B rows, 3 ANDed int/str filters, group-by 2 cols, BasicHist on 1 col",
    "c4": "1B rows time series, 256 time buckets + per-bucket BasicHist",
    "c5": "1B rows, group-by 1M distinct str keys, sum on 4 int cols",
}


def cpu_baseline(spec, store, q, nblocks, target_s=15.0):
    """The oracle (CPU restatement of the reference path) on a bounded sample of the same blocks."""
    from oracle.oracle_ffi import OracleTable, lib as olib
    threads = olib().orc_hardware_threads()
    d, keep = q.desc()
    probe = min(nblocks, max(threads, 8))
    ot = OracleTable(spec.key_table)
    for i in range(probe):
        ot.add_block(store.block(i))
    r = ot.query(d, q.aggs, nthreads=threads, details=False)
    rows_probe = sum(store.block(i).contents.num_records for i in range(probe))
    rate = rows_probe / max(r.seconds, 1e-6)
    want = int(min(nblocks, max(probe, rate * target_s / spec.block_rows)))
    for i in range(probe, want):
        ot.add_block(store.block(i))
    r = ot.query(d, q.aggs, nthreads=threads, details=False)
    rows = sum(store.block(i).contents.num_records for i in range(want))
    ot.close()
    return {"value": rows / r.seconds, "unit": "rows/s", "cores": threads, "kind": "port",
            "sample": "first %d blocks (%d rows) of the same table, %.1f s" % (want, rows, r.seconds)}


def run_reference(args):
    """--impl reference: the reference's CPU path (its C++ restatement: no Go toolchain in this
    image) on the host cores, same workload/metric, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sybil_b200 import synth
    from oracle.oracle_ffi import OracleTable, lib as olib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = synth.config(args.workload)
    per_gpu = rows_per_gpu(args, base.total_rows)
    spec = synth.config(args.workload, total_rows=per_gpu * world)
    threads = olib().orc_hardware_threads()
    q = make_query(spec, synth, None)
    q.set_flags()
    d, keep = q.desc()
    # calibrate a sample that keeps warmup+steps within a few minutes
    probe = min(spec.num_blocks(), max(threads, 8))
    store = synth.generate(spec, 0, probe)
    ot = OracleTable(spec.key_table)
    for i in range(probe):
        ot.add_block(store.block(i))
    r = ot.query(d, q.aggs, nthreads=threads, details=False)
    rate = probe * spec.block_rows / max(r.seconds, 1e-6)
    budget_s = 120.0 / max(args.steps + args.warmup, 1)
    want = int(min(spec.num_blocks(), max(probe, rate * min(budget_s, 10.0) / spec.block_rows)))
    if want > probe:
        store.close()
        ot.close()
        store = synth.generate(spec, 0, want)
        ot = OracleTable(spec.key_table)
        for i in range(want):
            ot.add_block(store.block(i))
    rows = sum(store.block(i).contents.num_records for i in range(want))
    for _ in range(args.warmup):
        ot.query(d, q.aggs, nthreads=threads, details=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ot.query(d, q.aggs, nthreads=threads, details=False)
    el = time.perf_counter() - t0
    value = rows * args.steps / el
    sample = "first %d blocks (%d rows) of the table per step" % (want, rows)
    out = {"impl": "reference", "metric": "rows/sec scanned (group-by sum+hist scan)", "value": value, "unit": "rows/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": {"workload": "%s: %s" % (args.workload, WORKLOADS.get(args.workload, "")), "sample": sample,
                      "block_rows": spec.block_rows},
           "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    import __graft_entry__
    __graft_entry__.build()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
