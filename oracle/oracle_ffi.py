"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  It returns plain Python structures shaped like the
engine's results so the parity tests can compare field by field.
"""
import ctypes as C
import os

import numpy as np

from sybil_b200 import _ffi as F

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
P = C.c_void_p
I64P = C.POINTER(C.c_int64)

SYMBOLS = {
    "orc_table_create": (P, [C.c_int32, C.POINTER(C.c_int32)]),
    "orc_table_free": (None, [P]),
    "orc_table_add_block": (C.c_int, [P, C.POINTER(F.sg_block_desc)]),
    "orc_table_set_str_replace": (C.c_int, [P, C.c_int32, C.c_char_p, C.c_char_p]),
    "orc_query": (P, [P, C.POINTER(F.sg_query_desc), C.c_int, C.c_int64]),
    "orc_result_free": (None, [P]),
    "orc_result_seconds": (C.c_double, [P]),
    "orc_result_matched_count": (C.c_int64, [P]),
    "orc_result_num_groups": (C.c_int64, [P]),
    "orc_result_num_broken": (C.c_int64, [P]),
    "orc_result_num_skipped": (C.c_int64, [P]),
    "orc_result_num_time_buckets": (C.c_int64, [P]),
    "orc_result_time_bucket": (C.c_int64, [P, C.c_int64]),
    "orc_result_time_num_groups": (C.c_int64, [P, C.c_int64]),
    "orc_result_group": (C.c_int, [P, C.c_int64, C.c_int64, C.POINTER(P), I64P, I64P, I64P]),
    "orc_result_hist": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int32, I64P, I64P, I64P, I64P, C.POINTER(C.c_double),
                                  I64P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), I64P]),
    "orc_result_hist_values": (C.c_int64, [P, C.c_int64, C.c_int64, C.c_int32, I64P, C.c_int64]),
    "orc_result_percentiles": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int32, I64P]),
    "orc_result_stddev": (C.c_double, [P, C.c_int64, C.c_int64, C.c_int32]),
    "orc_result_sparse_buckets": (C.c_int64, [P, C.c_int64, C.c_int64, C.c_int32, I64P, I64P, C.c_int64]),
    "orc_result_set_merged_view": (None, [P, C.c_int]),
    "orc_basic_layout": (None, [C.c_int64, C.c_int64, C.c_int32, I64P, I64P, I64P]),
    "orc_multi_layout": (C.c_int64, [C.c_int64, C.c_int64, I64P, C.c_int64]),
    "orc_basic_combine": (None, [C.c_int64, C.c_int64, C.c_int64, I64P, C.POINTER(C.c_double), I64P, C.c_int64, I64P,
                                 C.POINTER(C.c_double), I64P]),
    "orc_basic_percentiles": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, I64P, C.c_int64, I64P]),
    "orc_hardware_threads": (C.c_int, []),
}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/liboracle.so is not built (make -C oracle)")
        _lib = F.bind(C.CDLL(LIB_PATH), SYMBOLS)
    return _lib


class OHist:
    pass


class OResult:
    def __init__(self):
        self.GroupByKey, self.Count, self.Samples, self.Hists = "", 0, 0, {}


class OQuery:
    def __init__(self):
        self.Results, self.Sorted, self.TimeResults = {}, [], {}
        self.Cumulative, self.MatchedCount, self.BrokenBlocks, self.SkippedBlocks, self.seconds = None, 0, 0, 0, 0.0


class OracleTable:
    """The oracle's copy of a table: blocks are deep-copied from the same sg_block_desc."""

    def __init__(self, key_table):
        self.lib = lib()
        types = (C.c_int32 * len(key_table))(*[t for _, t in key_table])
        self.h = self.lib.orc_table_create(len(key_table), types)

    def add_block(self, blk):
        d = blk.desc() if hasattr(blk, "desc") else blk
        self.lib.orc_table_add_block(self.h, C.byref(d) if not isinstance(d, C._Pointer) else d)

    def set_str_replace(self, col_slot, pattern, replacement):
        """FLAGS.STR_REPLACE for one column (pattern None: off)."""
        rc = self.lib.orc_table_set_str_replace(self.h, col_slot, None if pattern is None else pattern.encode(),
                                                None if replacement is None else replacement.encode())
        if rc != 0:
            raise ValueError("bad pattern %r" % pattern)

    def close(self):
        if self.h:
            self.lib.orc_table_free(self.h)
            self.h = None

    def _group(self, r, tb, gi, agg_names):
        L = self.lib
        kb, kl, cnt, smp = P(), C.c_int64(), C.c_int64(), C.c_int64()
        if L.orc_result_group(r, tb, gi, C.byref(kb), C.byref(kl), C.byref(cnt), C.byref(smp)) != 0:
            return None
        o = OResult()
        o.GroupByKey = C.string_at(kb, kl.value).decode("utf-8", "replace")
        o.Count, o.Samples = cnt.value, smp.value
        for ai, name in enumerate(agg_names):
            c, s, mn, mx, sm = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
            avg = C.c_double()
            nb, bs, nv, ns = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            no = C.c_int64()
            rc = L.orc_result_hist(r, tb, gi, ai, C.byref(c), C.byref(s), C.byref(mn), C.byref(mx), C.byref(avg),
                                   C.byref(sm), C.byref(nb), C.byref(bs), C.byref(nv), C.byref(ns), C.byref(no))
            if rc != 1:
                continue
            h = OHist()
            h.Count, h.ExactSum, h.Min, h.Max, h.Avg, h.Samples = c.value, s.value, mn.value, mx.value, avg.value, sm.value
            h.NumBuckets, h.BucketSize, h.nsubhists, h.noutliers = nb.value, bs.value, ns.value, no.value
            vals = np.zeros(max(nv.value, 1), np.int64)
            n = L.orc_result_hist_values(r, tb, gi, ai, vals.ctypes.data_as(I64P), len(vals))
            h.Values = vals[:n].copy()
            # groups whose first-seen block result still holds Outliers (Q9): read the derived
            # values from the merged view (fresh clone + Combine), which is what the reference
            # reports whenever another block's result happens to be adopted first
            L.orc_result_set_merged_view(r, 1 if h.noutliers else 0)
            p = (C.c_int64 * 100)()
            n = L.orc_result_percentiles(r, tb, gi, ai, p)
            h.Percentiles = [p[i] for i in range(n)]
            h.StdDev = L.orc_result_stddev(r, tb, gi, ai)
            n = L.orc_result_sparse_buckets(r, tb, gi, ai, None, None, 0)
            e, cc = (C.c_int64 * max(n, 1))(), (C.c_int64 * max(n, 1))()
            L.orc_result_sparse_buckets(r, tb, gi, ai, e, cc, n)
            h.IntBuckets = {e[i]: cc[i] for i in range(n)}
            L.orc_result_set_merged_view(r, 0)
            o.Hists[name] = h
        return o

    def query(self, desc, agg_names, nthreads=1, max_blocks=0, details=True):
        """LoadAndQueryRecords on the CPU.  desc is the same sg_query_desc the GPU path gets."""
        L = self.lib
        r = L.orc_query(self.h, C.byref(desc), nthreads, max_blocks)
        q = OQuery()
        try:
            q.seconds = L.orc_result_seconds(r)
            q.MatchedCount = L.orc_result_matched_count(r)
            q.BrokenBlocks = L.orc_result_num_broken(r)
            q.SkippedBlocks = L.orc_result_num_skipped(r)
            if not details:
                return q
            q.Cumulative = self._group(r, -1, -1, agg_names)
            for gi in range(L.orc_result_num_groups(r)):
                g = self._group(r, -1, gi, agg_names)
                q.Sorted.append(g)
                q.Results[g.GroupByKey] = g
            for b in range(L.orc_result_num_time_buckets(r)):
                tb = L.orc_result_time_bucket(r, b)
                m = {}
                for gi in range(L.orc_result_time_num_groups(r, b)):
                    g = self._group(r, b, gi, agg_names)
                    m[g.GroupByKey] = g
                q.TimeResults[tb] = m
        finally:
            L.orc_result_free(r)
        return q
