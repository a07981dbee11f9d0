"""Synthetic tables of SURVEY.md §8(d) / BASELINE.json configs, in sybil's block format.

Row values are a counter-based function of (seed, column slot, row) — splitmix64 —
so any block can be generated independently, by the C++ generator
(csrc/blockgen.cpp, threaded, used for bench sizes) or by the numpy mirror below
(used by the tests to check the generator).  Table-level IntInfo (the histogram
extents, table_column_info.go:18-24) is the column's value range.
"""
import ctypes as C
import os

import numpy as np

from . import _ffi as F

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


class SynthCol:
    def __init__(self, name, col_slot, col_type, kind, lo=0, span=1, a=0, b=1, prefix="", nulls=0):
        self.name, self.col_slot, self.col_type, self.kind = name, col_slot, col_type, kind
        self.lo, self.span, self.a, self.b, self.prefix, self.nulls = lo, span, a, b, prefix, nulls

    def int_info(self):
        if self.kind == F.SBG_UNIFORM:
            return (self.lo, self.lo + self.span - 1)
        if self.kind == F.SBG_SUM4:
            return (self.lo, self.lo + self.span)
        return None


class SynthSpec:
    def __init__(self, name, key_table, cols, total_rows, seed, block_rows=F.SG_BLOCK_ROWS, threshold=5000):
        self.name, self.key_table, self.cols = name, key_table, cols
        self.total_rows, self.seed, self.block_rows, self.threshold = total_rows, seed, block_rows, threshold
        self.KeyTable = {n: i for i, (n, _) in enumerate(key_table)}
        self.KeyTypes = {i: t for i, (_, t) in enumerate(key_table)}
        self.IntInfo = {}
        for c in cols:
            ii = c.int_info()
            if ii and c.col_type == F.SG_COL_INT:
                self.IntInfo[c.name] = ii

    def num_blocks(self):
        return (self.total_rows + self.block_rows - 1) // self.block_rows

    def cell(self, c, row):
        """(value, valid) of one cell — pure-Python mirror of blockgen.cpp col_value()."""
        u = splitmix64(self.seed ^ (c.col_slot << 40) ^ row)
        valid = True
        if c.nulls > 0:
            valid = (splitmix64(u) & 1023) >= c.nulls
        if c.kind in (F.SBG_UNIFORM, F.SBG_STRKEY):
            v = c.lo + u % c.span
        elif c.kind == F.SBG_SUM4:
            s = (u & 0xffff) + ((u >> 16) & 0xffff) + ((u >> 32) & 0xffff) + ((u >> 48) & 0xffff)
            v = c.lo + (s * c.span) // (4 * 65535)
        else:
            v = c.lo + (row * c.a) // c.b + (u % c.span if c.span > 0 else 0)
        return v, valid

    def c_spec(self):
        arr = (F.sbg_col * len(self.cols))()
        for i, c in enumerate(self.cols):
            arr[i].col_slot, arr[i].col_type, arr[i].kind, arr[i].null_per_1024 = c.col_slot, c.col_type, c.kind, c.nulls
            arr[i].lo, arr[i].span, arr[i].a, arr[i].b = c.lo, c.span, c.a, c.b
            arr[i].prefix = c.prefix.encode()
        s = F.sbg_spec()
        s.seed, s.total_rows, s.block_rows, s.ncols = self.seed, self.total_rows, self.block_rows, len(self.cols)
        s.cardinality_threshold, s.num_col_slots = self.threshold, len(self.key_table)
        s.cols = C.cast(arr, C.POINTER(F.sbg_col))
        self._keep = arr
        return s


INT, STR = F.SG_COL_INT, F.SG_COL_STR


def config(name, total_rows=None, block_rows=F.SG_BLOCK_ROWS, seed=None):
    """The benchmark tables.  Only the columns a config's query references are generated
    (sybil loads only the files named in the LoadSpec, table_block_io.go:271-277)."""
    name = name.lower()
    if name == "c2":  # 100M rows: group-by 1 str col (64 values), avg on 3 int cols
        kt = [("g", STR), ("i0", INT), ("i1", INT), ("i2", INT)]
        cols = [SynthCol("g", 0, STR, F.SBG_STRKEY, 0, 64, prefix="k")] + [
            SynthCol("i%d" % i, 1 + i, INT, F.SBG_UNIFORM, 0, 1000000) for i in range(3)]
        rows, sd = 100_000_000, 0x5EB11 + 2
    elif name == "c3":  # 1B rows: 3 filters, group-by 2, BasicHist on lat
        kt = [("f0", INT), ("f1", INT), ("s0", STR), ("s1", STR), ("d", INT), ("lat", INT)]
        cols = [SynthCol("f0", 0, INT, F.SBG_UNIFORM, 0, 1000), SynthCol("f1", 1, INT, F.SBG_UNIFORM, 0, 1000000),
                SynthCol("s0", 2, STR, F.SBG_STRKEY, 0, 16, prefix="v"),
                SynthCol("s1", 3, STR, F.SBG_STRKEY, 0, 12, prefix="g"),
                SynthCol("d", 4, INT, F.SBG_UNIFORM, 0, 8), SynthCol("lat", 5, INT, F.SBG_SUM4, 30, 23470)]
        rows, sd = 1_000_000_000, 0x5EB11 + 3
    elif name == "c4":  # 1B rows time series: 256 buckets of B seconds + hist on lat
        B = 3600
        rows = total_rows or 1_000_000_000
        kt = [("time", INT), ("lat", INT)]
        cols = [SynthCol("time", 0, INT, F.SBG_TIME, 1_500_000_000 // B * B, B // 4, a=256 * B - B // 4, b=rows),
                SynthCol("lat", 1, INT, F.SBG_SUM4, 30, 23470)]
        sd = 0x5EB11 + 4
    elif name == "c5":  # high-cardinality group-by: 1M distinct str keys, sum on 4 int cols
        kt = [("k", STR), ("m0", INT), ("m1", INT), ("m2", INT), ("m3", INT)]
        cols = [SynthCol("k", 0, STR, F.SBG_STRKEY, 0, 1000000, prefix="key")] + [
            SynthCol("m%d" % i, 1 + i, INT, F.SBG_UNIFORM, 0, 10000) for i in range(4)]
        rows, sd = 1_000_000_000, 0x5EB11 + 5
    else:
        raise ValueError("unknown config %r" % name)
    spec = SynthSpec(name, kt, cols, total_rows or rows, seed if seed is not None else sd, block_rows)
    if name == "c4":
        B = 3600
        t0 = 1_500_000_000 // B * B
        spec.IntInfo["time"] = (t0, t0 + 256 * B - 1)
        spec.time_bucket = B
    return spec


def query_for(spec):
    """The config's query in the reference's vocabulary: (int_filters, str_filters, groups, aggs, op, time)."""
    n = spec.name
    if n == "c2":
        return dict(groups=["g"], aggs=["i0", "i1", "i2"], op="avg")
    if n == "c3":
        return dict(int_filters=[("f0", "gt", 100), ("f1", "lt", 900000)], str_filters=[("s0", "neq", "v3")],
                    groups=["s1", "d"], aggs=["lat"], op="hist")
    if n == "c4":
        return dict(aggs=["lat"], op="hist", time_col="time", time_bucket=spec.time_bucket)
    if n == "c5":
        return dict(groups=["k"], aggs=["m0", "m1", "m2", "m3"], op="avg")
    raise ValueError(n)


# bytes per row of SURVEY.md §8(d): 8 per distinct int column referenced, 4 per str column
def algorithmic_bytes_per_row(spec):
    return sum(8 if c.col_type == INT else 4 for c in spec.cols)


class Store:
    """Blocks [first, first+n) of a table, generated by the C++ generator into one arena."""

    def __init__(self, spec, h, arena_keep):
        self.spec, self.h, self._arena = spec, h, arena_keep
        self.g = F.gen()

    def num_blocks(self):
        return self.g.sbg_num_blocks(self.h)

    def block(self, i):
        return self.g.sbg_block(self.h, i)

    def block_ptrs(self):
        """(ctypes array of sg_block_desc*, n) for sg_table_add_blocks."""
        n = self.num_blocks()
        arr = (C.POINTER(F.sg_block_desc) * max(n, 1))()
        for i in range(n):
            arr[i] = self.block(i)
        return arr, n

    def encoded_bytes(self):
        return self.g.sbg_encoded_bytes(self.h)

    def close(self):
        if self.h:
            self.g.sbg_free(self.h)
            self.h = None


def generate(spec, first_block=0, nblocks=None, nthreads=None, arena_ptr=None, arena_bytes=None):
    g = F.gen()
    cs = spec.c_spec()
    total = spec.num_blocks()
    if nblocks is None:
        nblocks = total - first_block
    nthreads = nthreads or max(1, (os.cpu_count() or 1))
    if arena_bytes is None:
        per_row = sum(8 if c.col_type == INT else 4 for c in spec.cols) + 8
        arena_bytes = nblocks * spec.block_rows * per_row + nblocks * (1 << 21) + (1 << 20)
    h = g.sbg_generate(C.byref(cs), first_block, nblocks, nthreads, arena_ptr, arena_bytes)
    if not h:
        raise MemoryError("block generator arena too small")
    return Store(spec, h, arena_ptr)
