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
        # narrow arrays (uint16 ids, int32/int16 value deltas: sybilgpu.h) unless SG_WIDE=1 asks for Go's decoded types
        self.narrow = not os.environ.get("SG_WIDE")
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
        s.narrow = 1 if self.narrow else 0
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


# ---- ground truth straight from the row values (csrc/blockgen.cpp sbg_eval) --------------------------
class Expected:
    """The config's query evaluated on the generator's row values with plain loops: no block encoding,
    no decode, no oracle.  `groups`: {GroupByKey: slot}; `times`: list of time bucket starts (time mode);
    arrays indexed [slot] / [agg][slot] / [agg][slot][bucket] with slot = group slot + time index * ngroup_slots."""

    def __init__(self, spec, row0=0, row1=None, nthreads=None):
        q = query_for(spec)
        ci = {c.name: i for i, c in enumerate(spec.cols)}
        ev = F.sbg_eval_spec()
        fl = [(c, op, v) for c, op, v in q.get("int_filters", [])]
        for c, op, v in q.get("str_filters", []):
            col = spec.cols[ci[c]]
            assert v.startswith(col.prefix)
            fl.append((c, op, int(v[len(col.prefix):])))
        ev.nfilters = len(fl)
        for i, (c, op, v) in enumerate(fl):
            ev.filter_col[i], ev.filter_op[i], ev.filter_val[i] = ci[c], F.OPS[op], v
        gcols = [spec.cols[ci[g]] for g in q.get("groups", [])]
        ev.ngroups = len(gcols)
        for i, c in enumerate(gcols):
            ev.group_col[i] = ci[c.name]
        self.aggs = list(q.get("aggs", []))
        ev.naggs = len(self.aggs)
        self.hist = q.get("op") == "hist"
        self.nvals = 0
        for i, a in enumerate(self.aggs):
            ev.agg_col[i] = ci[a]
            mn, mx = spec.IntInfo[a]
            ev.info_min[i], ev.info_max[i] = mn, mx
            if self.hist:  # BasicHist.SetupBuckets (hist_basic.go:34-70), default bucket count
                size = mx - mn
                bs, nb = size // 1000, 1000  # NumBuckets stays 1000 unless the bucket size came out 0
                if bs == 0:
                    bs, nb = (1, size) if size < 100 else (size // 100, size // (size // 100))
                ev.bsize[i] = bs
                assert self.nvals in (0, nb + 2), "one bucket layout per query in this checker"
                self.nvals = nb + 2  # len(Values) = NumBuckets + 1 + 1
        ev.nvals = self.nvals
        ev.time_col = -1
        self.times = [None]
        if q.get("time_col"):
            tb = q["time_bucket"]
            mn, mx = spec.IntInfo[q["time_col"]]
            ev.time_col, ev.time_bucket, ev.time_first, ev.time_n = ci[q["time_col"]], tb, mn // tb, mx // tb - mn // tb + 1
            self.times = [(mn // tb + k) * tb for k in range(ev.time_n)]
        gslots = 1
        for c in gcols:
            gslots *= c.span
        self.gslots = gslots
        slots = gslots * len(self.times)
        na = max(len(self.aggs), 1)
        self.count = np.zeros(slots, np.uint64)
        self.hcount = np.zeros((na, slots), np.uint64)
        self.sum = np.zeros((na, slots), np.uint64)
        self.buckets = np.zeros((na, slots, max(self.nvals, 1)), np.uint64)
        cs = spec.c_spec()
        row1 = spec.total_rows if row1 is None else row1
        self.matched = F.gen().sbg_eval(C.byref(cs), C.byref(ev), row0, row1, nthreads or (os.cpu_count() or 1),
                                        self.count.ctypes.data, self.hcount.ctypes.data, self.sum.ctypes.data,
                                        self.buckets.ctypes.data)
        if self.matched < 0:
            raise ValueError("sbg_eval: row outside the time axis / unsupported column kind")
        self.gcols = gcols

    def key_of(self, gslot):
        """GroupByKey the engine renders for a group slot (translate_group_by, aggregate.go:284-324)."""
        if not self.gcols:
            return "total"
        parts = []
        for c in self.gcols:
            v = c.lo + gslot % c.span
            gslot //= c.span
            parts.append((c.prefix + str(v)) if c.col_type == STR else str(v))
        return "\t".join(parts) + "\t"

    def check(self, qs, max_groups=None):
        """Compare a filled QuerySpec (engine result) with the row-value evaluation: MatchedCount, the set of
        groups, every Count, hist Count, exact sum and bucket counter.  Returns the number of values compared."""
        n = 0
        assert qs.MatchedCount == self.matched, ("MatchedCount", qs.MatchedCount, self.matched)
        n += 1
        time_mode = self.times[0] is not None
        per_key = self.count.reshape(len(self.times), self.gslots).sum(0)
        live = [g for g in range(self.gslots) if per_key[g]]
        assert len(qs.Results) == len(live), ("number of groups", len(qs.Results), len(live))
        for g in live[:max_groups]:
            k = self.key_of(g)
            r = qs.Results.get(k)
            assert r is not None, ("missing group", k)
            assert r.Count == int(per_key[g]), ("Count", k, r.Count, int(per_key[g]))
            n += 1
            if not time_mode:
                n += self._check_hists(r, g, k)
        if time_mode:
            live_t = [ti for ti in range(len(self.times)) if self.count[ti * self.gslots:(ti + 1) * self.gslots].any()]
            assert sorted(qs.TimeResults) == [self.times[ti] for ti in live_t], "time buckets"
            for ti in live_t:
                m = qs.TimeResults[self.times[ti]]
                for g in range(self.gslots):
                    s = ti * self.gslots + g
                    if not self.count[s]:
                        continue
                    r = m.get(self.key_of(g))
                    assert r is not None and r.Count == int(self.count[s]), ("time Count", self.times[ti], g)
                    n += 1 + self._check_hists(r, s, (self.times[ti], g))
        return n

    def _check_hists(self, r, s, where):
        n = 0
        for ai, a in enumerate(self.aggs):
            h = r.Hists.get(a)
            if not self.hcount[ai, s]:
                assert h is None or h.TotalCount() == 0, ("unexpected hist", where, a)
                continue
            assert h is not None, ("missing hist", where, a)
            assert h.TotalCount() == int(self.hcount[ai, s]), ("hist Count", where, a)
            assert h.Sum() == int(self.sum[ai, s].astype(np.int64)), ("sum", where, a)
            n += 2
            if self.hist:
                assert np.array_equal(np.asarray(h.Values, np.uint64), self.buckets[ai, s]), ("bucket counters", where, a)
                n += self.nvals
        return n
