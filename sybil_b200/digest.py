"""The digest half of sybil's block writer: what happens between `sybil ingest` and the column
files `blockdir.write_block_dir` lays out — rows sorted by time, cut into 65,536-row blocks, each block encoded per
column, and the IntInfo book-keeping (with its outlier-skipping min/max rule) that the query path later builds
histograms and zone maps from.

    Table.SaveRecordsToColumns   src/lib/table_io.go:119-130      sort.Sort(SortRecordsByTime), FillPartialBlock, saveRecordList
    Table.saveRecordList         src/lib/table_io.go:79-117       CHUNK_SIZE chunks + the remainder
    SortRecordsByTime            src/lib/column_store.go:8-20     by Record.Timestamp (= the time column's value)
    Record.AddIntField           src/lib/record.go:118-128        table.update_int_info per ingested value
    SaveIntsToColumns            src/lib/column_store_io.go:83-95 block + table update_int_info once per DISTINCT value
    update_int_info              src/lib/table_column_info.go:75-131

Host-side and after/before the hot path: numpy + plain Python, no GPU.  Differences from the reference, all forced
by Go's unordered maps / unstable sort: rows with equal timestamps keep their ingestion order (Go's sort.Sort may
permute them); a block's distinct values are fed to update_int_info in ascending order (Go walks a map).  Partial
block filling (FillPartialBlock, table_block_io.go:48-110) appends to the last short block; here `digest` takes
the short block's rows back in front of the new ones, which is what the reference's reload + append amounts to.
"""
import math

import numpy as np

from . import _ffi as F
from .blocks import CARDINALITY_THRESHOLD, encode_block

CHUNK_SIZE = F.SG_BLOCK_ROWS  # table.go:44
STD_CUTOFF = 1000.0  # table_column_info.go:72
MIN_CUTOFF = 5       # table_column_info.go:73


class IntInfo:  # table_column_info.go:18-24
    __slots__ = ("Min", "Max", "Avg", "M2", "Count")

    def __init__(self):
        self.Min = self.Max = 0
        self.Avg = self.M2 = 0.0
        self.Count = 0


def update_int_info(table, name, val, skip_outliers=True):
    """table_column_info.go:75-131, statement for statement.  table: dict name -> IntInfo."""
    val = int(val)
    info = table.get(name)
    if info is None:
        info = IntInfo()
        table[name] = info
        info.Max = info.Min = val
        info.Avg = float(val)
        info.Count = 1
    delta = float(val) - info.Avg
    # Go: info.M2 / float64(info.Count-1) — with Count == 1 that is x/0: +Inf, -Inf or NaN, never a panic
    d = float(info.Count - 1)
    stddev = info.M2 / d if d != 0 else (math.nan if info.M2 == 0 else math.copysign(math.inf, info.M2))
    if stddev <= 1:  # (false for NaN, as in Go)
        stddev = max(info.Avg, 1.0)
    ignored = False
    if info.Max < val:
        delta_in_stddev = abs(delta) / stddev
        if (delta_in_stddev < STD_CUTOFF and info.Count > MIN_CUTOFF) or not skip_outliers:
            info.Max = val
        else:
            ignored = True
    if info.Min > val:
        delta_in_stddev = abs(delta) / stddev
        if (delta_in_stddev < STD_CUTOFF and info.Count > MIN_CUTOFF) or not skip_outliers:
            info.Min = val
        else:
            ignored = True
    if not ignored or info.Count < MIN_CUTOFF:
        info.Avg = info.Avg + delta / float(info.Count)
        info.M2 = info.M2 + delta * (float(val) - info.Avg)
    info.Count += 1


def digest(rows, key_table, time_col=None, valid=None, table_info=None, threshold=CARDINALITY_THRESHOLD,
           chunk_size=CHUNK_SIZE, skip_outliers=True, first_block_index=0, ingest_info=True, partial=None):
    """rows: {column: values in ingestion order}; valid: {column: bool mask} (a row lacking a column).
    Returns (blocks, table_info): SavedBlocks whose `info` is the block's IntInfoMap (Min, Max), and the table's
    IntInfo (dict name -> IntInfo) after AddIntField (per row, ingestion order; `ingest_info`) and the save-time
    updates.  time_col: the column Record.Timestamp is read from (OPTS.TIME_COL); None leaves the order alone
    (every Timestamp 0: the stable sort is the identity).
    partial: (rows, valid) of the table's last, short block (FillPartialBlock, table_block_io.go:48-110): its rows
    stay in front, in the order they were stored, the time-sorted new rows fill it up to chunk_size and the rest goes
    on in new blocks; the first returned block then REPLACES that block (give its index as first_block_index)."""
    slot = {n: i for i, (n, _) in enumerate(key_table)}
    typ = {n: t for n, t in key_table}
    valid = dict(valid or {})
    n = len(next(iter(rows.values()))) if rows else 0
    table_info = table_info if table_info is not None else {}
    if ingest_info:  # Record.AddIntField while the rows were ingested
        for name, vals in rows.items():
            if typ[name] != F.SG_COL_INT:
                continue
            va = valid.get(name)
            for i in range(n):
                if va is None or va[i]:
                    update_int_info(table_info, name, vals[i], skip_outliers)
    order = np.arange(n)
    if time_col is not None:
        ts = np.asarray(rows[time_col], np.int64)
        if time_col in valid:
            ts = np.where(np.asarray(valid[time_col], bool), ts, 0)  # a record without the column keeps Timestamp 0
        order = np.argsort(ts, kind="stable")
    if partial is not None:
        prow, pvalid = partial
        pn = len(next(iter(prow.values()))) if prow else 0
        if pn >= chunk_size:
            pn, prow = 0, {}
        # one row space: the partial block's rows first (indices 0..pn-1), then the new rows shifted by pn
        rows = {k: (list(prow.get(k, [None] * pn)) + list(v)) if typ[k] == F.SG_COL_SET
                else np.concatenate([np.asarray(prow[k]) if k in prow else np.zeros(pn, np.asarray(v).dtype), np.asarray(v)])
                for k, v in rows.items()}
        merged_valid = {}
        for k in rows:
            a = np.asarray((pvalid or {}).get(k, np.ones(pn, bool) if k in prow else np.zeros(pn, bool)), bool)
            b = np.asarray(valid.get(k, np.ones(n, bool)), bool)
            merged_valid[k] = np.concatenate([a, b])
        valid = merged_valid
        order = np.concatenate([np.arange(pn), order + pn])
        n += pn
    blocks = []
    for start in range(0, n, chunk_size):
        idx = order[start:start + chunk_size]
        cols = []
        for name, vals in rows.items():
            v = [vals[i] for i in idx] if typ[name] == F.SG_COL_SET else np.asarray(vals)[idx]
            va = valid.get(name)
            cols.append((slot[name], typ[name], v, None if va is None else np.asarray(va, bool)[idx]))
        blk = encode_block(first_block_index + len(blocks), len(idx), cols, threshold)
        blk.info = {}
        block_info = {}
        for s, t, v, va in cols:  # SaveIntsToColumns: once per distinct value, block and table (:93-94)
            if t != F.SG_COL_INT:
                continue
            name = key_table[s][0]
            vv = np.asarray(v, np.int64)
            if va is not None:
                vv = vv[va]
            for x in np.unique(vv):
                update_int_info(block_info, name, x, skip_outliers)
                update_int_info(table_info, name, x, skip_outliers)
            if name in block_info:
                blk.info[s] = (block_info[name].Min, block_info[name].Max)
        blocks.append(blk)
    return blocks, table_info
