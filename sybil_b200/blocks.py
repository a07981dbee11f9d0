"""Host-side column blocks: the post-gob form of sybil's per-column block files.

`SavedColumn` mirrors SavedIntColumn / SavedStrColumn (src/lib/column_store.go:46-64)
as flat numpy arrays; `encode_block` is the digest step restated in numpy
(SeparateRecordsIntoColumns / SaveIntsToColumns / SaveStrsToColumns,
src/lib/column_store_io.go:64-137,219-303,366-417) so tests can build blocks from
plain rows; `SavedBlock.desc()` builds the sg_block_desc the C ABI takes.
"""
import ctypes as C

import numpy as np

from . import _ffi as F

CARDINALITY_THRESHOLD = 5000  # column_store_io.go:18


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class SavedColumn:
    def __init__(self, col_slot, col_type, encoding=F.SG_ENC_ABSENT):
        self.col_slot = col_slot
        self.col_type = col_type
        self.encoding = encoding
        self.delta_ids = False  # DeltaEncodedIDs
        self.delta_values = False  # ValueEncoded
        self.bin_values = np.zeros(0, np.int64)  # Bins[i].Value
        self.bin_offsets = np.zeros(1, np.uint32)  # offsets into record_ids
        self.record_ids = np.zeros(0, np.uint32)  # concatenated Bins[i].Records
        self.values_i64 = np.zeros(0, np.int64)  # Values (int)
        self.values_i32 = np.zeros(0, np.int32)  # Values (str local ids)
        self.string_table = []  # StringTable (bytes objects)

    def fill(self, d):
        d.col_slot = self.col_slot
        d.col_type = self.col_type
        d.encoding = self.encoding
        d.delta_ids = int(self.delta_ids)
        d.delta_values = int(self.delta_values)
        self._keep = []
        if self.encoding == F.SG_ENC_BUCKET:
            self.bin_values = np.ascontiguousarray(self.bin_values, np.int64)
            self.bin_offsets = np.ascontiguousarray(self.bin_offsets, np.uint32)
            # narrow arrays (sybilgpu.h): a uint16 id array is passed as it is (id_bits = 16)
            if np.asarray(self.record_ids).dtype == np.uint16:
                self.record_ids = np.ascontiguousarray(self.record_ids)
                d.id_bits = 16
            else:
                self.record_ids = np.ascontiguousarray(self.record_ids, np.uint32)
            d.nbins = len(self.bin_values)
            d.nrecord_ids = len(self.record_ids)
            d.bin_values = _ptr(self.bin_values)
            d.bin_offsets = _ptr(self.bin_offsets)
            d.record_ids = _ptr(self.record_ids)
        elif self.encoding == F.SG_ENC_VALUES:
            if self.col_type == F.SG_COL_INT:
                # int32 / int16 arrays hold deltas relative to value_base (value_bits = 32 / 16)
                dt = np.asarray(self.values_i64).dtype
                if dt == np.int32 or dt == np.int16:
                    self.values_i64 = np.ascontiguousarray(self.values_i64)
                    d.value_bits = 32 if dt == np.int32 else 16
                    d.value_base = int(getattr(self, "value_base", 0))
                else:
                    self.values_i64 = np.ascontiguousarray(self.values_i64, np.int64)
                d.nvalues = len(self.values_i64)
                d.values_i64 = _ptr(self.values_i64)
            else:
                if np.asarray(self.values_i32).dtype == np.uint16:
                    self.values_i32 = np.ascontiguousarray(self.values_i32)
                    d.value_bits = 16
                else:
                    self.values_i32 = np.ascontiguousarray(self.values_i32, np.int32)
                d.nvalues = len(self.values_i32)
                d.values_i32 = _ptr(self.values_i32)
        if self.col_type == F.SG_COL_SET:
            d.nvalues = int(getattr(self, "set_nvalues", 0))  # len(Values) of the non-bucketed file form
        if self.col_type in (F.SG_COL_STR, F.SG_COL_SET):
            blob = b"".join(self.string_table)
            offs = np.zeros(len(self.string_table) + 1, np.uint32)
            if self.string_table:
                offs[1:] = np.cumsum([len(s) for s in self.string_table])
            self._blob = np.frombuffer(blob if blob else b"\0", dtype=np.uint8).copy()
            self._offs = offs
            d.ndict = len(self.string_table)
            d.dict_bytes = _ptr(self._blob)
            d.dict_offsets = _ptr(self._offs)


class SavedBlock:
    def __init__(self, block_index, num_records):
        self.block_index = block_index
        self.num_records = num_records
        self.cols = []  # SavedColumn
        self.info = {}  # col_slot -> (min, max): the block's info.db IntInfoMap

    def desc(self):
        n = len(self.cols)
        self._cols = (F.sg_column_desc * max(n, 1))()
        for i, c in enumerate(self.cols):
            c.fill(self._cols[i])
        self._info = (F.sg_int_info * max(len(self.info), 1))()
        for i, (slot, (mn, mx)) in enumerate(sorted(self.info.items())):
            self._info[i].col_slot = slot
            self._info[i].min = mn
            self._info[i].max = mx
        d = F.sg_block_desc()
        d.block_index = self.block_index
        d.num_records = self.num_records
        d.ncols = n
        d.cols = C.cast(self._cols, C.POINTER(F.sg_column_desc))
        d.ninfo = len(self.info)
        d.info = C.cast(self._info, C.POINTER(F.sg_int_info))
        self._desc = d
        return d


def _bucket_encode(codes, valid, col):
    """value -> ascending row ids, gaps delta-encoded (delta_encode_col, column_store_io.go:21-30)."""
    rows = np.nonzero(valid)[0].astype(np.int64)
    vals = codes[rows]
    order = np.lexsort((rows, vals))
    rows_s, vals_s = rows[order], vals[order]
    uniq, starts = np.unique(vals_s, return_index=True)
    offsets = np.append(starts, len(rows_s)).astype(np.uint32)
    gaps = rows_s.copy()
    gaps[1:] -= rows_s[:-1]
    gaps[starts] = rows_s[starts]  # first id of every bin is absolute
    col.encoding = F.SG_ENC_BUCKET
    col.delta_ids = True
    col.bin_values = uniq.astype(np.int64)
    col.bin_offsets = offsets
    col.record_ids = gaps.astype(np.uint32)


def encode_int_column(col_slot, values, valid, threshold=CARDINALITY_THRESHOLD):
    """SaveIntsToColumns (column_store_io.go:64-137) for one block."""
    values = np.asarray(values, np.int64)
    valid = np.asarray(valid, bool)
    col = SavedColumn(col_slot, F.SG_COL_INT)
    if not valid.any():
        return None  # the digest writes no file for a column no row has
    if len(np.unique(values[valid])) <= threshold:
        _bucket_encode(values, valid, col)
    else:
        max_r = int(np.nonzero(valid)[0][-1]) + 1
        v = np.where(valid[:max_r], values[:max_r], 0).astype(np.int64)
        d = v.copy()
        d[1:] = (v[1:].astype(np.uint64) - v[:-1].astype(np.uint64)).astype(np.int64)
        col.encoding = F.SG_ENC_VALUES
        col.delta_values = True
        col.values_i64 = d
    return col


def encode_str_column(col_slot, strings, valid, threshold=CARDINALITY_THRESHOLD):
    """SaveStrsToColumns (column_store_io.go:219-303): per-block ids in first-seen order."""
    valid = np.asarray(valid, bool)
    col = SavedColumn(col_slot, F.SG_COL_STR)
    if not valid.any():
        return None
    table, ids = {}, np.zeros(len(strings), np.int64)
    for i, s in enumerate(strings):
        if valid[i]:
            b = s if isinstance(s, bytes) else str(s).encode()
            ids[i] = table.setdefault(b, len(table))
    col.string_table = list(table.keys())
    if len(table) <= threshold:
        _bucket_encode(ids, valid, col)
    else:
        max_r = int(np.nonzero(valid)[0][-1]) + 1
        col.encoding = F.SG_ENC_VALUES
        col.values_i32 = np.where(valid[:max_r], ids[:max_r], 0).astype(np.int32)
    return col


def encode_set_column(col_slot, sets, valid, threshold=CARDINALITY_THRESHOLD):
    """SaveSetsToColumns (column_store_io.go:139-217) for one block.  sets: per row a list of tags (strings).
    At most `threshold` distinct tags: Bins[tag] = rows holding it.  More: the file holds Values [][]int32 (one id
    list per row up to the last populated row); the C ABI takes that form as bins too, plus len(Values) in
    `nvalues` (sybilgpu.h) — `set_values_to_bins` below is that conversion."""
    valid = np.asarray(valid, bool)
    col = SavedColumn(col_slot, F.SG_COL_SET)
    table, rows_of = {}, {}
    max_r = 0
    for r, tags in enumerate(sets):
        if not valid[r]:
            continue
        for tag in tags:
            b = tag if isinstance(tag, bytes) else str(tag).encode()
            k = table.setdefault(b, len(table))
            rows_of.setdefault(k, []).append(r)
            max_r = max(max_r, r + 1)
    if not table:
        return None  # a row with an empty set never reaches same_sets: no file without any tag
    col.string_table = list(table.keys())
    col.encoding = F.SG_ENC_BUCKET
    col.delta_ids = True
    vals, offs, ids = [], [0], []
    for k in sorted(rows_of):
        rows = np.asarray(sorted(set(rows_of[k])), np.int64)  # (a tag listed twice in one row counts once)
        gaps = rows.copy()
        gaps[1:] -= rows[:-1]
        vals.append(k)
        ids.append(gaps)
        offs.append(offs[-1] + len(rows))
    col.bin_values = np.asarray(vals, np.int64)
    col.bin_offsets = np.asarray(offs, np.uint32)
    col.record_ids = np.concatenate(ids).astype(np.uint32)
    if len(table) > threshold:
        col.set_nvalues = max_r  # was written as Values[:max_r]: every row below max_r reads back as populated
    return col


def set_values_to_bins(col_slot, values, string_table):
    """SavedSetColumn in its non-bucketed form (Values [][]int32) as the bins the C ABI takes: what a host does
    with a set_*.db file whose BucketEncoded is false."""
    col = SavedColumn(col_slot, F.SG_COL_SET)
    col.string_table = [s if isinstance(s, bytes) else str(s).encode() for s in string_table]
    rows_of = {}
    for r, tags in enumerate(values):
        for k in (tags or []):
            rows_of.setdefault(int(k), []).append(r)
    col.encoding = F.SG_ENC_BUCKET
    col.delta_ids = True
    vals, offs, ids = [], [0], []
    for k in sorted(rows_of):
        rows = np.asarray(sorted(set(rows_of[k])), np.int64)
        gaps = rows.copy()
        gaps[1:] -= rows[:-1]
        vals.append(k)
        ids.append(gaps)
        offs.append(offs[-1] + len(rows))
    col.bin_values = np.asarray(vals, np.int64)
    col.bin_offsets = np.asarray(offs, np.uint32)
    col.record_ids = (np.concatenate(ids) if ids else np.zeros(0)).astype(np.uint32)
    col.set_nvalues = len(values)
    return col


def encode_block(block_index, num_records, columns, threshold=CARDINALITY_THRESHOLD):
    """columns: list of (col_slot, col_type, values, valid).  Returns a SavedBlock."""
    blk = SavedBlock(block_index, num_records)
    for slot, typ, values, valid in columns:
        if valid is None:
            valid = np.ones(num_records, bool)
        if typ == F.SG_COL_INT:
            c = encode_int_column(slot, values, valid, threshold)
            v = np.asarray(values, np.int64)[np.asarray(valid, bool)]
            if len(v):
                blk.info[slot] = (int(v.min()), int(v.max()))
        elif typ == F.SG_COL_SET:
            c = encode_set_column(slot, values, valid, threshold)
        else:
            c = encode_str_column(slot, values, valid, threshold)
        if c is not None:
            blk.cols.append(c)
    return blk


def decode_column(col, num_records):
    """unpackIntCol / unpackStrCol in numpy (column_store_io.go:493-609,690-780): returns
    (values int64[num_records], populated bool[num_records]).  Test helper."""
    vals = np.zeros(num_records, np.int64)
    pop = np.zeros(num_records, bool)
    if col.encoding == F.SG_ENC_BUCKET:
        for b in range(len(col.bin_values)):
            ids = col.record_ids[col.bin_offsets[b]:col.bin_offsets[b + 1]].astype(np.int64)
            rows = np.cumsum(ids) if col.delta_ids else ids
            vals[rows] = col.bin_values[b]
            pop[rows] = True
    elif col.encoding == F.SG_ENC_VALUES:
        if col.col_type == F.SG_COL_INT:
            v = np.asarray(col.values_i64).astype(np.int64).astype(np.uint64)  # (narrow deltas sign-extend)
            if len(v) and np.asarray(col.values_i64).dtype != np.int64:
                v[0] += np.uint64(int(getattr(col, "value_base", 0)) & 0xFFFFFFFFFFFFFFFF)
            v = np.cumsum(v, dtype=np.uint64).astype(np.int64) if col.delta_values else v.astype(np.int64)
        else:
            v = col.values_i32.astype(np.int64)
        vals[:len(v)] = v
        pop[:len(v)] = True
    return vals, pop


def narrow_column(col):
    """The same column with its arrays as narrow as their values allow — what a decoder that keeps gob's
    varints narrow hands over (include/sybilgpu.h, sg_column_desc::id_bits / value_bits): uint16 record ids,
    int16 / int32 value deltas on top of value_base, uint16 local string ids."""
    import copy
    c = copy.copy(col)
    if c.encoding == F.SG_ENC_BUCKET:
        ids = np.asarray(c.record_ids)
        if len(ids) == 0 or int(ids.max()) < 65536:
            c.record_ids = ids.astype(np.uint16)
    elif c.encoding == F.SG_ENC_VALUES:
        if c.col_type == F.SG_COL_INT:
            d = np.asarray(c.values_i64, np.int64)
            if c.delta_values and len(d):
                rest = d[1:]
                lo, hi = (int(rest.min()), int(rest.max())) if len(rest) else (0, 0)
                dt = np.int16 if -32768 <= lo and hi <= 32767 else (np.int32 if -2**31 <= lo and hi < 2**31 else None)
                if dt is not None:
                    nd = np.zeros(len(d), dt)
                    nd[1:] = rest
                    c.values_i64, c.value_base = nd, int(d[0])
        else:
            ids = np.asarray(c.values_i32)
            if len(ids) == 0 or (int(ids.min()) >= 0 and int(ids.max()) < 65536):
                c.values_i32 = ids.astype(np.uint16)
    return c


def narrow_block(blk):
    """A SavedBlock with every column in its narrow form (see narrow_column)."""
    nb = SavedBlock(blk.block_index, blk.num_records)
    nb.info = dict(blk.info)
    nb.cols = [narrow_column(c) for c in blk.cols]
    return nb
