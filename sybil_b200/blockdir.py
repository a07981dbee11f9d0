"""sybil block directories on disk <-> `SavedBlock` (the flat arrays the C ABI takes).

A block directory holds one gob file per column, named by the column's type and name, and `info.db`:
    <block>/int_<col>.db   gob(SavedIntColumn)    column_store.go:46-54, column_store_io.go:117-134
    <block>/str_<col>.db   gob(SavedStrColumn)    column_store.go:56-64, column_store_io.go:280-299
    <block>/set_<col>.db   gob(SavedSetColumn)    column_store.go:66-74, column_store_io.go:139-217
    <block>/info.db        gob(SavedColumnInfo)   column_store.go:39-44, column_store_io.go:308-358
(`.db.gz` is accepted like `file_decoder.go:35-53` does.)  `read_block_dir` is what
`LoadBlockFromDir` + `unpack*Col` do up to the point where the reference starts scattering values
into rows (`table_block_io.go:225-310`): open the files the LoadSpec names, gob-decode them, keep
the arrays encoded.  `write_block_dir` is the matching writer (the save half of the digest step,
`SaveIntsToColumns` / `SaveStrsToColumns` / `SaveInfoToColumns`), used here to lay generated blocks
out the way sybil does.

The gob layer underneath (`gob.py`) is pinned to Go's own output by the reference's golden files;
the struct layouts below follow the reference's type declarations field for field.  No Go toolchain
exists in this image, so files written here have not been read back by sybil itself.
"""
import gzip
import os

import numpy as np

from . import _ffi as F
from . import gob
from .blocks import SavedBlock, SavedColumn

# Go type declarations as gob descriptors (field order = declaration order)
INT_BUCKET = ("struct", "SavedIntBucket", [("Value", "int"), ("Records", ("slice", "uint"))])
STR_BUCKET = ("struct", "SavedStrBucket", [("Value", "int"), ("Records", ("slice", "uint"))])
INT_COLUMN = ("struct", "SavedIntColumn", [("Name", "string"), ("DeltaEncodedIDs", "bool"), ("ValueEncoded", "bool"),
                                            ("BucketEncoded", "bool"), ("Bins", ("slice", INT_BUCKET)),
                                            ("Values", ("slice", "int")), ("VERSION", "int")])
STR_COLUMN = ("struct", "SavedStrColumn", [("Name", "string"), ("DeltaEncodedIDs", "bool"), ("BucketEncoded", "bool"),
                                            ("Bins", ("slice", STR_BUCKET)), ("Values", ("slice", "int")),
                                            ("StringTable", ("slice", "string")), ("VERSION", "int")])
SET_BUCKET = ("struct", "SavedSetBucket", [("Value", "int"), ("Records", ("slice", "uint"))])
SET_COLUMN = ("struct", "SavedSetColumn", [("Name", "string"), ("Bins", ("slice", SET_BUCKET)),
                                            ("Values", ("slice", ("slice", "int"))), ("StringTable", ("slice", "string")),
                                            ("DeltaEncodedIDs", "bool"), ("BucketEncoded", "bool"), ("VERSION", "int")])
INT_INFO = ("struct", "IntInfo", [("Min", "int"), ("Max", "int"), ("Avg", "float"), ("M2", "float"), ("Count", "int")])
STR_INFO = ("struct", "StrInfo", [("TopStringCount", ("map", "int", "int")), ("Cardinality", "int")])
COLUMN_INFO = ("struct", "SavedColumnInfo", [("NumRecords", "int"), ("StrInfoMap", ("map", "string", STR_INFO)),
                                              ("IntInfoMap", ("map", "string", INT_INFO))])


def _read(path):
    for p, op in ((path, open), (path + ".gz", gzip.open)):
        if os.path.exists(p):
            with op(p, "rb") as f:
                return f.read()
    return None


def _bins(col, bins):
    values, offs, ids = [], [0], []
    for b in bins:
        values.append(b.get("Value", 0))
        ids.extend(b.get("Records", []))
        offs.append(len(ids))
    col.bin_values = np.asarray(values, np.int64)
    col.bin_offsets = np.asarray(offs, np.uint32)
    col.record_ids = np.asarray(ids, np.uint32)


def int_column_from_gob(col_slot, v):
    """SavedIntColumn (decoded gob value) -> SavedColumn.  Zero-valued gob fields are absent."""
    c = SavedColumn(col_slot, F.SG_COL_INT)
    if v.get("BucketEncoded", False):
        c.encoding = F.SG_ENC_BUCKET
        c.delta_ids = bool(v.get("DeltaEncodedIDs", False))
        _bins(c, v.get("Bins", []))
    else:
        c.encoding = F.SG_ENC_VALUES
        c.delta_values = bool(v.get("ValueEncoded", False))
        c.values_i64 = np.asarray(v.get("Values", []), np.int64)
    return c


def str_column_from_gob(col_slot, v):
    c = SavedColumn(col_slot, F.SG_COL_STR)
    c.string_table = [s.encode("utf-8", "surrogateescape") for s in v.get("StringTable", [])]
    if v.get("BucketEncoded", False):
        c.encoding = F.SG_ENC_BUCKET
        c.delta_ids = bool(v.get("DeltaEncodedIDs", False))
        _bins(c, v.get("Bins", []))
    else:
        c.encoding = F.SG_ENC_VALUES
        c.values_i32 = np.asarray(v.get("Values", []), np.int32)
    return c


def set_column_from_gob(col_slot, v):
    """SavedSetColumn -> SavedColumn in the bucket form the C ABI takes (sybilgpu.h): the non-bucketed file form
    (Values [][]int32) is turned into bins, len(Values) kept as set_nvalues."""
    from .blocks import set_values_to_bins
    table = v.get("StringTable", [])
    if v.get("BucketEncoded", False):
        c = SavedColumn(col_slot, F.SG_COL_SET)
        c.string_table = [s.encode("utf-8", "surrogateescape") for s in table]
        c.encoding = F.SG_ENC_BUCKET
        c.delta_ids = bool(v.get("DeltaEncodedIDs", False))
        _bins(c, v.get("Bins", []))
        return c
    return set_values_to_bins(col_slot, v.get("Values", []), [s.encode("utf-8", "surrogateescape") for s in table])


_PREFIX = {F.SG_COL_INT: "int", F.SG_COL_STR: "str", F.SG_COL_SET: "set"}


def read_block_dir(dirname, key_table, columns=None, block_index=0):
    """The block at `dirname` as a SavedBlock.  key_table: [(name, SG_COL_INT|SG_COL_STR)] (the table's
    KeyTable/KeyTypes); columns: names to load (the LoadSpec; default all).  A column whose file is
    missing stays absent, as in the reference (`table_block_io.go:271-277`)."""
    raw = _read(os.path.join(dirname, "info.db"))
    if raw is None:
        raise FileNotFoundError(os.path.join(dirname, "info.db"))
    info = gob.decode(raw)
    blk = SavedBlock(block_index, int(info.get("NumRecords", 0)))
    slot = {n: i for i, (n, _) in enumerate(key_table)}
    for name, ii in info.get("IntInfoMap", {}).items():
        if name in slot:
            blk.info[slot[name]] = (int(ii.get("Min", 0)), int(ii.get("Max", 0)))
    for i, (name, typ) in enumerate(key_table):
        if columns is not None and name not in columns:
            continue
        raw = _read(os.path.join(dirname, "%s_%s.db" % (_PREFIX[typ], name)))
        if raw is None:
            continue
        v = gob.decode(raw)
        blk.cols.append({F.SG_COL_INT: int_column_from_gob, F.SG_COL_STR: str_column_from_gob,
                         F.SG_COL_SET: set_column_from_gob}[typ](i, v))
    return blk


def _bins_to_gob(c):
    out = []
    offs = c.bin_offsets
    for b in range(len(c.bin_values)):
        out.append({"Value": int(c.bin_values[b]), "Records": c.record_ids[int(offs[b]):int(offs[b + 1])].tolist()})
    return out


def column_to_gob(c, name):
    """SavedColumn -> (descriptor, value) of the SavedIntColumn / SavedStrColumn sybil would write."""
    v = {"Name": name, "VERSION": 1}
    if c.col_type == F.SG_COL_INT:
        if c.encoding == F.SG_ENC_BUCKET:
            v.update(BucketEncoded=True, DeltaEncodedIDs=bool(c.delta_ids), Bins=_bins_to_gob(c))
        else:
            v.update(ValueEncoded=bool(c.delta_values), Values=np.asarray(c.values_i64).tolist())
        return INT_COLUMN, v
    v["StringTable"] = [s.decode("utf-8", "surrogateescape") for s in c.string_table]
    if c.col_type == F.SG_COL_SET:
        nv = int(getattr(c, "set_nvalues", 0))
        if nv:  # written un-bucketed (more than CARDINALITY_THRESHOLD tags, column_store_io.go:183-192)
            rows = [[] for _ in range(nv)]
            for b in range(len(c.bin_values)):
                ids = np.asarray(c.record_ids[int(c.bin_offsets[b]):int(c.bin_offsets[b + 1])], np.int64)
                for r in (np.cumsum(ids) if c.delta_ids else ids):
                    rows[int(r)].append(int(c.bin_values[b]))
            v.update(DeltaEncodedIDs=True, Values=rows)
        else:
            v.update(BucketEncoded=True, DeltaEncodedIDs=bool(c.delta_ids), Bins=_bins_to_gob(c))
        return SET_COLUMN, v
    if c.encoding == F.SG_ENC_BUCKET:
        v.update(BucketEncoded=True, DeltaEncodedIDs=bool(c.delta_ids), Bins=_bins_to_gob(c))
    else:
        v["Values"] = np.asarray(c.values_i32).tolist()
    return STR_COLUMN, v


def write_block_dir(dirname, blk, key_table, compress=False):
    """Lay `blk` out as a sybil block directory (column files + info.db)."""
    os.makedirs(dirname, exist_ok=True)
    op = gzip.open if compress else open
    ext = ".gz" if compress else ""
    int_info, str_info = {}, {}
    for c in blk.cols:
        if c.encoding == F.SG_ENC_ABSENT:
            continue
        name, typ = key_table[c.col_slot]
        t, v = column_to_gob(c, name)
        with op(os.path.join(dirname, "%s_%s.db%s" % (_PREFIX[typ], name, ext)), "wb") as f:
            f.write(gob.encode(v, t))
        if typ == F.SG_COL_STR:
            str_info[name] = {"Cardinality": len(c.string_table)}
    for s, (mn, mx) in blk.info.items():
        int_info[key_table[s][0]] = {"Min": int(mn), "Max": int(mx)}
    with op(os.path.join(dirname, "info.db" + ext), "wb") as f:
        f.write(gob.encode({"NumRecords": int(blk.num_records), "StrInfoMap": str_info, "IntInfoMap": int_info}, COLUMN_INFO))
