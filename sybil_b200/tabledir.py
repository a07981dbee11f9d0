"""A sybil table directory: `<dir>/<table>/info.db` + one sub-directory per block.

    info.db = gob(Table{Name, KeyTable map[string]int16, KeyTypes map[int16]int8,
                        IntInfo map[int16]*IntInfo, StrInfo map[int16]*StrInfo})
              (`getSaveTable` / `SaveTableInfo`, src/lib/table_io.go:66-78; loaded by `LoadTableInfo` :128-175)
    blocks  = every sub-directory that `file_looks_like_block` accepts (table_io.go:213-239): not the
              ingest / cache / stomache directories, not names ending in info.db, old, broken, lock, export,
              partial — enumerated by `LoadAndQueryRecords` with ioutil.ReadDir, i.e. in name order
              (table_query.go:22,96-111).

`read_table` gives the key table (the table's KeyTable/KeyTypes by slot), the table-level IntInfo the
histograms take their extents from (`hist.go:27-38`) and the block directories; `load_blocks` feeds them
(through `blockdir.read_block_dir`) to anything with an `add_block` — `engine.Table` (GPU) or the oracle.
`write_table` is the save side for tables generated here.
"""
import os

from . import _ffi as F
from . import blockdir, gob

INT_INFO = blockdir.INT_INFO
STR_INFO = blockdir.STR_INFO
TABLE = ("struct", "Table", [("Name", "string"), ("KeyTable", ("map", "string", "int")), ("KeyTypes", ("map", "int", "int")),
                             ("StrInfo", ("map", "int", STR_INFO)), ("IntInfo", ("map", "int", INT_INFO))])

_NOT_BLOCKS = ("ingest", ".ingest.temp", "cache")
_BAD_SUFFIX = ("info.db", "old", "broken", "lock", "export", "partial")


def looks_like_block(name):
    """file_looks_like_block (table_io.go:213-239)."""
    if name in _NOT_BLOCKS or name.startswith("stomache"):
        return False
    return not name.endswith(_BAD_SUFFIX)


class TableInfo:
    def __init__(self, name, key_table, int_info, block_dirs):
        self.name = name
        self.key_table = key_table      # [(column name, SG_COL_INT | SG_COL_STR | SG_COL_SET | 0)] by slot
        self.IntInfo = int_info         # column name -> (Min, Max)
        self.block_dirs = block_dirs    # absolute paths, in the order the reference visits them


def read_table(dbdir, name):
    tdir = os.path.join(dbdir, name)
    raw = blockdir._read(os.path.join(tdir, "info.db"))
    if raw is None:
        raise FileNotFoundError(os.path.join(tdir, "info.db"))
    t = gob.decode(raw)
    by_slot = {int(s): n for n, s in t.get("KeyTable", {}).items()}
    types = {int(s): int(v) for s, v in t.get("KeyTypes", {}).items()}
    nslots = (max(by_slot) + 1) if by_slot else 0
    # record.go:14-19: INT_VAL = 1, STR_VAL = 2, SET_VAL = 3
    key_table = [(by_slot.get(s, "_unused_%d" % s),
                  types.get(s, 0) if types.get(s, 0) in (F.SG_COL_INT, F.SG_COL_STR, F.SG_COL_SET) else 0)
                 for s in range(nslots)]
    int_info = {by_slot[int(s)]: (int(ii.get("Min", 0)), int(ii.get("Max", 0)))
                for s, ii in t.get("IntInfo", {}).items() if int(s) in by_slot}
    blocks = sorted(e for e in os.listdir(tdir) if os.path.isdir(os.path.join(tdir, e)) and looks_like_block(e))
    return TableInfo(t.get("Name", name), key_table, int_info, [os.path.join(tdir, b) for b in blocks])


def load_blocks(info, sink, columns=None):
    """Stage every block of the table into `sink` (anything with add_block(SavedBlock)); returns the row count.
    A block directory that cannot be decoded is skipped, as LoadBlockFromDir returning nil is
    (table_query.go:134-139)."""
    rows = 0
    info.skipped_blocks = 0
    for i, d in enumerate(info.block_dirs):
        try:
            blk = blockdir.read_block_dir(d, info.key_table, columns, block_index=i)
            if blk.num_records <= 0:
                continue
            # a block the sink rejects (NumRecords above the block size, inconsistent bins, ids that do not fit
            # their array type) is one bad block, not a failed table load
            sink.add_block(blk)
        except (gob.GobError, EOFError, FileNotFoundError, OSError, OverflowError, ValueError, RuntimeError):
            info.skipped_blocks += 1
            continue
        rows += blk.num_records
    return rows


def write_table(dbdir, name, key_table, blocks, int_info, compress=False):
    """Write `<dbdir>/<name>/info.db` and one `block%05d` directory per SavedBlock."""
    tdir = os.path.join(dbdir, name)
    os.makedirs(tdir, exist_ok=True)
    slot = {n: i for i, (n, _) in enumerate(key_table)}
    v = {"Name": name, "KeyTable": {n: i for i, (n, _) in enumerate(key_table)},
         "KeyTypes": {i: int(t) for i, (_, t) in enumerate(key_table)},
         "IntInfo": {slot[n]: {"Min": int(lo), "Max": int(hi)} for n, (lo, hi) in int_info.items()}}
    with open(os.path.join(tdir, "info.db"), "wb") as f:
        f.write(gob.encode(v, TABLE))
    for i, b in enumerate(blocks):
        blockdir.write_block_dir(os.path.join(tdir, "block%05d" % i), b, key_table, compress=compress)
    return tdir
