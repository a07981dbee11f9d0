"""A whole sybil table directory (table info.db + block directories) read back and queried through the oracle."""
import os

import numpy as np

from sybil_b200 import tabledir
from tests.util import INT, STR, Q, Spec, run_oracle


class _Sink(Spec):
    def add_block(self, blk):
        self.blocks.append(blk)


def test_table_directory_round_trip(tmp_path):
    rng = np.random.default_rng(31)
    n = 5000
    s = Spec([("age", INT), ("lat", INT), ("host", STR)])
    s.add_rows({"age": rng.integers(10, 30, n), "lat": rng.integers(30, 9000, n),
                "host": np.array(["h%d" % x for x in rng.integers(0, 6, n)])}, {"age": rng.random(n) > 0.05}, block_rows=1200)
    tdir = tabledir.write_table(str(tmp_path), "t1", s.key_table, s.blocks, s.IntInfo)
    # directories the reference does not treat as blocks (table_io.go:213-239)
    for junk in ("ingest", "cache", "stomache123", "block00001.partial", "x.broken", "y.old"):
        os.makedirs(os.path.join(tdir, junk))
    info = tabledir.read_table(str(tmp_path), "t1")
    assert info.name == "t1" and info.key_table == s.key_table and info.IntInfo == s.IntInfo
    assert [os.path.basename(d) for d in info.block_dirs] == ["block%05d" % i for i in range(len(s.blocks))]
    back = _Sink(info.key_table)
    back.IntInfo = dict(info.IntInfo)
    assert tabledir.load_blocks(info, back) == n
    q = dict(int_filters=[("age", "gt", 12)], groups=["host"], aggs=["lat"], op="hist")
    a, b = run_oracle(s, Q(s, **q)), run_oracle(back, Q(back, **q))
    assert a.MatchedCount == b.MatchedCount and set(a.Results) == set(b.Results)
    for k in a.Results:
        assert a.Results[k].Count == b.Results[k].Count
        assert np.array_equal(a.Results[k].Hists["lat"].Values, b.Results[k].Hists["lat"].Values)
        assert a.Results[k].Hists["lat"].Percentiles == b.Results[k].Hists["lat"].Percentiles
    # a LoadSpec naming one column; a corrupt block directory is skipped like LoadBlockFromDir returning nil
    open(os.path.join(info.block_dirs[1], "info.db"), "wb").write(b"\x03junk")
    part = _Sink(info.key_table)
    assert tabledir.load_blocks(info, part, columns={"host"}) == n - s.blocks[1].num_records
    assert all([c.col_slot for c in blk.cols] == [2] for blk in part.blocks)
