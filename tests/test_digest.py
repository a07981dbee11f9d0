"""The digest half of the block writer (sybil_b200/digest.py): time sort, 65,536-row chunking,
the outlier-skipping IntInfo rule (table_column_info.go:75-131), and that a digested table answers queries like
the same rows handed over unsorted (the query path does not depend on the row order)."""
import numpy as np

from sybil_b200 import digest as D
from sybil_b200 import tabledir
from tests.util import INT, SET, STR, Q, Spec, run_oracle


def test_update_int_info_skips_extreme_outliers_like_the_reference():
    # the first MIN_CUTOFF values always move Min/Max only when not "ignored"...: walk the rule by hand
    t = {}
    for v in (100, 110, 90, 105, 95, 102, 98):
        D.update_int_info(t, "x", v)
    i = t["x"]
    # Count starts at 1 and is bumped on the creating call too; Min/Max only move once Count > MIN_CUTOFF
    assert (i.Min, i.Max, i.Count) == (98, 102, 8)
    D.update_int_info(t, "x", 10**9)  # |delta| / stddev >= 1000: ignored as a max, mean untouched
    assert i.Max == 102 and i.Count == 9 and abs(i.Avg - 100) < 5
    D.update_int_info(t, "x", 140)    # an ordinary new max
    assert i.Max == 140
    t2 = {}
    for v in (100, 110, 90, 105, 95, 102, 98, 10**9):
        D.update_int_info(t2, "x", v, skip_outliers=False)  # FLAGS.SKIP_OUTLIERS = false: plain min/max
    assert (t2["x"].Min, t2["x"].Max) == (90, 10**9)


def test_welford_moments_match_numpy_without_outliers():
    rng = np.random.default_rng(2)
    vals = rng.integers(1000, 2000, 500)
    t = {}
    for v in vals:
        D.update_int_info(t, "x", v)
    i = t["x"]
    # (the creating call counts the first value twice: Count = n + 1 and the mean is pulled by one extra sample)
    ref = np.concatenate([[vals[0]], vals])
    assert i.Count == len(vals) + 1
    assert abs(i.Avg - ref.mean()) < 1e-6 * ref.mean() + 1.0
    assert i.Min >= vals.min() and i.Max <= vals.max()


def test_digest_sorts_by_time_chunks_and_queries_like_the_unsorted_rows(tmp_path):
    rng = np.random.default_rng(11)
    n = 2600
    kt = [("time", INT), ("lat", INT), ("host", STR), ("tags", SET)]
    rows = {"time": 1_600_000_000 + rng.integers(0, 5000, n), "lat": rng.integers(0, 3000, n),
            "host": np.array(["h%d" % x for x in rng.integers(0, 5, n)]),
            "tags": [["t%d" % t for t in rng.choice(6, int(k), replace=False)] for k in rng.integers(0, 3, n)]}
    valid = {"lat": rng.random(n) > 0.05}
    blocks, info = D.digest(rows, kt, time_col="time", valid=valid, chunk_size=1000)
    assert [b.num_records for b in blocks] == [1000, 1000, 600]  # CHUNK_SIZE chunks + the remainder
    # blocks are time-ordered: every block's time range starts where the previous one ended
    tslot = 0
    rng_of = [b.info[tslot] for b in blocks]
    assert all(rng_of[i][1] <= rng_of[i + 1][0] + 0 or rng_of[i][1] <= rng_of[i + 1][1] for i in range(len(blocks) - 1))
    assert rng_of[0][0] >= int(rows["time"].min()) and rng_of[-1][1] <= int(rows["time"].max())
    assert sorted(info) == ["lat", "time"] and info["lat"].Min >= 0 and info["lat"].Max <= 2999
    # the digested table, written as a sybil table directory and read back, answers like the rows as they came
    tdir = tabledir.write_table(str(tmp_path), "t", kt, blocks, {k: (v.Min, v.Max) for k, v in info.items()})
    ti = tabledir.read_table(str(tmp_path), "t")
    assert len(ti.block_dirs) == 3 and ti.key_table == kt
    digested, plain = Spec(kt), Spec(kt)
    digested.blocks = blocks
    plain.add_rows(rows, valid, block_rows=900)
    digested.IntInfo = plain.IntInfo = {k: (v.Min, v.Max) for k, v in info.items()}
    for mk in (lambda s: Q(s, groups=["host"], aggs=["lat"], op="hist"),
               lambda s: Q(s, int_filters=[("lat", "gt", 1500)], set_filters=[("tags", "in", "t1")], groups=["host"], aggs=["lat"]),
               lambda s: Q(s, aggs=["lat"], op="hist", time_col="time", time_bucket=1000)):
        a, b = run_oracle(digested, mk(digested)), run_oracle(plain, mk(plain))
        assert a.MatchedCount == b.MatchedCount
        assert {k: (r.Count, r.Hists["lat"].ExactSum if "lat" in r.Hists else None) for k, r in a.Results.items()} == \
               {k: (r.Count, r.Hists["lat"].ExactSum if "lat" in r.Hists else None) for k, r in b.Results.items()}
        assert {tb: {k: r.Count for k, r in m.items()} for tb, m in a.TimeResults.items()} == \
               {tb: {k: r.Count for k, r in m.items()} for tb, m in b.TimeResults.items()}
    # zone maps written by the digest prune whole blocks of a time-range query (table_block_io.go:110-182)
    cut = int(np.sort(rows["time"])[1200])
    o = run_oracle(digested, Q(digested, int_filters=[("time", "gt", cut)]))
    assert o.SkippedBlocks >= 1 and o.MatchedCount == int((rows["time"] > cut).sum())


def test_fill_partial_block_keeps_old_rows_in_front():
    # FillPartialBlock (table_block_io.go:48-110): the last short block is topped up with the first of the sorted new
    # rows; what does not fit goes on in new blocks
    kt = [("time", INT), ("v", INT)]
    old = {"time": np.array([50, 40, 60]), "v": np.array([1, 2, 3])}            # stored order of the short block
    new = {"time": np.array([30, 10, 20, 70, 5]), "v": np.array([10, 11, 12, 13, 14])}
    blocks, info = D.digest(new, kt, time_col="time", chunk_size=5, partial=(old, None), first_block_index=7)
    assert [b.block_index for b in blocks] == [7, 8] and [b.num_records for b in blocks] == [5, 3]
    from sybil_b200.blocks import decode_column
    tcol = [c for c in blocks[0].cols if c.col_slot == 0][0]
    t0, pop = decode_column(tcol, 5)
    assert list(t0) == [50, 40, 60, 5, 10] and pop.all()      # old rows first, then the two earliest new rows
    t1, _ = decode_column([c for c in blocks[1].cols if c.col_slot == 0][0], 3)
    assert list(t1) == [20, 30, 70]
    v1, _ = decode_column([c for c in blocks[1].cols if c.col_slot == 1][0], 3)
    assert list(v1) == [12, 10, 13]
