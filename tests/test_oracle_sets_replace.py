"""The oracle's SetFilter / set-column decode (filter.go:252-285, column_store_io.go:611-688) and StrReplace
(column_store_io.go:515-549) against a row-by-row evaluation of the raw rows in plain Python: the reference holds
no test vector for either, so the restatement is pinned to its definition (parity otherwise unpinned)."""
import re
from collections import Counter

import numpy as np

from sybil_b200 import blocks as B
from sybil_b200 import _ffi as F
from tests.util import INT, SET, STR, Q, Spec, run_oracle


def _table(seed=5, nrows=4000, block_rows=1500, threshold=5000):
    rng = np.random.default_rng(seed)
    s = Spec([("v", INT), ("host", STR), ("tags", SET)])
    ntag = rng.integers(0, 4, nrows)
    rows = {
        "v": rng.integers(0, 1000, nrows),
        "host": np.array(["web-%02d.dc%d" % (a, b) for a, b in zip(rng.integers(0, 12, nrows), rng.integers(1, 4, nrows))]),
        "tags": [["t%d" % t for t in rng.choice(9, int(k), replace=False)] for k in ntag],
    }
    valid = {"host": rng.random(nrows) > 0.1}
    s.add_rows(rows, valid, threshold=threshold, block_rows=block_rows)
    return s, rows, valid


def test_set_filters_in_nin_absent_literal():
    s, rows, valid = _table()
    for op, tag in (("in", "t3"), ("nin", "t3"), ("in", "nope"), ("nin", "nope")):
        q = Q(s, set_filters=[("tags", op, tag)], groups=["host"], aggs=["v"])
        o = run_oracle(s, q)
        want = Counter()
        for i in range(len(rows["v"])):
            tags = rows["tags"][i]
            ok = (tag in tags) if op == "in" else (len(tags) > 0 and tag not in tags)  # no set: both ops say false
            if ok:
                want[(rows["host"][i] if valid["host"][i] else "") + "\t"] += 1
        assert o.MatchedCount == sum(want.values()), (op, tag)
        assert {k: r.Count for k, r in o.Results.items()} == dict(want), (op, tag)
    # two set filters and an int filter together
    q = Q(s, int_filters=[("v", "lt", 500)], set_filters=[("tags", "in", "t1"), ("tags", "nin", "t2")])
    o = run_oracle(s, q)
    n = sum(1 for i in range(len(rows["v"])) if rows["v"][i] < 500 and "t1" in rows["tags"][i] and "t2" not in rows["tags"][i])
    assert o.MatchedCount == n and n > 0


def test_set_column_values_form_marks_listed_rows_populated():
    # the non-bucketed file form: every row below len(Values) is populated, even with an empty set
    s = Spec([("v", INT), ("tags", SET)])
    n = 50
    vals = [[0], [], [1, 0], None, [1]] * 6  # 30 listed rows, 20 beyond len(Values)
    blk = B.SavedBlock(0, n)
    blk.cols.append(B.encode_int_column(0, np.arange(n), np.ones(n, bool)))
    blk.cols.append(B.set_values_to_bins(1, vals, ["a", "b"]))
    s.blocks.append(blk)
    s.IntInfo["v"] = (0, n - 1)
    assert run_oracle(s, Q(s, set_filters=[("tags", "nin", "a")])).MatchedCount == 18  # [], None, [1] x 6
    assert run_oracle(s, Q(s, set_filters=[("tags", "in", "a")])).MatchedCount == 12
    assert run_oracle(s, Q(s, set_filters=[("tags", "nin", "zzz")])).MatchedCount == 30


def test_str_replace_merges_groups_and_rewrites_filters():
    s, rows, valid = _table()
    pat, rep = r"^web-(\d+)\.dc\d$", "web-$1"
    py = lambda x: re.sub(pat, r"web-\1", x)
    q = Q(s, groups=["host"], aggs=["v"], str_replace={"host": (pat, rep)})
    o = run_oracle(s, q)
    want, sums = Counter(), Counter()
    for i in range(len(rows["v"])):
        k = (py(rows["host"][i]) if valid["host"][i] else "") + "\t"
        want[k] += 1
        sums[k] += int(rows["v"][i])
    assert {k: r.Count for k, r in o.Results.items()} == dict(want)
    assert {k: r.Hists["v"].ExactSum for k, r in o.Results.items()} == dict(sums)
    assert len(want) == 13  # 12 hosts (the dc suffix is gone) + rows without the column
    # filters see the rewritten strings: eq on a rewritten name, neq, and a regexp over it
    for op, lit in (("eq", "web-03"), ("neq", "web-03"), ("re", "^web-0[12]$")):
        o = run_oracle(s, Q(s, str_filters=[("host", op, lit)], str_replace={"host": (pat, rep)}))
        def ok(i):
            if not valid["host"][i]:
                return False
            h = py(rows["host"][i])
            return {"eq": h == lit, "neq": h != lit, "re": re.search(lit, h) is not None}[op]
        assert o.MatchedCount == sum(1 for i in range(len(rows["v"])) if ok(i)), (op, lit)


def test_str_replace_absent_literal_aliases_an_id_like_the_reference():
    # get_val_id gives an absent literal the id len(col.StringTable) (table_column.go:27-48).  After a rewrite that
    # merged strings the map is shorter than the block's string table, so that id belongs to another string and
    # `eq <absent literal>` matches ITS rows (column_store_io.go:536-546 + filter.go:205).  The oracle restates it;
    # the CUDA path answers "no row" (documented divergence).
    s = Spec([("v", INT), ("name", STR)])
    names = np.array(["a1", "a2", "b"] * 10)
    s.add_rows({"v": np.arange(30), "name": names})
    o = run_oracle(s, Q(s, str_filters=[("name", "eq", "zz")], str_replace={"name": (r"^a\d$", "a")}))
    assert o.MatchedCount == 10  # the rows of "b": local id 2 == len({"a": 0, "b": 2})
    assert run_oracle(s, Q(s, str_filters=[("name", "eq", "zz")])).MatchedCount == 0


def test_set_columns_round_trip_through_block_directories(tmp_path):
    # set_<col>.db (gob(SavedSetColumn), column_store_io.go:139-217) written and read back, in the bucketed form and
    # in the Values [][]int32 form (threshold 4 < 9 tags): same oracle answers as the in-memory blocks
    from sybil_b200 import blockdir
    for threshold in (5000, 4):
        s, rows, valid = _table(seed=8, nrows=2500, block_rows=900, threshold=threshold)
        s2 = Spec(s.key_table)
        s2.IntInfo = dict(s.IntInfo)
        for b in s.blocks:
            d = str(tmp_path / ("t%d_b%d" % (threshold, b.block_index)))
            blockdir.write_block_dir(d, b, s.key_table, compress=b.block_index % 2 == 1)
            assert any(f.startswith("set_tags.db") for f in __import__("os").listdir(d))
            s2.blocks.append(blockdir.read_block_dir(d, s.key_table, block_index=b.block_index))
        for op, tag in (("in", "t3"), ("nin", "t3"), ("nin", "nope")):
            q = lambda sp: Q(sp, set_filters=[("tags", op, tag)], groups=["host"], aggs=["v"])
            a, b2 = run_oracle(s, q(s)), run_oracle(s2, q(s2))
            assert a.MatchedCount == b2.MatchedCount > 0
            assert {k: r.Count for k, r in a.Results.items()} == {k: r.Count for k, r in b2.Results.items()}
        if threshold == 4:  # the Values form marks rows below len(Values) populated: "nin" also takes the empty sets
            n = sum(1 for r in s.blocks[0].cols if r.col_type == SET and getattr(r, "set_nvalues", 0) > 0)
            assert n == 1


def test_weights_in_the_oracle_match_a_row_by_row_evaluation():
    # OPTS.WEIGHT_COL (aggregate.go:68,100-102,202-203; hist_basic.go:111-116): Count += weight, Samples = rows,
    # hist Count / bucket counters += weight, exact sum += value * weight; and the carry-over of Q13
    rng = np.random.default_rng(4)
    n = 3000
    s = Spec([("v", INT), ("host", STR), ("w", INT)])
    rows = {"v": rng.integers(0, 1000, n), "host": np.array(["h%d" % x for x in rng.integers(0, 4, n)]),
            "w": rng.choice(np.asarray([1, 2, 5, 10]), n)}
    valid = {"w": rng.random(n) > 0.2}
    s.add_rows(rows, valid, block_rows=1100)
    o = run_oracle(s, Q(s, groups=["host"], aggs=["v"], op="hist", weight_col="w"))
    cnt, smp, sm = Counter(), Counter(), Counter()
    for start in range(0, n, 1100):
        weight = 1  # declared outside the row loop of one block: carries over to rows without the column
        for i in range(start, min(n, start + 1100)):
            if valid["w"][i]:
                weight = int(rows["w"][i])
            k = rows["host"][i] + "\t"
            cnt[k] += weight
            smp[k] += 1
            sm[k] += int(rows["v"][i]) * weight
    assert o.MatchedCount == n
    assert {k: (r.Count, r.Samples, r.Hists["v"].Count, r.Hists["v"].ExactSum, int(r.Hists["v"].Values.sum()))
            for k, r in o.Results.items()} == {k: (cnt[k], smp[k], cnt[k], sm[k], cnt[k]) for k in cnt}
