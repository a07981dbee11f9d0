"""The C++ oracle against the independent Python restatement (oracle/pyoracle.py) on small
random tables: both were written separately from the Go sources and must agree exactly."""
import numpy as np
import pytest

from oracle import pyoracle
from tests.util import INT, STR, Q, Spec, random_spec, run_oracle


def py_query(spec, q):
    filters = [(spec.KeyTable[c], "int", op, v) for c, op, v in q.int_filters]
    filters += [(spec.KeyTable[c], "str", op, v) for c, op, v in q.str_filters]
    filters += [(spec.KeyTable[c], "set", op, v) for c, op, v in q.set_filters]
    aggs = [(spec.KeyTable[a],) + tuple(spec.IntInfo.get(a, (0, 0))) for a in q.aggs]
    import re as _re
    # Go's $1 template as Python's \g<1>
    repl = {spec.KeyTable[c]: (p, _re.sub(r"\$(\d+)", r"\\g<\1>", r)) for c, (p, r) in q.str_replace.items()}
    return pyoracle.query(spec.blocks, spec.KeyTypes, filters, [spec.KeyTable[g] for g in q.groups], aggs,
                          op_hist=q.op == "hist", log_hist=q.loghist,
                          time_col=spec.KeyTable[q.time_col] if q.time_col else None, time_bucket=q.time_bucket,
                          hist_bucket=q.hist_bucket, weight_col=spec.KeyTable[q.weight_col] if q.weight_col else None,
                          str_replace=repl)


def check(spec, q):
    o = run_oracle(spec, q)
    res, tres, cum, matched, broken = py_query(spec, q)
    assert o.MatchedCount == matched
    assert o.BrokenBlocks == broken
    assert set(o.Results) == set(res)

    def cmp(og, pg):
        assert (og.Count, og.Samples) == (pg.Count, pg.Samples)
        for ai, a in enumerate(q.aggs):
            ph = pg.Hists.get(ai)
            oh = og.Hists.get(a)
            assert (ph is None) == (oh is None)
            if ph is None:
                continue
            assert oh.Count == ph.Count and oh.ExactSum == ph.ExactSum
            assert (oh.Min, oh.Max) == (ph.Min, ph.Max)
            assert oh.Avg == ph.Avg  # same operations in the same order: bit-identical
            assert list(oh.Values) == list(ph.Values)
            if q.op == "hist" and ph.Count:
                if oh.noutliers:  # Q9: the wrapper reports the merged view (fresh clone + Combine)
                    pm = ph.fresh()
                    pm.combine(ph)
                    ph = pm
                assert oh.Percentiles == ph.percentiles()
                assert oh.IntBuckets == ph.sparse()
                assert abs(oh.StdDev - ph.stddev()) <= 1e-12 * max(1.0, ph.stddev())

    for k, pg in res.items():
        cmp(o.Results[k], pg)
    assert (o.Cumulative.Count, o.Cumulative.GroupByKey) == (cum.Count, cum.GroupByKey)
    if not q.time_col:
        cmp(o.Cumulative, cum)
    assert set(o.TimeResults) == set(tres)
    for tb, m in tres.items():
        assert set(o.TimeResults[tb]) == set(m)
        for k, pg in m.items():
            cmp(o.TimeResults[tb][k], pg)


@pytest.mark.parametrize("seed", [1, 2])
def test_group_by_avg(seed):
    s = random_spec(seed, nrows=1500, block_rows=500)
    check(s, Q(s, groups=["host"], aggs=["age", "lat", "big"], op="avg"))


def test_filters_and_two_groups_hist():
    s = random_spec(3, nrows=1500, block_rows=400)
    check(s, Q(s, int_filters=[("age", "gt", 12), ("big", "lt", 900000)], str_filters=[("state", "neq", "s3")],
               groups=["host", "age"], aggs=["lat"], op="hist"))


def test_str_eq_absent_literal_and_regex():
    s = random_spec(4, nrows=900, block_rows=300)
    check(s, Q(s, str_filters=[("host", "eq", "nope")], groups=["state"], aggs=["age"], op="hist"))
    check(s, Q(s, str_filters=[("host", "neq", "nope")], groups=["state"], aggs=["age"], op="hist"))
    check(s, Q(s, str_filters=[("state", "re", "^s1")], groups=["state"], aggs=["age"], op="avg"))
    check(s, Q(s, str_filters=[("state", "nre", "^s1")], groups=["state"], aggs=["age"], op="avg"))


def test_loghist_and_time_series():
    s = random_spec(5, nrows=1200, block_rows=400)
    check(s, Q(s, groups=["host"], aggs=["big"], op="hist", loghist=True))
    check(s, Q(s, groups=["host"], aggs=["lat"], op="hist", time_col="time", time_bucket=600))


def test_value_array_encodings_and_int64_range():
    # threshold 50 forces the Values[] forms (column_store_test.go:143-211 style: values > 2^32)
    s = random_spec(6, nrows=800, block_rows=400, threshold=50, wide=True)
    check(s, Q(s, int_filters=[("big", "gt", 0)], groups=["host"], aggs=["big", "lat"], op="avg"))
    check(s, Q(s, groups=["uid"], aggs=["age"], op="avg"))


def test_no_groups_is_total_key():
    s = random_spec(7, nrows=500, block_rows=250)
    o = run_oracle(s, Q(s, aggs=["age"], op="hist"))
    assert list(o.Results) == ["total"]
    check(s, Q(s, aggs=["age"], op="hist"))


def test_set_filters_in_both_restatements():
    s = random_spec(8, nrows=1200, block_rows=400, sets=True)
    for op, tag in (("in", "t3"), ("nin", "t3"), ("nin", "nope"), ("in", "nope")):
        check(s, Q(s, set_filters=[("tags", op, tag)], groups=["host"], aggs=["lat"], op="hist"))
    check(s, Q(s, int_filters=[("age", "gt", 12)], set_filters=[("tags", "in", "t1"), ("tags", "nin", "t2")],
               groups=["state"], aggs=["lat"], op="avg"))


def test_weights_with_carry_over_in_both_restatements():
    # OPTS.WEIGHT_COL incl. the carry-over of Q13 (rows without the weight column), Count / Samples / weighted
    # histograms, the float running mean bit for bit
    import numpy as np
    rng = np.random.default_rng(9)
    n = 1500
    s = Spec([("lat", INT), ("host", STR), ("time", INT), ("w", INT)])
    s.add_rows({"lat": rng.integers(30, 9000, n), "host": np.array(["h%d" % x for x in rng.integers(0, 4, n)]),
                "time": 1500000000 + np.sort(rng.integers(0, 3600, n)), "w": rng.choice(np.asarray([1, 2, 5, 10]), n)},
               {"w": rng.random(n) > 0.15, "lat": rng.random(n) > 0.05}, block_rows=500)
    check(s, Q(s, groups=["host"], aggs=["lat"], op="hist", weight_col="w"))
    check(s, Q(s, groups=["host"], aggs=["lat"], op="avg", weight_col="w"))
    check(s, Q(s, groups=["host"], aggs=["lat"], op="hist", loghist=True, weight_col="w"))
    check(s, Q(s, aggs=["lat"], op="hist", weight_col="w", time_col="time", time_bucket=600))


def test_str_replace_in_both_restatements():
    s = random_spec(10, nrows=1200, block_rows=400)
    rep = {"state": (r"^s(\d)\d*$", "S$1")}
    check(s, Q(s, groups=["state"], aggs=["lat"], op="hist", str_replace=rep))
    check(s, Q(s, str_filters=[("state", "eq", "S1")], groups=["state", "host"], aggs=["lat"], op="avg", str_replace=rep))
    check(s, Q(s, str_filters=[("state", "neq", "S1")], groups=["state"], aggs=["lat"], op="avg", str_replace=rep))
