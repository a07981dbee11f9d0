"""The C++ oracle against the independent Python restatement (oracle/pyoracle.py) on small
random tables: both were written separately from the Go sources and must agree exactly."""
import numpy as np
import pytest

from oracle import pyoracle
from tests.util import INT, STR, Q, Spec, random_spec, run_oracle


def py_query(spec, q):
    filters = [(spec.KeyTable[c], "int", op, v) for c, op, v in q.int_filters]
    filters += [(spec.KeyTable[c], "str", op, v) for c, op, v in q.str_filters]
    aggs = [(spec.KeyTable[a],) + tuple(spec.IntInfo.get(a, (0, 0))) for a in q.aggs]
    return pyoracle.query(spec.blocks, spec.KeyTypes, filters, [spec.KeyTable[g] for g in q.groups], aggs,
                          op_hist=q.op == "hist", log_hist=q.loghist,
                          time_col=spec.KeyTable[q.time_col] if q.time_col else None, time_bucket=q.time_bucket,
                          hist_bucket=q.hist_bucket)


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
