"""NodeResults emitter: what the engine hands to an unmodified `sybil aggregate`."""
import os

import numpy as np

from sybil_b200 import gob, noderesults as NR
from tests.util import INT, STR, Q, Spec, run_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _covered(v, t):
    """v restricted to the fields descriptor t names (what this module emits), recursively."""
    if isinstance(t, str) or v is None:
        return v
    if t[0] == "struct":
        return {n: _covered(v[n], ft) for n, ft in t[2] if n in v}
    if t[0] == "slice":
        return [_covered(x, t[1]) for x in v]
    if t[0] == "map":
        return {k: _covered(x, t[2]) for k, x in v.items()}
    raise AssertionError(t)


def _as_interfaces(v):
    """decoded tree -> encoder input: interface values become (name, descriptor, value)."""
    if isinstance(v, dict):
        if v.get("__type__") == NR.HIST_NAME:
            return (NR.HIST_NAME, NR.HIST_COMPAT, _covered({k: x for k, x in v.items() if k != "__type__"}, NR.HIST_COMPAT))
        return {k: _as_interfaces(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_as_interfaces(x) for x in v]
    return v


def _strip(v):
    """Drop the decoder's type tags and every zero / empty field (gob does not send those)."""
    if isinstance(v, dict):
        out = {k: _strip(x) for k, x in v.items() if k != "__type__"}
        return {k: x for k, x in out.items() if x not in (0, 0.0, False, "", None) and x != {} and x != []}
    if isinstance(v, list):
        return [_strip(x) for x in v]
    return v


def test_reemitting_the_golden_node_results_preserves_every_covered_value():
    """Go's own NodeResults stream, decoded, re-emitted with this module's type layouts and decoded again:
    groups, keys, counts, all 1002 bucket counters, Avg (bit-equal), extents — identical."""
    gold = gob.decode(open(os.path.join(GOLD, "node_results.golden.gob"), "rb").read())
    want = _covered(gold, NR.NODE_RESULTS)
    raw = gob.encode(_as_interfaces(want), NR.NODE_RESULTS) + b"\n"
    back = gob.decode(raw)
    res = back["QuerySpec"]["QueryResults"]["Results"]
    assert len(res) == 12
    h = res["edge\tdesktop\t"]["Hists"]["pageload"]
    assert h["__type__"] == NR.HIST_NAME and len(h["BasicHist"]["BasicHistCachedInfo"]["Values"]) == 1002
    gold_h = gold["QuerySpec"]["QueryResults"]["Results"]["edge\tdesktop\t"]["Hists"]["pageload"]["BasicHist"]["BasicHistCachedInfo"]
    assert h["BasicHist"]["BasicHistCachedInfo"]["Avg"] == gold_h["Avg"]
    assert _strip(back) == _strip(want)


def test_oracle_results_round_trip_as_node_results():
    rng = np.random.default_rng(3)
    n = 3000
    s = Spec([("lat", INT), ("host", STR), ("dc", STR)])
    s.add_rows({"lat": rng.integers(30, 5000, n), "host": np.array(["h%d" % x for x in rng.integers(0, 4, n)]),
                "dc": np.array(["d%d" % x for x in rng.integers(0, 3, n)])}, block_rows=1000)
    q = Q(s, groups=["host", "dc"], aggs=["lat"], op="hist")
    o = run_oracle(s, q)
    raw = NR.encode_node_results(o, "t", ["host", "dc"], ["lat"], {"lat": s.IntInfo["lat"]})
    assert raw.endswith(b"\n")
    back = gob.decode(raw)
    qr = back["QuerySpec"]["QueryResults"]
    assert qr["MatchedCount"] == o.MatchedCount and set(qr["Results"]) == set(o.Results)
    assert [r["GroupByKey"] for r in qr["Sorted"]] == [r.GroupByKey for r in o.Sorted]
    assert [g["Name"] for g in back["QuerySpec"]["QueryParams"]["Groups"]] == ["host", "dc"]
    for k, r in o.Results.items():
        b = qr["Results"][k]
        assert b["Count"] == r.Count and b.get("Samples", 0) == r.Samples
        c = b["Hists"]["lat"]["BasicHist"]["BasicHistCachedInfo"]
        h = r.Hists["lat"]
        assert c["Count"] == h.Count and c["Avg"] == h.Avg and c["NumBuckets"] == h.NumBuckets and c["BucketSize"] == h.BucketSize
        assert c["Values"] == [int(v) for v in h.Values] and c["PercentileMode"] is True
        assert (c["Info"].get("Min", 0), c["Info"].get("Max", 0)) == tuple(s.IntInfo["lat"])
    assert qr["Cumulative"]["Count"] == o.Cumulative.Count


def test_binary_key_bytes_survive_the_wire():
    """BinaryByKey is 8 little-endian bytes per group column (aggregate.go:125-143); key words >= 128 hold bytes
    >= 0x80, which must stay single bytes in the gob string (latin-1 decoding doubled them)."""
    from sybil_b200 import noderesults as NR

    class R:
        BinaryByKey = [200, 0xFFFFFFFFFFFFFFFF, 5]
    key = NR._binary_key(R())
    assert len(key.encode("utf-8", "surrogateescape")) == 24
    assert key.encode("utf-8", "surrogateescape")[:8] == (200).to_bytes(8, "little")


def _halves(seed=4, n=4000):
    rng = np.random.default_rng(seed)
    rows = {"lat": rng.integers(30, 5000, n), "host": np.array(["h%d" % x for x in rng.integers(0, 4, n)])}
    whole, a, b = Spec([("lat", INT), ("host", STR)]), Spec([("lat", INT), ("host", STR)]), Spec([("lat", INT), ("host", STR)])
    whole.add_rows(rows, block_rows=1000)
    a.add_rows({k: v[:n // 2] for k, v in rows.items()}, block_rows=1000)
    b.add_rows({k: v[n // 2:] for k, v in rows.items()}, block_rows=1000)
    return whole, a, b


def test_stitch_with_equal_extents_reproduces_the_whole_table_counters():
    """`sybil aggregate` over two nodes' NodeResults (stitch.combine_node_results == CombineResults under
    MERGE_TABLE, fullMergeHist query_spec.go:118-135): with the same extents on both nodes every (bucket start,
    count) pair lands in its own bucket again, so counts and all bucket counters equal one node scanning it all."""
    from sybil_b200 import stitch
    whole, a, b = _halves()
    info = {"lat": whole.IntInfo["lat"]}
    for sp in (a, b):
        sp.IntInfo = dict(whole.IntInfo)  # both nodes build their histograms from the table's extents
    mk = lambda sp: Q(sp, groups=["host"], aggs=["lat"], op="hist")
    streams = [NR.encode_node_results(run_oracle(sp, mk(sp)), "t", ["host"], ["lat"], info) for sp in (a, b)]
    merged = stitch.combine_node_results(streams)
    want = run_oracle(whole, mk(whole))
    assert merged["MatchedCount"] == want.MatchedCount and set(merged["Results"]) == set(want.Results)
    for k, r in want.Results.items():
        g = merged["Results"][k]
        kind, h = g["Hists"]["lat"]
        assert kind == "basic" and g["Count"] == r.Count
        assert h.Count == r.Hists["lat"].Count and h.Values == [int(v) for v in r.Hists["lat"].Values]
        assert (h.NumBuckets, h.BucketSize) == (r.Hists["lat"].NumBuckets, r.Hists["lat"].BucketSize)
        # the merged mean is the mean of BUCKET STARTS (what the reference computes; values beyond the last bucket
        # were clamped into it): the weighted mean of the whole-table result's own sparse buckets
        ib = r.Hists["lat"].IntBuckets
        assert abs(h.Avg - sum(k2 * c for k2, c in ib.items()) / sum(ib.values())) < 1e-6
    assert merged["Cumulative"]["Count"] == want.Cumulative.Count


def test_stitch_rebuckets_histograms_of_different_extents():
    from sybil_b200 import stitch
    whole, a, b = _halves(seed=6)
    a.IntInfo, b.IntInfo = {"lat": (0, 2000)}, {"lat": (30, 6000)}
    mk = lambda sp: Q(sp, aggs=["lat"], op="hist")
    oa, ob = run_oracle(a, mk(a)), run_oracle(b, mk(b))
    streams = [NR.encode_node_results(o, "t", [], ["lat"], {"lat": sp.IntInfo["lat"]}) for o, sp in ((oa, a), (ob, b))]
    merged = stitch.combine_node_results(streams)
    kind, h = merged["Cumulative"]["Hists"]["lat"]
    assert (h.InfoMin, h.InfoMax) == (0, 6000)
    ref = stitch.BasicHist(0, 6000)
    assert (h.NumBuckets, h.BucketSize) == (ref.NumBuckets, ref.BucketSize) == (1001, 6)
    # every accepted value of both nodes is still counted, in the bucket of its node-side bucket start
    assert h.Count == oa.Cumulative.Hists["lat"].Count + ob.Cumulative.Hists["lat"].Count == sum(h.Values)
    want = [0] * len(h.Values)
    for o in (oa, ob):
        for start, cnt in o.Cumulative.Hists["lat"].IntBuckets.items():
            want[min(start // 6, len(want) - 1)] += cnt
    assert h.Values == want


def test_multihist_results_travel_as_multihistcompat():
    from sybil_b200 import stitch
    whole, a, b = _halves(seed=7)
    info = {"lat": whole.IntInfo["lat"]}
    q = Q(whole, groups=["host"], aggs=["lat"], op="hist", loghist=True)
    o = run_oracle(whole, q)
    raw = NR.encode_node_results(o, "t", ["host"], ["lat"], info, loghist=True)
    back = gob.decode(raw)
    assert back["QuerySpec"]["QueryParams"]["Aggregations"][0]["HistType"] == "multi"
    lay = stitch.multi_layout(*info["lat"])
    for k, r in o.Results.items():
        v = back["QuerySpec"]["QueryResults"]["Results"][k]["Hists"]["lat"]
        assert v["__type__"] == NR.MULTI_NAME and v["MultiHist"] == v["Histogram"]
        m = v["MultiHist"]
        assert m["Count"] == r.Hists["lat"].Count and m["Avg"] == r.Hists["lat"].Avg and len(m["Subhists"]) == len(lay)
        flat = []
        for sh, (lo, hi) in zip(m["Subhists"], lay):
            c = sh["BasicHist"]["BasicHistCachedInfo"]
            assert (c["Info"].get("Min", 0), c["Info"].get("Max", 0)) == (lo, hi)
            flat += c.get("Values", [])
        assert flat == [int(x) for x in r.Hists["lat"].Values] and sum(flat) == r.Hists["lat"].Count
    # and the aggregator merges two such streams into BasicHists over the union range
    merged = stitch.combine_node_results([raw, raw])
    for k, r in o.Results.items():
        kind, h = merged["Results"][k]["Hists"]["lat"]
        assert kind == "basic" and h.Count == 2 * r.Hists["lat"].Count == sum(h.Values)


def test_stitch_merges_time_results_per_bucket():
    from sybil_b200 import stitch
    rng = np.random.default_rng(12)
    n = 3000
    rows = {"lat": rng.integers(30, 3000, n), "host": np.array(["h%d" % x for x in rng.integers(0, 3, n)]),
            "time": 1500000000 + np.sort(rng.integers(0, 3600, n))}
    kt = [("lat", INT), ("host", STR), ("time", INT)]
    whole, a, b = Spec(kt), Spec(kt), Spec(kt)
    whole.add_rows(rows, block_rows=1000)
    a.add_rows({k: v[::2] for k, v in rows.items()}, block_rows=1000)   # every other row: both nodes see every bucket
    b.add_rows({k: v[1::2] for k, v in rows.items()}, block_rows=1000)
    for sp in (a, b):
        sp.IntInfo = dict(whole.IntInfo)
    mk = lambda sp: Q(sp, groups=["host"], aggs=["lat"], op="hist", time_col="time", time_bucket=900)
    info = {"lat": whole.IntInfo["lat"]}
    streams = [NR.encode_node_results(run_oracle(sp, mk(sp)), "t", ["host"], ["lat"], info, time_bucket=900) for sp in (a, b)]
    merged = stitch.combine_node_results(streams)
    want = run_oracle(whole, mk(whole))
    assert set(merged["TimeResults"]) == set(want.TimeResults) and len(want.TimeResults) == 5
    for tb, m in want.TimeResults.items():
        for k, r in m.items():
            g = merged["TimeResults"][tb][k]
            assert g["Count"] == r.Count
            assert g["Hists"]["lat"][1].Values == [int(v) for v in r.Hists["lat"].Values]
    assert {k: g["Count"] for k, g in merged["Results"].items()} == {k: r.Count for k, r in want.Results.items()}
