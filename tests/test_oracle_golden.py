"""Pins the oracle to the one golden vector the reference's tests hold for the hot path:
testdata/TestDecodeGoldenFiles/node_results.golden.json (src/lib/decoding_test.go:20-74),
committed here in reduced form by tests/golden/make_golden.py."""
import ctypes as C
import json
import os

import numpy as np

from oracle import oracle_ffi
from tests.util import INT, STR, Q, Spec, run_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "node_results_hist.json")))
ROWS = json.load(open(os.path.join(HERE, "golden", "node_results_rows.json")))
I64P = C.POINTER(C.c_int64)


def dense(h):
    v = np.zeros(h["nvalues"], np.int64)
    for k, c in h["Values"].items():
        v[int(k)] = c
    return v


def test_bucket_layout_matches_golden():
    # SetupBuckets, hist_basic.go:34-70: (23500-30)/1000 -> BucketSize 23, NumBuckets 1001, len(Values) 1002
    L = oracle_ffi.lib()
    nb, bs, nv = C.c_int64(), C.c_int64(), C.c_int64()
    h = G["Cumulative"]["hist"]
    L.orc_basic_layout(h["InfoMin"], h["InfoMax"], 0, C.byref(nb), C.byref(bs), C.byref(nv))
    assert (nb.value, bs.value, nv.value) == (h["NumBuckets"], h["BucketSize"], h["nvalues"]) == (1001, 23, 1002)
    for r in G["Results"].values():
        assert (r["hist"]["NumBuckets"], r["hist"]["BucketSize"], r["hist"]["nvalues"]) == (1001, 23, 1002)


def test_combine_reproduces_the_cumulative_row():
    # CombineResults' Cumulative = Result.Combine over every group (aggregate.go:422-436,
    # hist_basic.go:259-279): bucket counters and Count exactly, Avg to the last bits
    L = oracle_ffi.lib()
    keys = G["Sorted"]
    n, nv = len(keys), 1002
    counts = np.array([G["Results"][k]["hist"]["Count"] for k in keys], np.int64)
    avgs = np.array([G["Results"][k]["hist"]["Avg"] for k in keys], np.float64)
    vals = np.concatenate([dense(G["Results"][k]["hist"]) for k in keys])
    oc, oa = C.c_int64(), C.c_double()
    ov = np.zeros(nv, np.int64)
    cum = G["Cumulative"]["hist"]
    L.orc_basic_combine(cum["InfoMin"], cum["InfoMax"], n, counts.ctypes.data_as(I64P),
                        avgs.ctypes.data_as(C.POINTER(C.c_double)), vals.ctypes.data_as(I64P), nv, C.byref(oc),
                        C.byref(oa), ov.ctypes.data_as(I64P))
    assert oc.value == cum["Count"] == G["Cumulative"]["Count"] == 20000
    assert np.array_equal(ov, dense(cum))
    assert abs(oa.value - cum["Avg"]) <= 1e-12 * cum["Avg"]
    assert cum["Outliers"] == [] and cum["Underliers"] == []  # Combine drops them
    assert G["Cumulative"]["GroupByKey"] == "TOTAL\t"


def test_sorted_order_is_count_descending():
    counts = [G["Results"][k]["Count"] for k in G["Sorted"]]
    assert counts == sorted(counts, reverse=True)
    assert len(set(counts)) == len(counts)  # no ties in the golden file: the order is fully pinned


def golden_table():
    s = Spec([("browser", STR), ("device", STR), ("pageload", INT)])
    b, d, p = [], [], []
    for browser, device, v, rep in ROWS["rows"]:
        b += [browser] * rep
        d += [device] * rep
        p += [v] * rep
    # interleave deterministically so that both string dictionaries are built in mixed order
    order = np.random.default_rng(7).permutation(len(p))
    s.add_rows({"browser": np.array(b)[order], "device": np.array(d)[order], "pageload": np.array(p, np.int64)[order]},
               block_rows=8192)
    h = G["Cumulative"]["hist"]
    s.IntInfo["pageload"] = (h["InfoMin"], h["InfoMax"])
    return s


def test_full_pipeline_reproduces_golden_buckets():
    # rows rebuilt from the golden bucket averages, digested into blocks, decoded, grouped and
    # histogrammed by the oracle: every counter of the reference's result must come back
    s = golden_table()
    q = Q(s, groups=["browser", "device"], aggs=["pageload"], op="hist")
    o = run_oracle(s, q)
    assert o.MatchedCount == G["MatchedCount"]
    assert [r.GroupByKey for r in o.Sorted] == G["Sorted"]
    assert o.Cumulative.GroupByKey == "TOTAL\t"
    assert np.array_equal(o.Cumulative.Hists["pageload"].Values, dense(G["Cumulative"]["hist"]))
    for k, g in G["Results"].items():
        r = o.Results[k]
        h = r.Hists["pageload"]
        assert (r.Count, r.Samples) == (g["Count"], g["Samples"])
        assert h.Count == g["hist"]["Count"]
        assert (h.Min, h.Max) == (g["hist"]["Min"], g["hist"]["Max"])
        assert np.array_equal(h.Values, dense(g["hist"]))
        assert abs(h.Avg - g["hist"]["Avg"]) < 1.0  # rows carry floor(bucket average)
    # the one golden outlier (clamped into the spare last slot, hist_basic.go:134-137)
    assert G["Results"]["gecko\ttablet\t"]["hist"]["Values"]["1001"] == 1
    assert len(G["Results"]["gecko\ttablet\t"]["hist"]["Outliers"]) == 1


def test_group_key_bytes_layout():
    # BinaryByKey: 8 little-endian bytes per group column holding the per-block string id
    # (aggregate.go:125-143); the golden keys decode to two ids < 4
    for k, g in G["Results"].items():
        b = bytes(g["BinaryByKey"])
        assert len(b) == 16
        ids = [int.from_bytes(b[i:i + 8], "little") for i in (0, 8)]
        assert all(0 <= i < 4 for i in ids)
