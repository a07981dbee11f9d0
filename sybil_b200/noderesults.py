"""`NodeResults` in gob — what a sybil node hands to `sybil aggregate` (the step after the hot path).

`sybil query -encode-results` prints `gob(NodeResults{Table, Tables, QuerySpec, Samples})`
(`src/lib/printer.go:284-297`); `sybil aggregate` decodes one such stream per node and merges the
QuerySpecs with `CombineResults` (`src/lib/node_aggregator.go:147-177`), re-deriving percentiles
from the merged bucket counters.  `encode_node_results` emits that stream from a finished query of
this engine (or of the oracle: any object with Results / Sorted / Cumulative / MatchedCount whose
groups carry GroupByKey, Count, Samples and Hists[name] with Count, Avg, Min/Max, NumBuckets,
BucketSize, Values), so GPU results can be merged by an unmodified `sybil aggregate`
(SURVEY.md §8f N3).

Only what the merge reads is sent (gob receivers bind by field name and zero the rest): Groups,
Aggregations, OrderBy/Limit/TimeBucket; per group GroupByKey, BinaryByKey, Count, Samples and, per
aggregation, a `*sybil.HistCompat` (the name the reference registers, `query_cache.go:18-27`) holding
the BasicHist state `Combine` uses (`hist_basic.go:259-279`): NumBuckets, BucketSize, Values,
PercentileMode, Max, Min, Count, Avg, Info{Min, Max}.  Log-scale histograms (FLAGS.LOG_HIST) travel as
`*sybil.MultiHistCompat` with one such state per sub-histogram (`multi_hist_value`).  The receiving side —
`CombineResults` under `OPTS.MERGE_TABLE`, i.e. `fullMergeHist` — is restated in `stitch.py`.

Field layouts follow the type definitions Go itself sent in the reference's golden stream
(`testdata/TestDecodeGoldenFiles/node_results.golden.gob`); `tests/test_noderesults.py` decodes that
stream, re-emits it through this module and checks that every value this module covers comes back
identical.  No Go toolchain exists here, so the emitted stream has not been fed to `sybil aggregate`.
"""
from . import gob

INT_INFO = ("struct", "IntInfo", [("Min", "int"), ("Max", "int"), ("Avg", "float"), ("M2", "float"), ("Count", "int")])
BASIC_CACHED = ("struct", "BasicHistCachedInfo", [
    ("NumBuckets", "int"), ("BucketSize", "int"), ("Values", ("slice", "int")), ("Averages", ("slice", "float")),
    ("PercentileMode", "bool"), ("Outliers", ("slice", "int")), ("Underliers", ("slice", "int")), ("Max", "int"), ("Min", "int"),
    ("Samples", "int"), ("Count", "int"), ("Avg", "float"), ("Info", INT_INFO)])
BASIC_HIST = ("struct", "BasicHist", [("BasicHistCachedInfo", BASIC_CACHED)])
HIST_COMPAT = ("struct", "HistCompat", [("BasicHist", BASIC_HIST)])
RESULT = ("struct", "Result", [("Hists", ("map", "string", "interface")), ("GroupByKey", "string"), ("BinaryByKey", "string"),
                               ("Count", "int"), ("Samples", "int")])
GROUPING = ("struct", "Grouping", [("Name", "string")])
AGGREGATION = ("struct", "Aggregation", [("Op", "string"), ("Name", "string"), ("HistType", "string")])
QUERY_PARAMS = ("struct", "QueryParams", [("Groups", ("slice", GROUPING)), ("Aggregations", ("slice", AGGREGATION)),
                                          ("OrderBy", "string"), ("PruneBy", "string"), ("Limit", "int"), ("TimeBucket", "int")])
QUERY_RESULTS = ("struct", "QueryResults", [("Cumulative", RESULT), ("Results", ("map", "string", RESULT)),
                                            ("TimeResults", ("map", "int", ("map", "string", RESULT))), ("MatchedCount", "int"),
                                            ("Sorted", ("slice", RESULT))])
QUERY_SPEC = ("struct", "QuerySpec", [("QueryParams", QUERY_PARAMS), ("QueryResults", QUERY_RESULTS)])
TABLE = ("struct", "Table", [("Name", "string")])
NODE_RESULTS = ("struct", "NodeResults", [("Table", TABLE), ("Tables", ("slice", "string")), ("QuerySpec", QUERY_SPEC)])

HIST_NAME = "*sybil.HistCompat"
# hist_multi.go:6-18 (table is unexported); MultiHistCompat embeds *MultiHist and names it again as Histogram
# (hist_compat.go:50-54): gob flattens both pointers, the state travels twice
MULTI_HIST = ("struct", "MultiHist", [("Max", "int"), ("Min", "int"), ("Samples", "int"), ("Count", "int"), ("Avg", "float"),
                                      ("PercentileMode", "bool"), ("Subhists", ("slice", HIST_COMPAT)), ("Info", INT_INFO)])
MULTI_COMPAT = ("struct", "MultiHistCompat", [("MultiHist", MULTI_HIST), ("Histogram", MULTI_HIST)])
MULTI_NAME = "*sybil.MultiHistCompat"


def hist_value(count, avg, vmin, vmax, num_buckets, bucket_size, values, info_min, info_max, samples=0):
    """The BasicHist state `Combine` reads, as the interface value of Result.Hists[name]."""
    cached = {"NumBuckets": int(num_buckets), "BucketSize": int(bucket_size), "Values": [int(v) for v in values],
              "PercentileMode": len(values) > 0, "Max": int(vmax), "Min": int(vmin), "Samples": int(samples), "Count": int(count),
              "Avg": float(avg), "Info": {"Min": int(info_min), "Max": int(info_max)}}
    return (HIST_NAME, HIST_COMPAT, {"BasicHist": {"BasicHistCachedInfo": cached}})


def multi_hist_value(count, avg, vmin, vmax, values, info_min, info_max, samples=0, hist_bucket=0):
    """A log-scale histogram (FLAGS.LOG_HIST) as `*sybil.MultiHistCompat`: `values` are the bucket counters of all
    sub-histograms concatenated in Subhists order (the engine's layout, sg_hist.h::make_layout ==
    TrackPercentiles, hist_multi.go:223-257).  Each sub-histogram is sent with its own layout, its counters and
    Count = their sum; the engine keeps no per-sub-histogram mean, so their Avg is sent as 0 — nothing the
    reference prints or merges reads it (percentiles, buckets and stddev of a MultiHist come from the counters and
    the outer Avg, hist_multi.go:90-158)."""
    from .stitch import BasicHist, multi_layout
    subs, at = [], 0
    for lo, hi in multi_layout(int(info_min), int(info_max)):
        lay = BasicHist(lo, hi, hist_bucket)
        n = len(lay.Values)
        vals = [int(v) for v in values[at:at + n]]
        at += n
        subs.append(hist_value(sum(vals), 0.0, lo, hi, lay.NumBuckets, lay.BucketSize, vals, lo, hi)[2])
    if at != len(values):
        raise ValueError("MultiHist: %d counters for a layout of %d" % (len(values), at))
    m = {"Max": int(vmax), "Min": int(vmin), "Samples": int(samples), "Count": int(count), "Avg": float(avg),
         "PercentileMode": len(values) > 0, "Subhists": subs, "Info": {"Min": int(info_min), "Max": int(info_max)}}
    return (MULTI_NAME, MULTI_COMPAT, {"MultiHist": m, "Histogram": m})


def _binary_key(r):
    b = getattr(r, "BinaryByKey", "")
    if isinstance(b, str):
        return b
    # key words (uint64) -> the 8-byte little-endian fields of aggregate.go:125-143, one per group column
    # (the gob writer encodes str as utf-8 with surrogateescape: decoding the raw bytes the same way makes every byte,
    # including those >= 0x80, come out as ONE byte on the wire)
    return b"".join(int(w).to_bytes(8, "little") for w in b).decode("utf-8", "surrogateescape")


def _result(r, agg_names, int_info, ngroups):
    hists = {}
    for name in agg_names:
        h = r.Hists.get(name)
        if h is None:
            continue
        mn = h.Min() if callable(getattr(h, "Min", None)) else h.Min
        mx = h.Max() if callable(getattr(h, "Max", None)) else h.Max
        lo, hi = int_info.get(name, (mn, mx))
        if getattr(h, "nsubhists", 0) >= 1:  # FLAGS.LOG_HIST: *sybil.MultiHistCompat
            hists[name] = multi_hist_value(h.Count, h.Avg, mn, mx, list(h.Values), lo, hi, getattr(h, "Samples", 0))
            continue
        hists[name] = hist_value(h.Count, h.Avg, mn, mx, h.NumBuckets, h.BucketSize, list(h.Values), lo, hi, getattr(h, "Samples", 0))
    key = _binary_key(r)
    if not isinstance(getattr(r, "BinaryByKey", ""), str):
        raw = key.encode("utf-8", "surrogateescape")[:8 * ngroups]  # GROUP_BY_WIDTH bytes per group column
        key = raw.decode("utf-8", "surrogateescape")
    return {"Hists": hists, "GroupByKey": r.GroupByKey, "BinaryByKey": key, "Count": int(r.Count), "Samples": int(r.Samples)}


def node_results_value(qs, table_name, groups, aggs, int_info, op="hist", order_by="$COUNT", limit=100, time_bucket=0,
                       loghist=False):
    """The NodeResults value (a dict matching NODE_RESULTS) of a finished query.
    groups / aggs: column names; int_info: name -> (Min, Max) of the table (hist.go:27-38)."""
    res = lambda r: _result(r, aggs, int_info, len(groups))  # noqa: E731
    qr = {"Results": {k: res(r) for k, r in qs.Results.items()}, "MatchedCount": int(qs.MatchedCount),
          "Sorted": [res(r) for r in qs.Sorted]}
    if qs.Cumulative is not None:
        qr["Cumulative"] = res(qs.Cumulative)
    if getattr(qs, "TimeResults", None):
        qr["TimeResults"] = {int(tb): {k: res(r) for k, r in m.items()} for tb, m in qs.TimeResults.items()}
    qp = {"Groups": [{"Name": g} for g in groups],
          "Aggregations": [{"Op": op, "Name": a, "HistType": ("multi" if loghist else "basic") if op == "hist" else ""} for a in aggs],
          "OrderBy": order_by, "Limit": int(limit), "TimeBucket": int(time_bucket)}
    return {"Table": {"Name": table_name}, "Tables": [table_name], "QuerySpec": {"QueryParams": qp, "QueryResults": qr}}


def encode_node_results(qs, table_name, groups, aggs, int_info, **kw):
    """gob(NodeResults) + the trailing newline PrintBytes appends (printer.go:272-282)."""
    return gob.encode(node_results_value(qs, table_name, groups, aggs, int_info, **kw), NODE_RESULTS) + b"\n"
