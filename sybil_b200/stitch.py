"""Cross-node merge of `NodeResults` — what `sybil aggregate` does with the streams its nodes printed
(`src/lib/node_aggregator.go:147-177`): `CombineResults` with `OPTS.MERGE_TABLE` set, so two histograms of
one (group, aggregation) are not added counter by counter but re-bucketed by `fullMergeHist`
(`src/lib/query_spec.go:118-135`): a fresh histogram over the union of both ranges receives every
(bucket start, count) pair of both inputs through `AddWeightedValue` — nodes may have built their
histograms from different table extents.  This is the step AFTER the hot path, host-side
and small: a few thousand counters per group; it lets results of this engine (emitted by
`noderesults.encode_node_results`) be stitched without Go, and is the restatement against which a real
`sybil aggregate` run can be compared.

Faithful to the reference's arithmetic, including what it loses: the re-bucketed histogram only knows bucket
starts, so its Avg is the mean of bucket starts (`hist_basic.go:117`), Min/Max become the union range, and
Samples counts buckets, not rows (`hist_basic.go:111-113`: weight > 1 => Samples++).
"""
from . import gob

NUM_BUCKETS = 1000  # hist.go:3


def _wrap(v):
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


class BasicHist:
    """BasicHist in percentile mode (hist_basic.go:10-151)."""

    def __init__(self, info_min, info_max, hist_bucket=0):
        # SetupBuckets (hist_basic.go:34-70)
        self.InfoMin, self.InfoMax = int(info_min), int(info_max)
        self.Avg, self.Count, self.Samples = 0.0, 0, 0
        self.Min, self.Max = self.InfoMin, self.InfoMax
        size = self.InfoMax - self.InfoMin
        nb, bs = NUM_BUCKETS, size // NUM_BUCKETS if size >= 0 else -((-size) // NUM_BUCKETS)
        if hist_bucket > 0:
            bs = hist_bucket
        if bs == 0:
            if size < 100:
                bs, nb = 1, size
            else:
                bs = size // 100
                nb = size // bs
        nb += 1
        self.NumBuckets, self.BucketSize = nb, bs
        self.Values = [0] * (nb + 1)
        self.Averages = [0.0] * (nb + 1)
        self.Outliers, self.Underliers = [], []

    def AddWeightedValue(self, value, weight, weight_col=False):  # hist_basic.go:101-151
        if value > _wrap(self.InfoMax * 10) or value < self.InfoMin:
            return
        if weight_col or weight > 1:
            self.Samples += 1
            self.Count += weight
        else:
            self.Count += 1
        self.Avg = self.Avg + ((float(value) - self.Avg) / float(self.Count)) * float(weight)
        self.Max = max(self.Max, value)
        self.Min = min(self.Min, value)
        q = value - self.Min
        b = abs(q) // self.BucketSize * (1 if q >= 0 else -1)  # Go's integer division truncates
        if b >= len(self.Values):
            self.Outliers.append(value)
            b = len(self.Values) - 1
        if b < 0:
            self.Underliers.append(value)
            b = 0
        partial = self.Averages[b]
        self.Values[b] += weight
        self.Averages[b] = partial + ((float(value) - partial) / float(self.Values[b]) * float(weight))

    def GetIntBuckets(self):  # GetSparseBuckets, hist_basic.go:221-239
        ret = {}
        for k, v in enumerate(self.Values):
            if v > 0:
                ret[k * self.BucketSize + self.Min] = v
        for v in self.Outliers + self.Underliers:
            ret[v] = ret.get(v, 0) + 1
        return ret

    def Range(self):
        return self.InfoMin, self.InfoMax

    def Combine(self, o):  # hist_basic.go:259-279 (same layout on both sides)
        for k, v in enumerate(o.Values):
            self.Values[k] += v
        total = self.Count + o.Count
        if total:
            self.Avg = self.Avg * (float(self.Count) / float(total)) + o.Avg * (float(o.Count) / float(total))
        self.Min, self.Max = min(self.Min, o.Min), max(self.Max, o.Max)
        self.Samples += o.Samples
        self.Count = total


def multi_layout(info_min, info_max):
    """Sub-ranges of a MultiHist, in Subhists order (TrackPercentiles, hist_multi.go:223-257)."""
    bucket = info_max - info_min
    num, t = 0, bucket
    while t > NUM_BUCKETS:
        num += 1
        t >>= 1
    out, right = [], info_max
    for _ in range(num):
        bucket >>= 1
        out.append((right - bucket, right))
        right -= bucket
    out.append((info_min, right))
    return out


def hist_from_gob(name, v):
    """A decoded Result.Hists[...] interface value -> (kind, state): BasicHist for `*sybil.HistCompat`; for
    `*sybil.MultiHistCompat` a list of BasicHists (its Subhists) plus the outer Info."""
    if name.endswith("MultiHistCompat"):
        m = v.get("MultiHist") or v.get("Histogram") or {}
        info = m.get("Info") or {}
        subs = [_basic_from_cached(((s or {}).get("BasicHist") or {}).get("BasicHistCachedInfo") or {}) for s in m.get("Subhists", [])]
        return "multi", {"subs": subs, "InfoMin": int(info.get("Min", 0)), "InfoMax": int(info.get("Max", 0)),
                         "Count": int(m.get("Count", 0)), "Avg": float(m.get("Avg", 0.0))}
    return "basic", _basic_from_cached(((v.get("BasicHist") or {}).get("BasicHistCachedInfo")) or {})


def _basic_from_cached(c):
    info = c.get("Info") or {}
    h = BasicHist.__new__(BasicHist)
    h.InfoMin, h.InfoMax = int(info.get("Min", 0)), int(info.get("Max", 0))
    h.NumBuckets, h.BucketSize = int(c.get("NumBuckets", 0)), int(c.get("BucketSize", 0))
    h.Values = [int(x) for x in c.get("Values", [])]
    h.Averages = [float(x) for x in c.get("Averages", [])] or [0.0] * len(h.Values)
    h.Outliers, h.Underliers = list(c.get("Outliers", [])), list(c.get("Underliers", []))
    h.Min, h.Max = int(c.get("Min", 0)), int(c.get("Max", 0))
    h.Count, h.Samples, h.Avg = int(c.get("Count", 0)), int(c.get("Samples", 0)), float(c.get("Avg", 0.0))
    return h


def _int_buckets(kind, st):
    if kind == "basic":
        return st.GetIntBuckets(), st.Range()
    out = {}
    for s in st["subs"]:  # MultiHist.GetSparseBuckets (hist_multi.go:184-200): every sub-histogram's buckets
        for k, v in s.GetIntBuckets().items():
            out[k] = out.get(k, 0) + v
    return out, (st["InfoMin"], st["InfoMax"])


def fullMergeHist(a, b, hist_bucket=0):
    """query_spec.go:118-135 with the aggregator's FLAGS.LOG_HIST unset (NewHist gives a BasicHist).  a, b:
    (kind, state) pairs from hist_from_gob.  Buckets are fed in ascending order (Go's map order is random; the
    counters do not depend on it, the float Avg does in its last bits)."""
    ba, (l1, r1) = _int_buckets(*a)
    bb, (l2, r2) = _int_buckets(*b)
    nh = BasicHist(min(l1, l2), max(r1, r2), hist_bucket)
    for src in (ba, bb):
        for bucket in sorted(src):
            nh.AddWeightedValue(bucket, src[bucket])
    return "basic", nh


def combine_node_results(streams, hist_bucket=0):
    """AggregateSpecs (node_aggregator.go:147-177): gob(NodeResults) byte strings, one per node -> the merged
    {"Results": {key: group}, "Cumulative": group, "MatchedCount": n, "TimeResults": {...}} with group =
    {"GroupByKey", "Count", "Samples", "Hists": {name: (kind, state)}}."""
    def lift(r):
        return {"GroupByKey": r.get("GroupByKey", ""), "Count": int(r.get("Count", 0)), "Samples": int(r.get("Samples", 0)),
                "Hists": {k: hist_from_gob(*_iface(v)) for k, v in (r.get("Hists") or {}).items()}}

    def _iface(v):  # gob.decode renders an interface value as (registered name, value)
        return (v[0], v[1]) if isinstance(v, tuple) else (v.get("__type__", "*sybil.HistCompat"), v)

    def combine(into, nxt):  # Result.Combine with MERGE_TABLE set (query_spec.go:138-181)
        if nxt["Count"] == 0:
            return
        for k, h in nxt["Hists"].items():
            into["Hists"][k] = fullMergeHist(h, into["Hists"][k], hist_bucket) if k in into["Hists"] else h
        into["Samples"] += nxt["Samples"]
        into["Count"] += nxt["Count"]

    def combine_map(master, m):  # ResultMap.Combine (query_spec.go:107-116)
        for k, v in m.items():
            if k in master:
                combine(master[k], v)
            else:
                master[k] = v

    out = {"Results": {}, "TimeResults": {}, "Cumulative": None, "MatchedCount": 0}
    for raw in streams:
        nr = gob.decode(raw)
        qr = (nr.get("QuerySpec") or {}).get("QueryResults") or {}
        out["MatchedCount"] += int(qr.get("MatchedCount", 0))
        combine_map(out["Results"], {k: lift(v) for k, v in (qr.get("Results") or {}).items()})
        for tb, m in (qr.get("TimeResults") or {}).items():
            combine_map(out["TimeResults"].setdefault(int(tb), {}), {k: lift(v) for k, v in m.items()})
        cum = qr.get("Cumulative")
        if cum:
            c = lift(cum)
            if out["Cumulative"] is None:
                out["Cumulative"] = c
            else:
                combine(out["Cumulative"], c)
    return out
