"""Extracts the hot-path golden vector from the reference's own test data.

Source: /root/reference/src/lib/testdata/TestDecodeGoldenFiles/node_results.golden.json
(the fixture of src/lib/decoding_test.go:20-74): a real `group browser,device / hist
pageload` result over 20,000 rows — 12 groups with full BasicHist state, the
Cumulative "TOTAL" row and the Sorted order.

Writes tests/golden/node_results_hist.json (bucket counters stored sparsely) and
tests/golden/node_results_rows.json: one representative row per counted value
(floor of the bucket's running average, which lies inside the bucket), so that the
whole pipeline can be driven to reproduce the golden bucket counters.

Also copies the reference's gob fixtures (`node_results.golden.gob`, `flag_defs.golden.gob` and the JSON
rendering of the latter) next to them: test DATA of `decoding_test.go:20-74`, used by tests/test_gob.py to pin
the gob reader to Go's own output.

Run in the build container only (the GPU box has no /root/reference).
"""
import json
import math
import os
import shutil

SRC = "/root/reference/src/lib/testdata/TestDecodeGoldenFiles/node_results.golden.json"
HERE = os.path.dirname(os.path.abspath(__file__))


def hist(h):
    return {
        "NumBuckets": h["NumBuckets"], "BucketSize": h["BucketSize"], "nvalues": len(h["Values"]),
        "Values": {str(k): v for k, v in enumerate(h["Values"]) if v},
        "Outliers": h["Outliers"] or [], "Underliers": h["Underliers"] or [],
        "Max": h["Max"], "Min": h["Min"], "Samples": h["Samples"], "Count": h["Count"], "Avg": h["Avg"],
        "InfoMin": h["Info"]["Min"], "InfoMax": h["Info"]["Max"],
    }


def main():
    qs = json.load(open(SRC))["QuerySpec"]
    out = {
        "source": "src/lib/testdata/TestDecodeGoldenFiles/node_results.golden.json",
        "Groups": [g["Name"] for g in qs["Groups"]],
        "Aggregation": qs["Aggregations"][0]["Name"],
        "MatchedCount": qs["MatchedCount"],
        "Cumulative": {"GroupByKey": qs["Cumulative"]["GroupByKey"], "Count": qs["Cumulative"]["Count"],
                       "Samples": qs["Cumulative"]["Samples"], "hist": hist(qs["Cumulative"]["Hists"]["pageload"])},
        "Sorted": [r["GroupByKey"] for r in qs["Sorted"]],
        "Results": {},
    }
    rows = []
    for k, r in qs["Results"].items():
        h = r["Hists"]["pageload"]
        out["Results"][k] = {"Count": r["Count"], "Samples": r["Samples"],
                             "BinaryByKey": [ord(c) for c in r["BinaryByKey"]], "hist": hist(h)}
        browser, device = k.split("\t")[:2]
        for b, c in enumerate(h["Values"]):
            if c:
                v = int(math.floor(h["Averages"][b]))
                lo = h["Min"] + b * h["BucketSize"]
                if b < len(h["Values"]) - 1:
                    assert lo <= v < lo + h["BucketSize"], (k, b, v, lo)
                rows.append([browser, device, v, c])
    json.dump(out, open(os.path.join(HERE, "node_results_hist.json"), "w"), separators=(",", ":"))
    for f in ("node_results.golden.gob", "flag_defs.golden.gob", "flag_defs.golden.json"):
        shutil.copyfile(os.path.join(os.path.dirname(SRC), f), os.path.join(HERE, f))
    json.dump({"columns": ["browser", "device", "pageload", "repeat"], "rows": rows},
              open(os.path.join(HERE, "node_results_rows.json"), "w"), separators=(",", ":"))
    print("groups", len(out["Results"]), "row classes", len(rows), "rows", sum(r[3] for r in rows))


if __name__ == "__main__":
    main()
