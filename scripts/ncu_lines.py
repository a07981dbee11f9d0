#!/usr/bin/env python
"""Per-source-line view of an ncu capture:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | this
Prints the hottest CUDA source lines by warp-stall samples with their share of executed instructions and
the dominant stall reasons."""
import csv, sys
rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 50
hdr = None
lines = {}
for r in rows:
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = lines.setdefault(ln, {"src": r[1], "samples": 0, "inst": 0, "stalls": {}})
    d["samples"] += int(r[hdr.index("# Samples")] or 0)
    d["inst"] += int(r[hdr.index("Instructions Executed")] or 0)
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            d["stalls"][h[6:]] = d["stalls"].get(h[6:], 0) + int(r[i] or 0)
tot = sum(d["samples"] for d in lines.values()) or 1
toti = sum(d["inst"] for d in lines.values()) or 1
print("total samples %d, warp instructions %d" % (tot, toti))
for ln, d in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% smp %5.1f%% inst  L%-5d %-22s %s" % (100 * d["samples"] / tot, 100 * d["inst"] / toti, ln,
          ",".join("%s:%d%%" % (k, 100 * v / max(d["samples"], 1)) for k, v in st if v), d["src"].strip()[:90]))
