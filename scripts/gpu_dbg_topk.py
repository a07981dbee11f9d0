"""debug: the device-side top-K path against the host path (SG_NO_GPU_TOPK=1) on the test's tables"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import Q, Spec, run_gpu, INT, STR

rng = np.random.default_rng(53)
n = 400_000
s = Spec([("k", STR), ("m", INT)])
s.add_rows({"k": np.array(["key%d" % v for v in rng.integers(0, 300_000, n)]), "m": rng.integers(0, 10000, n)}, block_rows=65536)
keys = (rng.pareto(1.1, n) * 40).astype(np.int64) % 300_000
s2 = Spec([("k", STR), ("m", INT), ("w", INT)])
s2.add_rows({"k": np.array(["key%d" % v for v in keys]), "m": rng.integers(0, 10000, n), "w": rng.integers(-500, 500, n)}, block_rows=65536)
cases = [(s, ["m"], dict(limit=100)), (s, ["m"], dict(limit=100, order_by="m")),
         (s2, ["m", "w"], dict(limit=100)), (s2, ["m", "w"], dict(limit=1000, order_by="w")), (s2, ["m", "w"], dict(limit=7, order_by="m"))]
for sp, aggs, kw in cases:
    q = Q(sp, groups=["k"], aggs=aggs, op="avg", **kw)
    os.environ.pop("SG_NO_GPU_TOPK", None)
    a = run_gpu(sp, q)
    os.environ["SG_NO_GPU_TOPK"] = "1"
    b = run_gpu(sp, q)
    ka, kb = [r.GroupByKey for r in a.Sorted], [r.GroupByKey for r in b.Sorted]
    print(kw, "same" if ka == kb else "DIFF", len(ka), len(kb), a.NumGroups, b.NumGroups, a.MatchedCount, b.MatchedCount)
    if ka != kb:
        ob = kw.get("order_by", "$COUNT")
        for i, (x, y) in enumerate(zip(a.Sorted, b.Sorted)):
            if x.GroupByKey != y.GroupByKey:
                def val(r):
                    return r.Count if ob == "$COUNT" else (r.Hists[ob].Mean(), r.Hists[ob].TotalCount())
                print("  first diff at", i, repr(x.GroupByKey), val(x), "| host:", repr(y.GroupByKey), val(y))
                print("  gpu list around:", [(r.GroupByKey, val(r)) for r in a.Sorted[max(0, i - 2):i + 3]])
                print("  host list around:", [(r.GroupByKey, val(r)) for r in b.Sorted[max(0, i - 2):i + 3]])
                break
        print("  in gpu not host:", len(set(ka) - set(kb)), " in host not gpu:", len(set(kb) - set(ka)))
