"""diagnostic: config at reduced size on the GPU vs the row-value evaluation, per block"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sybil_b200 import engine as E, synth, _ffi as F
from tests.util import Q, Spec

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
spec = synth.config(cfg, total_rows=rows)
store = synth.generate(spec)
s = Spec(spec.key_table); s.IntInfo = dict(spec.IntInfo)
q = Q(s, **synth.query_for(spec))

def run(blocks):
    t = E.Table(cfg, spec.key_table); t.IntInfo = dict(spec.IntInfo)
    for i in blocks:
        t.add_block_desc_ptr(store.block(i))
    qs = q.query_spec()
    ls = t.NewLoadSpec()
    for c in spec.cols:
        (ls.Int if c.col_type == F.SG_COL_INT else ls.Str)(c.name)
    t.LoadAndQueryRecords(ls, qs)
    t.close()
    return qs

nb = store.num_blocks()
for i in range(nb):
    qs = run([i])
    r0, r1 = i * spec.block_rows, min(rows, (i + 1) * spec.block_rows)
    exp = synth.Expected(spec, r0, r1)
    try:
        n = exp.check(qs)
        print("block", i, "ok", n, "values; matched", qs.MatchedCount)
    except AssertionError as e:
        print("block", i, "MISMATCH", repr(e)[:300], "gpu matched", qs.MatchedCount, "expected", exp.matched)
qs = run(range(nb))
exp = synth.Expected(spec)
try:
    print("all blocks ok", exp.check(qs))
except AssertionError as e:
    print("all blocks MISMATCH", repr(e)[:300], qs.MatchedCount, exp.matched)
