"""The gob layer either side of the hot path: reader pinned to Go's own output by
the reference's golden files, writer checked byte-for-byte against the same files where the type is
expressible, block directories round-tripped and queried through the oracle."""
import json
import os

import numpy as np
import pytest

from sybil_b200 import _ffi as F
from sybil_b200 import blockdir, gob
from sybil_b200.blocks import encode_block
from tests.util import INT, STR, Q, Spec, compare, run_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EMBEDDED = ("QueryParams", "QueryResults", "BasicHist", "BasicHistCachedInfo")  # anonymous struct fields: JSON flattens them


def flat(d):
    out = {}
    for k, v in d.items():
        if k in EMBEDDED and isinstance(v, dict):
            out.update(flat(v))
        else:
            out[k] = v
    return out


def test_decodes_flag_defs_golden_like_decoding_test_go():
    v = gob.decode(open(os.path.join(GOLD, "flag_defs.golden.gob"), "rb").read())
    j = json.load(open(os.path.join(GOLD, "flag_defs.golden.json")))
    assert len(v) > 5
    for k, x in v.items():
        assert j[k] == x, k
    # gob omits zero values: whatever the JSON holds beyond the decoded fields must be a zero value
    for k, x in j.items():
        if k not in v:
            assert x in (0, False, "", None) or x == [] or x == {}, k


def test_decodes_node_results_golden_gob_to_the_golden_json_values():
    v = gob.decode(open(os.path.join(GOLD, "node_results.golden.gob"), "rb").read())
    want = json.load(open(os.path.join(GOLD, "node_results_hist.json")))  # reduced from node_results.golden.json
    qs = flat(v["QuerySpec"])
    assert qs["MatchedCount"] == want["MatchedCount"]
    assert [g["Name"] for g in qs["Groups"]] == want["Groups"]
    assert qs["Aggregations"][0]["Name"] == want["Aggregation"]
    assert [r["GroupByKey"] for r in qs["Sorted"]] == want["Sorted"]
    assert set(qs["Results"]) == set(want["Results"]) and len(qs["Results"]) == 12

    def check(r, w):
        assert r["Count"] == w["Count"] and r["Samples"] == w["Samples"]
        h = flat(r["Hists"][want["Aggregation"]])
        assert h["__type__"] == "*sybil.HistCompat"
        wh = w["hist"]
        assert h["NumBuckets"] == wh["NumBuckets"] and h["BucketSize"] == wh["BucketSize"]
        assert len(h["Values"]) == wh["nvalues"]
        assert {str(k): c for k, c in enumerate(h["Values"]) if c} == wh["Values"]
        assert h["Max"] == wh["Max"] and h["Min"] == wh["Min"] and h["Count"] == wh["Count"]
        assert h["Avg"] == wh["Avg"]  # float64, bit for bit
        assert h["Info"]["Min"] == wh["InfoMin"] and h["Info"]["Max"] == wh["InfoMax"]

    for k, w in want["Results"].items():
        r = qs["Results"][k]
        assert [ord(c) for c in r["BinaryByKey"]] == w["BinaryByKey"]
        check(r, w)
    assert qs["Cumulative"]["GroupByKey"] == want["Cumulative"]["GroupByKey"]
    check(qs["Cumulative"], want["Cumulative"])


def _descriptor(dec, tid):
    """Encoder descriptor of a type the decoder learnt from the stream."""
    basic = {gob.T_BOOL: "bool", gob.T_INT: "int", gob.T_UINT: "uint", gob.T_FLOAT: "float", gob.T_BYTES: "bytes", gob.T_STRING: "string"}
    if tid in basic:
        return basic[tid]
    t = dec.types[tid]
    if t[0] == "struct":
        return ("struct", t[2], [(n, _descriptor(dec, i)) for n, i in t[1]])
    if t[0] == "slice":
        return ("slice", _descriptor(dec, t[1]))
    if t[0] == "map":
        return ("map", _descriptor(dec, t[1]), _descriptor(dec, t[2]))
    raise AssertionError(t)


def test_writer_reproduces_go_bytes_for_flag_defs():
    """Re-encoding the decoded golden value with the type the stream declared gives Go's bytes back:
    type-definition message, field deltas, zero-value omission, varints — all byte for byte."""
    raw = open(os.path.join(GOLD, "flag_defs.golden.gob"), "rb").read()
    dec = gob.Decoder(raw)
    v = dec.decode()
    top = min(t for t in dec.types if t >= gob.FIRST_USER_ID)  # the first user type is the top-level struct
    again = gob.encode(v, _descriptor(dec, top))
    assert raw.endswith(b"\n")  # PrintBytes appends one (printer.go:272-282)
    assert again == raw[:-1]


def test_writer_reader_round_trip_all_kinds():
    rng = np.random.default_rng(5)
    inner = ("struct", "Inner", [("A", "int"), ("B", ("slice", "uint")), ("S", "string")])
    t = ("struct", "Outer", [("Flag", "bool"), ("N", "int"), ("U", "uint"), ("X", "float"), ("Name", "string"), ("Raw", "bytes"),
                             ("Ints", ("slice", "int")), ("Items", ("slice", inner)), ("M", ("map", "string", inner)),
                             ("Counts", ("map", "int", "int")), ("Empty", ("slice", "int"))])
    v = {"Flag": True, "N": -(1 << 62), "U": (1 << 64) - 1, "X": -1234.5e-7, "Name": "héllo\tworld", "Raw": b"\x00\xff\x80",
         "Ints": [0, -1, 1, 127, 128, -129, 1 << 40, -(1 << 63)], "Items": [{"A": 5, "B": [1, 2, 300000], "S": "x"}, {"A": -7}],
         "M": {"k1": {"A": 1}, "k2": {"B": [7], "S": "z"}}, "Counts": {-5: 10, 70000: -3}}
    got = gob.decode(gob.encode(v, t))
    assert got == v  # zero-valued / empty fields are not sent and so not present
    big = [int(x) for x in rng.integers(-(1 << 62), 1 << 62, 5000)]
    assert gob.decode(gob.encode(big, ("slice", "int"))) == big
    assert gob.decode(gob.encode(3.25, "float")) == 3.25 and gob.decode(gob.encode("s", "string")) == "s"


@pytest.mark.parametrize("compress", [False, True])
def test_block_directory_round_trip_and_query(tmp_path, compress):
    """Blocks digested here, written as sybil block directories (int_*.db / str_*.db / info.db), read back
    and queried: arrays identical, query results identical to the in-memory blocks."""
    rng = np.random.default_rng(11)
    n = 3000
    s = Spec([("age", INT), ("big", INT), ("host", STR), ("uid", STR)])
    s.add_rows({"age": rng.integers(10, 30, n), "big": rng.integers(-(1 << 40), 1 << 45, n),
                "host": np.array(["h%d" % x for x in rng.integers(0, 5, n)]),
                "uid": np.array(["u%d" % x for x in rng.integers(0, 4 * n, n)])},
               {"age": rng.random(n) > 0.1, "host": rng.random(n) > 0.1}, threshold=50, block_rows=1000)
    back = Spec(s.key_table)
    back.IntInfo = dict(s.IntInfo)
    for i, b in enumerate(s.blocks):
        d = str(tmp_path / ("block%d" % i))
        blockdir.write_block_dir(d, b, s.key_table, compress=compress)
        names = sorted(os.listdir(d))
        assert ("info.db.gz" if compress else "info.db") in names and any(x.startswith("int_age.db") for x in names)
        r = blockdir.read_block_dir(d, s.key_table, block_index=b.block_index)
        assert r.num_records == b.num_records and r.info == b.info
        for c0, c1 in zip(b.cols, r.cols):
            assert (c0.col_slot, c0.col_type, c0.encoding, c0.delta_ids, c0.delta_values) == (
                c1.col_slot, c1.col_type, c1.encoding, c1.delta_ids, c1.delta_values)
            for f in ("bin_values", "bin_offsets", "record_ids", "values_i64", "values_i32"):
                assert np.array_equal(getattr(c0, f), getattr(c1, f)), f
            assert c0.string_table == c1.string_table
        back.blocks.append(r)
    # a LoadSpec that names only some columns: the others stay absent
    part = blockdir.read_block_dir(str(tmp_path / "block0"), s.key_table, columns={"age", "host"})
    assert sorted(c.col_slot for c in part.cols) == [0, 2]
    q = Q(s, int_filters=[("age", "gt", 12)], groups=["host"], aggs=["big", "age"], op="hist")
    a, b = run_oracle(s, q), run_oracle(back, Q(back, int_filters=[("age", "gt", 12)], groups=["host"], aggs=["big", "age"], op="hist"))
    assert a.MatchedCount == b.MatchedCount and set(a.Results) == set(b.Results)
    for k in a.Results:
        assert a.Results[k].Count == b.Results[k].Count
        for name in ("big", "age"):
            assert a.Results[k].Hists[name].Count == b.Results[k].Hists[name].Count
            assert a.Results[k].Hists[name].ExactSum == b.Results[k].Hists[name].ExactSum


def test_writer_reader_round_trip_property():
    """Random values of random shapes survive encode -> decode (zero / empty fields excepted: gob drops them)."""
    from hypothesis import given, settings, strategies as st

    ints = st.integers(min_value=-(1 << 63), max_value=(1 << 63) - 1)
    inner_t = ("struct", "In", [("A", "int"), ("U", "uint"), ("S", "string"), ("F", "float"), ("L", ("slice", "int"))])
    outer_t = ("struct", "Out", [("I", inner_t), ("Items", ("slice", inner_t)), ("M", ("map", "string", inner_t)),
                                 ("K", ("map", "int", ("slice", "uint"))), ("B", "bool"), ("Raw", "bytes")])
    inner = st.fixed_dictionaries({"A": ints, "U": st.integers(0, (1 << 64) - 1), "S": st.text(max_size=12),
                                   "F": st.floats(allow_nan=False), "L": st.lists(ints, max_size=6)})
    outer = st.fixed_dictionaries({"I": inner, "Items": st.lists(inner, max_size=4), "M": st.dictionaries(st.text(max_size=5), inner, max_size=3),
                                   "K": st.dictionaries(ints, st.lists(st.integers(0, 1 << 40), min_size=1, max_size=4), max_size=3),
                                   "B": st.booleans(), "Raw": st.binary(max_size=9)})

    def drop_zero_fields(v, t):
        if isinstance(t, str):
            return v
        if t[0] == "struct":
            out = {}
            for n, ft in t[2]:
                x = drop_zero_fields(v[n], ft)
                zero = (x in (0, 0.0, False, "", b"") and isinstance(ft, str)) or (not isinstance(ft, str) and ft[0] != "struct" and len(x) == 0)
                if not zero:
                    out[n] = x
            return out
        if t[0] == "slice":
            return [drop_zero_fields(x, t[1]) for x in v]
        return {k: drop_zero_fields(x, t[2]) for k, x in v.items()}

    @settings(max_examples=150, deadline=None)
    @given(outer)
    def check(v):
        assert gob.decode(gob.encode(v, outer_t)) == drop_zero_fields(v, outer_t)

    check()
