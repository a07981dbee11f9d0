"""Host-side logic of the query mirror that needs no GPU: descriptors for the ABI v4 features and the StrReplace
template conversion (Go's Expand syntax -> Python's)."""
import ctypes as C

from sybil_b200 import _ffi as F
from sybil_b200 import engine as E
from tests.util import INT, SET, STR, Q, Spec


def _spec():
    return Spec([("v", INT), ("w", INT), ("host", STR), ("tags", SET)])


def test_query_desc_carries_set_filters_and_the_weight_column():
    s = _spec()
    q = Q(s, int_filters=[("v", "gt", 3)], str_filters=[("host", "re", "^h")], set_filters=[("tags", "nin", "t1")],
          groups=["host"], aggs=["v"], weight_col="w", order_by="v", order_asc=True, limit=7)
    d, keep = q.desc()
    assert d.abi_version == F.SG_ABI_VERSION == 4
    assert (d.nfilters, d.ngroups, d.naggs, d.weight_col_slot) == (3, 1, 1, s.KeyTable["w"])
    f = [d.filters[i] for i in range(3)]
    assert [(x.col_type, x.op) for x in f] == [(F.SG_COL_INT, F.SG_OP_GT), (F.SG_COL_STR, F.SG_OP_RE), (F.SG_COL_SET, F.SG_OP_NIN)]
    assert C.string_at(f[2].str_value, f[2].str_len) == b"t1" and f[2].col_slot == s.KeyTable["tags"]
    assert (d.order_by_agg, d.order_asc, d.limit) == (0, 1, 7)
    q2 = Q(s, aggs=["v"])
    d2, _ = q2.desc()
    assert d2.weight_col_slot == -1  # FLAGS.WEIGHT_COL is reset per query


def test_str_replace_templates_follow_go_expand():
    # regexp.ReplaceAllString (column_store_io.go:531): $1, ${1}, no group, a literal backslash
    assert E.StrReplace(r"^web-(\d+)\.dc(\d)$", "dc$2/web-$1").apply("web-07.dc3") == "dc3/web-07"
    assert E.StrReplace(r"(\d+)", "<${1}>").apply("a12b3") == "a<12>b<3>"
    assert E.StrReplace(r"[aeiou]", "").apply("sybil-engine") == "sybl-ngn"
    assert E.StrReplace(r"x", r"a\b").apply("xyx") == "a\\bya\\b"
    assert E.StrReplace(r"^nomatch$", "z").apply("kept") == "kept"


def test_load_spec_names_the_files_like_the_reference():
    s = _spec()
    t = type("T", (), {"KeyTable": s.KeyTable})()
    ls = E.LoadSpec(t)
    ls.Int("v"); ls.Str("host"); ls.Set("tags")
    assert sorted(ls.files) == ["int_v.db", "set_tags.db", "str_host.db"]  # table_load_spec.go:59-72
