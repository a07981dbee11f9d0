"""The native (C++) block-directory reader, include/sybilgob.h: the same decode the Python reader does
(sybil_b200/gob.py, pinned to Go's output by the reference's golden gob files), producing the sg_block_desc
the C ABI takes.  Checked array for array against the Python reader and through the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from sybil_b200 import _ffi as F
from sybil_b200 import blockdir
from tests.util import INT, STR, Q, Spec, run_oracle


def _read_native(d, key_table, mask=None, block_index=0):
    g = F.gobread()
    names = (C.c_char_p * len(key_table))(*[n.encode() for n, _ in key_table])
    types = (C.c_int32 * len(key_table))(*[t for _, t in key_table])
    m = (C.c_uint8 * len(key_table))(*mask) if mask is not None else None
    err = C.create_string_buffer(512)
    h = g.sgob_read_block_dir(d.encode(), names, types, len(key_table), m, block_index, err, len(err))
    return g, h, err.value.decode()


def _arr(ptr, n, dtype):
    """numpy copy of n items at a raw address (the descriptor's pointer fields are void*)."""
    if not n:
        return np.zeros(0, dtype)
    ct = {np.int64: C.c_int64, np.uint32: C.c_uint32, np.int32: C.c_int32}[dtype]
    return np.frombuffer((ct * int(n)).from_address(int(ptr)), dtype=dtype).copy()


@pytest.mark.parametrize("compress", [False, True])
def test_native_reader_matches_python_reader_and_oracle(tmp_path, compress):
    rng = np.random.default_rng(21)
    n = 4000
    s = Spec([("age", INT), ("big", INT), ("host", STR), ("uid", STR), ("never", INT)])
    s.add_rows({"age": rng.integers(10, 30, n), "big": rng.integers(-(1 << 40), 1 << 45, n),
                "host": np.array(["h%d" % x for x in rng.integers(0, 5, n)]),
                "uid": np.array(["u%d" % x for x in rng.integers(0, 4 * n, n)])},
               {"age": rng.random(n) > 0.1, "host": rng.random(n) > 0.1}, threshold=50, block_rows=1500)
    from oracle.oracle_ffi import OracleTable
    ot = OracleTable(s.key_table)
    keep = []
    try:
        for i, b in enumerate(s.blocks):
            d = str(tmp_path / ("block%d" % i))
            blockdir.write_block_dir(d, b, s.key_table, compress=compress)
            want = blockdir.read_block_dir(d, s.key_table, block_index=b.block_index)
            g, h, err = _read_native(d, s.key_table, block_index=b.block_index)
            assert h, err
            keep.append((g, h))
            assert g.sgob_block_bytes(h) > 0
            desc = g.sgob_block_desc(h).contents
            assert desc.num_records == b.num_records and desc.block_index == b.block_index
            assert desc.ncols == len(want.cols)
            assert {desc.info[k].col_slot: (desc.info[k].min, desc.info[k].max) for k in range(desc.ninfo)} == want.info
            for k, wc in enumerate(want.cols):
                cd = desc.cols[k]
                assert (cd.col_slot, cd.col_type, cd.encoding, bool(cd.delta_ids), bool(cd.delta_values)) == (
                    wc.col_slot, wc.col_type, wc.encoding, wc.delta_ids, wc.delta_values)
                if wc.encoding == F.SG_ENC_BUCKET:
                    assert np.array_equal(_arr(cd.bin_values, cd.nbins, np.int64), wc.bin_values)
                    assert np.array_equal(_arr(cd.bin_offsets, cd.nbins + 1, np.uint32), wc.bin_offsets)
                    assert np.array_equal(_arr(cd.record_ids, cd.nrecord_ids, np.uint32), wc.record_ids)
                elif wc.col_type == F.SG_COL_INT:
                    assert np.array_equal(_arr(cd.values_i64, cd.nvalues, np.int64), wc.values_i64)
                else:
                    assert np.array_equal(_arr(cd.values_i32, cd.nvalues, np.int32), wc.values_i32)
                if wc.col_type == F.SG_COL_STR:
                    offs = _arr(cd.dict_offsets, cd.ndict + 1, np.uint32)
                    blob = C.string_at(cd.dict_bytes, int(offs[-1])) if cd.ndict else b""
                    assert [blob[offs[j]:offs[j + 1]] for j in range(cd.ndict)] == wc.string_table
            ot.add_block(g.sgob_block_desc(h))
        # the oracle over the natively decoded descriptors == the oracle over the in-memory blocks
        q = Q(s, int_filters=[("age", "gt", 12)], groups=["host"], aggs=["big", "age"], op="hist")
        d, _keep = q.desc()
        o2 = ot.query(d, q.aggs, nthreads=2)
        o1 = run_oracle(s, q)
        assert o1.MatchedCount == o2.MatchedCount and set(o1.Results) == set(o2.Results)
        for k in o1.Results:
            assert o1.Results[k].Count == o2.Results[k].Count
            for name in ("big", "age"):
                assert o1.Results[k].Hists[name].ExactSum == o2.Results[k].Hists[name].ExactSum
                assert np.array_equal(o1.Results[k].Hists[name].Values, o2.Results[k].Hists[name].Values)
    finally:
        for g, h in keep:
            g.sgob_block_free(h)
        ot.close()


def test_native_reader_load_mask_and_errors(tmp_path):
    rng = np.random.default_rng(22)
    s = Spec([("a", INT), ("b", STR)])
    s.add_rows({"a": rng.integers(0, 9, 500), "b": np.array(["x%d" % v for v in rng.integers(0, 3, 500)])}, block_rows=500)
    d = str(tmp_path / "blk")
    blockdir.write_block_dir(d, s.blocks[0], s.key_table)
    g, h, err = _read_native(d, s.key_table, mask=[0, 1])
    assert h and g.sgob_block_desc(h).contents.ncols == 1 and g.sgob_block_desc(h).contents.cols[0].col_slot == 1
    g.sgob_block_free(h)
    g, h, err = _read_native(str(tmp_path / "nope"), s.key_table)
    assert not h and "info.db" in err
    raw = open(os.path.join(d, "int_a.db"), "rb").read()
    open(os.path.join(d, "int_a.db"), "wb").write(raw[:len(raw) // 2])  # truncated column file
    g, h, err = _read_native(d, s.key_table)
    assert not h and "gob" in err


def test_native_reader_exports_every_declared_symbol():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "sybilgob.h")).read()
    declared = set(re.findall(r"\b(sgob_[a-z_]+)\s*\(", hdr))
    assert declared == set(F.GOB_SYMBOLS), (declared, set(F.GOB_SYMBOLS))
    lib = C.CDLL(F.GOB_PATH)
    for name in declared:
        getattr(lib, name)


def test_native_table_directory_matches_python_reader(tmp_path):
    from sybil_b200 import tabledir
    rng = np.random.default_rng(41)
    n = 3000
    s = Spec([("age", INT), ("lat", INT), ("host", STR)])
    s.add_rows({"age": rng.integers(10, 30, n), "lat": rng.integers(30, 9000, n),
                "host": np.array(["h%d" % x for x in rng.integers(0, 6, n)])}, block_rows=1000)
    tdir = tabledir.write_table(str(tmp_path), "tbl", s.key_table, s.blocks, s.IntInfo)
    for junk in ("ingest", "cache", "stomache9", "b.partial", "c.broken"):
        os.makedirs(os.path.join(tdir, junk))
    want = tabledir.read_table(str(tmp_path), "tbl")
    g = F.gobread()
    err = C.create_string_buffer(256)
    t = g.sgob_table_open(str(tmp_path).encode(), b"tbl", err, len(err))
    assert t, err.value
    try:
        nc = g.sgob_table_num_cols(t)
        assert [(g.sgob_table_col_name(t, i).decode(), g.sgob_table_col_type(t, i)) for i in range(nc)] == want.key_table
        for i, (name, _) in enumerate(want.key_table):
            mn, mx = C.c_int64(), C.c_int64()
            has = g.sgob_table_int_info(t, i, C.byref(mn), C.byref(mx))
            assert (has == 1) == (name in want.IntInfo)
            if has:
                assert (mn.value, mx.value) == want.IntInfo[name]
        assert [g.sgob_table_block_dir(t, i).decode() for i in range(g.sgob_table_num_blocks(t))] == want.block_dirs
    finally:
        g.sgob_table_free(t)
    assert not g.sgob_table_open(str(tmp_path).encode(), b"missing", err, len(err)) and b"info.db" in err.value


def test_native_example_host_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/query_dir.cpp (table dir -> gob decode -> stage -> query, no Go) builds against both headers;
    on a machine without a CUDA device it must stop at sg_create with an error, never fall back."""
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "query_dir")
    csrc = os.path.join(root, "sybil_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "query_dir.cpp"),
                           "-L", csrc, "-lsybilgob", "-lsybilgpu", "-lz", "-Wl,-rpath," + csrc, "-o", exe])
    if not torch.cuda.is_available():
        p = subprocess.run([exe, str(tmp_path), "nosuchtable", "g", "a"], capture_output=True, text=True)
        assert p.returncode != 0 and "info.db" in p.stderr


@pytest.mark.parametrize("threshold", [5000, 4])
def test_native_reader_set_columns(tmp_path, threshold):
    # set_<col>.db in both file forms (bucketed; Values [][]int32 turned into bins + nvalues): the native reader's
    # descriptor gives the oracle the same answers as the in-memory blocks
    from tests.util import SET
    from oracle.oracle_ffi import OracleTable
    rng = np.random.default_rng(3)
    n = 2500
    s = Spec([("v", INT), ("tags", SET)])
    s.forms = "wide"  # (write_block_dir lays out the post-gob arrays as they are)
    s.add_rows({"v": rng.integers(0, 100, n),
                "tags": [["t%d" % t for t in rng.choice(9, int(k), replace=False)] for k in rng.integers(0, 4, n)]},
               threshold=threshold, block_rows=900)
    ot = OracleTable(s.key_table)
    keep = []
    try:
        for i, b in enumerate(s.blocks):
            d = str(tmp_path / ("block%d" % i))
            blockdir.write_block_dir(d, b, s.key_table, compress=i % 2 == 0)
            g, h, err = _read_native(d, s.key_table, block_index=b.block_index)
            assert h, err
            keep.append((g, h))
            desc = g.sgob_block_desc(h).contents
            cd = [desc.cols[k] for k in range(desc.ncols) if desc.cols[k].col_type == SET][0]
            assert cd.encoding == F.SG_ENC_BUCKET and cd.ndict == 9
            assert (cd.nvalues > 0) == (threshold == 4)
            ot.add_block(g.sgob_block_desc(h))
        for op, tag in (("in", "t3"), ("nin", "t3"), ("nin", "nope")):
            q = Q(s, set_filters=[("tags", op, tag)], aggs=["v"])
            want = run_oracle(s, q)
            d, _keep = q.desc()
            got = ot.query(d, q.aggs)
            assert got.MatchedCount == want.MatchedCount > 0
            assert got.Cumulative.Hists["v"].ExactSum == want.Cumulative.Hists["v"].ExactSum
    finally:
        ot.close()
        for g, h in keep:
            g.sgob_block_free(h)
