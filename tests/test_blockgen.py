"""The C++ block generator against the numpy restatement of the digest's encoding rules, and
encode -> decode round trips (the codec TESTPLAN.md:55-62 lists as untested in the reference)."""
import ctypes as C

import numpy as np
import pytest

from sybil_b200 import _ffi as F
from sybil_b200.blocks import decode_column, encode_int_column, encode_str_column, narrow_column
from sybil_b200 import synth


def col_arrays(cd):
    """The descriptor's arrays in the element types its id_bits / value_bits name (sybilgpu.h)."""
    def arr(p, n, t):
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), (n,)).copy() if p and n else np.zeros(0, t)
    idt = C.c_uint16 if cd.id_bits == 16 else C.c_uint32
    vit = {16: C.c_int16, 32: C.c_int32}.get(cd.value_bits, C.c_int64)
    vst = C.c_uint16 if cd.value_bits == 16 else C.c_int32
    return dict(enc=cd.encoding, delta_ids=cd.delta_ids, delta_values=cd.delta_values,
                bin_values=arr(cd.bin_values, cd.nbins, C.c_int64),
                bin_offsets=arr(cd.bin_offsets, cd.nbins + 1, C.c_uint32),
                record_ids=arr(cd.record_ids, cd.nrecord_ids, idt),
                values_i64=arr(cd.values_i64, cd.nvalues if cd.col_type == F.SG_COL_INT else 0, vit),
                values_i32=arr(cd.values_i32, cd.nvalues if cd.col_type == F.SG_COL_STR else 0, vst),
                value_base=cd.value_base)


@pytest.mark.parametrize("narrow", [False, True])
@pytest.mark.parametrize("cfg", ["c2", "c3", "c5"])
def test_generator_matches_numpy_encoder(cfg, narrow):
    spec = synth.config(cfg, total_rows=3 * 4096 + 1000, block_rows=4096)
    spec.narrow = narrow
    store = synth.generate(spec, 0, None, nthreads=2)
    spec.threshold = spec.threshold
    g = F.gen()
    assert g.sbg_num_blocks(store.h) == 4
    for bi in range(4):
        bd = g.sbg_block(store.h, bi).contents
        n = bd.num_records
        assert n == (4096 if bi < 3 else 1000)
        for ci in range(bd.ncols):
            cd = bd.cols[ci]
            sc = [c for c in spec.cols if c.col_slot == cd.col_slot][0]
            vals = np.array([spec.cell(sc, bi * 4096 + r)[0] for r in range(n)], np.int64)
            got = col_arrays(cd)
            if cd.col_type == F.SG_COL_INT:
                ref = encode_int_column(cd.col_slot, vals, np.ones(n, bool), spec.threshold)
            else:
                strs = [sc.prefix + str(v) for v in vals]
                ref = encode_str_column(cd.col_slot, strs, np.ones(n, bool), spec.threshold)
                offs = np.ctypeslib.as_array(C.cast(cd.dict_offsets, C.POINTER(C.c_uint32)), (cd.ndict + 1,))
                blob = C.string_at(cd.dict_bytes, int(offs[-1]))
                table = [blob[offs[i]:offs[i + 1]] for i in range(cd.ndict)]
                assert table == ref.string_table
            if narrow:
                ref = narrow_column(ref)
            assert got["enc"] == ref.encoding
            for k in ("record_ids", "values_i64", "values_i32"):  # same element types as the numpy narrowing picks
                assert got[k].dtype == np.asarray(getattr(ref, k)).dtype or len(got[k]) == 0, (k, got[k].dtype)
            assert got["value_base"] == getattr(ref, "value_base", 0)
            if ref.encoding == F.SG_ENC_BUCKET:
                assert np.array_equal(got["bin_values"], ref.bin_values)
                assert np.array_equal(got["bin_offsets"], ref.bin_offsets)
                assert np.array_equal(got["record_ids"], ref.record_ids)
            elif cd.col_type == F.SG_COL_INT:
                assert np.array_equal(got["values_i64"], ref.values_i64)
            else:
                assert np.array_equal(got["values_i32"], ref.values_i32)


@pytest.mark.parametrize("threshold", [5000, 10])
def test_encode_decode_round_trip_with_missing_rows(threshold):
    rng = np.random.default_rng(0)
    n = 2000
    v = rng.integers(-(1 << 50), 1 << 50, n) if threshold == 10 else rng.integers(0, 300, n)
    valid = rng.random(n) > 0.1
    c = encode_int_column(0, v, valid, threshold)
    dv, pop = decode_column(c, n)
    dn, popn = decode_column(narrow_column(c), n)  # the narrow form decodes to the same rows
    assert np.array_equal(dn, dv) and np.array_equal(popn, pop)
    if c.encoding == F.SG_ENC_BUCKET:
        assert np.array_equal(pop, valid)
        assert np.array_equal(dv[valid], v[valid])
    else:
        # value arrays populate every row up to the last populated one, missing rows read 0 (Q6)
        max_r = np.nonzero(valid)[0][-1] + 1
        assert pop[:max_r].all() and not pop[max_r:].any()
        assert np.array_equal(dv[:max_r], np.where(valid[:max_r], v[:max_r], 0))
