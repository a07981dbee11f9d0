"""The C-ABI library loads without a GPU, exports every symbol include/sybilgpu.h declares, and the
ctypes mirrors have the header's struct sizes.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

from sybil_b200 import _ffi as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sybilgpu.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(F.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libsybilgpu.so does not export %s" % n
        assert n in F.SYMBOLS, "%s is declared in the header but not bound in _ffi.py" % n
    for n in F.SYMBOLS:
        assert n in names, "%s is bound but not declared in include/sybilgpu.h" % n


def test_struct_sizes_match_the_header():
    structs = ["sg_filter_desc", "sg_group_desc", "sg_agg_desc", "sg_query_desc", "sg_column_desc", "sg_int_info",
               "sg_block_desc", "sg_hist_view", "sg_stats"]
    prog = '#include <stdio.h>\n#include "sybilgpu.h"\nint main(){%s return 0;}' % "".join(
        'printf("%s %%zu\\n", sizeof(%s));' % (s, s) for s in structs)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")], text=True)
    for line in out.strip().splitlines():
        name, size = line.split()
        assert C.sizeof(getattr(F, name)) == int(size), name


def test_abi_version_and_loud_failure_without_gpu():
    lib = F.lib()
    assert lib.sg_abi_version() == F.SG_ABI_VERSION
    import torch
    if torch.cuda.is_available():
        return
    # no device: a context comes back only to carry the message; nothing can be created from it
    st = C.c_int(0)
    ctx = lib.sg_create(0, C.byref(st))
    assert st.value == F.SG_ERR_CUDA
    assert lib.sg_last_error(ctx)
    types = (C.c_int32 * 1)(F.SG_COL_INT)
    assert not lib.sg_table_create(ctx, 1, types)
    assert b"no CPU path" in lib.sg_last_error(ctx)
    lib.sg_destroy(ctx)


def test_blockgen_and_oracle_libraries_load():
    from oracle import oracle_ffi
    assert oracle_ffi.lib().orc_hardware_threads() >= 1
    assert F.gen() is not None
