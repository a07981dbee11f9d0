"""The cgo shim (go/sybilgpu/*.go) cannot be compiled here (no Go toolchain), but it can be kept in step with
the header: every C function, type and constant it names must exist in include/sybilgpu.h / sybilgob.h, and
every struct field it assigns through a descriptor variable must be a field of that struct.  Checked by
generating a C translation unit from the Go sources and running gcc -fsyntax-only on it."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def test_cgo_names_and_fields_exist_in_the_headers(tmp_path):
    srcs = sorted(glob.glob(os.path.join(ROOT, "go", "sybilgpu", "*.go")))
    assert srcs
    names, fields = set(), set()
    for path in srcs:
        text = open(path).read()
        names |= set(re.findall(r"\bC\.((?:sg|sgob|SG)_\w+)", text))
        # fields assigned through `var x C.sg_..._desc` variables, and through the per-item pointer `d`
        for var, struct in re.findall(r"\bvar (\w+) C\.(sg_\w+_desc)\b", text):
            for f in re.findall(r"\b%s\.(\w+)\s*(?:,[^=\n]*)?=[^=]" % re.escape(var), text):
                fields.add(("any_desc", f))  # (the shim reuses short names like `d` for several descriptors)
        for f in re.findall(r"\bd\.(\w+)\s*(?:,\s*d\.\w+\s*)*=", text):
            fields.add(("any_desc", f))
    header = open(os.path.join(ROOT, "include", "sybilgpu.h")).read() + open(os.path.join(ROOT, "include", "sybilgob.h")).read()
    lines = ['#include <stddef.h>', '#include "sybilgpu.h"', '#include "sybilgob.h"', "void check(void) {"]
    for n in sorted(names):
        if n.startswith("SG_"):
            lines.append("  (void)(%s);" % n)  # an enumerator / macro
        elif re.search(r"\b%s\s*\(" % n, header) and not re.search(r"typedef struct %s\b" % n, header):
            lines.append("  (void)&%s;" % n)  # a function
        elif False:
            lines.append("  (void)(%s);" % n)  # an enumerator / macro
        else:
            lines.append("  { %s* p = 0; (void)p; }" % n)  # a type (possibly opaque)
    descs = ["sg_column_desc", "sg_query_desc", "sg_block_desc", "sg_filter_desc", "sg_group_desc", "sg_agg_desc", "sg_int_info"]
    for struct, f in sorted(fields):
        if struct == "any_desc":  # `d.` is used for several descriptors: the field must exist in one of them
            ok = [s for s in descs if re.search(r"\b%s\b[^;]*;" % f, header[header.index("typedef struct " + s):header.index("} " + s)])]
            assert ok, "go shim assigns d.%s: no descriptor struct has that field" % f
            lines.append("  (void)offsetof(%s, %s);" % (ok[0], f))
        else:
            lines.append("  (void)offsetof(%s, %s);" % (struct, f))
    lines.append("}")
    c = tmp_path / "shim_check.c"
    c.write_text("\n".join(lines) + "\n")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert len(names) > 30 and len(fields) > 20
