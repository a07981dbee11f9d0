"""Build recipes (in-tree, so the .so files travel to the GPU box with the snapshot)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_gpu(force=False):
    """libsybilgpu.so: hand-written sm_100a kernels + runtime, cudart linked statically.  The kernel unit is
    compiled twice (sg_internal.h): 16-warp CTAs, one per SM, and 8-warp CTAs, two per SM; the three compilations
    run in parallel."""
    out = os.environ.get("SG_LIB_OUT") or os.path.join(CSRC, "libsybilgpu.so")
    srcs = [os.path.join(CSRC, f) for f in ("sg_kernels.cu", "sg_runtime.cu")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("sg_internal.h", "sg_hist.h")] + [
        os.path.join(ROOT, "include", "sybilgpu.h")]
    if not force and not _newer(out, deps):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    common = [nvcc] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-diag-suppress", "186"] + \
        os.environ.get("SG_NVCC_FLAGS", "").split()
    jobs = [(common + ["-DSG_THREADS=512", "-c", srcs[0], "-o", os.path.join(bdir, "sg_kernels_w16.o")]),
            (common + ["-DSG_THREADS=256", "-c", srcs[0], "-o", os.path.join(bdir, "sg_kernels_w8.o")]),
            (common + ["-c", srcs[1], "-o", os.path.join(bdir, "sg_runtime.o")])]
    # an object whose own sources are older than it is kept (the two kernel units take minutes each); the flags
    # it was built with are part of the check
    hdrs = [os.path.join(CSRC, "sg_internal.h"), os.path.join(ROOT, "include", "sybilgpu.h")]
    obj_deps = [[srcs[0]] + hdrs, [srcs[0]] + hdrs, [srcs[1], os.path.join(CSRC, "sg_hist.h")] + hdrs]
    stale = []
    for cmd, deps_o in zip(jobs, obj_deps):
        stamp = cmd[-1] + ".cmd"
        same_cmd = os.path.exists(stamp) and open(stamp).read() == " ".join(cmd)
        stale.append(force or not same_cmd or _newer(cmd[-1], deps_o))
    todo = [cmd for cmd, st in zip(jobs, stale) if st]
    procs = [(cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for cmd in todo]
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), log))
        if os.environ.get("SG_BUILD_VERBOSE"):
            print(log)
        with open(cmd[-1] + ".cmd", "w") as f:
            f.write(" ".join(cmd))
    _run([nvcc] + NVCC_ARCH + ["-shared", "-o", out] + [j[-1] for j in jobs] + ["-ldl"])
    return out


def build_blockgen(force=False):
    out = os.path.join(CSRC, "libsybilblockgen.so")
    src = os.path.join(CSRC, "blockgen.cpp")
    if not force and not _newer(out, [src, os.path.join(ROOT, "include", "sybilgpu.h")]):
        return out
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


def build_oracle(force=False):
    """oracle/liboracle.so — the CPU restatement (test infrastructure, never loaded by the product)."""
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "liboracle.so")
    src = os.path.join(odir, "oracle.cpp")
    if not force and not _newer(out, [src, os.path.join(ROOT, "include", "sybilgpu.h")]):
        return out
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out, src])
    return out


def build_gobread(force=False):
    """libsybilgob.so: sybil block directory (gob column files) -> sg_block_desc, host-side C++ (include/sybilgob.h)."""
    out = os.path.join(CSRC, "libsybilgob.so")
    src = os.path.join(CSRC, "gobread.cpp")
    if not force and not _newer(out, [src, os.path.join(ROOT, "include", "sybilgob.h"), os.path.join(ROOT, "include", "sybilgpu.h")]):
        return out
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src, "-lz"])
    return out


def build_all(force=False):
    return [build_gpu(force), build_blockgen(force), build_gobread(force), build_oracle(force)]
