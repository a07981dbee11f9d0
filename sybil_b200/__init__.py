"""sybil_b200 — B200-native scan-and-aggregate engine behind sybil's LoadAndQueryRecords.

csrc/           CUDA kernels (sm_100a) + host runtime -> libsybilgpu.so (C ABI: include/sybilgpu.h)
engine.py       host-side mirror of the reference's query interface (ctypes over the C ABI)
blocks.py       post-gob column blocks + the digest's encoding rules (test/bench input)
_build.py       nvcc / g++ build recipes
"""
from . import _ffi  # noqa: F401
