"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/effort_b200.h declares, and fails loudly (error codes, never a CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import pytest

from effort_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "effort_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(effort_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    L = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/effort_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in effort_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_strerror():
    L = _lib.load()
    assert L.effort_version() == 200
    assert L.effort_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5, -6):
        assert len(L.effort_strerror(code)) > 3


def test_no_torch_types_in_header():
    src = open(os.path.join(ROOT, "include", "effort_b200.h")).read()
    assert "torch" not in src.replace("No torch", "").replace("no torch", "") and "at::" not in src and "std::" not in src


def test_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    L = _lib.load()
    h = C.c_void_p()
    rc = L.effort_ctx_create(-1, C.byref(h))
    assert rc == -2 and not h.value           # EFFORT_ECUDA, and the error text names the CUDA failure
    assert len(L.effort_last_cuda_error()) > 0
    from effort_b200 import ops
    with pytest.raises(_lib.EffortError):
        ops.Context()


def test_argument_validation_needs_no_gpu():
    L = _lib.load()
    assert L.effort_bucket_mul(None, None, None, None, None, 0.25, None) == -1
    assert L.effort_expert_mul_batch(None, None, 0, None) == -1
    h = C.c_void_p()
    # no buckets and no core: "buckets not loaded" (loader.swift:105-108) with nothing to fall back on
    assert L.effort_weights_create(None, None, None, None, 0, None, 4096, 4096, 1, 16, 0, 0, None, C.byref(h)) == -6
    assert L.effort_weights_create(None, None, None, None, 0, None, 4096, 4096, 1, 16, 7, 0, None, C.byref(h)) == -1
    # convert.swift:210-215 preconditions
    one = C.c_void_p(16)
    assert L.effort_bucketize(one, 4096, 2048, one, one, one, None) == -1      # in < 4096
    assert L.effort_bucketize(one, 3008, 4096, one, one, one, None) == -1      # out < 4096 and 4096 % out != 0
    assert L.effort_bucketize(one, 32016, 4096, one, one, one, None) == -1     # > 32000


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under effort_b200/ or include/ may import, load or link it."""
    import re
    pkg = os.path.join(ROOT, "effort_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                continue
            text = open(os.path.join(base, f), errors="replace").read()
            if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/|import_module\(.oracle", text, re.M):
                offenders.append(os.path.join(base, f))
    assert not offenders, offenders
    # and the library itself has no dependency on it
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libeffort_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
