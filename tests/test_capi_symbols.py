"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/zigzag_gpu.h declares
(no compute calls here).  The product has no CPU fallback: creating a handle without a CUDA device must fail loudly."""
import ctypes
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "zigzag_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 39
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_binding_covers_the_header(built_lib):
    from zigzag_b200 import _capi
    assert sorted(_capi.EXPORTS) == declared_symbols()
    _capi.load()


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C11 with no CUDA or C++ in sight."""
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "zigzag_gpu.h")])


def test_sass_is_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built_lib):
    import torch
    from zigzag_b200 import _capi
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is not reachable")
    with pytest.raises(_capi.ZigzagGpuError, match="no CPU fallback"):
        _capi.Handle(100, 2, 2)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under zigzag_b200/ may import, load or link it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "zigzag_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "libzzoracle" not in txt and "zz_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_the_library_is_not_older_than_its_sources():
    """A failed nvcc run leaves the previous library in place and every later measurement silently uses it: the CPU suite (which
    the driver runs right after build()) refuses a library that is older than any file of csrc/ or the ABI header."""
    from zigzag_b200 import build as zb
    assert os.path.exists(zb.OUT), "build the library first: python -m zigzag_b200.build"
    assert not zb.needs_build(), "zigzag_b200/libzigzag_gpu.so is older than its sources: the last build failed or was not run"
