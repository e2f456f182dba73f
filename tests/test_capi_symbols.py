"""CPU-only: the C-ABI library loads and exports every symbol include/myrio_b200.h
declares (no compute calls without a GPU), and the product package does not
touch the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "myrio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(myrio_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from myrio_b200 import _lib
    L = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert L.myrio_version() >= 100


def test_no_gpu_fails_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from myrio_b200 import api
    with pytest.raises(api.MyrioError) as e:
        api.Context(0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "myrio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", text, re.M), f
                assert "libmyrio_oracle" not in text, f
                assert "myrio_oracle.h" not in text, f


def test_header_is_plain_c99():
    """The boundary is a C ABI: include/myrio_b200.h must compile as C (what cgo / bindgen / ctypes
    authors read), with no C++ or CUDA types in any signature."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "myrio_b200.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
