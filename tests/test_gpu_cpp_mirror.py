"""Builds and runs tests/cpp/test_host_mirror.cpp: the reference's own unit tests for the
hot path through the C++ host mirror (include/myrio_b200.hpp) on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def build_cpp_mirror_test() -> str:
    src = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
    lib_dir = os.path.join(ROOT, "myrio_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                    src, "-o", EXE, "-L", lib_dir, "-lmyrio_b200", f"-Wl,-rpath,{lib_dir}",
                    "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    return EXE


def test_cpp_mirror_compiles():
    # CPU: the header-only mirror and its test build against the C ABI library
    assert os.path.exists(build_cpp_mirror_test())


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_tests_on_gpu():
    exe = build_cpp_mirror_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout
