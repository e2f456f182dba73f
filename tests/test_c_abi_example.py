"""examples/c_abi_example.c: the hot path driven from plain C through include/myrio_b200.h --
what the reference-side binding (cgo / Rust -sys / JNI) does.  CPU: it compiles as C99, links
against libmyrio_b200.so and fails loudly without a GPU; GPU: it runs and every read's best
leaf is the leaf it was cut from."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    lib_dir = os.path.join(ROOT, "myrio_b200")
    if not os.path.exists(os.path.join(lib_dir, "libmyrio_b200.so")):
        pytest.skip("libmyrio_b200.so not built")
    exe = str(tmp_path / "c_abi_example")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_example.c"), "-o", exe, "-L", lib_dir, "-lmyrio_b200",
                    f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_c_example_links_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "read 0: best leaf 1" in r.stdout and "read 1: best leaf 2" in r.stdout
