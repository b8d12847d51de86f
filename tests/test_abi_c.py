"""include/ace_b200.h compiles as C99 and the library binds through real C linkage ("exercised from
Python ctypes and from a C harness").  The harness only calls host-side entry points unless a GPU is present."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from additivecausalexpansion_b200 import _lib

    _lib.lib()  # raises if the library has not been built
    exe = str(tmp_path / "abi_harness")
    libdir = os.path.dirname(_lib.SO_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c", "abi_harness.c"), "-o", exe, "-L", libdir, "-l:libace_b200.so", "-lm",
           f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_header_is_c99_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("OK ace_b200")


@pytest.mark.gpu
def test_c_harness_runs_the_hot_path_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "GPU residual" in out.stdout
