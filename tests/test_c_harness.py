"""The C ABI from plain C: include/radiorust_b200.h must be valid C99 and the shared library must link and run
from a C program (the situation of every FFI: Rust extern "C", cgo, JNI...)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_harness", "abi_check.c")


def build(tmp_path):
    from radiorust_b200 import _ffi, build as b

    if not os.path.exists(_ffi.LIB_PATH):
        b.build()
    gcc = shutil.which("gcc")
    assert gcc, "gcc not found"
    exe = str(tmp_path / "abi_check")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-lradiorust_b200", "-lm", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_c99_and_host_entry_points_work(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_check ok" in r.stdout


@pytest.mark.gpu
def test_block_task_call_sequence_from_c(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_check ok" in r.stdout and "kernel launches" in r.stdout
