"""include/warpsense_b200.hpp: the reference's classes as C++ shims over the C ABI.

CPU: the header and its test program compile and link against libwarpsense_b200.so.
GPU: the program runs (G1 known-answer vector + a registration round trip + a device-side map shift)."""
import os
import subprocess

import pytest

from warpsense_b200 import build as wb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def _compile():
    wb.build_library()
    libdir = os.path.dirname(wb.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", libdir, "-lwarpsense_b200", "-Wl,-rpath," + libdir, "-lpthread"]
    subprocess.check_call(cmd)


def test_cpp_shim_compiles_and_links():
    _compile()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_shim_runs_on_gpu():
    _compile()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp shim ok" in out.stdout
