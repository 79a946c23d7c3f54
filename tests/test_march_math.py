"""warpsense_b200/csrc/march_math.cuh on the host: the fast 32-bit arithmetic of the ray march (DDA
projection, 32-bit magic divisions, FP64-estimated quotients) equals the literal formulas of
update_tsdf.cpp:450-506 for every ray the fast path admits.  No GPU needed (nvcc host compile)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_march_math.cu")
EXE = os.path.join(ROOT, "tests", "cpp", "test_march_math")


def test_march_math_host():
    subprocess.check_call(["nvcc", "-std=c++17", "-O2", "-fmad=false", "-Wno-deprecated-gpu-targets", SRC, "-o", EXE])
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "march math ok" in out.stdout
