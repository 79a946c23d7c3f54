"""The merge kernels compact their share of the touched-brick flags in rounds of BRICK_SHARE_CAP map bricks per CTA
(update_tsdf.cu collect_bricks).  At 513^3 one round covers a CTA's share; maps with more than ~450,000 bricks per rank
(BASELINE configs[3]: 2049^3 over 8 GPUs) need several.  This test builds the library with a capacity of 256 -- two
rounds at 513^3 -- and runs the full-size scan against the oracle and a slice of the soak with it, in a fresh process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_merges_with_several_flag_compaction_rounds():
    lib = os.path.join(ROOT, "build", "variants", "libws_cap256.so")
    out = subprocess.run(["bash", os.path.join(ROOT, "tools", "mkvariant.sh"), "cap256", "-DBRICK_SHARE_CAP=256"], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and os.path.exists(lib), out.stdout[-2000:] + out.stderr[-2000:]
    env = dict(os.environ, WS_LIB_PATH=lib)
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_full_size.py"),
                          "-k", "oracle"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and "1 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "soak.py"), "8", "31"], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
