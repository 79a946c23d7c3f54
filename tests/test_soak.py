"""Randomised soak of the CUDA path against the oracle, a small slice of tests/soak.py in the GPU suite:
random map sizes, resolutions, tau, sensor poses, ring offsets, slab / stripe sharding, cloud sizes.
`python tests/soak.py 300` runs the long version by hand."""
import numpy as np
import pytest

import soak

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 1234])
def test_soak_random_cases(seed):
    rng = np.random.default_rng(seed)
    for case in range(12):
        soak.one_case(rng, case)


def test_soak_with_the_optional_march_launch_shapes():
    """The march's optional launch shapes (the near field marched in two parts, a far-field launch that joins after the
    replay's record pass, three far-field CTAs per SM; update_tsdf.cu ws_update_enqueue) are read from the
    environment when the library first enqueues an update: a fresh process runs a slice of the soak and the
    full-size scan with them turned on -- results must not depend on how the work items are handed out."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WS_NEAR_SPLIT="5", WS_LS_GRID_F2="1", WS_LS_GRID_F="3", WS_LS_GRID_N="2")
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "soak.py"), "10", "99"], env=env, cwd=root,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(root, "tests", "test_full_size.py"),
                          "-k", "oracle"], env=env, cwd=root, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and "1 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
