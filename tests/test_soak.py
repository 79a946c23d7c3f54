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
