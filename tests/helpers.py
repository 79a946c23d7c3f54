"""Shared helpers of the parity tests: build the same map in the oracle and on the device."""
import numpy as np

from oracle import oracle as orc
from warpsense_b200 import api


def make_pair(size, tau, max_weight, res, device=0):
    """(oracle LocalMap, TSDFCuda) holding identical, default-filled maps of side `size` (3-tuple)."""
    om = orc.LocalMap(size[0], size[1], size[2], tau, 0)
    hm = api.HostLocalMap(size[0], size[1], size[2], tau, 0)
    assert (hm.size == om.size).all()
    tsdf = api.TSDFCuda(api.DeviceMap(hm), tau, max_weight, res, device=device)
    return om, hm, tsdf


def device_grid(tsdf, hm):
    """Download the device map into the host ring array and return it (uint32, reference layout)."""
    tsdf.avg_map().to_host(api.DeviceMap(hm))
    return hm.data


def assert_same_grid(om, tsdf, hm, what=""):
    got = device_grid(tsdf, hm)
    want = om.data
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        sz = om.size
        i = int(bad[0])
        rx, ry, rz = i // (sz[1] * sz[2]), (i // sz[2]) % sz[1], i % sz[2]
        gv = (np.int16(got[i] & 0xFFFF), np.int16(got[i] >> 16))
        wv = (np.int16(want[i] & 0xFFFF), np.int16(want[i] >> 16))
        raise AssertionError("%s: %d/%d voxels differ; first at ring (%d,%d,%d): device %s oracle %s"
                             % (what, len(bad), len(got), rx, ry, rz, gv, wv))


def random_cloud(rng, n, lo, hi):
    return rng.integers(lo, hi, size=(n, 3), dtype=np.int64).astype(np.int32)
