"""BASELINE.json's full sizes (128x1024 OS1-128 scan, 513^3 @ 5 cm): one scan against the oracle bit for bit
(about 6 s of single-threaded CPU work), then size-independent properties on the following scans --
results independent of the number of slabs the map is sharded into, registration sums equal to the oracle's on
the device-built map, counter identities, and run-to-run determinism."""
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as orc
from warpsense_b200 import api, fixedpoint as fp
from warpsense_b200.synth import ScanStream

pytestmark = pytest.mark.gpu

RES, SIDE, TAU, MW = 50, 512, 1000, 640


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).view(np.uint8))


@pytest.fixture(scope="module")
def stream():
    return ScanStream(128, 1024, SIDE, RES)


def test_full_size_update_and_registration_match_oracle(stream):
    s = stream
    size = (SIDE + 1,) * 3
    om = orc.LocalMap(*size, TAU, 0)
    hm = api.HostLocalMap(*size, TAU, 0)
    tsdf = api.TSDFCuda(api.DeviceMap(hm), TAU, MW, RES)
    reg = api.RegistrationCuda(tsdf)
    f0 = s.frame(0)
    assert len(f0["points_map"]) == 128 * 1024
    pos, up = fp.convert_pose_to_gpu(f0["pose"], RES)
    st = orc.update_tsdf(om, f0["points_map"], pos, up, TAU, MW, RES)            # the reference path, 1 thread
    tsdf.update_tsdf(f0["points_map"], pos, up)
    c = tsdf.counters()
    assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"])
    assert c["n_written"] <= c["n_touched"] <= c["n_candidates"]
    tsdf.avg_map().to_host(api.DeviceMap(hm))
    assert np.array_equal(hm.data, om.data), "513^3 grid after one full scan differs from the oracle"
    # registration of the next scan against that map: 20 iterations, every int64 sum and the cloud bit-exact
    cloud = s.frame(1, prior_pose=s.pose(0))["points_prior"].copy()
    ocloud = cloud.copy()
    I = np.eye(4, dtype=np.float32)
    oT, oit, otr = orc.register_cloud(om, ocloud, I, 20, 0.1, 0.0, RES, trace=True)
    T, it = reg.register_cloud(cloud, I, 20, 0.1, 0.0, RES)
    assert it == oit == 20
    assert np.array_equal(reg.trace()[:it], otr[:it])
    assert np.abs(T[:3, :3] - oT[:3, :3]).max() <= 1e-4 and np.abs(T[:3, 3] - oT[:3, 3]).max() <= 0.1   # mm
    assert np.array_equal(cloud, ocloud)
    tsdf.close()


def test_full_size_properties_sharding_and_determinism(stream):
    s = stream
    size = (SIDE + 1,) * 3

    class _View:
        size_ = np.array(size, np.int32)
        offset_ = np.array([v // 2 for v in size], np.int32)
        pos_ = np.zeros(3, np.int32)
        data_ = None

    def run(world, stripe=0):
        """three scans through `world` slab-sharded handles on this GPU; returns per-scan counters, the grids'
        owned slabs stitched together (crc) and the registration traces"""
        if stripe:
            os.environ["WS_STRIPE_COLS"] = str(stripe)
        try:
            hs = [api.TSDFCuda(_View(), TAU, MW, RES, rank=r, world=world, upload=False) for r in range(world)]
        finally:
            os.environ.pop("WS_STRIPE_COLS", None)
        counters = []
        for k in range(3):
            f = s.frame(k)
            pos, up = fp.convert_pose_to_gpu(f["pose"], RES)
            cs = []
            for t in hs:
                t.update_tsdf(f["points_map"], pos, up)
                cs.append(t.counters())
            counters.append(cs)
        row = size[1] * size[2]
        whole = np.zeros((size[0], row), np.uint32)
        back = api.HostLocalMap(*size, TAU, 0)
        for r, t in enumerate(hs):
            rows = t.owned_rows()
            t.avg_map().to_host(api.DeviceMap(back))
            whole[rows] = back.data.reshape(-1, row)[rows]
        crc = zlib.crc32(whole.view(np.uint8))
        # registration sums of the slabs add up (host-side sum stands in for the exchange)
        cloud = s.frame(3, prior_pose=s.pose(2))["points_prior"]
        regs = [api.RegistrationCuda(t) for t in hs]
        sums = np.zeros(29, np.int64)
        import ctypes as C
        I16 = fp.colmajor16(np.eye(4, dtype=np.float32))
        for rg in regs:
            rg.prepare_registration(cloud)
            hd = rg._hd
            hd.check(hd.L.ws_reg_begin(hd.h, I16.ctypes.data_as(C.POINTER(C.c_float))))
            hd.check(hd.L.ws_reg_accumulate(hd.h, RES))
            sums += rg.sums_get()
        for t in hs:
            t.close()
        return counters, crc, sums

    c1, crc1, sums1 = run(1)
    c1b, crc1b, sums1b = run(1)
    assert crc1 == crc1b and np.array_equal(sums1, sums1b), "two identical runs differ"
    assert [c[0]["n_candidates"] for c in c1] == [c[0]["n_candidates"] for c in c1b]
    c2, crc2, sums2 = run(2)
    assert crc2 == crc1, "the map depends on the number of slabs"
    assert np.array_equal(sums2, sums1), "registration sums depend on the number of slabs"
    c3, crc3, sums3 = run(3, stripe=6)                      # round-robin stripes of 6 brick columns
    assert crc3 == crc1 and np.array_equal(sums3, sums1), "the result depends on the partition (stripes)"
    for k in range(3):
        one = c1[k][0]
        assert one["n_written"] <= one["n_touched"] <= one["n_candidates"]
        assert sum(c["n_touched"] for c in c2[k]) >= one["n_touched"]          # halo columns are held twice
        assert max(c["n_candidates"] for c in c2[k]) < one["n_candidates"]      # each rank marches less
    assert sums1[28] > 60000                                                    # most points found a seen voxel
