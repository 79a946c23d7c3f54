"""Scan preprocessing: App::preprocess, src/warpsense/app.cpp:118-148.

CPU: the oracle's restatement against a literal numpy/python evaluation of the reference's formulas.
GPU: ws_preprocess_scan against the oracle, bit-exact and in the same (scan) order; then the preprocessed
scan goes through update_tsdf on both sides."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from warpsense_b200 import api, fixedpoint as fp
from warpsense_b200.synth import ScanStream

from helpers import assert_same_grid, make_pair


def _literal(cloud, pose, res):
    """app.cpp:118-148 evaluated point by point in numpy float32 / python ints; returns the SET of points."""
    M = orc.to_int_mat(pose)
    out = set()
    f = np.float32
    for x, y, z in np.asarray(cloud, np.float32)[:, :3]:
        if float(x) < 0.3 and float(y) < 0.3 and float(z) < 0.3:
            continue
        c = []
        for v in (x, y, z):
            mm = f(v) * f(1000.0)
            c.append(int(f(f(np.floor(f(mm / f(res)))) * f(res)) + f(res // 2)))
        p = orc.transform_point(np.array(c, np.int32), M)
        out.add((int(p[0]), int(p[1]), int(p[2])))
    return out


def _pose(yaw_deg, t):
    a = math.radians(yaw_deg)
    P = np.eye(4, dtype=np.float32)
    P[:3, :3] = np.array([[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]], np.float32)
    P[:3, 3] = t
    return P


def _cloud(n, seed, spread=12.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-spread, spread, size=(n, 3)).astype(np.float32)
    c[::7] = c[1::7][: len(c[::7])]          # exact duplicates
    c[5] = [0.1, 0.2, -0.5]                  # dropped: x, y and z all below 0.3 (app.cpp:128)
    c[6] = [0.1, 5.0, 0.2]                   # kept: only two of the three are below 0.3
    c[8] = c[9] + np.float32(0.001)          # same voxel, different return
    return c


def test_oracle_preprocess_matches_literal_formulas():
    for res, seed in ((64, 1), (50, 2), (100, 3)):
        cloud = _cloud(1500, seed)
        pose = _pose(33.0, [1234.5, -777.25, 90.0])
        got = orc.preprocess(cloud, pose, res)
        assert len(got) == len({tuple(p) for p in got.tolist()}), "duplicates survived"
        assert {tuple(p) for p in got.tolist()} == _literal(cloud, pose, res)
        # scan order: first occurrences, in order
        M = orc.to_int_mat(pose)
        seen, want = set(), []
        for x, y, z in cloud:
            if x < 0.3 and y < 0.3 and z < 0.3:
                continue
            s = tuple(_literal(np.array([[x, y, z]]), pose, res))[0]
            if s not in seen:
                seen.add(s); want.append(s)
        assert [tuple(p) for p in got.tolist()] == want


def test_oracle_cpu_node_preprocess_matches_literal_formulas():
    """src/cpu/fastsense.cpp:143-163: key = voxel of the untransformed mm point, value = transform_point(point)."""
    res = 64
    cloud = _cloud(1200, 9)
    pose = _pose(-20.0, [500.0, 20.5, -77.0])
    M = orc.to_int_mat(pose)
    seen, want = set(), []
    for x, y, z in cloud:
        if x < 0.3 and y < 0.3 and z < 0.3:
            continue
        pt = [int(np.float32(v) * np.float32(1000)) for v in (x, y, z)]
        key = tuple(int(c / res) for c in pt)                  # C++ integer division truncates toward zero
        if key in seen:
            continue
        seen.add(key)
        want.append(tuple(int(v) for v in orc.transform_point(np.array(pt, np.int32), M)))
    got = orc.preprocess(cloud, pose, res, cpu_node=True)
    assert [tuple(p) for p in got.tolist()] == want


def test_oracle_preprocess_strided_and_empty():
    cloud = _cloud(300, 4)
    wide = np.zeros((300, 6), np.float32)      # PointCloud2 with extra fields (point_step 24)
    wide[:, :3] = cloud
    wide[:, 3:] = 7.0
    pose = _pose(-10.0, [0, 0, 0])
    assert np.array_equal(orc.preprocess(wide, pose, 64), orc.preprocess(cloud, pose, 64))
    assert len(orc.preprocess(np.zeros((0, 3), np.float32), pose, 64)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n,res", [(1, 64), (1000, 64), (40000, 50), (131072, 100)])
def test_preprocess_scan_device_matches_oracle(n, res):
    om, hm, tsdf = make_pair((65, 65, 65), 600, 640, res)
    cloud = _cloud(max(n, 10), n)[:n]
    pose = _pose(12.5, [300.0, -120.0, 40.0])
    want = orc.preprocess(cloud, pose, res)
    got, m = tsdf.preprocess_scan(cloud, pose, res)
    assert m == len(want)
    assert np.array_equal(got, want), "device preprocessing differs from the oracle (values or order)"
    wide = np.zeros((len(cloud), 8), np.float32)
    wide[:, :3] = cloud
    got2, _ = tsdf.preprocess_scan(wide, pose, res)
    assert np.array_equal(got2, want)
    got3, m3 = tsdf.preprocess_scan(np.zeros((0, 3), np.float32), pose, res)
    assert m3 == 0
    want4 = orc.preprocess(cloud, pose, res, cpu_node=True)          # the CPU node's variant, its own order
    got4, m4 = tsdf.preprocess_scan(cloud, pose, res, cpu_node=True)
    assert m4 == len(want4) and np.array_equal(got4, want4)
    # NaN / inf returns (is_dense == false) are dropped, on the device as in the oracle, in both variants
    bad = cloud.copy()
    bad[::5, 0] = np.nan
    bad[1::7, 2] = np.inf
    bad[2::11, 1] = -np.inf
    for cpu_node in (False, True):
        want5 = orc.preprocess(bad, pose, res, cpu_node=cpu_node)
        keep = np.isfinite(bad).all(axis=1)
        assert np.array_equal(want5, orc.preprocess(np.ascontiguousarray(bad[keep]), pose, res, cpu_node=cpu_node))
        got5, m5 = tsdf.preprocess_scan(bad, pose, res, cpu_node=cpu_node)
        assert m5 == len(want5) and np.array_equal(got5, want5)
    tsdf.close()


@pytest.mark.gpu
def test_preprocess_then_update_on_device_matches_oracle():
    """The reference's per-scan flow (app.cpp:65-112): preprocess -> update_tsdf with the same points."""
    res, tau, mw, side = 100, 1000, 640, 96
    s = ScanStream(32, 256, side, res)
    om, hm, tsdf = make_pair((side + 1,) * 3, tau, mw, res)
    for k in range(3):
        f = s.frame(k)
        pts = orc.preprocess(f["points_sensor_m"], f["pose"], res)
        _, m = tsdf.preprocess_scan(f["points_sensor_m"], f["pose"], res, fetch=False)
        assert m == len(pts) and m < len(f["points_sensor_m"])
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        st = orc.update_tsdf(om, pts, pos, up, tau, mw, res)
        ptr, n = tsdf.scan_points_device()
        tsdf.update_tsdf_device(ptr, n, pos, up)
        c = tsdf.counters()
        assert (c["n_candidates"], c["n_touched"]) == (st["n_candidates"], st["n_touched"])
        assert_same_grid(om, tsdf, hm, "frame %d" % k)
    tsdf.close()
