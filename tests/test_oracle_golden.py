"""The oracle against every known-answer test the reference holds for the hot path.

G1 test/map.cpp:9-90 (= test/cuda.cpp:268-347)   one-point update_tsdf, 8 voxels
G2 test/map.cpp:240-365                          ring-buffer shift across chunk borders
G3 test/cuda.cpp:416-532                         H/g/e reduction of k x (1..6)
G4 test/cuda.cpp:760-827                         fixed-point +-90 degree transform
G5 test/params_a.cpp:26-41                       parameter scaling
J*J^T test/cuda.cpp:837-923
(all paths relative to /root/reference)
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

WR = orc.WEIGHT_RESOLUTION
MR = orc.MATRIX_RESOLUTION


def g1_setup():
    max_distance = 3
    res = 1000
    tau = int(max_distance * 1000)
    max_weight = 10 * WR
    size = int(20 * 1000 / res)
    m = orc.LocalMap(size, size, size, tau, 0)
    pose = np.eye(4, dtype=np.float32)
    # test/map.cpp:38-48
    pos = np.array([int(math.floor(pose[i, 3])) * 1000 // res for i in range(3)], np.int32)
    rot = np.eye(4, dtype=np.int32)
    rot[:3, :3] = orc.to_int_mat(pose)[:3, :3]
    up = orc.transform_point([0, 0, MR], rot)
    return m, pos, up, tau, max_weight, res


def test_g1_tsdf_write():
    m, pos, up, tau, max_weight, res = g1_setup()
    assert list(m.size) == [21, 21, 21] and list(m.offset) == [10, 10, 10]
    assert list(up) == [0, 0, MR]
    stats = orc.update_tsdf(m, [[5500, 500, 500]], pos, up, tau, max_weight, res)
    eps = tau // 10
    expect = {1: tau, 2: tau, 3: 2000, 4: 1000, 5: 0, 6: -1000, 7: -2000}
    for k, v in expect.items():
        val, w = m.value(k, 0, 0)
        assert val == v
        assert w == orc.calc_weight(val, tau, eps)
    assert m.value(8, 0, 0) == (3000, 0)  # default value, weight == 0
    assert stats["n_candidates"] == 8 and stats["n_touched"] == 8


def test_g1_read_back_through_chunks():
    """test/map.cpp:92-238: the same values after write_back, read from chunk 0_0_0."""
    m, pos, up, tau, max_weight, res = g1_setup()
    orc.update_tsdf(m, [[5500, 500, 500]], pos, up, tau, max_weight, res)
    m.write_back()
    chunk = m.chunk(0, 0, 0)
    assert chunk is not None
    for k, v in {1: tau, 2: tau, 3: 2000, 4: 1000, 5: 0, 6: -1000, 7: -2000}.items():
        raw = int(chunk[k * 64 * 64])
        assert raw == orc.make_entry(v, orc.calc_weight(v, tau, tau // 10))
        assert m.global_value(k, 0, 0) == (v, orc.calc_weight(v, tau, tau // 10))


def test_g2_map_raw_shift():
    DV, DW = 4, 6
    m = orc.LocalMap(5, 5, 5, DV, DW)
    entries = {(-2, 2, 0): (0, 0), (-1, 2, 0): (1, 1), (-2, 1, 0): (2, 1),
               (-1, 1, 0): (3, 2), (-2, 0, 0): (4, 3), (-1, 0, 0): (5, 5)}
    for p, e in entries.items():
        m.set_value(*p, *e)
    assert list(m.pos) == [0, 0, 0] and list(m.size) == [5, 5, 5] and list(m.offset) == [2, 2, 2]
    assert m.in_bounds(0, 2, -2) and not m.in_bounds(22, 0, 0)
    assert m.value(0, 0, 0) == (DV, DW)
    assert m.value(-1, 2, 0) == (1, 1)

    for x in (5, 10, 15, 20, 24):
        m.shift([x, 0, 0])
    assert list(m.pos) == [24, 0, 0] and list(m.offset) == [26 % 5, 2, 2]
    assert not m.in_bounds(0, 2, -2) and m.in_bounds(22, 0, 0)
    assert m.value(24, 0, 0) == (DV, DW)

    m.set_value(24, 0, 0, 24, 0)
    m.shift([24, 5, 0]); m.set_value(24, 5, 0, 24, 5)
    m.shift([19, 5, 0]); m.set_value(19, 5, 0, 19, 5)
    m.shift([19, 0, 0]); m.set_value(19, 0, 0, 19, 0)
    m.shift([24, 0, 0]); assert m.value(24, 0, 0) == (24, 0)
    m.shift([19, 0, 0]); assert m.value(19, 0, 0) == (19, 0)
    m.shift([24, 5, 0]); assert m.value(24, 5, 0) == (24, 5)
    m.shift([19, 5, 0]); assert m.value(19, 5, 0) == (19, 5)
    m.shift([24, 0, 0]); assert m.value(24, 0, 0) == (24, 0)

    for x in (19, 14, 9, 4, 0):
        m.shift([x, 0, 0])
    assert list(m.pos) == [0, 0, 0] and list(m.offset) == [2, 2, 2]
    assert m.in_bounds(0, 2, -2) and not m.in_bounds(22, 0, 0)
    assert m.value(0, 0, 0) == (DV, DW)
    assert m.value(-1, 2, 0) == (1, 1)
    with pytest.raises(IndexError):
        m.value(22, 0, 0)


@pytest.mark.parametrize("n", [1, 1024, 2000])
def test_g3_hge_reduction_shape(n):
    """k copies of J=(1..6), value J[0]=1: H = k*J*J^T, g = k*J, e = k (test/cuda.cpp:416-532)."""
    J = np.arange(1, 7, dtype=np.int64)
    H = np.zeros((6, 6), np.int64)
    g = np.zeros(6, np.int64)
    for _ in range(n):
        H += orc.jacobi_2_h(J)
        g += J * J[0]
    assert (H == n * np.outer(J, J)).all()
    assert (g == n * J).all()


def test_g4_transform_point():
    for theta, expect in ((math.pi / 2, [0, 1, 0]), (-math.pi / 2, [0, -1, 0])):
        T = np.eye(4, dtype=np.float32)
        T[0, 0] = math.cos(theta); T[0, 1] = -math.sin(theta)
        T[1, 0] = math.sin(theta); T[1, 1] = math.cos(theta)
        M = orc.to_int_mat(T)
        assert list(orc.transform_point([1, 0, 0], M)) == expect


def test_g5_param_scaling():
    tau, mw, size = orc.scale_params(3.0, 10, [20, 20, 20], 1000)
    assert tau == 3000 and mw == 10 * WR and list(size) == [20, 20, 20]
    tau, mw, size = orc.scale_params(1.0, 10, [40, 40, 25], 64)  # params/params.yaml
    assert tau == 1000 and mw == 640 and list(size) == [625, 625, 390]
    m = orc.LocalMap(*size, tau, 0)
    assert list(m.size) == [625, 625, 391]


@pytest.mark.parametrize("J", [[0, 1, 2, 3, 4, 5], [-1, 1, 2, 3, 4, -5], [0, 1, -2, 12, 4, 5],
                               [0, -1, 20, 3, -4, 5]])
def test_jacobi_2_h(J):
    Jv = np.array(J, np.int64)
    assert (orc.jacobi_2_h(J) == np.outer(Jv, Jv)).all()


def test_entry_packing():
    # include/map/tsdf.h:16-19 : value in the low half-word, weight in the high one
    assert orc.make_entry(-1, 2) == (0xFFFF | (2 << 16))
    assert orc.make_entry(3000, 0) == 3000
    assert orc.make_entry(5, -64) == (5 | (0xFFC0 << 16))


def test_update_tsdf_omp_variant():
    """src/cpu/update_tsdf.cpp:566-724, the OpenMP overload (timing only): with one thread it is the
    sequential variant; with several it reaches the same voxels with the same candidates, and differs only where
    two threads met on a voxel."""
    rng = np.random.default_rng(17)
    size, res, tau = 65, 64, 600
    d = rng.normal(size=(3000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (d * rng.uniform(400, 1900, (3000, 1))).astype(np.int32)
    pos, up = (0, 0, 0), (0, 0, 1024)
    maps, stats = [], []
    for kind in ("seq", 1, 4):
        m = orc.LocalMap(size, size, size, tau, 0)
        if kind == "seq":
            st = orc.update_tsdf(m, pts, pos, up, tau, 640, res)
        else:
            st = orc.update_tsdf_omp(m, pts, pos, up, tau, 640, res, kind)
        maps.append(np.array(m.data, copy=True))
        stats.append(st)
    assert np.array_equal(maps[0], maps[1]) and stats[0] == stats[1]
    assert stats[2]["n_candidates"] == stats[0]["n_candidates"] and stats[2]["n_touched"] >= stats[0]["n_touched"]
    touched = lambda a: (a >> 16).astype(np.int16) != 0                      # weight != 0
    assert np.array_equal(touched(maps[0]), touched(maps[2]))
    assert (maps[0] != maps[2]).mean() < 0.05
