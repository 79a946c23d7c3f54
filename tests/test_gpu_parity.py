"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): integer results bit-exact -- voxel indices, stored entries,
H/g/err/count; recovered pose within 1e-4 (relative to the matrix entry scale, stated per test).
"""
import math

import numpy as np
import pytest

from helpers import assert_same_grid, device_grid, make_pair, random_cloud
from oracle import oracle as orc
from warpsense_b200 import api, fixedpoint as fp, lib
from warpsense_b200.synth import ScanStream, frame_pose

pytestmark = pytest.mark.gpu

MR = fp.MATRIX_RESOLUTION
WR = fp.WEIGHT_RESOLUTION


# ----------------------------------------------------------------------------- golden vectors
def test_g1_tsdf_write_device():
    """test/map.cpp:9-90 == test/cuda.cpp:268-347 on the device map."""
    tau, res, mw = 3000, 1000, 10 * WR
    om, hm, tsdf = make_pair((20, 20, 20), tau, mw, res)
    pos, up = fp.convert_pose_to_gpu(np.eye(4, dtype=np.float32), res)
    assert list(up) == [0, 0, MR]
    tsdf.update_tsdf(np.array([[5500, 500, 500]], np.int32), pos, up)
    eps = tau // 10
    for k, v in {1: tau, 2: tau, 3: 2000, 4: 1000, 5: 0, 6: -1000, 7: -2000}.items():
        assert tsdf.voxel(k, 0, 0) == (v, orc.calc_weight(v, tau, eps))
    assert tsdf.voxel(8, 0, 0) == (3000, 0)
    c = tsdf.counters()
    assert c["n_candidates"] == 8 and c["n_touched"] == 8 and c["n_written"] == 8
    orc.update_tsdf(om, [[5500, 500, 500]], pos, up, tau, mw, res)
    assert_same_grid(om, tsdf, hm, "G1")
    with pytest.raises(lib.WarpsenseError):
        tsdf.voxel(22, 0, 0)


@pytest.mark.parametrize("n", [1, 1024, 2000])
def test_g3_reduction(n):
    """test/cuda.cpp:416-532: n copies of J=(1..6) with value 1."""
    _, _, tsdf = make_pair((9, 9, 9), 600, 640, 64)
    reg = api.RegistrationCuda(tsdf)
    J = np.tile(np.arange(1, 7, dtype=np.int64), (n, 1))
    H, g, e, c = reg.test_reduce(J, np.ones(n, np.int32))
    assert (H == n * np.outer(J[0], J[0])).all()
    assert (g == n * J[0]).all()
    assert e == n and c == n


@pytest.mark.parametrize("J", [[0, 1, 2, 3, 4, 5], [-1, 1, 2, 3, 4, -5], [0, 1, -2, 12, 4, 5], [0, -1, 20, 3, -4, 5]])
def test_jacobi_2_h_device(J):
    """test/cuda.cpp:837-923"""
    _, _, tsdf = make_pair((9, 9, 9), 600, 640, 64)
    reg = api.RegistrationCuda(tsdf)
    H, g, e, c = reg.test_reduce([J], [-7])
    Jv = np.array(J, np.int64)
    assert (H == np.outer(Jv, Jv)).all() and (g == -7 * Jv).all() and e == 7 and c == 1


def test_g4_transform_point_device():
    """test/cuda.cpp:760-827 through the in-place cloud transform of register_cloud (0 iterations)."""
    _, _, tsdf = make_pair((9, 9, 9), 600, 640, 64)
    reg = api.RegistrationCuda(tsdf)
    for theta, expect in ((math.pi / 2, [0, 1, 0]), (-math.pi / 2, [0, -1, 0])):
        T = np.eye(4, dtype=np.float32)
        T[0, 0] = math.cos(theta); T[0, 1] = -math.sin(theta)
        T[1, 0] = math.sin(theta); T[1, 1] = math.cos(theta)
        cloud = np.array([[1, 0, 0]], np.int32)
        out, it = reg.register_cloud(cloud, T, 0, 0.1, 0.0, 64)
        assert it == 0 and list(cloud[0]) == expect
        assert np.array_equal(out, T)
    rng = np.random.default_rng(5)
    cloud = random_cloud(rng, 5000, -30000, 30000)
    T = frame_pose(37)
    want = orc.transform_points(cloud, orc.to_int_mat(T))
    reg.register_cloud(cloud, T, 0, 0.1, 0.0, 64)
    assert np.array_equal(cloud, want)


# ----------------------------------------------------------------------------- map residency
def test_upload_download_roundtrip_with_ring_offsets():
    rng = np.random.default_rng(11)
    hm = api.HostLocalMap(12, 21, 9, 600, 0)
    hm.data[:] = rng.integers(0, 2 ** 32, size=hm.data.shape, dtype=np.uint64).astype(np.uint32)
    hm.pos[:] = [5, -3, 2]
    hm.offset[:] = [7, 20, 0]
    tsdf = api.TSDFCuda(api.DeviceMap(hm), 600, 640, 64)
    back = api.HostLocalMap(12, 21, 9, 600, 0)
    tsdf.avg_map().to_host(api.DeviceMap(back))
    assert np.array_equal(back.data, hm.data)
    assert list(back.pos) == [5, -3, 2] and list(back.offset) == [7, 20, 0]
    for _ in range(50):
        x, y, z = (int(hm.pos[a] + rng.integers(-(hm.size[a] // 2), hm.size[a] // 2 + 1)) for a in range(3))
        assert tsdf.voxel(x, y, z) == hm.value(x, y, z)
    tsdf.set_voxel(5, -3, 2, -123, -64)
    assert tsdf.voxel(5, -3, 2) == (-123, -64)


# ----------------------------------------------------------------------------- update_tsdf parity
CASES = [
    # size, res, tau, n, spread(mm), scanner voxel
    ((33, 33, 33), 64, 600, 400, 900, (0, 0, 0)),
    ((41, 41, 21), 50, 1000, 600, 950, (3, -2, 1)),
    ((65, 65, 33), 100, 1000, 800, 3000, (-5, 4, 0)),
    ((129, 129, 65), 20, 300, 500, 1200, (0, 0, 0)),      # fine voxels: long fans of interpolated candidates
    ((65, 65, 65), 33, 400, 500, 1000, (2, 2, -3)),        # odd resolution
    ((21, 21, 21), 1000, 3000, 300, 9000, (0, 0, 0)),
]


@pytest.mark.parametrize("size,res,tau,n,spread,spos", CASES)
def test_update_tsdf_random_clouds(size, res, tau, n, spread, spos):
    rng = np.random.default_rng(hash((size, res, n)) & 0xFFFF)
    mw = 10 * WR
    om, hm, tsdf = make_pair(size, tau, mw, res)
    up = np.array([0, 0, MR], np.int32)
    centre = np.array(spos, np.int64) * res
    for scan in range(3):
        pts = (centre + rng.integers(-spread, spread, size=(n, 3))).astype(np.int32)
        pts[0] = centre                       # distance 0 -> skipped (update_tsdf.cpp:424)
        pts[1] = centre + 10 * spread         # own cell out of bounds -> skipped (:430)
        pts[2] = pts[3]                       # duplicate ray
        st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
        tsdf.update_tsdf(pts, spos, up)
        c = tsdf.counters()
        assert c["n_candidates"] == st["n_candidates"], "candidate count (scan %d)" % scan
        assert c["n_touched"] == st["n_touched"], "touched voxels (scan %d)" % scan
        assert c["n_written"] == st["n_written"], "written voxels (scan %d)" % scan
        assert_same_grid(om, tsdf, hm, "scan %d" % scan)


@pytest.mark.parametrize("case", [0, 2, 3])
def test_update_tsdf_general_arithmetic_path(case, monkeypatch):
    """WS_MARCH_GENERAL=1 switches the march's 32-bit fast path (march_math.cuh) off: the literal wrapping
    arithmetic must give the same grids as the oracle (and therefore as the fast path)."""
    monkeypatch.setenv("WS_MARCH_GENERAL", "1")
    test_update_tsdf_random_clouds(*CASES[case])


def test_update_tsdf_far_from_origin_and_long_rays():
    """Rays the fast path must refuse: a map 3,000 km from the origin (|coordinate| * res >= 2^31) and rays
    long enough for d * len to wrap int32, which the reference's arithmetic does too (update_tsdf.cpp:452)."""
    rng = np.random.default_rng(21)
    up = np.array([0, 0, MR], np.int32)
    # (a) far from the origin, coarse voxels
    tau, res, mw = 3000, 1000, 640
    om, hm, tsdf = make_pair((21, 21, 21), tau, mw, res)
    spos = [3000, -2999, 2500]
    om.set_state(spos, [10, 10, 10])
    hm.pos[:] = spos
    tsdf.avg_map().update_params(api.DeviceMap(hm))
    centre = np.array(spos, np.int64) * res
    for scan in range(2):
        pts = (centre + rng.integers(-9000, 9000, size=(300, 3))).astype(np.int32)
        st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
        tsdf.update_tsdf(pts, spos, up)
        c = tsdf.counters()
        assert (c["n_candidates"], c["n_touched"]) == (st["n_candidates"], st["n_touched"])
        assert_same_grid(om, tsdf, hm, "far map, scan %d" % scan)
    tsdf.close()
    # (b) 45 m rays through a long thin map: d * len exceeds 2^31 near the far end (d^2 still fits int32)
    tau, res, mw = 1000, 100, 640
    om, hm, tsdf = make_pair((513, 17, 17), tau, mw, res)
    spos = (-230, 0, 0)
    pts = np.stack([rng.integers(22000, 23000, 200), rng.integers(-700, 700, 200), rng.integers(-700, 700, 200)], 1).astype(np.int32)
    st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
    tsdf.update_tsdf(pts, spos, up)
    c = tsdf.counters()
    assert (c["n_candidates"], c["n_touched"]) == (st["n_candidates"], st["n_touched"])
    assert_same_grid(om, tsdf, hm, "long rays")


def test_update_tsdf_tilted_up_vector_and_shifted_ring():
    """Non-trivial up vector (rolled sensor) and a map whose ring origin is not at the array origin."""
    rng = np.random.default_rng(77)
    tau, res, mw = 600, 40, 5 * WR
    om, hm, tsdf = make_pair((81, 81, 41), tau, mw, res)
    om.set_state([10, -7, 3], [5, 60, 17])
    hm.pos[:] = [10, -7, 3]
    hm.offset[:] = [5, 60, 17]
    tsdf.avg_map().update_params(api.DeviceMap(hm))
    a = math.radians(25.0)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = np.array([[1, 0, 0], [0, math.cos(a), -math.sin(a)], [0, math.sin(a), math.cos(a)]], np.float32)
    pose[:3, 3] = [10 * res + 7, -7 * res + 3, 3 * res + 11]
    pos, up = fp.convert_pose_to_gpu(pose, res)
    opos, oup = orc.convert_pose(pose, res)
    assert (pos == opos).all() and (up == oup).all()
    for scan in range(2):
        pts = (np.array(pos, np.int64) * res + rng.integers(-1500, 1500, size=(700, 3))).astype(np.int32)
        orc.update_tsdf(om, pts, pos, up, tau, mw, res)
        tsdf.update_tsdf(pts, pos, up)
        assert_same_grid(om, tsdf, hm, "tilted scan %d" % scan)


def test_update_tsdf_interpolated_chains_replay_rounds():
    """Far, dense, repeated rays at a fine resolution: many voxels whose winner is an interpolated
    candidate below tau, which forces the replay rounds (DESIGN.md 'Collision rule')."""
    rng = np.random.default_rng(3)
    tau, res, mw = 400, 10, 10 * WR
    om, hm, tsdf = make_pair((257, 65, 65), tau, mw, res)
    up = np.array([0, 0, MR], np.int32)
    base = np.stack([rng.integers(900, 1250, 3000), rng.integers(-250, 250, 3000), rng.integers(-250, 250, 3000)], 1)
    pts = np.concatenate([base, base[::-1], base[::3]]).astype(np.int32)
    spos = (-120, 0, 0)
    st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
    assert st["n_neg_candidates"] > 0
    tsdf.update_tsdf(pts, spos, up)
    c = tsdf.counters()
    assert c["n_candidates"] == st["n_candidates"] and c["n_touched"] == st["n_touched"]
    assert c["n_parked"] > 0 and c["n_rounds"] >= 1, "the case is meant to exercise the replay rounds: %r" % c
    assert_same_grid(om, tsdf, hm, "replay rounds")


def test_update_tsdf_record_overflow_regenerates(monkeypatch):
    """A candidate record far too small for the scan: the library must grow it, regenerate it from the
    far part of the rays and still settle every parked voxel exactly."""
    monkeypatch.setenv("WS_RECORD_CAP", "64")
    rng = np.random.default_rng(4)
    tau, res, mw = 400, 10, 10 * WR
    om, hm, tsdf = make_pair((257, 65, 65), tau, mw, res)
    monkeypatch.delenv("WS_RECORD_CAP")
    up = np.array([0, 0, MR], np.int32)
    spos = (-120, 0, 0)
    for scan in range(2):
        base = np.stack([rng.integers(900, 1250, 2500), rng.integers(-250, 250, 2500), rng.integers(-250, 250, 2500)], 1)
        pts = np.concatenate([base, base[::-1]]).astype(np.int32)
        st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
        tsdf.update_tsdf(pts, spos, up)
        c = tsdf.counters()
        assert c["n_candidates"] == st["n_candidates"] and c["n_touched"] == st["n_touched"]
        assert c["n_parked"] > 0
        assert_same_grid(om, tsdf, hm, "record overflow, scan %d" % scan)


def test_update_tsdf_64bit_address_path(monkeypatch):
    """Maps beyond 2^32 voxels take the kernels' 64-bit address path (x table holds address / 64, record chunks carry
    the addresses' high words).  WS_FORCE_WIDE puts a small map on that path: replay rounds, general rays and two
    scans, against the oracle."""
    monkeypatch.setenv("WS_FORCE_WIDE", "1")
    rng = np.random.default_rng(14)
    tau, res, mw = 400, 10, 10 * WR
    om, hm, tsdf = make_pair((257, 65, 65), tau, mw, res)
    monkeypatch.delenv("WS_FORCE_WIDE")
    up = np.array([0, 0, MR], np.int32)
    spos = (-120, 0, 0)
    for scan in range(2):
        base = np.stack([rng.integers(900, 1250, 2500), rng.integers(-250, 250, 2500), rng.integers(-250, 250, 2500)], 1)
        pts = np.concatenate([base, base[::-1]]).astype(np.int32)
        st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
        tsdf.update_tsdf(pts, spos, up)
        c = tsdf.counters()
        assert c["n_candidates"] == st["n_candidates"] and c["n_touched"] == st["n_touched"]
        assert c["n_parked"] > 0
        assert_same_grid(om, tsdf, hm, "64-bit address path, scan %d" % scan)


def test_update_tsdf_empty_and_over_capacity():
    tau, res, mw = 600, 64, 640
    om, hm, tsdf = make_pair((17, 17, 17), tau, mw, res)
    tsdf.update_tsdf(np.zeros((0, 3), np.int32), (0, 0, 0), (0, 0, MR))
    assert tsdf.counters()["n_candidates"] == 0
    assert_same_grid(om, tsdf, hm, "empty scan")
    hd = tsdf.device_map()
    import ctypes as C
    z = np.zeros(3, np.int32)
    rc = hd.L.ws_update_tsdf(hd.h, None, lib.WS_MAX_POINTS + 1, z.ctypes.data_as(C.POINTER(C.c_int32)),
                             z.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == lib.WS_ERR_CAPACITY                      # reference: prints and returns (update_tsdf.cu:146-150)
    assert_same_grid(om, tsdf, hm, "after rejected scan")


def _stream_setup(H, W, side, res, tau=1000, mw=640):
    s = ScanStream(H, W, side, res)
    size = (side, side, side)
    om, hm, tsdf = make_pair(size, tau, mw, res)
    return s, om, hm, tsdf, tau, mw


def test_update_tsdf_synthetic_stream_small():
    """A few frames of the benchmark's scan generator at a reduced size, map accumulated over frames."""
    res = 100
    s, om, hm, tsdf, tau, mw = _stream_setup(32, 256, 128, res)
    for k in range(4):
        f = s.frame(k)
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        st = orc.update_tsdf(om, f["points_map"], pos, up, tau, mw, res)
        tsdf.update_tsdf(f["points_map"], pos, up)
        c = tsdf.counters()
        assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"])
        assert_same_grid(om, tsdf, hm, "frame %d" % k)


# ----------------------------------------------------------------------------- registration parity
def _registration_scene(res=100, side=128, H=32, W=256, frames=2):
    s, om, hm, tsdf, tau, mw = _stream_setup(H, W, side, res)
    for k in range(frames):
        f = s.frame(k)
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        orc.update_tsdf(om, f["points_map"], pos, up, tau, mw, res)
        tsdf.update_tsdf(f["points_map"], pos, up)
    nxt = s.frame(frames, prior_pose=s.pose(frames - 1))
    return s, om, hm, tsdf, nxt["points_prior"].copy(), res


def test_reg_step_sums_bit_exact():
    s, om, hm, tsdf, cloud, res = _registration_scene()
    reg = api.RegistrationCuda(tsdf)
    reg.prepare_registration(cloud)
    for T in (np.eye(4, dtype=np.float32), s.pose(1) @ np.linalg.inv(s.pose(0)).astype(np.float32)):
        T = np.asarray(T, np.float32)
        H, g, e, c = reg.perform_registration(T, res)
        oH, og, oe, oc = orc.reg_step(om, cloud, T, res)
        assert c == oc and e == oe and c > 1000
        assert np.array_equal(H, oH) and np.array_equal(g, og)
        assert np.array_equal(H, H.T)


@pytest.mark.parametrize("host_solve", [True, False])
def test_register_cloud_matches_oracle(host_solve):
    s, om, hm, tsdf, cloud, res = _registration_scene()
    reg = api.RegistrationCuda(tsdf)
    ocloud = cloud.copy()
    oT, oit, otrace = orc.register_cloud(om, ocloud, np.eye(4, dtype=np.float32), 20, 0.1, 0.0, res, trace=True)
    T, it = reg.register_cloud(cloud, np.eye(4, dtype=np.float32), 20, 0.1, 0.0, res, host_solve=host_solve)
    assert it == oit == 20
    trace = reg.trace()
    assert trace.shape == otrace.shape
    # iteration 0 starts from the same transform: its sums are exact whatever the solver does afterwards
    assert np.array_equal(trace[0], otrace[0])
    # recovered pose: 1e-4 on the rotation entries, 1e-4 relative (to the metre scale, 1000 mm) on translation
    assert np.abs(T[:3, :3] - oT[:3, :3]).max() <= 1e-4
    assert np.abs(T[:3, 3] - oT[:3, 3]).max() <= 1e-4 * 1000.0
    # the sums of all later iterations and the transformed cloud agree exactly when every intermediate
    # fixed-point matrix agrees, which holds here for both solvers
    assert np.array_equal(trace, otrace)
    assert np.array_equal(cloud, ocloud)
    assert np.array_equal(T, oT) or np.abs(T - oT).max() < 1e-5


def test_register_cloud_epsilon_stops_early_like_oracle():
    s, om, hm, tsdf, cloud, res = _registration_scene()
    reg = api.RegistrationCuda(tsdf)
    ocloud = cloud.copy()
    oT, oit = orc.register_cloud(om, ocloud, np.eye(4, dtype=np.float32), 200, 0.1, 0.03, res)
    T, it = reg.register_cloud(cloud, np.eye(4, dtype=np.float32), 200, 0.1, 0.03, res)
    assert it == oit and it < 200
    assert np.abs(T - oT).max() <= 1e-4 * 1000.0
    assert np.abs(T[:3, :3] - oT[:3, :3]).max() <= 1e-4


def test_register_then_update_on_device_cloud():
    """The chained flow of the benchmark: register, keep the transformed cloud on the device, update."""
    s, om, hm, tsdf, cloud, res = _registration_scene()
    reg = api.RegistrationCuda(tsdf)
    tau, mw = 1000, 640
    ocloud = cloud.copy()
    oT, _ = orc.register_cloud(om, ocloud, np.eye(4, dtype=np.float32), 10, 0.1, 0.0, res)
    reg.prepare_registration(cloud)
    T, it = reg.register_cloud(None, np.eye(4, dtype=np.float32), 10, 0.1, 0.0, res, keep_on_device=True)
    pose = (oT @ s.pose(1)).astype(np.float32)
    pos, up = fp.convert_pose_to_gpu(pose, res)
    orc.update_tsdf(om, ocloud, pos, up, tau, mw, res)
    ptr, n = reg.points_device()
    assert n == len(cloud)
    tsdf.update_tsdf_device(ptr, n, pos, up)
    assert_same_grid(om, tsdf, hm, "register -> update chain")


# ----------------------------------------------------------------------------- shift (G2 + random)
def test_g2_shift_on_device():
    """test/map.cpp:240-365 with the map resident on the device."""
    DV, DW = 4, 6
    hm = api.HostLocalMap(5, 5, 5, DV, DW)
    tsdf = api.TSDFCuda(api.DeviceMap(hm), DV, 640, 64)
    tsdf.device_map().check(tsdf.device_map().L.ws_map_fill(tsdf.device_map().h, DV, DW))
    om = orc.LocalMap(5, 5, 5, DV, DW)
    entries = {(-2, 2, 0): (0, 0), (-1, 2, 0): (1, 1), (-2, 1, 0): (2, 1), (-1, 1, 0): (3, 2), (-2, 0, 0): (4, 3), (-1, 0, 0): (5, 5)}
    for p, e in entries.items():
        tsdf.set_voxel(*p, *e)
        om.set_value(*p, *e)

    def both_shift(p):
        tsdf.shift(p)
        om.shift(p)
        s, o, q = tsdf.avg_map().params()
        assert list(q) == list(om.pos) and list(o) == list(om.offset)
        assert_same_grid(om, tsdf, hm, "G2 shift to %r" % (p,))

    for x in (5, 10, 15, 20, 24):
        both_shift([x, 0, 0])
    s, o, q = tsdf.avg_map().params()
    assert list(q) == [24, 0, 0] and list(o) == [26 % 5, 2, 2]
    tsdf.set_voxel(24, 0, 0, 24, 0); om.set_value(24, 0, 0, 24, 0)
    both_shift([24, 5, 0]); tsdf.set_voxel(24, 5, 0, 24, 5); om.set_value(24, 5, 0, 24, 5)
    both_shift([19, 5, 0]); tsdf.set_voxel(19, 5, 0, 19, 5); om.set_value(19, 5, 0, 19, 5)
    both_shift([19, 0, 0]); tsdf.set_voxel(19, 0, 0, 19, 0); om.set_value(19, 0, 0, 19, 0)
    both_shift([24, 0, 0]); assert tsdf.voxel(24, 0, 0) == (24, 0)
    both_shift([19, 0, 0]); assert tsdf.voxel(19, 0, 0) == (19, 0)
    both_shift([24, 5, 0]); assert tsdf.voxel(24, 5, 0) == (24, 5)
    both_shift([19, 5, 0]); assert tsdf.voxel(19, 5, 0) == (19, 5)
    both_shift([24, 0, 0]); assert tsdf.voxel(24, 0, 0) == (24, 0)
    for x in (19, 14, 9, 4, 0):
        both_shift([x, 0, 0])
    assert tsdf.voxel(-1, 2, 0) == (1, 1)
    for p, e in entries.items():
        assert tsdf.voxel(*p) == e
    assert_same_grid(om, tsdf, hm, "G2 final")


def test_shift_random_walk_matches_oracle():
    rng = np.random.default_rng(21)
    tau, res, mw = 600, 64, 640
    om, hm, tsdf = make_pair((33, 25, 17), tau, mw, res)
    up = np.array([0, 0, MR], np.int32)
    pos = np.zeros(3, np.int64)
    for step in range(6):
        pts = (pos * res + rng.integers(-700, 700, size=(300, 3))).astype(np.int32)
        orc.update_tsdf(om, pts, pos, up, tau, mw, res)
        tsdf.update_tsdf(pts, pos, up)
        pos = pos + rng.integers(-9, 10, size=3)
        om.shift(pos)
        tsdf.shift(pos)
        s, o, q = tsdf.avg_map().params()
        assert list(q) == list(om.pos) and list(o) == list(om.offset)
        assert_same_grid(om, tsdf, hm, "after shift %d" % step)
    om.write_back()
    tsdf.write_back()
    assert sorted(tsdf.chunk_list()) == sorted(om.chunk_list())
    for c in om.chunk_list():
        assert np.array_equal(tsdf.chunk(*c), om.chunk(*c)), "chunk %r" % (c,)


def test_shift_walk_with_bounded_chunk_store_and_file_round_trip(tmp_path):
    """hdf5_global_map.cpp:59-137: only a few chunks stay in memory (here 3), the least recently used one is spilled
    and read back on demand -- a long walk must still agree with the oracle's unbounded store; then the file the
    writer leaves is read back by ws_import_hdf5 into a fresh handle and reloaded into its local map."""
    rng = np.random.default_rng(33)
    tau, res, mw = 600, 64, 640
    om, hm, tsdf = make_pair((33, 25, 17), tau, mw, res)
    tsdf.store_configure(3)
    up = np.array([0, 0, MR], np.int32)
    pos = np.zeros(3, np.int64)
    for step in range(10):
        pts = (pos * res + rng.integers(-700, 700, size=(300, 3))).astype(np.int32)
        orc.update_tsdf(om, pts, pos, up, tau, mw, res)
        tsdf.update_tsdf(pts, pos, up)
        pos = pos + rng.integers(-16, 17, size=3)              # crosses 64-voxel chunk borders often
        om.shift(pos)
        tsdf.shift(pos)
        assert_same_grid(om, tsdf, hm, "after shift %d" % step)
    # walk back to the start: everything that was stored must come back
    while np.abs(pos).max() > 0:
        pos = pos - np.clip(pos, -12, 12)                      # a shift may not exceed the map size (:63-65)
        om.shift(pos)
        tsdf.shift(pos)
        assert_same_grid(om, tsdf, hm, "walking back to %r" % (list(pos),))
    assert tsdf.store_evictions() > 0, "the walk was meant to overflow the 3-chunk store"
    om.write_back()
    tsdf.write_back()
    assert sorted(tsdf.chunk_list()) == sorted(om.chunk_list())
    for c in om.chunk_list():
        assert np.array_equal(tsdf.chunk(*c), om.chunk(*c)), "chunk %r" % (c,)
    # file round trip
    path = str(tmp_path / "walk.h5")
    poses = np.round(rng.normal(size=(5, 7)).astype(np.float32) * 1000) / 1000
    tsdf.export_hdf5(path, tau, (33, 25, 17), 0.6, res, mw, poses)
    hm2 = api.HostLocalMap(33, 25, 17, tau, 0)
    fresh = api.TSDFCuda(api.DeviceMap(hm2), tau, mw, res)
    fresh.store_configure(2)
    meta, poses_back, n_chunks = fresh.import_hdf5(path)
    assert n_chunks == len(om.chunk_list()) and np.array_equal(poses_back, poses)
    assert meta == {"tau": tau, "map_size": (33, 25, 17), "max_distance": np.float32(0.6), "map_resolution": res, "max_weight": mw}
    for c in om.chunk_list():
        assert np.array_equal(fresh.chunk(*c), om.chunk(*c)), "imported chunk %r" % (c,)
    fresh.reload()                                              # the window around pos = 0 out of the imported chunks
    assert_same_grid(om, fresh, hm2, "reloaded from the file")
    fresh.close()
    tsdf.close()


# ----------------------------------------------------------------------------- fused per-scan pipeline
def _compose_f32(X, prior):
    """pose = X * prior in float32, accumulating over k = 0..3 in order (ws_track_scan's definition)."""
    X = np.asarray(X, np.float32); prior = np.asarray(prior, np.float32)
    out = np.zeros((4, 4), np.float32)
    for r in range(4):
        for c in range(4):
            acc = np.float32(X[r, 0] * prior[0, c])
            for k in range(1, 4):
                acc = np.float32(acc + np.float32(X[r, k] * prior[k, c]))
            out[r, c] = acc
    return out


def test_track_scan_matches_register_then_update():
    """ws_track_scan (registration, pose and update chained on the device, one host sync) against the oracle
    doing register_cloud -> pose = X * prior -> convert_pose -> update_tsdf step by step."""
    res = 100
    s, om, hm, tsdf, tau, mw = _stream_setup(32, 256, 128, res)
    reg = api.RegistrationCuda(tsdf)
    f0 = s.frame(0)
    pos, up = fp.convert_pose_to_gpu(f0["pose"], res)
    orc.update_tsdf(om, f0["points_map"], pos, up, tau, mw, res)
    tsdf.update_tsdf(f0["points_map"], pos, up)
    I = np.eye(4, dtype=np.float32)
    for k in range(1, 4):
        prior = s.pose(k - 1)
        cloud = s.frame(k, prior_pose=prior)["points_prior"]
        ocloud = cloud.copy()
        oX, oit = orc.register_cloud(om, ocloud, I, 10, 0.1, 0.0, res)
        opose = _compose_f32(oX, prior)
        opos, oup = orc.convert_pose(opose, res)
        st = orc.update_tsdf(om, ocloud, opos, oup, tau, mw, res)
        X, pose, it = reg.track_scan(cloud, prior, 10, 0.1, 0.0, res)
        assert it == oit
        assert np.array_equal(X, oX), "transform differs from the oracle"
        assert np.array_equal(pose, opose), "pose product differs"
        c = tsdf.counters()
        assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"])
        assert_same_grid(om, tsdf, hm, "frame %d" % k)
    tsdf.close()


def _yaw(deg, tx=0.0):
    T = np.eye(4, dtype=np.float32)
    a = math.radians(deg)
    T[0, 0] = math.cos(a); T[0, 1] = -math.sin(a); T[1, 0] = math.sin(a); T[1, 1] = math.cos(a)
    T[0, 3] = tx
    return T


def test_track_scan_reference_pose_update_and_pretransform():
    """WS_TRACK_REFERENCE_POSE: the pose is composed as App::update_pose_estimate does (app.cpp:172-176:
    R = X.R * R, t += X.t -- no X.R * t term) and register_cloud starts from an IMU-like pretransform
    (app.cpp:97-102).  Oracle: register_cloud -> orc.update_pose_estimate -> convert_pose -> update_tsdf."""
    res = 100
    s, om, hm, tsdf, tau, mw = _stream_setup(32, 256, 128, res)
    reg = api.RegistrationCuda(tsdf)
    f0 = s.frame(0)
    pos, up = fp.convert_pose_to_gpu(f0["pose"], res)
    orc.update_tsdf(om, f0["points_map"], pos, up, tau, mw, res)
    tsdf.update_tsdf(f0["points_map"], pos, up)
    pre = _yaw(0.3)
    for k in range(1, 4):
        prior = s.pose(k - 1)
        assert np.abs(prior[:3, 3]).max() > 0 or k == 1      # prior.t != 0: the two compositions differ
        cloud = s.frame(k, prior_pose=prior)["points_prior"]
        ocloud = cloud.copy()
        oX, oit = orc.register_cloud(om, ocloud, pre, 10, 0.1, 0.0, res)
        opose = orc.update_pose_estimate(prior, oX)
        opos, oup = orc.convert_pose(opose, res)
        st = orc.update_tsdf(om, ocloud, opos, oup, tau, mw, res)
        X, pose, it = reg.track_scan(cloud, prior, 10, 0.1, 0.0, res, pretransform=pre, reference_pose=True)
        assert it == oit and np.array_equal(X, oX)
        assert np.array_equal(pose, opose), "reference pose composition differs"
        if k > 1:
            assert not np.array_equal(pose, _compose_f32(oX, prior)), "the two compositions should differ here"
        c = tsdf.counters()
        assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"])
        assert_same_grid(om, tsdf, hm, "frame %d" % k)
    tsdf.close()


def test_track_submit_wait_two_in_flight_chained_pose():
    """ws_track_submit / ws_track_wait: scan k+1 is enqueued (host copy on the second stream, prior pose chained
    on the device) before scan k is collected; results equal the blocking calls and the oracle."""
    res = 100
    s, om, hm, tsdf, tau, mw = _stream_setup(32, 256, 128, res)
    reg = api.RegistrationCuda(tsdf)
    f0 = s.frame(0)
    pos, up = fp.convert_pose_to_gpu(f0["pose"], res)
    orc.update_tsdf(om, f0["points_map"], pos, up, tau, mw, res)
    tsdf.update_tsdf(f0["points_map"], pos, up)
    I = np.eye(4, dtype=np.float32)
    n_scans = 5
    # oracle: every scan is placed with the ground-truth pose of the scan before; the tracked pose chains
    clouds = [s.frame(k, prior_pose=s.pose(k - 1))["points_prior"] for k in range(1, n_scans + 1)]
    want = []
    opose = s.pose(0)
    for k in range(n_scans):
        ocloud = clouds[k].copy()
        oX, oit = orc.register_cloud(om, ocloud, I, 8, 0.1, 0.0, res)
        opose = _compose_f32(oX, opose)
        opos, oup = orc.convert_pose(opose, res)
        st = orc.update_tsdf(om, ocloud, opos, oup, tau, mw, res)
        want.append((oX, opose.copy(), oit, st))
    tickets = []
    got = []
    for k in range(n_scans):
        tickets.append(reg.track_submit(clouds[k], s.pose(0) if k == 0 else None, 8, 0.1, 0.0, res))
        if len(tickets) == 2:
            got.append(reg.track_wait(tickets.pop(0)) + (tsdf.counters(),))
    while tickets:
        got.append(reg.track_wait(tickets.pop(0)) + (tsdf.counters(),))
    for k in range(n_scans):
        X, pose, it, c = got[k]
        oX, op, oit, st = want[k]
        assert it == oit and np.array_equal(X, oX), "scan %d: transform" % k
        assert np.array_equal(pose, op), "scan %d: chained pose" % k
        assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"])
    assert_same_grid(om, tsdf, hm, "after %d pipelined scans" % n_scans)
    with pytest.raises(lib.WarpsenseError):
        reg.track_wait(0)                                   # nothing in flight
    tsdf.close()
