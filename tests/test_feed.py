"""featsense feed: the TSDF-facing half of Mapping::thread_run, src/featsense/mapping.cpp:39-147."""
import numpy as np
import pytest

from warpsense_b200 import api, fixedpoint as fp
from warpsense_b200.params import MapParams, Params


class _StubMapping:
    """Records what the feed hands to TSDFMapping::update_tsdf_from_ros."""

    def __init__(self, res=100, update_distance=0.5):
        self.params_ = Params(map=MapParams(resolution=res, update_distance=update_distance))
        self.calls = []
        self.shifting = False

    def is_shifting(self):
        return self.shifting

    def subsample(self, cloud, res_m):
        from oracle import oracle as orc                     # CPU stand-in for the device voxel grid
        return orc.voxelgrid(cloud, res_m)[1]

    def update_tsdf_from_ros(self, cloud, pose):
        self.calls.append((np.array(cloud, np.float32), np.array(pose, np.float64)))


def _pose(x):
    P = np.eye(4)
    P[0, 3] = x
    return P


def test_feed_gates_on_distance_and_accumulates_during_shift():
    m = _StubMapping()
    feed = api.MappingFeed(m)
    rng = np.random.default_rng(0)
    clouds = [rng.uniform(-5, 5, (200, 3)).astype(np.float32) for _ in range(6)]
    assert feed.push(clouds[0], _pose(0.0))                   # first cloud initialises (:66-75)
    assert len(m.calls) == 1 and len(m.calls[0][0]) <= 200     # ... subsampled
    assert not feed.push(clouds[1], _pose(0.3))               # moved 0.3 m <= update_distance (:81)
    assert feed.push(clouds[2], _pose(0.6))
    assert len(m.calls) == 2 and np.array_equal(m.calls[1][0], clouds[2])     # used as is (no subsample, :126)
    m.shifting = True
    assert not feed.push(clouds[3], _pose(1.2))               # accumulated (:115-119)
    assert not feed.push(clouds[4], _pose(1.8))
    assert len(m.calls) == 2 and len(feed.accumulated_) == 2
    m.shifting = False
    assert feed.push(clouds[5], _pose(2.4))                   # flush: concatenated + subsampled (:121-126)
    flushed = m.calls[2][0]
    from oracle import oracle as orc
    want = orc.voxelgrid(np.concatenate([clouds[5], clouds[3], clouds[4]]), 0.1)[1]
    assert np.array_equal(flushed, want) and not feed.accumulated_
    assert len(feed.poses) == 4 and feed.updates == 3         # every used pose is recorded (:137)


def test_oracle_voxelgrid_is_pcl_voxelgrid():
    """The restated pcl::VoxelGrid (filters/impl/voxel_grid.hpp): one float centroid per occupied leaf, leaves in
    ascending ijk0 + ijk1 * div0 + ijk2 * div0 * div1 (x fastest), leaf = floor(x * inverse_leaf) - min_b."""
    from oracle import oracle as orc
    pts = np.array([[0.01, 0.01, 0.01], [0.03, 0.05, 0.07], [0.25, 0.01, 0.01], [-0.01, 0.0, 0.0],
                    [0.01, 0.15, 0.01], [np.nan, 0.0, 0.0]], np.float32)
    mm, xyz = orc.voxelgrid(pts, 0.1)
    assert len(xyz) == 4                                        # the NaN point is dropped (!is_dense)
    # leaf order: x fastest -> (-1,0,0), (0,0,0), (2,0,0), then y = 1
    assert np.allclose(xyz[0], [-0.01, 0, 0]) and np.allclose(xyz[1], [0.02, 0.03, 0.04], atol=1e-7)
    assert np.allclose(xyz[2], [0.25, 0.01, 0.01]) and np.allclose(xyz[3], [0.01, 0.15, 0.01])
    assert np.array_equal(mm[1], np.trunc(xyz[1] * np.float32(1000.0)).astype(np.int32))     # tsdf_mapping.cpp:156-157
    # float accumulation in point order, then one division (AccumulatorXYZ)
    rng = np.random.default_rng(3)
    big = rng.uniform(0.0, 0.0999, (50, 3)).astype(np.float32)
    _, c = orc.voxelgrid(big, 0.1)
    acc = np.zeros(3, np.float32)
    for p in big:
        acc = (acc + p).astype(np.float32)
    assert len(c) == 1 and np.array_equal(c[0], (acc / np.float32(50)).astype(np.float32))


def test_mm_pose_from_isometry_matches_oracle():
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    for _ in range(20):
        a, b, c = rng.uniform(-3, 3, 3)
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(c), -np.sin(c)], [0, np.sin(c), np.cos(c)]])
        P = np.eye(4)
        P[:3, :3] = Rz @ Ry @ Rx
        P[:3, 3] = rng.uniform(-20, 20, 3)
        assert np.array_equal(api.mm_pose_from_isometry(P), orc.mm_pose_from_isometry(P))


@pytest.mark.gpu
def test_voxelgrid_on_device_matches_oracle():
    """ws_voxelgrid_subsample against the restated pcl::VoxelGrid: same leaves, same order, bit-identical float
    centroids and millimetre points; host and device input; NaN / inf returns; an extent too large for the leaf."""
    from oracle import oracle as orc
    from warpsense_b200.synth import ScanStream
    hm = api.HostLocalMap(33, 33, 33, 600, 0)
    tsdf = api.TSDFCuda(api.DeviceMap(hm), 600, 640, 50)
    rng = np.random.default_rng(7)
    s = ScanStream(64, 512, 256, 50)
    clouds = [rng.uniform(-4, 4, (20000, 3)).astype(np.float32),
              (s.frame(3)["points_map"].astype(np.float32) / np.float32(1000.0)),
              rng.normal(0, 0.02, (5000, 3)).astype(np.float32),          # many points per leaf
              np.zeros((0, 3), np.float32), np.array([[1.5, -2.5, 0.25]], np.float32)]
    bad = rng.uniform(-2, 2, (3000, 3)).astype(np.float32)
    bad[::7, 1] = np.nan
    bad[5::11, 0] = np.inf
    clouds.append(bad)
    for ci, cl in enumerate(clouds):
        for leaf in (0.05, 0.1, 0.033):
            omm, oxyz = orc.voxelgrid(cl, leaf)
            mm, xyz, m = tsdf.voxelgrid_subsample(cl, leaf, want_xyz=True)
            assert m == len(omm), "cloud %d leaf %g: %d leaves vs %d" % (ci, leaf, m, len(omm))
            assert np.array_equal(xyz, oxyz), "cloud %d leaf %g: centroids" % (ci, leaf)
            assert np.array_equal(mm, omm), "cloud %d leaf %g: millimetre points" % (ci, leaf)
    # strided input (PointXYZI: 16 bytes per point)
    xyzi = np.zeros((4000, 4), np.float32)
    xyzi[:, :3] = rng.uniform(-3, 3, (4000, 3))
    xyzi[:, 3] = 7.0
    mm, m = tsdf.voxelgrid_subsample(xyzi, 0.05)
    omm, _ = orc.voxelgrid(xyzi, 0.05)
    assert np.array_equal(mm, omm)
    # leaf too small for the extent: PCL gives the cloud back unfiltered
    wide = rng.uniform(-3000, 3000, (500, 3)).astype(np.float32)
    omm, _ = orc.voxelgrid(wide, 0.001)
    mm, m = tsdf.voxelgrid_subsample(wide, 0.001)
    assert m == 500 and np.array_equal(mm, omm)
    tsdf.close()


@pytest.mark.gpu
def test_feed_on_device_matches_oracle():
    from oracle import oracle as orc
    from warpsense_b200.synth import ScanStream
    res, side = 100, 96
    params = Params(map=MapParams(resolution=res, max_distance=1.0, update_distance=0.15, shift=100.0,
                                  size_m=(side * res / 1000.0,) * 3))
    hm = api.HostLocalMap(*params.map.grid_size, params.map.tau, 0)
    om = orc.LocalMap(*params.map.grid_size, params.map.tau, 0)
    mapping = api.TSDFMapping(params, hm)
    feed = api.MappingFeed(mapping)
    s = ScanStream(32, 256, side, res)
    used = 0
    for k in range(5):
        f = s.frame(k)
        pose_m = f["pose"].astype(np.float64)
        pose_m[:3, 3] /= 1000.0
        cloud_m = (f["points_map"].astype(np.float32) / np.float32(1000.0))
        if feed.push(cloud_m, pose_m):
            used += 1
            # the oracle's own chain: pcl::VoxelGrid (restated) on the first cloud (mapping.cpp:70), then
            # preprocess_from_ros = VoxelGrid at the map resolution + metres -> millimetres + pose conversion
            c = orc.voxelgrid(cloud_m, res / 1000.0)[1] if k == 0 else cloud_m
            pts, _ = orc.voxelgrid(c, np.float32(res) / np.float32(1000.0))
            mm_pose = orc.mm_pose_from_isometry(pose_m)
            pos, up = orc.convert_pose(mm_pose, res)
            orc.update_tsdf(om, pts, pos, up, params.map.tau, params.map.max_weight, res)
    assert used >= 3
    got = mapping.get_tsdf_map()
    assert np.array_equal(got.data, om.data)
    mapping.close()
