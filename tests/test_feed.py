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
    want = api.MappingFeed.subsample(np.concatenate([clouds[5], clouds[3], clouds[4]]), 0.1)
    assert np.array_equal(flushed, want) and not feed.accumulated_
    assert len(feed.poses) == 4 and feed.updates == 3         # every used pose is recorded (:137)


def test_subsample_is_one_centroid_per_leaf():
    pts = np.array([[0.01, 0.01, 0.01], [0.03, 0.05, 0.07], [0.25, 0.01, 0.01], [-0.01, 0.0, 0.0]], np.float32)
    out = api.MappingFeed.subsample(pts, 0.1)
    assert len(out) == 3
    assert any(np.allclose(o, [0.02, 0.03, 0.04], atol=1e-6) for o in out)


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
            c = api.MappingFeed.subsample(cloud_m, res / 1000.0) if k == 0 else cloud_m
            pts, mm_pose = mapping.preprocess_from_ros(c, pose_m)
            pos, up = orc.convert_pose(mm_pose, res)
            orc.update_tsdf(om, pts, pos, up, params.map.tau, params.map.max_weight, res)
    assert used >= 3
    got = mapping.get_tsdf_map()
    assert np.array_equal(got.data, om.data)
    mapping.close()
