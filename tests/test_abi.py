"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every symbol the
public header declares; the host-side mirrors agree with the oracle.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from warpsense_b200 import api, build, fixedpoint as fp, lib
from warpsense_b200.params import MapParams
from warpsense_b200.synth import ScanStream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    path = build.build_library()
    assert os.path.exists(path)
    hdr = open(os.path.join(ROOT, "include", "warpsense_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    L = C.CDLL(path)
    for name in sorted(declared):
        assert hasattr(L, name), "header declares %s but the library does not export it" % name
    # and the python binding binds exactly the header's surface
    assert declared == set(lib.EXPORTED_SYMBOLS)
    assert b"sm_100a" in lib.load().ws_version()


def test_create_fails_loudly_without_gpu_or_with_bad_args():
    L = lib.load()
    h = C.c_void_p()
    size = np.array([21, 21, 21], np.int32)
    p = size.ctypes.data_as(C.POINTER(C.c_int32))
    assert L.ws_create(p, 3000, 640, 1000, 0, None) == lib.WS_ERR_INVALID
    even = np.array([20, 21, 21], np.int32)
    assert L.ws_create(even.ctypes.data_as(C.POINTER(C.c_int32)), 3000, 640, 1000, 0, C.byref(h)) == lib.WS_ERR_INVALID
    assert L.ws_create(p, 0, 640, 1000, 0, C.byref(h)) == lib.WS_ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        assert L.ws_create(p, 3000, 640, 1000, 0, C.byref(h)) == lib.WS_ERR_CUDA
        with pytest.raises(lib.WarpsenseError):
            api.TSDFCuda(api.DeviceMap(api.HostLocalMap(20, 20, 20, 3000)), 3000, 640, 1000)
    assert L.ws_sync(None) == lib.WS_ERR_INVALID
    assert b"null" in L.ws_last_error(None)


def test_fixedpoint_host_mirror_matches_oracle():
    rng = np.random.default_rng(0)
    s = ScanStream(16, 64, 128, 100)
    for k in (0, 3, 17):
        pose = s.pose(k)
        assert (fp.to_int_mat(pose) == orc.to_int_mat(pose)).all()
        pts = rng.integers(-20000, 20000, size=(500, 3)).astype(np.int32)
        assert (fp.transform_points(pts, fp.to_int_mat(pose)) == orc.transform_points(pts, orc.to_int_mat(pose))).all()
        pos, up = fp.convert_pose_to_gpu(pose, 100)
        opos, oup = orc.convert_pose(pose, 100)
        assert (pos == opos).all() and (up == oup).all()
    neg = np.eye(4, dtype=np.float32)
    neg[:3, 3] = [-30.0, -64.0, 63.9]
    assert list(fp.to_map(neg, 64)) == list(orc.to_map(neg, 64)) == [-1, -1, 0]


def test_host_local_map_matches_oracle_indexing():
    om = orc.LocalMap(8, 10, 6, 600, 0)
    hm = api.HostLocalMap(8, 10, 6, 600, 0)
    assert (hm.size == om.size).all() and (hm.offset == om.offset).all()
    om.set_state([3, -2, 1], [1, 7, 4])
    hm.pos[:] = [3, -2, 1]
    hm.offset[:] = [1, 7, 4]
    rng = np.random.default_rng(1)
    for _ in range(200):
        x, y, z = (int(v) for v in rng.integers(-8, 9, size=3))
        assert hm.in_bounds(x, y, z) == om.in_bounds(x, y, z)
        if om.in_bounds(x, y, z):
            assert hm.get_index(x, y, z) == om.index(x, y, z)
    assert api.make_entry(-5, -64) == orc.make_entry(-5, -64)
    assert (api.entry_value(orc.make_entry(-5, -64)), api.entry_weight(orc.make_entry(-5, -64))) == (-5, -64)
    with pytest.raises(IndexError):
        hm.value(100, 0, 0)


def test_param_scaling_g5():
    """test/params_a.cpp:26-41 through the product's own MapParams."""
    p = MapParams(resolution=1000, max_distance=3.0, max_weight=10, size_m=(20, 20, 20))
    assert p.tau == 3000 and p.max_weight == 640 and p.size == (20, 20, 20) and p.grid_size == (21, 21, 21)
    tau, mw, size = orc.scale_params(1.0, 10, [40, 40, 25], 64)
    q = MapParams(resolution=64, max_distance=1.0, max_weight=10, size_m=(40, 40, 25))
    assert (q.tau, q.max_weight, q.size) == (tau, mw, tuple(int(v) for v in size))


def test_synthetic_stream_is_deterministic():
    a = ScanStream(16, 64, 128, 100).frame(3)
    b = ScanStream(16, 64, 128, 100).frame(3)
    assert np.array_equal(a["points_map"], b["points_map"])
    assert a["points_map"].shape == (16 * 64, 3)
    assert a["points_map"].dtype == np.int32
