"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py with the oracle).

CPU: the oracle still reproduces them (guards the restatement against drift).
GPU: the CUDA path reproduces them through the C ABI -- bit-exact entries, counters, Gauss-Newton sums, cloud."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from warpsense_b200 import api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


def _run_update(update, new_map):
    g = _load("update_small.npz")
    res, tau, mw, sx, sy, sz = (int(v) for v in g["params"])
    m = new_map((sx, sy, sz), tau, mw, res)
    counters = []
    for pts in (g["scan0"], g["scan1"]):
        counters.append(update(m, pts, g["scanner_pos"], g["up"], tau, mw, res))
    return g, m, np.array(counters, np.int64), (res, tau, mw)


def _check_grid(g, data):
    want = np.full(data.shape, g["default_entry"], np.uint32)
    want[g["index"]] = g["entry"]
    assert np.array_equal(data, want), "grid differs from tests/golden/update_small.npz"


def test_oracle_reproduces_golden_update_and_registration():
    def upd(m, pts, pos, up, tau, mw, res):
        st = orc.update_tsdf(m, pts, pos, up, tau, mw, res)
        return [st["n_candidates"], st["n_touched"], st["n_written"]]

    g, m, counters, (res, tau, mw) = _run_update(upd, lambda size, tau, mw, res: orc.LocalMap(*size, tau, 0))
    assert np.array_equal(counters, g["counters"])
    _check_grid(g, m.data)
    r = _load("register_small.npz")
    c = r["cloud"].copy()
    T, it, trace = orc.register_cloud(m, c, np.eye(4, dtype=np.float32), 8, 0.1, 0.0, res, trace=True)
    assert it == int(r["iterations"]) and np.array_equal(trace[:it], r["trace"])
    assert np.array_equal(T, r["T"]) and np.array_equal(c, r["transformed"])
    p = _load("preprocess_small.npz")
    assert np.array_equal(orc.preprocess(p["cloud"], p["pose"], int(p["res"])), p["points"])


@pytest.mark.gpu
def test_cuda_path_reproduces_golden_fixtures():
    hms = {}

    def new_map(size, tau, mw, res):
        hm = api.HostLocalMap(*size, tau, 0)
        t = api.TSDFCuda(api.DeviceMap(hm), tau, mw, res)
        hms[id(t)] = hm
        return t

    def upd(t, pts, pos, up, tau, mw, res):
        t.update_tsdf(pts, pos, up)
        c = t.counters()
        return [c["n_candidates"], c["n_touched"], c["n_written"]]

    g, tsdf, counters, (res, tau, mw) = _run_update(upd, new_map)
    assert np.array_equal(counters, g["counters"])
    hm = hms[id(tsdf)]
    tsdf.avg_map().to_host(api.DeviceMap(hm))
    _check_grid(g, hm.data)
    r = _load("register_small.npz")
    reg = api.RegistrationCuda(tsdf)
    c = r["cloud"].copy()
    T, it = reg.register_cloud(c, np.eye(4, dtype=np.float32), 8, 0.1, 0.0, res)
    assert it == int(r["iterations"]) and np.array_equal(reg.trace()[:it], r["trace"])
    assert np.array_equal(c, r["transformed"])
    assert np.abs(T[:3, :3] - r["T"][:3, :3]).max() <= 1e-4 and np.abs(T[:3, 3] - r["T"][:3, 3]).max() <= 0.1   # mm
    p = _load("preprocess_small.npz")
    got, n = tsdf.preprocess_scan(p["cloud"], p["pose"], int(p["res"]))
    assert np.array_equal(got, p["points"])
    tsdf.close()
