#!/usr/bin/env python
"""Randomised soak of the CUDA path against the oracle (not collected by pytest; run by hand on a GPU box):
random map sizes, resolutions, tau, sensor poses, ring offsets, slab / stripe sharding, cloud sizes.

    python tests/soak.py [n_cases] [seed]
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from warpsense_b200 import api, fixedpoint as fp  # noqa: E402


def one_case(rng, case):
    res = int(rng.choice([20, 33, 50, 64, 100, 250]))
    tau = int(rng.integers(4, 30)) * res // 2 if rng.random() < 0.7 else int(rng.integers(200, 1500))
    tau = max(2, min(tau, 32000))
    mw = int(rng.integers(1, 20)) * 64
    size = tuple(int(rng.integers(8, 48)) * 2 + 1 for _ in range(3))
    world = int(rng.choice([1, 1, 2, 3]))
    stripe = int(rng.choice([0, 0, 1, 2])) if world > 1 else 0
    nbx = (size[0] + 7) // 8
    if nbx < world:
        world = 1
    pos0 = [int(rng.integers(-50, 50)) for _ in range(3)]
    off = [int(rng.integers(0, s)) for s in size]
    om = orc.LocalMap(*size, tau, 0)
    om.set_state(pos0, off)
    hm = api.HostLocalMap(*size, tau, 0)
    hm.pos[:] = pos0
    hm.offset[:] = off
    if stripe:
        os.environ["WS_STRIPE_COLS"] = str(stripe)
    try:
        ranks = [api.TSDFCuda(api.DeviceMap(hm), tau, mw, res, rank=r, world=world) for r in range(world)]
    finally:
        os.environ.pop("WS_STRIPE_COLS", None)
    ext = np.array(size) * res / 2.0
    centre = np.array(pos0, np.int64) * res
    for scan in range(2):
        a, b = rng.uniform(-0.5, 0.5, 2)
        pose = np.eye(4, dtype=np.float32)
        ca, sa, cb, sb = math.cos(a), math.sin(a), math.cos(b), math.sin(b)
        pose[:3, :3] = (np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, cb, -sb], [0, sb, cb]])).astype(np.float32)
        pose[:3, 3] = centre + rng.uniform(-0.3, 0.3, 3) * ext
        spos, up = fp.convert_pose_to_gpu(pose, res)
        n = int(rng.integers(50, 3000))
        pts = (centre + rng.uniform(-1.1, 1.1, size=(n, 3)) * ext).astype(np.int32)     # some fall outside the map
        st = orc.update_tsdf(om, pts, spos, up, tau, mw, res)
        tot_c = 0
        for t in ranks:
            t.update_tsdf(pts, spos, up)
            c = t.counters()
            tot_c += c["n_candidates"]
            if world == 1:
                assert (c["n_candidates"], c["n_touched"], c["n_written"]) == (st["n_candidates"], st["n_touched"], st["n_written"]), (case, scan)
        assert tot_c >= st["n_candidates"], (case, scan)
    row = size[1] * size[2]
    covered = np.zeros(size[0], np.int32)
    for t in ranks:
        rows = t.owned_rows()
        covered += rows
        back = api.HostLocalMap(*size, tau, 0)
        t.avg_map().to_host(api.DeviceMap(back))
        assert np.array_equal(back.data.reshape(-1, row)[rows], om.data.reshape(-1, row)[rows]), \
            "case %d: grid differs (size %s res %d tau %d world %d stripe %d)" % (case, size, res, tau, world, stripe)
    assert (covered == 1).all()
    if world == 1:
        cloud = (centre + rng.uniform(-0.9, 0.9, size=(int(rng.integers(100, 2000)), 3)) * ext).astype(np.int32)
        oc = cloud.copy()
        I = np.eye(4, dtype=np.float32)
        oT, oit, otr = orc.register_cloud(om, oc, I, 6, 0.1, 0.0, res, trace=True)
        if otr[:oit, 28].min() > 0:          # the reference divides by the match count: skip degenerate clouds
            reg = api.RegistrationCuda(ranks[0])
            T, it = reg.register_cloud(cloud, I, 6, 0.1, 0.0, res)
            assert it == oit and np.array_equal(reg.trace()[:it], otr[:it]), "case %d: registration sums differ" % case
            assert np.array_equal(cloud, oc), "case %d: transformed cloud differs" % case
    for t in ranks:
        t.close()
    return size, res, tau, world, stripe


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    rng = np.random.default_rng(seed)
    for case in range(n):
        info = one_case(rng, case)
        print("case %3d ok  size %s res %d tau %d world %d stripe %d" % ((case,) + info), flush=True)
    print("soak ok: %d cases, seed %d" % (n, seed))


if __name__ == "__main__":
    main()
