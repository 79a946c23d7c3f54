#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ with the CPU oracle (oracle/ws_oracle.c, itself
pinned to the reference's known-answer tests G1-G5 by tests/test_oracle_golden.py).

    python tests/golden/make_golden.py

update_small.npz   : two seeded scans into a 33^3 @ 64 mm map -> the non-default voxels (ring index, raw entry)
                     and the work counters after each scan (src/cpu/update_tsdf.cpp:397-564)
register_small.npz : a seeded cloud registered against that map, 8 Gauss-Newton iterations -> the 29 int64 sums
                     per iteration, the total transform and the transformed cloud (src/cpu/registration.cpp:14-177)
preprocess_small.npz : App::preprocess (src/warpsense/app.cpp:118-148) of a seeded float cloud
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle as orc  # noqa: E402

RES, TAU, MW, SIZE = 64, 600, 640, (33, 33, 33)


def inputs():
    rng = np.random.default_rng(20261017)
    spos = np.array([1, -2, 0], np.int32)
    centre = spos.astype(np.int64) * RES
    scans = [(centre + rng.integers(-950, 950, size=(500, 3))).astype(np.int32) for _ in range(2)]
    cloud = (centre + rng.integers(-800, 800, size=(700, 3))).astype(np.int32)
    fcloud = rng.uniform(-6.0, 6.0, size=(900, 3)).astype(np.float32)
    fcloud[::9] = fcloud[1::9][: len(fcloud[::9])]
    pose = np.eye(4, dtype=np.float32)
    a = np.float32(0.3)
    pose[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
    pose[:3, 3] = [120.5, -40.25, 15.0]
    return spos, scans, cloud, fcloud, pose


def main():
    spos, scans, cloud, fcloud, pose = inputs()
    up = np.array([0, 0, orc.MATRIX_RESOLUTION], np.int32)
    m = orc.LocalMap(*SIZE, TAU, 0)
    default = int(m.data[0])
    counters = []
    for pts in scans:
        st = orc.update_tsdf(m, pts, spos, up, TAU, MW, RES)
        counters.append([st["n_candidates"], st["n_touched"], st["n_written"]])
    idx = np.nonzero(m.data != default)[0].astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "update_small.npz"), scanner_pos=spos, up=up, scan0=scans[0], scan1=scans[1],
                        index=idx, entry=m.data[idx].copy(), default_entry=np.uint32(default),
                        counters=np.array(counters, np.int64), params=np.array([RES, TAU, MW, *SIZE], np.int32))
    c = cloud.copy()
    T, it, trace = orc.register_cloud(m, c, np.eye(4, dtype=np.float32), 8, 0.1, 0.0, RES, trace=True)
    np.savez_compressed(os.path.join(HERE, "register_small.npz"), cloud=cloud, transformed=c, T=T, iterations=np.int32(it),
                        trace=trace[:it].astype(np.int64))
    np.savez_compressed(os.path.join(HERE, "preprocess_small.npz"), cloud=fcloud, pose=pose, res=np.int32(RES),
                        points=orc.preprocess(fcloud, pose, RES))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
