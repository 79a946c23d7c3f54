"""x-slab sharding of the ring over ranks.

CPU (`not gpu`): the slab layout is a partition with one-row halos; a world_size-2 gloo run of the
registration exchange (oracle partial sums of each rank's slab -> all_reduce(int64[29]) -> identical solve)
reproduces the single-process oracle trace bit for bit.
GPU: the same map held as 2 or 3 sharded handles (one per "rank", all on cuda:0) gives bit-identical voxels,
sums and poses to the oracle -- the NCCL all-reduce is replaced by a host-side sum of ws_reg_sums_get()."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from warpsense_b200 import api, fixedpoint as fp
from warpsense_b200.synth import ScanStream

MR = fp.MATRIX_RESOLUTION


# ----------------------------------------------------------------------------- host logic
@pytest.mark.parametrize("size_x,world", [(5, 1), (21, 2), (129, 2), (129, 3), (513, 2), (513, 4), (513, 8), (2049, 8), (65, 8)])
def test_slab_layout_is_a_partition_with_halos(size_x, world):
    owned = np.zeros(size_x, np.int32)
    for r in range(world):
        lo, hi, cols = api.slab_layout(size_x, r, world)
        assert 0 <= lo < hi <= size_x and lo % 8 == 0
        owned[lo:hi] += 1
        resident = set()
        for c in cols:
            resident.update(range(c * 8, min(c * 8 + 8, size_x)))
        need = set(range(lo, hi))
        if world > 1:
            need |= {(lo - 1) % size_x, hi % size_x}        # the registration stencil reads x+-1 (ring wrap)
        assert need <= resident
    assert (owned == 1).all(), "every ring-x row is owned by exactly one rank"


@pytest.mark.parametrize("size_x,world,stripe", [(97, 2, 2), (97, 3, 1), (513, 2, 8), (513, 4, 4), (513, 8, 4), (513, 8, 8),
                                                 (2049, 8, 16), (129, 3, 0), (65, 8, 1)])
def test_stripe_layout_is_a_partition_with_halos(size_x, world, stripe):
    """Round-robin stripes: every brick column has exactly one owner, and every rank keeps resident the
    columns that hold the rows just outside each of its stripes (registration reads x +- 1)."""
    nbx = (size_x + 7) // 8
    owners = np.zeros(nbx, np.int32)
    for r in range(world):
        own, res = api.shard_layout(size_x, r, world, stripe)
        owners += own
        assert own.any() and (res | ~own).all()
        for c in np.nonzero(own)[0]:
            lo_row, hi_row = c * 8, min(c * 8 + 8, size_x) - 1
            assert res[((lo_row - 1) % size_x) // 8] and res[((hi_row + 1) % size_x) // 8]
        if stripe and nbx // stripe >= world:
            assert all(((c // stripe) % world == r) == bool(own[c]) for c in range(nbx))
    assert (owners == 1).all()


def test_slab_layout_rejects_more_ranks_than_columns():
    with pytest.raises(ValueError):
        api.slab_layout(17, 0, 4)      # 3 brick columns, 4 ranks: rank 0 gets none


def _owned_mask(points, T, res, om, lo, hi):
    """Points whose centre voxel's ring-x row lies in [lo, hi) (the rank that sums them)."""
    q = fp.transform_points(points, fp.to_int_mat(T))
    vx = (np.sign(q[:, 0]) * (np.abs(q[:, 0]) // res)).astype(np.int64)       # trunc toward zero (registration.cpp:65)
    size, pos, off = int(om.size[0]), int(om.pos[0]), int(om.offset[0])
    inb = np.abs(vx - pos) <= size // 2
    rx = (vx - pos + off + size) % size
    return inb & (rx >= lo) & (rx < hi)


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res, tau, mw, side = 100, 1000, 640, 96
        s = ScanStream(16, 128, side, res)
        om = orc.LocalMap(side, side, side, tau, 0)
        for k in range(2):
            f = s.frame(k)
            pos, up = fp.convert_pose_to_gpu(f["pose"], res)
            orc.update_tsdf(om, f["points_map"], pos, up, tau, mw, res)
        om.set_state(list(om.pos), [9, 40, 48])              # ring origin away from the array origin
        cloud = s.frame(2, prior_pose=s.pose(1))["points_prior"]
        lo, hi, _ = api.slab_layout(int(om.size[0]), rank, world)
        T, alpha = np.eye(4, dtype=np.float32), 0.0
        trace = []
        for it in range(6):
            mask = _owned_mask(cloud, T, res, om, lo, hi)
            H, g, e, c = orc.reg_step(om, cloud[mask], T, res)
            # but points are transformed by the FULL cloud's T: reg_step does that per point, so a subset is exact
            sums = np.concatenate([H[np.triu_indices(6)], g, [e, c]]).astype(np.int64)
            t = torch.from_numpy(sums)
            dist.all_reduce(t)                                # the one exchange step of the path
            tot = t.numpy()
            Hf = np.zeros((6, 6), np.int64)
            Hf[np.triu_indices(6)] = tot[:21]
            Hf = Hf + Hf.T - np.diag(np.diag(Hf))
            T, _, _ = orc.reg_solve(Hf, tot[21:27], int(tot[27]), int(tot[28]), alpha, T)
            alpha += 0.1
            trace.append(tot.copy())
        # single-process oracle for comparison (same on every rank)
        oT, oit, otr = orc.register_cloud(om, cloud.copy(), np.eye(4, dtype=np.float32), 6, 0.1, 0.0, res, trace=True)
        ok = np.array_equal(np.array(trace), otr) and np.array_equal(T, oT)
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, T.tobytes()))
        if rank == 0:
            out.put((all(g[0] for g in gathered), len({g[1] for g in gathered})))
    finally:
        dist.destroy_process_group()


def test_registration_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    all_ok, distinct_T = out.get(timeout=10)
    assert all_ok, "all-reduced partial sums differ from the single-process oracle trace"
    assert distinct_T == 1, "ranks disagree on the transform"


# ----------------------------------------------------------------------------- GPU: sharded handles
def _slab_rows(hm_size, lo, hi):
    row = int(hm_size[1]) * int(hm_size[2])
    return slice(lo * row, hi * row)


@pytest.mark.gpu
@pytest.mark.parametrize("world,ring_offset,stripe", [(2, None, 0), (3, None, 0), (4, (30, 5, 60), 0), (2, (93, 0, 11), 0),
                                                      (2, None, 2), (3, (30, 5, 60), 1), (4, None, 1), (2, (93, 0, 11), 3)])
def test_sharded_handles_match_oracle_on_one_gpu(world, ring_offset, stripe, monkeypatch):
    """stripe = 0: contiguous x slabs; stripe = k: stripes of k brick columns dealt round-robin (WS_STRIPE_COLS)."""
    if stripe:
        monkeypatch.setenv("WS_STRIPE_COLS", str(stripe))
    res, tau, mw, side = 100, 1000, 640, 96
    s = ScanStream(32, 256, side, res)
    om = orc.LocalMap(side, side, side, tau, 0)
    hm = api.HostLocalMap(side, side, side, tau, 0)
    if ring_offset is not None:
        # the ring seam (hdf5_local_map.h:140-151) falls inside a slab: its resident columns are two x intervals
        om.set_state([0, 0, 0], list(ring_offset))
        hm.offset[:] = ring_offset
    ranks = []
    for r in range(world):
        t = api.TSDFCuda(api.DeviceMap(hm), tau, mw, res, device=0, rank=r, world=world)
        ranks.append((t, api.RegistrationCuda(t)))
    # ---- update: no exchange; every rank marches the ray segments that can reach its slab ----
    for k in range(3):
        f = s.frame(k)
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        st = orc.update_tsdf(om, f["points_map"], pos, up, tau, mw, res)
        touched = cands = 0
        for t, _ in ranks:
            t.update_tsdf(f["points_map"], pos, up)
            c = t.counters()
            assert 0 < c["n_candidates"] <= st["n_candidates"]   # candidates that land in resident columns
            cands += c["n_candidates"]
            touched += c["n_touched"]
        assert touched >= st["n_touched"]                        # halo columns are updated twice
        assert cands >= st["n_candidates"]
    covered = np.zeros(int(hm.size[0]), np.int32)
    row = int(hm.size[1]) * int(hm.size[2])
    for r, (t, _) in enumerate(ranks):
        rows = t.owned_rows()
        covered += rows
        back = api.HostLocalMap(side, side, side, tau, 0)
        t.avg_map().to_host(api.DeviceMap(back))
        got = back.data.reshape(-1, row)[rows]
        assert np.array_equal(got, om.data.reshape(-1, row)[rows]), "rank %d: owned rows differ from the oracle" % r
    assert (covered == 1).all(), "the owned rows of the ranks are not a partition of the ring"
    # ---- registration: host-side sum of the 29 int64 stands in for ncclAllReduce ----
    cloud = s.frame(3, prior_pose=s.pose(2))["points_prior"].copy()
    oT, oit, otr = orc.register_cloud(om, cloud.copy(), np.eye(4, dtype=np.float32), 8, 0.1, 0.0, res, trace=True)
    for _, reg in ranks:
        reg.prepare_registration(cloud)
    hds = [reg._hd for _, reg in ranks]
    I16 = fp.colmajor16(np.eye(4, dtype=np.float32))
    import ctypes as C
    f32p = C.POINTER(C.c_float)
    for hd in hds:
        hd.check(hd.L.ws_reg_begin(hd.h, I16.ctypes.data_as(f32p)))
    for it in range(8):
        for hd in hds:
            hd.check(hd.L.ws_reg_accumulate(hd.h, res))
        tot = sum(reg.sums_get() for _, reg in ranks)
        assert np.array_equal(tot, otr[it]), "iteration %d: summed slab sums differ from the oracle" % it
        for (_, reg), hd in zip(ranks, hds):
            reg.sums_set(tot)
            hd.check(hd.L.ws_reg_solve(hd.h, 0.1, 0.0))
    out = np.zeros(16, np.float32)
    Ts = []
    for hd in hds:
        it_, fin_ = C.c_int32(), C.c_int32()
        hd.check(hd.L.ws_reg_finish(hd.h, out.ctypes.data_as(f32p), C.byref(it_), C.byref(fin_)))
        assert it_.value == 8
        Ts.append(fp.from_colmajor16(out).copy())
    for T in Ts:
        assert np.array_equal(T, Ts[0])
        assert np.abs(T[:3, :3] - oT[:3, :3]).max() <= 1e-4 and np.abs(T[:3, 3] - oT[:3, 3]).max() <= 0.1
    for t, _ in ranks:
        t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,stripe", [(2, 0), (3, 0), (4, 0), (2, 2)])
def test_shift_on_sharded_handles(world, stripe, monkeypatch):
    """HDF5LocalMap::shift (src/map/hdf5_local_map.cpp:53-118) on a map sharded over `world` handles: slabs are
    in ring (storage) coordinates, so a shift never re-partitions -- after every shift each rank's owned rows equal
    the oracle's, and the ranks' chunk stores, merged by row owner, equal the oracle's chunk store."""
    if stripe:
        monkeypatch.setenv("WS_STRIPE_COLS", str(stripe))
    rng = np.random.default_rng(5 + world)
    tau, res, mw = 600, 64, 640
    size = (41, 25, 17)
    om = orc.LocalMap(*size, tau, 0)
    hm = api.HostLocalMap(*size, tau, 0)
    ranks = [api.TSDFCuda(api.DeviceMap(hm), tau, mw, res, device=0, rank=r, world=world) for r in range(world)]
    up = np.array([0, 0, MR], np.int32)
    pos = np.zeros(3, np.int64)
    row = int(hm.size[1]) * int(hm.size[2])

    def check(what):
        covered = np.zeros(int(hm.size[0]), np.int32)
        for r, t in enumerate(ranks):
            rows = t.owned_rows()
            covered += rows
            back = api.HostLocalMap(*size, tau, 0)
            t.avg_map().to_host(api.DeviceMap(back))
            assert list(back.pos) == list(om.pos) and list(back.offset) == list(om.offset)
            assert np.array_equal(back.data.reshape(-1, row)[rows], om.data.reshape(-1, row)[rows]), \
                "%s: rank %d owned rows differ from the oracle" % (what, r)
        assert (covered == 1).all()

    for step in range(7):
        pts = (pos * res + rng.integers(-900, 900, size=(400, 3))).astype(np.int32)
        orc.update_tsdf(om, pts, pos, up, tau, mw, res)
        for t in ranks:
            t.update_tsdf(pts, pos, up)
        pos = pos + np.array([int(rng.integers(-30, 31)), int(rng.integers(-20, 21)), int(rng.integers(-5, 6))])
        om.shift(pos)
        for t in ranks:
            t.shift(pos)
        check("after shift %d to %r" % (step, list(pos)))
    # back to the start: what left the window comes back from each rank's own store
    while np.abs(pos).max() > 0:
        pos = pos - np.clip(pos, -12, 12)                      # a shift may not exceed the map size (:63-65)
        om.shift(pos)
        for t in ranks:
            t.shift(pos)
    check("back at the origin")
    # chunk stores: merge the ranks' chunks row by row (ring-x owner of the world row) and compare with the oracle's
    om.write_back()
    for t in ranks:
        t.write_back()
    owned = [t.owned_rows() for t in ranks]
    sx = int(hm.size[0])
    for c in om.chunk_list():
        want = om.chunk(*c).reshape(64, 64 * 64)
        got = np.full_like(want, want[0, 0])
        have = [t.chunk(*c) for t in ranks]
        for i in range(64):
            x = c[0] * 64 + i
            ring = (x - int(om.pos[0]) + int(om.offset[0]) + 2 * sx) % sx
            in_window = abs(x - int(om.pos[0])) <= sx // 2
            # a world row outside the final window was saved by the rank that owned its ring row when it left
            owners = [r for r in range(world) if owned[r][ring] and have[r] is not None]
            assert owners or not in_window
            if owners:
                got[i] = have[owners[0]].reshape(64, 64 * 64)[i]
            else:
                got[i] = want[i]
        assert np.array_equal(got, want), "merged chunk %r differs from the oracle" % (c,)
    for t in ranks:
        t.close()


# ----------------------------------------------------------------------------- GPU: fused peer exchange
def _build_sharded_scene(world, device_of_rank, frames=2):
    res, tau, mw, side = 100, 1000, 640, 96
    s = ScanStream(32, 256, side, res)
    om = orc.LocalMap(side, side, side, tau, 0)
    hm = api.HostLocalMap(side, side, side, tau, 0)
    ranks = []
    for r in range(world):
        t = api.TSDFCuda(api.DeviceMap(hm), tau, mw, res, device=device_of_rank(r), rank=r, world=world)
        ranks.append((t, api.RegistrationCuda(t)))
    for k in range(frames):
        f = s.frame(k)
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        orc.update_tsdf(om, f["points_map"], pos, up, tau, mw, res)
        for t, _ in ranks:
            t.update_tsdf(f["points_map"], pos, up)
    cloud = s.frame(frames, prior_pose=s.pose(frames - 1))["points_prior"].copy()
    return s, om, ranks, cloud, res


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_fused_peer_registration_one_process(world, monkeypatch):
    """register_cloud on slab-sharded handles with the in-kernel mailbox exchange (registration.cu): `world`
    persistent kernels side by side on ONE GPU (WS_REG_BLOCKS keeps them co-resident), one host thread per
    rank.  Sums per iteration and the pose must equal the single-map oracle bit for bit."""
    import threading
    monkeypatch.setenv("WS_REG_BLOCKS", "24")
    s, om, ranks, cloud, res = _build_sharded_scene(world, lambda r: 0)
    iters = 9
    oT, oit, otr = orc.register_cloud(om, cloud.copy(), np.eye(4, dtype=np.float32), iters, 0.1, 0.0, res, trace=True)
    ptrs = [reg.peer_local_ptr() for _, reg in ranks]
    for _, reg in ranks:
        reg.peer_attach_ptrs(ptrs)
        reg.peer_set_timeout(4.0)
    for rep in range(2):                      # twice: the epoch parity of the mailbox slots alternates
        results, errors = [None] * world, []

        def run(r):
            try:
                c = cloud.copy()
                results[r] = ranks[r][1].register_cloud(c, np.eye(4, dtype=np.float32), iters, 0.1, 0.0, res) + (c,)
            except Exception as exc:          # noqa: BLE001
                errors.append((r, repr(exc)))

        ths = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=60)
        assert not errors, errors
        ocloud = cloud.copy()
        orc.register_cloud(om, ocloud, np.eye(4, dtype=np.float32), iters, 0.1, 0.0, res)
        for r in range(world):
            T, it, c = results[r]
            assert it == oit
            assert np.array_equal(ranks[r][1].trace()[:it], otr[:it]), "rank %d: summed sums differ from the oracle" % r
            assert np.array_equal(T, results[0][0]), "ranks disagree on the transform"
            assert np.abs(T[:3, :3] - oT[:3, :3]).max() <= 1e-4 and np.abs(T[:3, 3] - oT[:3, 3]).max() <= 0.1
            assert np.array_equal(c, ocloud), "transformed cloud differs from the oracle"
    for t, _ in ranks:
        t.close()


@pytest.mark.gpu
def test_fused_peer_registration_times_out_without_peer(monkeypatch):
    """A rank whose peer never shows up must return an error, not hang."""
    monkeypatch.setenv("WS_REG_BLOCKS", "24")
    s, om, ranks, cloud, res = _build_sharded_scene(2, lambda r: 0, frames=1)
    ptrs = [reg.peer_local_ptr() for _, reg in ranks]
    reg0 = ranks[0][1]
    reg0.peer_attach_ptrs(ptrs)
    reg0.peer_set_timeout(0.3)
    with pytest.raises(Exception) as ei:
        reg0.register_cloud(cloud.copy(), np.eye(4, dtype=np.float32), 4, 0.1, 0.0, res)
    assert "peer" in str(ei.value)
    with pytest.raises(Exception):            # sharded map without attached peers
        ranks[1][1].register_cloud(cloud.copy(), np.eye(4, dtype=np.float32), 4, 0.1, 0.0, res)
    for t, _ in ranks:
        t.close()


@pytest.mark.gpu
def test_fused_peer_registration_two_processes():
    """Two processes, one GPU each, mailboxes exchanged as CUDA IPC handles (tests/mp_peer_check.py)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(root, "tests", "mp_peer_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "mp peer check ok" in out.stdout
