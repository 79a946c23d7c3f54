#!/usr/bin/env python
"""Multi-process check of the fused multi-GPU path: one process per GPU (torchrun), x-slab sharded maps,
update_tsdf with per-rank step culling, register_cloud with the in-kernel NVLink mailbox exchange.
Every rank compares its slab and the registration trace with the CPU oracle (test infrastructure)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from warpsense_b200 import api, fixedpoint as fp
    from warpsense_b200.synth import ScanStream

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    res, tau, mw, side = 100, 1000, 640, 96
    s = ScanStream(32, 256, side, res)
    om = orc.LocalMap(side, side, side, tau, 0)
    hm = api.HostLocalMap(side, side, side, tau, 0)
    tsdf = api.TSDFCuda(api.DeviceMap(hm), tau, mw, res, device=local, rank=rank, world=world)
    reg = api.RegistrationCuda(tsdf)
    handles = [None] * world
    dist.all_gather_object(handles, reg.peer_export())
    reg.peer_attach_ipc(handles)
    reg.peer_set_timeout(10.0)
    I = np.eye(4, dtype=np.float32)
    for k in range(4):
        f = s.frame(k, prior_pose=s.pose(max(k - 1, 0)))
        cloud = f["points_prior"].copy() if k > 0 else f["points_map"].copy()
        pose = f["pose"]
        if k > 0:
            ocloud = cloud.copy()
            oT, oit, otr = orc.register_cloud(om, ocloud, I, 12, 0.1, 0.0, res, trace=True)
            dist.barrier()
            if k >= 2:
                # fused per-scan pipeline on the sharded map: registration -> pose -> update on the device
                prior = s.pose(k - 1)
                T, pose, it = reg.track_scan(cloud, prior, 12, 0.1, 0.0, res)
                assert it == oit and np.array_equal(T, oT), "rank %d frame %d: fused transform differs" % (rank, k)
                assert np.array_equal(reg.trace()[:it], otr[:it])
                opos, oup = orc.convert_pose(pose, res)
                orc.update_tsdf(om, ocloud, opos, oup, tau, mw, res)
            else:
                T, it = reg.register_cloud(cloud, I, 12, 0.1, 0.0, res)
                assert it == oit, (it, oit)
                assert np.array_equal(reg.trace()[:it], otr[:it]), "rank %d frame %d: sums differ from the oracle" % (rank, k)
                assert np.array_equal(cloud, ocloud), "transformed cloud differs"
                pose = (T @ s.pose(k - 1)).astype(np.float32)
        if k < 2:
            pos, up = fp.convert_pose_to_gpu(pose, res)
            orc.update_tsdf(om, cloud, pos, up, tau, mw, res)
            tsdf.update_tsdf(cloud, pos, up)
        rows = tsdf.owned_rows()
        back = api.HostLocalMap(side, side, side, tau, 0)
        tsdf.avg_map().to_host(api.DeviceMap(back))
        row = int(hm.size[1]) * int(hm.size[2])
        assert np.array_equal(back.data.reshape(-1, row)[rows], om.data.reshape(-1, row)[rows]), "rank %d: owned rows differ" % rank
    dist.barrier()
    tsdf.close()
    if rank == 0:
        print("mp peer check ok (world %d)" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
