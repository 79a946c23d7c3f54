#!/usr/bin/env python
"""bench.py -- OS1-128 scans/s through the TSDF hot path (register_cloud 20 GN iterations + update_tsdf).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line on rank 0.
Workload (BASELINE.json configs[2], which contains configs[1]): synthetic 128x1024 OS1-128 stream into a
512^3 @ 5 cm local TSDF (513^3 voxels); per step = one scan: Point-to-TSDF registration (exactly 20
Gauss-Newton iterations, epsilon 0) against the map of the previous frames, then update_tsdf with the
registered cloud.
  value : scans/s with the scan already resident in HBM when the timed region starts;
  e2e   : the same through the public API with the scan in pinned HOST memory (H2D of the points and D2H
          of the pose + counters inside the timed region);
  roofline : update_tsdf kernels (ray march + merge + replay), algorithmic bytes 12*N + 8*T per scan
           over their CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth;
  cpu_baseline : the CPU oracle (a C port of the reference's src/cpu path) on this box's host cores.
N > 1: the grid is sharded in x-slabs over the ranks (one process per GPU); every rank marches only the ray
segments that can reach its slab, and the 29 int64 Gauss-Newton sums are exchanged once per iteration INSIDE
the persistent registration kernel through NVLink peer mailboxes (CUDA IPC); NCCL only carries the timing
reduction and the barrier.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "OS1-128 scans/sec (update_tsdf+reg)"
UNIT = "scans/s"
GN_ITERS = 20
IT_WEIGHT = 0.1
EPSILON = 0.0          # |.| < 0 never holds: exactly GN_ITERS iterations
TAU, MAX_WEIGHT = 1000, 640
# DRAM bytes of one update_tsdf on the default workload, from the committed ncu capture (profiles/r01p_*):
# setup 1.6 + 12.5 + march 224.9 + 541.9 + brick_list 1.1 + merge 353.2 + 380.3 + replay 623.9 + 35.5 MB
NCU_TRAFFIC_BYTES_PER_SCAN = 2174.9e6
REF_SUBSAMPLE = 8      # --impl reference: every 8th ray per step (bounded sample)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--grid", type=int, default=512, help="local map side in voxels")
    ap.add_argument("--res", type=int, default=50, help="voxel size in mm")
    ap.add_argument("--beams", type=int, default=128)
    ap.add_argument("--cols", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the reference's own CUDA kernels")
    ap.add_argument("--update-only", action="store_true", help="BASELINE configs[1]: update_tsdf without registration")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every few
    milliseconds from a thread (the timed region of the default run lasts tens of milliseconds -- too short for
    `nvidia-smi -lms`), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap",
            }
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(mx)
                r = int(get_reasons(h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.008)
        finally:
            nv.nvmlShutdown()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.gpu])
            except (ValueError, IndexError):
                pass
        return self.gpu

    def _smi_loop(self):
        proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q,
                                 "--format=csv,noheader,nounits", "-lms", "20"],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in proc.stdout:
                parts = [p.strip() for p in ln.split(",")]
                if len(parts) >= 9:
                    try:
                        self.sm.append(float(parts[1])); self.mx.append(float(parts[2]))
                    except ValueError:
                        continue
                    for name, val in zip(names, parts[5:9]):
                        if val.lower().startswith("active"):
                            self.reasons.add(name)
                if self.stop_flag.is_set():
                    break
        finally:
            proc.terminate()

    def _run(self):
        try:
            self.source = "nvml"
            self._nvml_loop()
        except Exception:
            try:
                self.source = "nvidia-smi"
                self._smi_loop()
            except Exception:
                self.source = None

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        time.sleep(0.02)           # let the first sample land before the timed region begins

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=3)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx), "samples": len(self.sm),
                "reasons": sorted(self.reasons), "source": self.source}


def make_frames(args, count):
    from warpsense_b200.synth import ScanStream
    s = ScanStream(args.beams, args.cols, args.grid, args.res)
    frames = [s.frame(0)]
    for k in range(1, count + 1):
        frames.append(s.frame(k, prior_pose=s.pose(k - 1)))
    if os.environ.get("WS_BENCH_TRANSPOSE"):      # experiment: column-major scan order (vertical neighbours adjacent)
        for f in frames:
            for key in ("points_map", "points_prior"):
                if key in f and len(f[key]) == args.beams * args.cols:
                    f[key] = np.ascontiguousarray(f[key].reshape(args.beams, args.cols, 3).transpose(1, 0, 2).reshape(-1, 3))
    return s, frames


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port of src/cpu; the reference itself needs
    Eigen/HDF5/PCL/ROS and cannot be built here) on this box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from warpsense_b200 import fixedpoint as fp
    K, W = args.steps, args.warmup
    s, frames = make_frames(args, K + W)
    side, res = args.grid, args.res
    om = orc.LocalMap(side, side, side, TAU, 0)
    # all the host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        orc.set_num_threads(os.cpu_count() or 1)
    cores = orc.num_threads()
    pos, up = fp.convert_pose_to_gpu(frames[0]["pose"], res)
    orc.update_tsdf(om, frames[0]["points_map"][::REF_SUBSAMPLE], pos, up, TAU, MAX_WEIGHT, res)
    I = np.eye(4, dtype=np.float32)
    times = []
    for k in range(1, K + W + 1):
        f = frames[k]
        cloud = np.ascontiguousarray(f["points_prior"][::REF_SUBSAMPLE])
        t0 = time.perf_counter()
        if not args.update_only:
            T, it = orc.register_cloud(om, cloud, I, GN_ITERS, IT_WEIGHT, EPSILON, res)
            pose = (T @ s.pose(k - 1)).astype(np.float32)
        else:
            pose = f["pose"]
            cloud = np.ascontiguousarray(f["points_map"][::REF_SUBSAMPLE])
        pos, up = fp.convert_pose_to_gpu(pose, res)
        orc.update_tsdf(om, cloud, pos, up, TAU, MAX_WEIGHT, res)
        dt = time.perf_counter() - t0
        if k > W:
            times.append(dt)
    n_full = len(frames[1]["points_prior"])
    n_s = len(frames[1]["points_prior"][::REF_SUBSAMPLE])
    total = sum(times)
    # a step covers n_s of the n_full rays of a scan: scale to whole scans
    value = (K * n_s / n_full) / total
    sample = ("every %dth ray (%d of %d) of each scan per step; update_tsdf single-threaded as launched by the "
              "reference (update_tsdf.cpp:405), register_cloud %d iterations on %d OpenMP threads; scans/s scaled by "
              "rays processed" % (REF_SUBSAMPLE, n_s, n_full, GN_ITERS, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1000.0 * total / K * (n_full / n_s),
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "int32+int64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    side = args.grid if args.grid % 2 == 1 else args.grid + 1
    return {
        "workload": ("synthetic OS1-128 %dx%d stream, %d^3 @%dmm local TSDF (%d^3 voxels), "
                     % (args.beams, args.cols, args.grid, args.res, side)
                     + ("update_tsdf only" if args.update_only else
                        "Point-to-TSDF reg %d GN iters + update_tsdf" % GN_ITERS)),
        "baseline_config": "configs[1]" if args.update_only else "configs[2]",
        "points_per_scan": args.beams * args.cols,
        "tau_mm": TAU, "max_weight": MAX_WEIGHT,
        "l2_policy": "working set larger than L2: every step is a different scan and touches ~21 M voxels "
                     "(12 B key+entry each, ~255 MB) of a 1.6 GB grid+scratch; no explicit flush",
        "parallelism": "x-slab spatial sharding of the ring; per-rank culling of march steps; int64[29] sums exchanged per GN iteration inside the kernel over NVLink peer memory",
    }


# ------------------------------------------------------------------------------------------------
class SumsView:
    """__cuda_array_interface__ view of the handle's int64[29] Gauss-Newton sums (for NCCL)."""

    def __init__(self, ptr):
        self.__cuda_array_interface__ = {"shape": (29,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def run_native(args):
    import torch
    import torch.distributed as dist
    from warpsense_b200 import api, fixedpoint as fp, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the CUDA library is the only compute path)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W = args.steps, args.warmup
    if W < 3:
        W = 3
    s, frames = make_frames(args, K + W)
    side, res = args.grid, args.res
    N = len(frames[1]["points_prior"])

    size = tuple(v if v % 2 == 1 else v + 1 for v in (side, side, side))

    class _View:   # DeviceMap-like view without a 540 MB host array
        size_ = np.array(size, np.int32)
        offset_ = np.array([v // 2 for v in size], np.int32)
        pos_ = np.zeros(3, np.int32)
        data_ = None

    tsdf = api.TSDFCuda(_View(), TAU, MAX_WEIGHT, res, device=local_rank, rank=rank, world=world, upload=False)
    reg = api.RegistrationCuda(tsdf)
    hd = tsdf.device_map()
    stream = torch.cuda.Stream()
    tsdf.set_stream(stream.cuda_stream)
    if world > 1:
        # fused exchange: every rank maps every rank's mailbox (CUDA IPC over NVLink); the Gauss-Newton sums
        # then travel inside the persistent registration kernel -- no NCCL call per iteration
        handles = [None] * world
        dist.all_gather_object(handles, reg.peer_export())
        reg.peer_attach_ipc(handles)
        reg.peer_set_timeout(20.0)

    I16 = fp.colmajor16(np.eye(4, dtype=np.float32))
    import ctypes as C
    f32p = C.POINTER(C.c_float)
    Tout = np.zeros(16, np.float32)
    it_out, fin_out = C.c_int32(), C.c_int32()

    def register(n):
        """20 GN iterations on the cloud already staged with prepare_registration*; returns T (4x4)."""
        T, _ = reg.register_cloud(None, np.eye(4, dtype=np.float32), GN_ITERS, IT_WEIGHT, EPSILON, res,
                                  keep_on_device=True)
        return T

    def step_device(k, dev_cloud):
        f = frames[k]
        if args.update_only:
            pos, up = fp.convert_pose_to_gpu(f["pose"], res)
            tsdf.update_tsdf_device(dev_cloud.data_ptr(), counts[k], pos, up)
            return
        # one fused call per scan: 20 GN iterations -> pose -> update_tsdf, chained on the device
        reg.track_scan(None, priors[k], GN_ITERS, IT_WEIGHT, EPSILON, res, device_ptr=dev_ptrs[k], n=counts[k])

    def step_host(k, host_cloud):
        f = frames[k]
        if args.update_only:
            pos, up = fp.convert_pose_to_gpu(f["pose"], res)
            tsdf.update_tsdf(host_cloud, pos, up)
            return
        # H2D of the scan, D2H of the transform + pose + work counters, all inside the call
        reg.track_scan(host_cloud, priors[k], GN_ITERS, IT_WEIGHT, EPSILON, res)

    key = "points_map" if args.update_only else "points_prior"
    with torch.cuda.stream(stream):
        # seed the map with frame 0 (untimed)
        pos, up = fp.convert_pose_to_gpu(frames[0]["pose"], res)
        tsdf.update_tsdf(frames[0]["points_map"], pos, up)

        dev_clouds = [None] + [torch.from_numpy(frames[k][key]).to("cuda") for k in range(1, K + W + 1)]
        pinned = [None] + [torch.from_numpy(frames[k][key]).pin_memory() for k in range(1, K + W + 1)]
        host_clouds = [None] + [p.numpy() for p in pinned[1:]]
        # step inputs prepared outside the timed region: the odometry prior of every scan in the ABI layout
        priors = [None] + [fp.colmajor16(s.pose(k - 1)) for k in range(1, K + W + 1)]
        dev_ptrs = [None] + [c.data_ptr() for c in dev_clouds[1:]]
        counts = [0] + [int(c.shape[0]) for c in dev_clouds[1:]]

        def barrier():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(step_fn, inputs):
            for k in range(1, W + 1):
                step_fn(k, inputs[k])
            barrier()
            tsdf.profile(True)
            tsdf.profile_reset()
            launches0 = tsdf.launch_count()
            sampler = ClockSampler(local_rank)
            sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(W + 1, W + K + 1):
                step_fn(k, inputs[k])          # every call ends with the D2H of its transform, pose and counters
            e1.record(stream)
            last = tsdf.counters()             # the last step's work counters (already on the host)
            barrier()
            clocks = sampler.stop()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            kern = {name: tsdf.profile_get(kind) for name, kind in
                    (("march", lib.TIMER_MARCH), ("merge", lib.TIMER_MERGE), ("reg", lib.TIMER_REG),
                     ("replay", lib.TIMER_REPLAY))}
            tsdf.profile(False)
            return ms, clocks, kern, tsdf.launch_count() - launches0, last

        ms_dev, clocks, kern, launches, counters = timed(step_device, dev_clouds)
        # the end-to-end pass replays the same scans (the map keeps accumulating; the march work is identical)
        ms_e2e, clocks_e2e, kern_e2e, launches_e2e, counters_e2e = timed(step_host, host_clouds)

    value = K / (ms_dev / 1000.0)
    e2e_value = K / (ms_e2e / 1000.0)
    if world > 1:
        # work counters: sum over the slabs (halo columns count twice); kernel times: slowest rank
        wk = torch.tensor([counters["n_touched"], counters["n_candidates"], counters["n_touched_bricks"],
                           counters["n_parked"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(wk)
        counters = dict(counters)
        counters["n_touched"], counters["n_candidates"], counters["n_touched_bricks"], counters["n_parked"] = \
            (int(v) for v in wk.tolist())
        km = torch.tensor([kern[k][0] for k in ("march", "merge", "reg", "replay")], dtype=torch.float64, device="cuda")
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        kern = {k: (float(v), kern[k][1]) for k, v in zip(("march", "merge", "reg", "replay"), km.tolist())}

    if rank == 0:
        # T of this rank's slab; for the roofline at N=1 it is the whole scan's T
        T_vox = counters["n_touched"]
        C_cand = counters["n_candidates"]
        march_ms = kern["march"][0] / max(1, K)
        merge_ms = kern["merge"][0] / max(1, K)
        reg_ms = kern["reg"][0] / max(1, K)
        replay_ms = kern["replay"][0] / max(1, K)
        upd_bytes = 12 * N + 8 * T_vox
        peak, peak_src = peaks()
        upd_ms = march_ms + merge_ms + replay_ms
        achieved = upd_bytes / (upd_ms / 1000.0) / 1e9 if upd_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "int32+int64", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": 12 * N, "d2h_bytes_per_step": 352 + 64,
                    "note": "ws_track_scan with the scan in pinned host memory: H2D -> 20 GN iterations -> pose on the "
                            "device -> update_tsdf -> transform + pose + counters D2H, one host synchronisation"},
            "gpu_launches": launches,
            "roofline": {
                "bound": "hbm", "kernel": "update_tsdf = march_kernel + brick_list/merge_kernel + replay_kernel (per scan)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                "traffic": NCU_TRAFFIC_BYTES_PER_SCAN if (world == 1 and not args.update_only and args.grid == 512 and args.res == 50
                                                          and args.beams == 128 and args.cols == 1024) else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of setup + march + brick_list + merge + replay, one "
                                  "launch each, ncu --set full capture of this workload (profiles/r01p_summary.md)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_scan": upd_bytes,
                "formula": "12*N + 8*T (SURVEY.md 8d), N=%d points, T=%d touched voxels, C=%d candidates" % (N, T_vox, C_cand),
                "kernel_ms_per_scan": {"march": march_ms, "merge": merge_ms, "replay": replay_ms, "reg_20_iterations": reg_ms,
                                       "step_total": ms_dev / K},
                "reg_bytes_per_scan": None,
            },
            "work": {"N": N, "C": C_cand, "T": T_vox, "touched_bricks": counters["n_touched_bricks"],
                     "parked": counters["n_parked"], "replay_rounds": counters["n_rounds"],
                     "replay_list": counters["n_list"], "record_chunks": counters["n_record_chunks"],
                     "replay_phase_us": [v / 1000.0 for v in counters["replay_phase_ns"]],
                     "iterations": GN_ITERS, "V": int(np.prod(np.array(size, np.int64)))},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, frames, s)
    tsdf.close()
    if rank == 0:
        if world == 1 and not args.no_ref_cuda:
            try:
                line["reference_cuda_same_gpu"] = reference_cuda(args, frames, s)
            except Exception as exc:   # the secondary baseline must never take the bench line down
                line["reference_cuda_same_gpu"] = {"unavailable": repr(exc)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_cuda(args, frames, s, scans=5):
    """Secondary baseline: the REFERENCE's own CUDA kernels (src/warpsense/cuda/*.cu, recompiled unmodified
    for sm_100a into oracle/_ref/libws_refcuda.so) on this same GPU, through the reference's own classes
    (blocking copies, default stream).  Not the parity target: different collision rule, sensor origin and
    bounds, and its registration kernels only look at the first 65,536 points."""
    from oracle import refcuda
    from oracle import oracle as orc
    from warpsense_b200 import fixedpoint as fp
    if not refcuda.available():
        return {"unavailable": "oracle/_ref/libws_refcuda.so not built (needs /root/reference at build time)"}
    side, res = args.grid, args.res
    r = refcuda.RefCuda((side, side, side), TAU, MAX_WEIGHT, res)
    pos, up = fp.convert_pose_to_gpu(frames[0]["pose"], res)
    r.set_points(frames[0]["points_map"])
    r.update(pos, up)
    t_upd, t_reg = [], []
    I = np.eye(4, dtype=np.float32)
    for k in range(1, min(scans, len(frames) - 1) + 1):
        f = frames[k]
        r.set_points(f["points_prior"])
        if not args.update_only:
            t0 = time.perf_counter()
            r.reg_prepare()
            T, alpha = I.copy(), 0.0
            for _ in range(GN_ITERS):
                H, g, e, c = r.reg_step(fp.colmajor16(T))
                if c > 0:
                    T, _, _ = orc.reg_solve(H.reshape(6, 6).T, g, e, c, alpha, T)
                alpha += IT_WEIGHT
            t_reg.append(time.perf_counter() - t0)
        pos, up = fp.convert_pose_to_gpu(f["pose"], res)
        r.set_points(f["points_map"])
        t0 = time.perf_counter()
        r.update(pos, up)
        t_upd.append(time.perf_counter() - t0)
    r.close()
    upd = statistics.median(t_upd)
    reg = statistics.median(t_reg) if t_reg else 0.0
    return {"value": 1.0 / (upd + reg), "unit": UNIT, "update_ms": 1000.0 * upd, "reg_20_iterations_ms": 1000.0 * reg,
            "kind": "reference CUDA kernels recompiled for sm_100a, reference classes, host buffers (e2e-like)",
            "note": "not the parity target; registration covers only the first 65,536 points; the 6x6 solve "
                    "between iterations runs on the host as in tsdf_registration.cpp:67-69", "scans": len(t_upd)}


def cpu_baseline(args, frames, s):
    """The oracle on this box's host cores: ONE full scan (update_tsdf of frame 0 single-threaded, as the
    reference launches it, + register_cloud of frame 1, 20 iterations on all cores)."""
    from oracle import oracle as orc
    from warpsense_b200 import fixedpoint as fp
    side, res = args.grid, args.res
    om = orc.LocalMap(side, side, side, TAU, 0)
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        orc.set_num_threads(os.cpu_count() or 1)
    pos, up = fp.convert_pose_to_gpu(frames[0]["pose"], res)
    t0 = time.perf_counter()
    st = orc.update_tsdf(om, frames[0]["points_map"], pos, up, TAU, MAX_WEIGHT, res)
    t_upd = time.perf_counter() - t0
    t_reg = 0.0
    if not args.update_only:
        cloud = frames[1]["points_prior"].copy()
        t0 = time.perf_counter()
        orc.register_cloud(om, cloud, np.eye(4, dtype=np.float32), GN_ITERS, IT_WEIGHT, EPSILON, res)
        t_reg = time.perf_counter() - t0
    return {"value": 1.0 / (t_upd + t_reg), "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": "1 full scan: update_tsdf of frame 0 single-threaded (%.2f s, C=%d, T=%d) + register_cloud of "
                      "frame 1, %d iterations on %d OpenMP threads (%.2f s)"
                      % (t_upd, st["n_candidates"], st["n_touched"], GN_ITERS, orc.num_threads(), t_reg),
            "update_s": t_upd, "reg_s": t_reg}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
