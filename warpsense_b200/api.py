"""Host-side mirror of the reference's class interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (paths under /root/reference):

  HostLocalMap      ~ HDF5LocalMap's in-memory part   include/map/hdf5_local_map.h
  DeviceMap         ~ cuda::DeviceMap                 include/warpsense/cuda/device_map.h:32-164
  DeviceMapMemWrapper ~ cuda::DeviceMapMemWrapper     include/warpsense/cuda/device_map_wrapper.h:10-36
  TSDFCuda          ~ cuda::TSDFCuda                  include/warpsense/cuda/update_tsdf.h:9-34
  RegistrationCuda  ~ cuda::RegistrationCuda          include/warpsense/cuda/registration.h:10-45
  TSDFMapping       ~ cuda::TSDFMapping               include/warpsense/tsdf_mapping.h:17-60
  TSDFRegistration  ~ cuda::TSDFRegistration          include/warpsense/tsdf_registration.h:17-27

Everything that touches voxels or points runs in libwarpsense_b200.so on the GPU; this file only
marshals arguments.  (The same shims exist as C++ in include/warpsense_b200.hpp.)
"""
import ctypes as C
import threading

import math

import numpy as np

from . import lib as _lib
from .fixedpoint import colmajor16, convert_pose_to_gpu, from_colmajor16, to_map
from .params import Params

_i32p = C.POINTER(C.c_int32)


def _i3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(3))


def _pts(points):
    p = np.asarray(points)
    if p.dtype != np.int32 or not p.flags.c_contiguous:
        p = np.ascontiguousarray(p, dtype=np.int32)
    return p.reshape(-1, 3)


def make_entry(value, weight):
    """TSDFEntry(value, weight).raw()  (include/map/tsdf.h:16-57)."""
    return (int(value) & 0xFFFF) | ((int(weight) & 0xFFFF) << 16)


def entry_value(raw):
    v = int(raw) & 0xFFFF
    return v - 0x10000 if v >= 0x8000 else v


def entry_weight(raw):
    w = (int(raw) >> 16) & 0xFFFF
    return w - 0x10000 if w >= 0x8000 else w


class HostLocalMap:
    """The in-memory part of HDF5LocalMap: ring array + size/pos/offset (hdf5_local_map.cpp:5-20).

    Only storage and index arithmetic live here (the data contract of cuda::DeviceMap); updates,
    lookups in bulk and shifts happen on the device."""

    def __init__(self, sx, sy, sz, default_value, default_weight=0):
        self.size = np.array([s if s % 2 == 1 else s + 1 for s in (int(sx), int(sy), int(sz))], np.int32)
        self.pos = np.zeros(3, np.int32)
        self.offset = (self.size // 2).astype(np.int32)
        self.default_entry = make_entry(default_value, default_weight)
        self.data = np.full(int(np.prod(self.size.astype(np.int64))), self.default_entry, dtype=np.uint32)

    def get_size(self):
        return self.size

    def get_pos(self):
        return self.pos

    def get_offset(self):
        return self.offset

    def get_data(self):
        return self.data

    def in_bounds(self, x, y, z):
        """hdf5_local_map.h:275-279"""
        d = np.abs(np.array([x, y, z], np.int64) - self.pos)
        return bool((d <= self.size // 2).all())

    def get_index(self, x, y, z):
        """hdf5_local_map.h:140-151 (64-bit)"""
        r = (np.array([x, y, z], np.int64) - self.pos + self.offset + self.size) % self.size
        return int((r[0] * self.size[1] + r[1]) * self.size[2] + r[2])

    def value(self, x, y, z):
        if not self.in_bounds(x, y, z):
            raise IndexError("Index out of bounds: %d; %d; %d" % (x, y, z))   # std::out_of_range
        raw = self.data[self.get_index(x, y, z)]
        return entry_value(raw), entry_weight(raw)

    def set_value(self, x, y, z, value, weight):
        if not self.in_bounds(x, y, z):
            raise IndexError("Index out of bounds: %d; %d; %d" % (x, y, z))
        self.data[self.get_index(x, y, z)] = make_entry(value, weight)


class DeviceMap:
    """cuda::DeviceMap: a non-owning view {size, offset, data, pos} of a host ring array."""

    def __init__(self, size_or_map, offset=None, data=None, pos=None):
        if offset is None:
            m = size_or_map
            self.size_, self.offset_, self.data_, self.pos_ = m.get_size(), m.get_offset(), m.get_data(), m.get_pos()
        else:
            self.size_, self.offset_, self.data_, self.pos_ = size_or_map, offset, data, pos

    def get_size(self):
        return self.size_

    def get_offset(self):
        return self.offset_

    def get_pos(self):
        return self.pos_


def slab_layout(size_x, rank, world):
    """Host-only: (own_lo, own_hi, resident brick columns) of rank's ring-x slab (ws_slab_layout)."""
    L = _lib.load()
    lo, hi = C.c_int32(), C.c_int32()
    cols = np.zeros(512, np.int32)
    n = L.ws_slab_layout(int(size_x), int(rank), int(world), C.byref(lo), C.byref(hi), cols.ctypes.data_as(_i32p), 512)
    if n < 0:
        raise ValueError("no slab for rank %d of %d with ring side %d" % (rank, world, size_x))
    return lo.value, hi.value, [int(c) for c in cols[:n]]


def shard_layout(size_x, rank, world, stripe_cols=0):
    """Host-only: (owned, resident) boolean arrays over the ring-x brick columns (ws_shard_layout)."""
    L = _lib.load()
    own, res = (C.c_uint8 * 512)(), (C.c_uint8 * 512)()
    nbx = L.ws_shard_layout(int(size_x), int(rank), int(world), int(stripe_cols), own, res, 512)
    if nbx < 0:
        raise _lib.WarpsenseError(nbx, "ws_shard_layout")
    return np.frombuffer(own, np.uint8, nbx).astype(bool), np.frombuffer(res, np.uint8, nbx).astype(bool)


class _Handle:
    """Owns one ws_handle (RAII like the reference's device wrappers; copy is not supported)."""

    def __init__(self, size, tau, max_weight, map_resolution, device=0, rank=0, world=1):
        self.L = _lib.load()
        self.h = C.c_void_p()
        s = _i3(size)
        if world == 1:
            rc = self.L.ws_create(s.ctypes.data_as(_i32p), int(tau), int(max_weight), int(map_resolution),
                                  int(device), C.byref(self.h))
        else:
            rc = self.L.ws_create_sharded(s.ctypes.data_as(_i32p), int(tau), int(max_weight), int(map_resolution),
                                          int(device), int(rank), int(world), C.byref(self.h))
        if rc != _lib.WS_OK:
            raise _lib.WarpsenseError(rc, "ws_create failed (bad arguments or no usable CUDA device)")
        self.size = s.copy()
        self.tau, self.max_weight, self.res = int(tau), int(max_weight), int(map_resolution)
        self.device, self.rank, self.world = int(device), int(rank), int(world)

    def check(self, rc):
        if rc < 0:
            raise _lib.WarpsenseError(rc, (self.L.ws_last_error(self.h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.ws_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMapMemWrapper:
    """cuda::DeviceMapMemWrapper: whole-grid host<->device transfers and the param-only update."""

    def __init__(self, handle):
        self._hd = handle

    def to_device(self, existing_map):
        """device_map_wrapper.cu:64-75"""
        d = existing_map.data_
        assert d.dtype == np.uint32 and d.flags.c_contiguous
        s, o, p = _i3(existing_map.size_), _i3(existing_map.offset_), _i3(existing_map.pos_)
        hd = self._hd
        hd.check(hd.L.ws_map_upload(hd.h, d.ctypes.data, s.ctypes.data_as(_i32p), o.ctypes.data_as(_i32p),
                                    p.ctypes.data_as(_i32p)))

    def to_host(self, existing_map):
        """device_map_wrapper.cu:77-92 (also returns the device-side pos/offset into the view)"""
        d = existing_map.data_
        assert d.dtype == np.uint32 and d.flags.c_contiguous and d.flags.writeable
        hd = self._hd
        hd.check(hd.L.ws_map_download(hd.h, d.ctypes.data))
        s, o, p = np.zeros(3, np.int32), np.zeros(3, np.int32), np.zeros(3, np.int32)
        hd.check(hd.L.ws_map_get_params(hd.h, s.ctypes.data_as(_i32p), o.ctypes.data_as(_i32p), p.ctypes.data_as(_i32p)))
        existing_map.offset_[:] = o
        existing_map.pos_[:] = p

    def update_params(self, existing_map):
        """device_map_wrapper.cu:35-43"""
        o, p = _i3(existing_map.offset_), _i3(existing_map.pos_)
        hd = self._hd
        hd.check(hd.L.ws_map_set_params(hd.h, o.ctypes.data_as(_i32p), p.ctypes.data_as(_i32p)))

    def params(self):
        s, o, p = np.zeros(3, np.int32), np.zeros(3, np.int32), np.zeros(3, np.int32)
        hd = self._hd
        hd.check(hd.L.ws_map_get_params(hd.h, s.ctypes.data_as(_i32p), o.ctypes.data_as(_i32p), p.ctypes.data_as(_i32p)))
        return s, o, p


class TSDFCuda:
    """cuda::TSDFCuda(existing_map, tau, max_weight, map_resolution) -- update_tsdf.h:9-34."""

    n_max_points_ = _lib.WS_MAX_POINTS

    def __init__(self, existing_map, tau, max_weight, map_resolution, device=0, rank=0, world=1, upload=True):
        self._hd = _Handle(existing_map.size_, tau, max_weight, map_resolution, device, rank, world)
        self._avg = DeviceMapMemWrapper(self._hd)
        if upload:
            self._avg.to_device(existing_map)
        else:
            self._avg.update_params(existing_map)

    # -- reference surface --
    def update_tsdf(self, scan_points, scanner_pos, up):
        """update_tsdf.cu:169-183: ray-march the scan into the device map.  Blocking."""
        p = _pts(scan_points)
        sp, u = _i3(scanner_pos), _i3(up)
        hd = self._hd
        hd.check(hd.L.ws_update_tsdf(hd.h, p.ctypes.data, len(p), sp.ctypes.data_as(_i32p), u.ctypes.data_as(_i32p)))

    def update_tsdf_result(self, result, scan_points, scanner_pos, up):
        """update_tsdf.cu:143-166 overload: update, then copy the map into `result` (a DeviceMap view)."""
        self.update_tsdf(scan_points, scanner_pos, up)
        self._avg.to_host(result)

    def update_tsdf_device(self, device_ptr, n, scanner_pos, up):
        sp, u = _i3(scanner_pos), _i3(up)
        hd = self._hd
        hd.check(hd.L.ws_update_tsdf_device(hd.h, C.c_void_p(int(device_ptr)), int(n), sp.ctypes.data_as(_i32p),
                                            u.ctypes.data_as(_i32p)))

    def preprocess_scan(self, cloud_xyz_m, pose_mm, map_resolution, point_step_bytes=None, fetch=True, cpu_node=False):
        """App::preprocess (app.cpp:118-148) on the device: float metres [n, >=3] -> unique int32 mm points in
        the map frame, scan order (first occurrence stays).  The points stay on the device (scan_points_device)
        and are returned as an array when `fetch`.  cpu_node=True: the CPU node's variant (fastsense.cpp:143-163)."""
        a = np.ascontiguousarray(cloud_xyz_m, dtype=np.float32)
        a = a.reshape(-1, 3) if a.ndim == 1 else a
        step = int(point_step_bytes or a.shape[1] * 4)
        n = a.shape[0]
        out = np.zeros((max(n, 1), 3), np.int32) if fetch else None
        n_out = C.c_int64()
        hd = self._hd
        T = colmajor16(pose_mm)
        hd.check(hd.L.ws_preprocess_scan(hd.h, a.ctypes.data, n, step, 2 if cpu_node else 0, T.ctypes.data_as(C.POINTER(C.c_float)),
                                         int(map_resolution), out.ctypes.data if fetch else None, C.byref(n_out)))
        return (out[:n_out.value].copy() if fetch else None), n_out.value

    def voxelgrid_subsample(self, cloud_xyz_m, leaf_m, fetch=True, want_xyz=False, device_ptr=None, n=None, point_step_bytes=12):
        """pcl::VoxelGrid with a cubic leaf + metres -> int millimetres on the device (ws_voxelgrid_subsample,
        tsdf_mapping.cpp:147-158).  The points stay on the device (scan_points_device); returned as arrays when
        `fetch`: (int32 [m,3] mm[, float32 [m,3] centroids]), m."""
        hd = self._hd
        if device_ptr is None:
            a = np.ascontiguousarray(cloud_xyz_m, dtype=np.float32)
            a = a.reshape(-1, 3) if a.ndim == 1 else a
            ptr, cnt, step, dev = a.ctypes.data, a.shape[0], a.shape[1] * 4, 0
        else:
            ptr, cnt, step, dev = C.c_void_p(int(device_ptr)), int(n), int(point_step_bytes), 1
        out_mm = np.zeros((max(cnt, 1), 3), np.int32) if fetch else None
        out_xyz = np.zeros((max(cnt, 1), 3), np.float32) if (fetch and want_xyz) else None
        n_out = C.c_int64()
        hd.check(hd.L.ws_voxelgrid_subsample(hd.h, ptr, cnt, step, dev, float(leaf_m),
                                             out_mm.ctypes.data if out_mm is not None else None,
                                             out_xyz.ctypes.data if out_xyz is not None else None, C.byref(n_out)))
        m = n_out.value
        if not fetch:
            return None, m
        if want_xyz:
            return out_mm[:m].copy(), out_xyz[:m].copy(), m
        return out_mm[:m].copy(), m

    def update_tsdf_from_ros(self, cloud_xyz_m, pose_m, device_ptr=None, n=None, point_step_bytes=12):
        """TSDFMapping::update_tsdf_from_ros (tsdf_mapping.cpp:165-173) in one call: voxel grid at the map
        resolution, metres -> millimetres, pose (4x4, metres) -> mm Matrix4f, update_tsdf.  Returns (mm_pose, n_points)."""
        hd = self._hd
        if device_ptr is None:
            a = np.ascontiguousarray(cloud_xyz_m, dtype=np.float32)
            a = a.reshape(-1, 3) if a.ndim == 1 else a
            ptr, cnt, step, dev = a.ctypes.data, a.shape[0], a.shape[1] * 4, 0
        else:
            ptr, cnt, step, dev = C.c_void_p(int(device_ptr)), int(n), int(point_step_bytes), 1
        P = np.ascontiguousarray(np.asarray(pose_m, dtype=np.float64).reshape(4, 4).T).reshape(16)
        out = np.zeros(16, np.float32)
        n_out = C.c_int64()
        hd.check(hd.L.ws_update_tsdf_from_ros(hd.h, ptr, cnt, step, dev, P.ctypes.data_as(C.POINTER(C.c_double)),
                                              out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n_out)))
        return from_colmajor16(out), n_out.value

    def scan_points_device(self):
        n = C.c_int64()
        ptr = self._hd.L.ws_scan_points_device(self._hd.h, C.byref(n))
        return ptr, n.value

    def owned_rows(self):
        """Boolean mask over the ring-x rows this handle owns (contiguous slab or round-robin stripes)."""
        hd = self._hd
        own = (C.c_uint8 * 512)()
        nbx = hd.check(hd.L.ws_shard_columns(hd.h, own, None, 512))
        size_x = int(self._avg.params()[0][0])
        rows = np.repeat(np.frombuffer(own, np.uint8, nbx).astype(bool), 8)[:size_x]
        return rows

    def checksum(self, x_lo=0, x_hi=None, owned_only=False):
        """Order-free 64-bit checksum of the entries of ring-x rows [x_lo, x_hi), computed on the device
        (equal for any sharding of equal map contents)."""
        hd = self._hd
        if x_hi is None:
            x_hi = int(hd.size[0])
        out = C.c_uint64()
        hd.check(hd.L.ws_map_checksum(hd.h, int(x_lo), int(x_hi), 1 if owned_only else 0, C.byref(out)))
        return int(out.value)

    def avg_map(self):
        return self._avg

    def new_map(self):
        # the reference keeps a dense scratch grid (new_map_); here scratch keys live beside the map and
        # only pos/offset matter to callers (tsdf_mapping.cpp:123)
        return self._avg

    def device_map(self):
        return self._hd

    # -- extras --
    def counters(self):
        c = _lib.UpdateCounters()
        self._hd.check(self._hd.L.ws_get_update_counters(self._hd.h, C.byref(c)))
        return c.as_dict()

    def voxel(self, x, y, z):
        e = C.c_uint32()
        self._hd.check(self._hd.L.ws_map_get_voxel(self._hd.h, int(x), int(y), int(z), C.byref(e)))
        return entry_value(e.value), entry_weight(e.value)

    def set_voxel(self, x, y, z, value, weight):
        self._hd.check(self._hd.L.ws_map_set_voxel(self._hd.h, int(x), int(y), int(z), make_entry(value, weight)))

    def shift(self, new_pos):
        p = _i3(new_pos)
        self._hd.check(self._hd.L.ws_shift(self._hd.h, p.ctypes.data_as(_i32p)))

    def write_back(self):
        self._hd.check(self._hd.L.ws_write_back(self._hd.h))

    def chunk_list(self):
        hd = self._hd
        n = int(hd.L.ws_store_num_chunks(hd.h))
        out = np.zeros((max(n, 1), 3), np.int32)
        k = hd.check(hd.L.ws_store_chunk_list(hd.h, out.ctypes.data_as(_i32p), n))
        return [tuple(int(v) for v in out[i]) for i in range(k)]

    def chunk(self, cx, cy, cz):
        out = np.zeros(64 ** 3, np.uint32)
        hd = self._hd
        rc = hd.L.ws_store_get_chunk(hd.h, int(cx), int(cy), int(cz), out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out if rc == _lib.WS_OK else None

    def export_hdf5(self, path, tau, map_size, max_distance, map_resolution, max_weight, poses7=None):
        """HDF5GlobalMap's file (hdf5_global_map.cpp): stored chunks, /map attributes, /poses/<n>/pose.
        `poses7`: [n,7] float32 rows x y z qx qy qz qw as write_pose rounds them.  write_back() first to
        include the local map."""
        meta = _lib.MapMeta(int(tau), (C.c_int32 * 3)(*[int(v) for v in map_size]), float(max_distance),
                            int(map_resolution), int(max_weight))
        p = np.zeros((0, 7), np.float32) if poses7 is None else np.ascontiguousarray(poses7, np.float32).reshape(-1, 7)
        self._hd.check(self._hd.L.ws_export_hdf5(self._hd.h, str(path).encode(), C.byref(meta),
                                                 p.ctypes.data_as(C.POINTER(C.c_float)), len(p)))

    def import_hdf5(self, path, max_poses=4096):
        """Read a global-map file back into the chunk store (ws_import_hdf5).  Returns (meta dict, poses [n,7], n_chunks);
        `reload()` then brings the stored chunks into the device-resident local map."""
        meta = _lib.MapMeta()
        poses = np.zeros((max_poses, 7), np.float32)
        nc, npz = C.c_int64(), C.c_int64()
        self._hd.check(self._hd.L.ws_import_hdf5(self._hd.h, str(path).encode(), C.byref(meta),
                                                 poses.ctypes.data_as(C.POINTER(C.c_float)), max_poses, C.byref(nc), C.byref(npz)))
        m = {"tau": meta.tau, "map_size": tuple(meta.map_size), "max_distance": meta.max_distance,
             "map_resolution": meta.map_resolution, "max_weight": meta.max_weight}
        return m, poses[:min(npz.value, max_poses)].copy(), nc.value

    def reload(self):
        self._hd.check(self._hd.L.ws_map_reload(self._hd.h))

    def store_configure(self, max_chunks_in_memory):
        self._hd.check(self._hd.L.ws_store_configure(self._hd.h, int(max_chunks_in_memory)))

    def store_evictions(self):
        return int(self._hd.L.ws_store_evictions(self._hd.h))

    def set_chunk(self, cx, cy, cz, data):
        a = np.ascontiguousarray(data, np.uint32).reshape(64 ** 3)
        self._hd.check(self._hd.L.ws_store_set_chunk(self._hd.h, int(cx), int(cy), int(cz), a.ctypes.data_as(C.POINTER(C.c_uint32))))

    def sync(self):
        self._hd.check(self._hd.L.ws_sync(self._hd.h))

    def set_stream(self, cuda_stream):
        self._hd.check(self._hd.L.ws_set_stream(self._hd.h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    def launch_count(self):
        return int(self._hd.L.ws_launch_count(self._hd.h))

    def profile(self, on=True):
        """False / 0: off; True / 1: every range (about 20 event records per scan); 2: the registration loop and the
        whole update only (what a timed run can afford)."""
        level = int(on) if not isinstance(on, bool) else (1 if on else 0)
        self._hd.check(self._hd.L.ws_profile_enable(self._hd.h, level))

    def profile_reset(self):
        self._hd.check(self._hd.L.ws_profile_reset(self._hd.h))

    def profile_get(self, kind):
        ms, n = C.c_double(), C.c_int64()
        self._hd.check(self._hd.L.ws_profile_get(self._hd.h, int(kind), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_timeline(self, cap=64):
        """(kind, start_ms, stop_ms) of every timed range of the last update_tsdf, relative to its start."""
        buf = (C.c_double * (3 * cap))()
        n = self._hd.check(self._hd.L.ws_profile_timeline(self._hd.h, buf, cap))
        return [(int(buf[3 * i]), buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]

    def close(self):
        self._hd.close()


class RegistrationCuda:
    """cuda::RegistrationCuda(map) -- registration.h:10-45.  Shares the device map of a TSDFCuda."""

    def __init__(self, tsdf):
        self._hd = tsdf.device_map() if isinstance(tsdf, TSDFCuda) else tsdf
        self.curr_n_points = 0

    def prepare_registration(self, points):
        """registration.cu:303-308"""
        p = _pts(points)
        hd = self._hd
        hd.check(hd.L.ws_reg_prepare(hd.h, p.ctypes.data, len(p)))
        self.curr_n_points = len(p)

    def prepare_registration_device(self, device_ptr, n):
        hd = self._hd
        hd.check(hd.L.ws_reg_prepare_device(hd.h, C.c_void_p(int(device_ptr)), int(n)))
        self.curr_n_points = int(n)

    def perform_registration(self, pretransform, map_resolution):
        """registration.cu:347-368 -> (H 6x6 int64, g int64[6], e, c) for transform `pretransform`."""
        T = colmajor16(pretransform)
        H = np.zeros(36, np.int64)
        g = np.zeros(6, np.int64)
        e, c = C.c_int32(), C.c_int32()
        hd = self._hd
        hd.check(hd.L.ws_reg_step(hd.h, T.ctypes.data_as(C.POINTER(C.c_float)), int(map_resolution),
                                  H.ctypes.data_as(C.POINTER(C.c_int64)), g.ctypes.data_as(C.POINTER(C.c_int64)),
                                  C.byref(e), C.byref(c)))
        return H.reshape(6, 6).T.copy(), g, e.value, c.value

    def register_cloud(self, cloud, pretransform, max_iterations, it_weight_gradient, epsilon, map_resolution,
                       host_solve=False, keep_on_device=False):
        """The whole Gauss-Newton loop (src/cpu/registration.cpp:14-177).  `cloud` (int32 [n,3]) is
        transformed IN PLACE; returns (total_transform 4x4 float32, iterations)."""
        T = colmajor16(pretransform)
        out = np.zeros(16, np.float32)
        it = C.c_int32()
        hd = self._hd
        if keep_on_device:
            ptr, n = None, self.curr_n_points
        else:
            assert cloud.dtype == np.int32 and cloud.flags.c_contiguous and cloud.flags.writeable
            ptr, n = cloud.ctypes.data, len(cloud)
            self.curr_n_points = n
        hd.check(hd.L.ws_register_cloud(hd.h, ptr, int(n), T.ctypes.data_as(C.POINTER(C.c_float)),
                                        int(max_iterations), float(it_weight_gradient), float(epsilon),
                                        int(map_resolution),
                                        _lib.WS_REG_HOST_SOLVE if host_solve else _lib.WS_REG_DEVICE_SOLVE,
                                        out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(it)))
        return from_colmajor16(out), it.value

    def _track_bufs(self):
        buf = getattr(self, "_track_buf", None)
        if buf is None:                              # marshalling buffers, reused from call to call
            f32p = C.POINTER(C.c_float)
            buf = self._track_buf = (np.zeros(16, np.float32), np.zeros(16, np.float32), C.c_int32(), C.c_int32())
            self._track_ptrs = (buf[0].ctypes.data_as(f32p), buf[1].ctypes.data_as(f32p), C.byref(buf[2]), C.byref(buf[3]))
        return buf, self._track_ptrs

    @staticmethod
    def _mat16(m):
        if m is None:
            return None
        if isinstance(m, np.ndarray) and m.dtype == np.float32 and m.shape == (16,) and m.flags.c_contiguous:
            return m
        return colmajor16(m)

    def track_submit(self, points, prior_pose, max_iterations, it_weight_gradient, epsilon, map_resolution,
                     device_ptr=None, n=None, pretransform=None, reference_pose=False):
        """Enqueue one scan of the per-scan pipeline (ws_track_submit): register against the map, pose update,
        update_tsdf with the registered cloud.  Returns a ticket for track_wait; at most two scans in flight.
        `points`: int32 [n,3] host array (must stay alive until track_wait), or None with `device_ptr`/`n`.
        `prior_pose`: 4x4 (or float32[16] column-major); None = the pose of the previous tracked scan (on the device).
        `reference_pose`: compose the pose as App::update_pose_estimate (app.cpp:172-176) instead of X @ prior."""
        hd = self._hd
        f32p = C.POINTER(C.c_float)
        pp, pre = self._mat16(prior_pose), self._mat16(pretransform)
        _, ptrs = self._track_bufs()
        if points is not None:
            p = _pts(points)
            ptr, cnt, dev = p.ctypes.data, len(p), 0
            self._track_keep = getattr(self, "_track_keep", {})
        else:
            p = None
            ptr, cnt, dev = C.c_void_p(int(device_ptr)), int(n), 1
        hd.check(hd.L.ws_track_submit(hd.h, ptr, cnt, dev, pp.ctypes.data_as(f32p) if pp is not None else None,
                                      pre.ctypes.data_as(f32p) if pre is not None else None, int(max_iterations),
                                      float(it_weight_gradient), float(epsilon), int(map_resolution),
                                      _lib.WS_TRACK_REFERENCE_POSE if reference_pose else 0, ptrs[3]))
        ticket = self._track_buf[3].value
        if p is not None:
            self._track_keep[ticket] = p             # the copy is asynchronous: keep the array alive
        self.curr_n_points = cnt
        return ticket

    def track_wait(self, ticket):
        """Collect a submitted scan: (X, pose, iterations)."""
        hd = self._hd
        buf, ptrs = self._track_bufs()
        try:
            hd.check(hd.L.ws_track_wait(hd.h, int(ticket), ptrs[0], ptrs[1], ptrs[2]))
        finally:
            getattr(self, "_track_keep", {}).pop(ticket, None)
        return from_colmajor16(buf[0]), from_colmajor16(buf[1]), buf[2].value

    def track_scan(self, points, prior_pose, max_iterations, it_weight_gradient, epsilon, map_resolution,
                   device_ptr=None, n=None, pretransform=None, reference_pose=False):
        """One scan through the per-scan pipeline, blocking (ws_track_scan_ex = submit + wait).
        Returns (X, pose, iterations)."""
        t = self.track_submit(points, prior_pose, max_iterations, it_weight_gradient, epsilon, map_resolution,
                              device_ptr=device_ptr, n=n, pretransform=pretransform, reference_pose=reference_pose)
        return self.track_wait(t)

    # -- multi-GPU: one handle per rank, the caller supplies the exchange step --
    def sums_device_ptr(self):
        """Device address of this rank's int64[29] Gauss-Newton sums (what the all-reduce operates on)."""
        return int(self._hd.L.ws_reg_sums_device(self._hd.h))

    def sums_get(self):
        out = np.zeros(29, np.int64)
        self._hd.check(self._hd.L.ws_reg_sums_get(self._hd.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def sums_set(self, sums):
        a = np.ascontiguousarray(sums, dtype=np.int64).reshape(29)
        self._hd.check(self._hd.L.ws_reg_sums_set(self._hd.h, a.ctypes.data_as(C.POINTER(C.c_int64))))

    # -- fused multi-GPU registration: in-kernel exchange over NVLink peer memory (registration.cu) --
    def peer_export(self):
        """64-byte IPC handle of this rank's mailbox (bytes); all-gather it and pass the list to peer_attach_ipc."""
        buf = (C.c_uint8 * 64)()
        self._hd.check(self._hd.L.ws_peer_export(self._hd.h, buf))
        return bytes(buf)

    def peer_attach_ipc(self, handles):
        """handles: the peer_export() of every rank, in rank order (other processes, one GPU each)."""
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._hd.check(self._hd.L.ws_peer_attach_ipc(self._hd.h, arr, len(handles)))

    def peer_local_ptr(self):
        return int(self._hd.L.ws_peer_local_ptr(self._hd.h) or 0)

    def peer_attach_ptrs(self, ptrs, devices=None):
        """Ranks that live in ONE process: raw mailbox pointers (peer_local_ptr of every rank, rank order)."""
        a = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
        d = None if devices is None else np.ascontiguousarray(devices, np.int32).ctypes.data_as(_i32p)
        self._hd.check(self._hd.L.ws_peer_attach_ptrs(self._hd.h, a, d, len(ptrs)))

    def peer_set_timeout(self, seconds):
        self._hd.check(self._hd.L.ws_peer_set_timeout(self._hd.h, float(seconds)))

    def register_cloud_sharded(self, pretransform, max_iterations, it_weight_gradient, epsilon, map_resolution,
                               allreduce, check_every=8):
        """register_cloud over an x-slab sharded map: every rank calls this with the same cloud (staged with
        prepare_registration*) and an `allreduce(self)` that sums the int64[29] at sums_device_ptr() over all
        ranks (e.g. torch.distributed.all_reduce on a view of it).  Integer sums: bit-identical on every
        rank for any rank count, so each rank runs the identical solve.  Returns (T, iterations)."""
        hd = self._hd
        f32p = C.POINTER(C.c_float)
        T0 = colmajor16(pretransform)
        hd.check(hd.L.ws_reg_begin(hd.h, T0.ctypes.data_as(f32p)))
        it, fin = C.c_int32(), C.c_int32()
        for i in range(int(max_iterations)):
            hd.check(hd.L.ws_reg_accumulate(hd.h, int(map_resolution)))
            allreduce(self)
            hd.check(hd.L.ws_reg_solve(hd.h, float(it_weight_gradient), float(epsilon)))
            if epsilon > 0 and (i + 1) % check_every == 0 and i + 1 < max_iterations:
                hd.check(hd.L.ws_reg_peek(hd.h, C.byref(it), C.byref(fin)))
                if fin.value:
                    break
        out = np.zeros(16, np.float32)
        hd.check(hd.L.ws_reg_finish(hd.h, out.ctypes.data_as(f32p), C.byref(it), C.byref(fin)))
        return from_colmajor16(out), it.value

    def trace(self, max_iterations=256):
        out = np.zeros((max_iterations, 29), np.int64)
        hd = self._hd
        n = hd.check(hd.L.ws_reg_get_trace(hd.h, out.ctypes.data_as(C.POINTER(C.c_int64)), int(max_iterations)))
        return out[:n].copy()

    def points_device(self):
        n = C.c_int64()
        ptr = self._hd.L.ws_reg_points_device(self._hd.h, C.byref(n))
        return ptr, n.value

    def test_reduce(self, jacobis, values):
        """Reduction-shape hook (test/cuda.cpp:416-532)."""
        J = np.ascontiguousarray(np.asarray(jacobis, np.int64).reshape(-1, 6))
        v = np.ascontiguousarray(np.asarray(values, np.int32).reshape(-1))
        H = np.zeros(36, np.int64)
        g = np.zeros(6, np.int64)
        e, c = C.c_int32(), C.c_int32()
        hd = self._hd
        hd.check(hd.L.ws_test_reduce(hd.h, J.ctypes.data_as(C.POINTER(C.c_int64)), v.ctypes.data_as(_i32p), len(J),
                                     H.ctypes.data_as(C.POINTER(C.c_int64)), g.ctypes.data_as(C.POINTER(C.c_int64)),
                                     C.byref(e), C.byref(c)))
        return H.reshape(6, 6).T.copy(), g, e.value, c.value


def mm_pose_from_isometry(pose_m):
    """tsdf_mapping.cpp:160-161: Eigen::Quaterniond(pose.rotation()).toRotationMatrix().cast<float>() and
    translation * 1000 -- Isometry3d (4x4, metres) -> Matrix4f (millimetres)."""
    m = np.asarray(pose_m, dtype=np.float64).reshape(4, 4)
    q = np.zeros(4)                      # x y z w (Eigen QuaternionBase::operator=(MatrixBase))
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0.0:
        t = math.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t
        q[1] = (m[0, 2] - m[2, 0]) * t
        q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    tx, ty, tz = 2.0 * q[0], 2.0 * q[1], 2.0 * q[2]
    twx, twy, twz = tx * q[3], ty * q[3], tz * q[3]
    txx, txy, txz = tx * q[0], ty * q[0], tz * q[0]
    tyy, tyz, tzz = ty * q[1], tz * q[1], tz * q[2]
    out = np.eye(4, dtype=np.float32)
    R = np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                  [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                  [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])
    out[:3, :3] = R.astype(np.float32)
    out[:3, 3] = (m[:3, 3] * 1000.0).astype(np.float32)
    return out


class TSDFMapping:
    """cuda::TSDFMapping(params, local_map) -- tsdf_mapping.h:17-60, tsdf_mapping.cpp:13-205 (no ROS).

    The reference serialises callers with a shared_mutex (update = exclusive, registration = shared);
    a plain lock gives the same exclusion here.  The map-shift thread becomes an explicit, synchronous
    `map_shift(pose)` that runs the shift ON THE DEVICE (ws_shift) instead of a D2H/H2D round trip."""

    def __init__(self, params: Params, local_map: HostLocalMap, device=0, rank=0, world=1):
        self.params_ = params
        self.hdf5_local_map_ = local_map
        self.cuda_map_ = DeviceMap(local_map)
        self.tsdf_ = TSDFCuda(self.cuda_map_, params.map.tau, params.map.max_weight, params.map.resolution,
                              device=device, rank=rank, world=world)
        self.mutex_ = threading.RLock()
        self.shifted_ = False
        self.is_shifting_ = False
        self._last_shift_pose = np.eye(4, dtype=np.float32)

    def tsdf(self):
        return self.tsdf_

    def shifted(self):
        return self.shifted_

    def is_shifting(self):
        return self.is_shifting_

    def join_mapping_thread(self):
        pass

    def convert_pose_to_gpu(self, pose):
        """tsdf_mapping.cpp:77-85"""
        return convert_pose_to_gpu(pose, self.params_.map.resolution)

    def update_tsdf(self, scan_points, pose_or_pos, up=None):
        """update_tsdf(points, pose) / update_tsdf(points, pos, up) -- tsdf_mapping.cpp:62-75."""
        if up is None:
            pos, up = self.convert_pose_to_gpu(pose_or_pos)
        else:
            pos = pose_or_pos
        with self.mutex_:
            self.tsdf_.update_tsdf(scan_points, pos, up)

    def preprocess_from_ros(self, cloud_xyz_m, pose):
        """tsdf_mapping.cpp:145-163 without PCL: voxel-grid subsample at the map resolution ON THE DEVICE
        (ws_voxelgrid_subsample: one float centroid per occupied leaf in pcl::VoxelGrid's leaf order), metres ->
        int millimetres, pose -> mm Matrix4f (through Eigen::Quaterniond, :160-161)."""
        res_m = np.float32(self.params_.map.resolution) / np.float32(1000.0)
        with self.mutex_:
            points_rm, _ = self.tsdf_.voxelgrid_subsample(cloud_xyz_m, float(res_m))
        return points_rm, mm_pose_from_isometry(pose)

    def subsample(self, cloud_xyz_m, resolution_m):
        """pcl::VoxelGrid with a cubic leaf on the device: float32 [m,3] centroids in PCL's leaf order."""
        with self.mutex_:
            _, xyz, _ = self.tsdf_.voxelgrid_subsample(cloud_xyz_m, float(resolution_m), want_xyz=True)
        return xyz

    def update_tsdf_from_ros(self, cloud_xyz_m, pose):
        """tsdf_mapping.cpp:165-173: cloud in metres (sensor points already in the map frame), pose in
        metres (Isometry3d as a 4x4)."""
        points_rm, mm_pose = self.preprocess_from_ros(cloud_xyz_m, pose)
        self.map_shift(mm_pose)
        self.update_tsdf(points_rm, mm_pose)
        return points_rm, mm_pose

    def map_shift(self, current_pose):
        """tsdf_mapping.cpp:97-136, one turn of the loop body, synchronous."""
        d = (self._last_shift_pose[:3, 3] / 1000.0) - (np.asarray(current_pose, np.float32)[:3, 3] / 1000.0)
        if float(np.linalg.norm(d)) < self.params_.map.shift:
            return False
        self.is_shifting_ = True
        self._last_shift_pose = np.array(current_pose, dtype=np.float32)
        pos = to_map(current_pose, self.params_.map.resolution)
        with self.mutex_:
            self.tsdf_.shift(pos)
            _, o, p = self.tsdf_.avg_map().params()
            self.hdf5_local_map_.offset[:] = o
            self.hdf5_local_map_.pos[:] = p
        self.shifted_ = True
        self.is_shifting_ = False
        return True

    def get_tsdf_map(self):
        """tsdf_mapping.cpp:138-143: bring the device map back into the host local map."""
        with self.mutex_:
            self.tsdf_.avg_map().to_host(self.cuda_map_)
        return self.hdf5_local_map_

    def close(self):
        self.tsdf_.close()


class TSDFRegistration(TSDFMapping):
    """cuda::TSDFRegistration -- tsdf_registration.h:17-27, tsdf_registration.cpp:29-96."""

    def __init__(self, params: Params, local_map: HostLocalMap, device=0, rank=0, world=1):
        super().__init__(params, local_map, device=device, rank=rank, world=world)
        self.reg_ = RegistrationCuda(self.tsdf_)

    def register_cloud(self, cloud, pretransform, host_solve=False):
        r = self.params_.registration
        with self.mutex_:
            T, _ = self.reg_.register_cloud(cloud, pretransform, r.max_iterations, r.it_weight_gradient, r.epsilon,
                                            self.params_.map.resolution, host_solve=host_solve)
        return T


def write_map_hdf5(path, chunks, tau, map_size, max_distance, map_resolution, max_weight, poses7=None):
    """HDF5GlobalMap's file from host data (no GPU): `chunks` maps (cx, cy, cz) -> uint32[64^3]."""
    L = _lib.load()
    meta = _lib.MapMeta(int(tau), (C.c_int32 * 3)(*[int(v) for v in map_size]), float(max_distance),
                        int(map_resolution), int(max_weight))
    keys = list(chunks.keys())
    xyz = np.ascontiguousarray(np.array(keys, np.int32).reshape(-1, 3))
    data = np.ascontiguousarray(np.stack([np.asarray(chunks[k], np.uint32).reshape(64 ** 3) for k in keys])
                                if keys else np.zeros((0, 64 ** 3), np.uint32))
    p = np.zeros((0, 7), np.float32) if poses7 is None else np.ascontiguousarray(poses7, np.float32).reshape(-1, 7)
    rc = L.ws_hdf5_write_chunks(str(path).encode(), C.byref(meta), xyz.ctypes.data_as(_i32p),
                                data.ctypes.data_as(C.POINTER(C.c_uint32)), len(keys),
                                p.ctypes.data_as(C.POINTER(C.c_float)), len(p))
    if rc != _lib.WS_OK:
        raise _lib.WarpsenseError(rc, "ws_hdf5_write_chunks failed")


class MappingFeed:
    """The TSDF-facing half of featsense's Mapping::thread_run (src/featsense/mapping.cpp:39-147) without
    ROS/PCL: F-LOAM/VGICP stay outside -- the caller hands over the cloud already placed in
    the map frame (metres) and its pose (4x4, metres).

      * the first cloud initialises the map (:66-75);
      * later clouds are used only after the sensor moved more than `map.update_distance` since the last
        used pose (:79-81);
      * while the local map is shifting the clouds are accumulated, and flushed -- concatenated and
        subsampled at the map resolution -- with the next update (:115-129);
      * every used pose is offered to the map-shift logic (:140, update_map_shift)."""

    def __init__(self, mapping: TSDFMapping):
        self.gpu_ = mapping
        self.initialized_ = False
        self.last_pose_ = np.eye(4)
        self.accumulated_ = []
        self.poses = []                      # what hdf5_global_map_->write_pose would have stored (:137)
        self.updates = 0

    def subsample(self, cloud_xyz_m, resolution_m):
        """mapping.cpp `subsample`: pcl::VoxelGrid with a cubic leaf -- one float centroid per occupied leaf
        (TSDFMapping.subsample: on the device)."""
        pts = np.asarray(cloud_xyz_m, dtype=np.float32).reshape(-1, 3)
        if not len(pts):
            return pts
        return self.gpu_.subsample(pts, resolution_m)

    def push(self, cloud_xyz_m, pose_m):
        """One (cloud, pose) pair from the odometry front end.  Returns True if the map was updated."""
        pose = np.asarray(pose_m, dtype=np.float64).reshape(4, 4)
        res_m = self.gpu_.params_.map.resolution / 1000.0
        cloud = np.asarray(cloud_xyz_m, dtype=np.float32).reshape(-1, 3)
        if not self.initialized_:
            self.last_pose_ = pose.copy()
            self.gpu_.update_tsdf_from_ros(self.subsample(cloud, res_m), pose)
            self.initialized_ = True
            self.updates += 1
            return True
        distance = float(np.linalg.norm(self.last_pose_[:3, 3] - pose[:3, 3]))
        if not distance > self.gpu_.params_.map.update_distance:
            return False
        updated = False
        if self.gpu_.is_shifting():
            self.accumulated_.append(cloud)
        else:
            if self.accumulated_:
                cloud = self.subsample(np.concatenate([cloud] + self.accumulated_), res_m)
                self.accumulated_ = []
            self.gpu_.update_tsdf_from_ros(cloud, pose)
            self.updates += 1
            updated = True
        self.last_pose_ = pose.copy()
        self.poses.append(pose.copy())
        return updated
