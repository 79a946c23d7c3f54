"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (warpsense_b200) never does.

The oracle itself is oracle/ws_oracle.c, a C restatement of the reference's CPU path
(/root/reference/src/cpu); see ws_oracle.h for the parity-pinning statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libws_oracle.so")

MATRIX_RESOLUTION = 32768
WEIGHT_RESOLUTION = 64
CHUNK_SIZE = 64


def build(force=False):
    """Compile oracle/ws_oracle.c -> oracle/libws_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "ws_oracle.c")
    hdr = os.path.join(_HERE, "ws_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libws_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class UpdateStats(C.Structure):
    _fields_ = [("n_points", C.c_int64), ("n_marched", C.c_int64), ("n_candidates", C.c_int64),
                ("n_touched", C.c_int64), ("n_written", C.c_int64), ("n_neg_candidates", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    i32p = C.POINTER(C.c_int32)
    i64p = C.POINTER(C.c_int64)
    f32p = C.POINTER(C.c_float)
    f64p = C.POINTER(C.c_double)
    u32p = C.POINTER(C.c_uint32)
    vp = C.c_void_p
    sig = {
        "orc_make_entry": (C.c_uint32, [C.c_int, C.c_int]),
        "orc_entry_value": (C.c_int, [C.c_uint32]),
        "orc_entry_weight": (C.c_int, [C.c_uint32]),
        "orc_map_create": (vp, [C.c_int] * 5),
        "orc_map_destroy": (None, [vp]),
        "orc_map_clone": (vp, [vp]),
        "orc_map_get_size": (None, [vp, i32p]),
        "orc_map_get_pos": (None, [vp, i32p]),
        "orc_map_get_offset": (None, [vp, i32p]),
        "orc_map_data": (u32p, [vp]),
        "orc_map_num_voxels": (C.c_int64, [vp]),
        "orc_map_in_bounds": (C.c_int, [vp, C.c_int, C.c_int, C.c_int]),
        "orc_map_index": (C.c_int64, [vp, C.c_int, C.c_int, C.c_int]),
        "orc_map_get": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, u32p]),
        "orc_map_set": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_uint32]),
        "orc_map_set_state": (None, [vp, i32p, i32p]),
        "orc_map_shift": (C.c_int, [vp, i32p]),
        "orc_map_write_back": (None, [vp]),
        "orc_store_num_chunks": (C.c_int64, [vp]),
        "orc_store_chunk_list": (C.c_int, [vp, i32p, C.c_int64]),
        "orc_store_chunk": (u32p, [vp, C.c_int, C.c_int, C.c_int]),
        "orc_store_get_value": (C.c_uint32, [vp, C.c_int, C.c_int, C.c_int]),
        "orc_to_int_mat": (None, [f32p, i32p]),
        "orc_transform_points": (None, [vp, C.c_int64, i32p, vp]),
        "orc_preprocess": (C.c_int64, [vp, C.c_int64, C.c_int, f32p, C.c_int, vp]),
        "orc_preprocess_cpu_node": (C.c_int64, [vp, C.c_int64, C.c_int, f32p, C.c_int, vp]),
        "orc_to_map": (None, [f32p, C.c_int, i32p]),
        "orc_convert_pose": (None, [f32p, C.c_int, i32p, i32p]),
        "orc_transform_point_cloud": (None, [vp, C.c_int64, f32p]),
        "orc_scale_params": (None, [C.c_float, C.c_int, f32p, C.c_int, i32p, i32p, i32p]),
        "orc_calc_weight": (C.c_int, [C.c_int, C.c_int, C.c_int]),
        "orc_update_tsdf": (None, [vp, vp, C.c_int64, i32p, i32p, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(UpdateStats)]),
        "orc_update_tsdf_omp": (None, [vp, vp, C.c_int64, i32p, i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(UpdateStats)]),
        "orc_reg_step": (None, [vp, vp, C.c_int64, f32p, C.c_int, i64p, i64p, i32p, i32p]),
        "orc_reg_solve": (C.c_float, [i64p, i64p, C.c_int32, C.c_int32, C.c_float, f32p, f64p]),
        "orc_xi_to_transform": (None, [f64p, i32p, f32p]),
        "orc_register_cloud": (C.c_int, [vp, vp, C.c_int64, f32p, C.c_int, C.c_float, C.c_float,
                                         C.c_int, f32p, i64p, C.c_int]),
        "orc_jacobi_2_h": (None, [i64p, i64p]),
        "orc_num_threads": (C.c_int, []),
        "orc_set_num_threads": (None, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(3))


def _colmajor(m):
    """4x4 (row/col indexable) numpy matrix -> column-major float32[16] (Eigen storage)."""
    m = np.asarray(m, dtype=np.float32).reshape(4, 4)
    return np.ascontiguousarray(m.T).reshape(16)


def _from_colmajor(a):
    return np.array(a, dtype=np.float32).reshape(4, 4).T.copy()


def _points(pts):
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.int32).reshape(-1, 3))
    return pts


def make_entry(value, weight):
    return int(lib().orc_make_entry(int(value), int(weight)))


def to_int_mat(m):
    out = np.zeros(16, np.int32)
    cm = _colmajor(m)
    lib().orc_to_int_mat(_p(cm, C.c_float), _p(out, C.c_int32))
    return out.reshape(4, 4).T.copy()


def transform_points(pts, int_mat):
    """include/util/util.h:13-18 applied to an [n,3] int32 array; int_mat is a 4x4 int32 matrix."""
    pts = _points(pts)
    m = np.ascontiguousarray(np.asarray(int_mat, dtype=np.int32).reshape(4, 4).T).reshape(16)
    out = np.zeros_like(pts)
    lib().orc_transform_points(pts.ctypes.data, len(pts), _p(m, C.c_int32), out.ctypes.data)
    return out


def preprocess(cloud_xyz_m, pose_mm, map_resolution, cpu_node=False):
    """App::preprocess (src/warpsense/app.cpp:118-148): float metres [n, >=3] -> unique int32 mm points in the
    map frame, in scan order (the reference's std::unordered_set order is implementation defined).
    cpu_node=True: the CPU node's variant (src/cpu/fastsense.cpp:143-163), whose own order is scan order."""
    a = np.ascontiguousarray(cloud_xyz_m, dtype=np.float32)
    a = a.reshape(-1, 3) if a.ndim == 1 else a
    out = np.zeros((max(len(a), 1), 3), np.int32)
    cm = _colmajor(pose_mm)
    fn = lib().orc_preprocess_cpu_node if cpu_node else lib().orc_preprocess
    n = fn(a.ctypes.data, len(a), a.shape[1], _p(cm, C.c_float), int(map_resolution), out.ctypes.data)
    return out[:n].copy()


def transform_point(p, int_mat):
    return transform_points(np.asarray(p, np.int32).reshape(1, 3), int_mat)[0]


def to_map(pose, res):
    out = np.zeros(3, np.int32)
    cm = _colmajor(pose)
    lib().orc_to_map(_p(cm, C.c_float), int(res), _p(out, C.c_int32))
    return out


def convert_pose(pose, res):
    """tsdf_mapping.cpp:77-85 -> (pos voxel, up)."""
    pos = np.zeros(3, np.int32)
    up = np.zeros(3, np.int32)
    cm = _colmajor(pose)
    lib().orc_convert_pose(_p(cm, C.c_float), int(res), _p(pos, C.c_int32), _p(up, C.c_int32))
    return pos, up


def update_pose_estimate(pose, transform):
    """App::update_pose_estimate (src/warpsense/app.cpp:172-176; the CPU node does the same,
    src/cpu/fastsense.cpp:219-221): R = T.R * R, t += T.t, in float32 (Eigen Matrix4f blocks)."""
    p = np.array(pose, dtype=np.float32, copy=True)
    t = np.asarray(transform, dtype=np.float32)
    R = np.zeros((3, 3), np.float32)
    for r in range(3):
        for c in range(3):
            acc = np.float32(t[r, 0] * p[0, c])
            acc = np.float32(acc + np.float32(t[r, 1] * p[1, c]))
            acc = np.float32(acc + np.float32(t[r, 2] * p[2, c]))
            R[r, c] = acc
    p[:3, :3] = R
    p[:3, 3] = (p[:3, 3] + t[:3, 3]).astype(np.float32)
    return p


def voxelgrid(cloud_xyz_m, leaf_m):
    """pcl::VoxelGrid + metres -> millimetres as preprocess_from_ros does (tsdf_mapping.cpp:145-159).
    Returns (int32 [m,3] millimetre points, float32 [m,3] centroids)."""
    a = np.ascontiguousarray(cloud_xyz_m, dtype=np.float32)
    a = a.reshape(-1, 3) if a.ndim == 1 else a
    n, stride = a.shape[0], a.shape[1]
    out_xyz = np.zeros((max(n, 1), 3), np.float32)
    out_mm = np.zeros((max(n, 1), 3), np.int32)
    L = lib()
    L.orc_voxelgrid.restype = C.c_int64
    m = L.orc_voxelgrid(a.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int(stride), C.c_float(leaf_m),
                        out_xyz.ctypes.data_as(C.c_void_p), out_mm.ctypes.data_as(C.c_void_p))
    return out_mm[:m].copy(), out_xyz[:m].copy()


def mm_pose_from_isometry(pose_m):
    """tsdf_mapping.cpp:160-161: Isometry3d (4x4, metres) -> Matrix4f in millimetres through Eigen::Quaterniond."""
    P = np.ascontiguousarray(np.asarray(pose_m, dtype=np.float64).reshape(4, 4).T).reshape(16)
    out = np.zeros(16, np.float32)
    lib().orc_mm_pose_from_isometry(P.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return _from_colmajor(out)


def transform_point_cloud(pts, m):
    pts = _points(pts).copy()
    cm = _colmajor(m)
    lib().orc_transform_point_cloud(pts.ctypes.data, len(pts), _p(cm, C.c_float))
    return pts


def scale_params(max_distance, max_weight, size_m, resolution):
    tau = C.c_int32()
    mw = C.c_int32()
    size = np.zeros(3, np.int32)
    s = np.asarray(size_m, np.float32)
    lib().orc_scale_params(float(max_distance), int(max_weight), _p(s, C.c_float), int(resolution),
                           C.byref(tau), C.byref(mw), _p(size, C.c_int32))
    return tau.value, mw.value, size


def calc_weight(value, tau, weight_epsilon):
    return int(lib().orc_calc_weight(int(value), int(tau), int(weight_epsilon)))


def jacobi_2_h(J):
    J = np.ascontiguousarray(np.asarray(J, np.int64).reshape(6))
    H = np.zeros(36, np.int64)
    lib().orc_jacobi_2_h(_p(J, C.c_int64), _p(H, C.c_int64))
    return H.reshape(6, 6).T.copy()


def xi_to_transform(xi, center):
    xi = np.ascontiguousarray(np.asarray(xi, np.float64).reshape(6))
    c = _i3(center)
    out = np.zeros(16, np.float32)
    lib().orc_xi_to_transform(_p(xi, C.c_double), _p(c, C.c_int32), _p(out, C.c_float))
    return _from_colmajor(out)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


class LocalMap:
    """HDF5LocalMap + in-memory HDF5GlobalMap stand-in (include/map/hdf5_local_map.h)."""

    def __init__(self, sx, sy, sz, default_value, default_weight=0, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.orc_map_create(
            int(sx), int(sy), int(sz), int(default_value), int(default_weight))

    def __del__(self):
        try:
            if self._h:
                self._L.orc_map_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def clone(self):
        return LocalMap(0, 0, 0, 0, 0, _handle=self._L.orc_map_clone(self._h))

    def _get3(self, fn):
        out = np.zeros(3, np.int32)
        fn(self._h, _p(out, C.c_int32))
        return out

    @property
    def size(self):
        return self._get3(self._L.orc_map_get_size)

    @property
    def pos(self):
        return self._get3(self._L.orc_map_get_pos)

    @property
    def offset(self):
        return self._get3(self._L.orc_map_get_offset)

    @property
    def num_voxels(self):
        return int(self._L.orc_map_num_voxels(self._h))

    @property
    def data(self):
        """uint32 raw entries, ring layout [x][y][z] (a view onto the oracle's memory)."""
        n = self.num_voxels
        ptr = self._L.orc_map_data(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def set_state(self, pos, offset):
        p, o = _i3(pos), _i3(offset)
        self._L.orc_map_set_state(self._h, _p(p, C.c_int32), _p(o, C.c_int32))

    def in_bounds(self, x, y, z):
        return bool(self._L.orc_map_in_bounds(self._h, int(x), int(y), int(z)))

    def index(self, x, y, z):
        return int(self._L.orc_map_index(self._h, int(x), int(y), int(z)))

    def value(self, x, y, z):
        """-> (value, weight); raises IndexError like std::out_of_range (hdf5_local_map.h:172-181)."""
        e = C.c_uint32()
        if self._L.orc_map_get(self._h, int(x), int(y), int(z), C.byref(e)) != 0:
            raise IndexError("Index out of bounds: %d; %d; %d" % (x, y, z))
        return self._L.orc_entry_value(e.value), self._L.orc_entry_weight(e.value)

    def set_value(self, x, y, z, value, weight):
        if self._L.orc_map_set(self._h, int(x), int(y), int(z), make_entry(value, weight)) != 0:
            raise IndexError("Index out of bounds: %d; %d; %d" % (x, y, z))

    def shift(self, new_pos):
        p = _i3(new_pos)
        if self._L.orc_map_shift(self._h, _p(p, C.c_int32)) != 0:
            raise ValueError("shift larger than the map size")

    def write_back(self):
        self._L.orc_map_write_back(self._h)

    # ---- global chunk store ----
    def chunk_list(self):
        n = int(self._L.orc_store_num_chunks(self._h))
        out = np.zeros((max(n, 1), 3), np.int32)
        k = self._L.orc_store_chunk_list(self._h, _p(out, C.c_int32), n)
        return [tuple(int(v) for v in out[i]) for i in range(k)]

    def chunk(self, cx, cy, cz):
        ptr = self._L.orc_store_chunk(self._h, int(cx), int(cy), int(cz))
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(CHUNK_SIZE ** 3,)).copy()

    def global_value(self, x, y, z):
        e = self._L.orc_store_get_value(self._h, int(x), int(y), int(z))
        return self._L.orc_entry_value(e), self._L.orc_entry_weight(e)


def update_tsdf(m, points, scanner_pos, up, tau, max_weight, map_resolution):
    """src/cpu/update_tsdf.cpp:397-564.  Returns the work counters as a dict."""
    pts = _points(points)
    sp, u = _i3(scanner_pos), _i3(up)
    st = UpdateStats()
    lib().orc_update_tsdf(m._h, pts.ctypes.data, len(pts), _p(sp, C.c_int32), _p(u, C.c_int32),
                          int(tau), int(max_weight), int(map_resolution), C.byref(st))
    return st.as_dict()


def update_tsdf_omp(m, points, scanner_pos, up, tau, max_weight, map_resolution, threads=0):
    """src/cpu/update_tsdf.cpp:566-724, the OpenMP overload (threads = 0: all cores).  For timing: with more than
    one thread its result differs from update_tsdf's wherever two threads reach the same voxel."""
    pts = _points(points)
    sp, u = _i3(scanner_pos), _i3(up)
    st = UpdateStats()
    lib().orc_update_tsdf_omp(m._h, pts.ctypes.data, len(pts), _p(sp, C.c_int32), _p(u, C.c_int32),
                              int(tau), int(max_weight), int(map_resolution), int(threads), C.byref(st))
    return st.as_dict()


def reg_step(m, points, T, map_resolution):
    """One accumulation pass (registration.cpp:52-118) -> (H 6x6 int64, g int64[6], err, cnt)."""
    pts = _points(points)
    cm = _colmajor(T)
    H = np.zeros(36, np.int64)
    g = np.zeros(6, np.int64)
    e = C.c_int32()
    c = C.c_int32()
    lib().orc_reg_step(m._h, pts.ctypes.data, len(pts), _p(cm, C.c_float), int(map_resolution),
                       _p(H, C.c_int64), _p(g, C.c_int64), C.byref(e), C.byref(c))
    return H.reshape(6, 6).T.copy(), g, e.value, c.value


def reg_solve(H, g, err, cnt, alpha, T):
    """registration.cpp:128-145 -> (new T, err/count, xi)."""
    Hc = np.ascontiguousarray(np.asarray(H, np.int64).reshape(6, 6).T).reshape(36)
    gc = np.ascontiguousarray(np.asarray(g, np.int64).reshape(6))
    cm = _colmajor(T)
    xi = np.zeros(6, np.float64)
    e = lib().orc_reg_solve(_p(Hc, C.c_int64), _p(gc, C.c_int64), int(err), int(cnt), float(alpha),
                            _p(cm, C.c_float), _p(xi, C.c_double))
    return _from_colmajor(cm), float(e), xi


def register_cloud(m, cloud, pretransform, max_iterations, it_weight_gradient, epsilon,
                   map_resolution, trace=False):
    """src/cpu/registration.cpp:14-177.  `cloud` (int32 [n,3]) is transformed IN PLACE.

    Returns (total_transform 4x4 float32, iterations[, trace int64[iters,29]])."""
    assert cloud.dtype == np.int32 and cloud.flags.c_contiguous
    cm = _colmajor(pretransform)
    out = np.zeros(16, np.float32)
    tr = np.zeros((max(int(max_iterations), 1), 29), np.int64)
    it = lib().orc_register_cloud(m._h, cloud.ctypes.data, len(cloud), _p(cm, C.c_float),
                                  int(max_iterations), float(it_weight_gradient), float(epsilon),
                                  int(map_resolution), _p(out, C.c_float),
                                  _p(tr, C.c_int64), int(max_iterations))
    if trace:
        return _from_colmajor(out), it, tr[:it].copy()
    return _from_colmajor(out), it
