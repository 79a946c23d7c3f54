"""ctypes loader for oracle/_ref/libws_refcuda.so: the REFERENCE's own CUDA kernels recompiled for sm_100a
(TEST / BENCH INFRASTRUCTURE ONLY -- a same-hardware secondary baseline, never the parity target and never
on the product path).  Built by `make -C oracle _ref/libws_refcuda.so` where /root/reference exists; the
.so travels to the GPU box, the reference sources do not."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libws_refcuda.so")
REFERENCE = "/root/reference"


def build(force=False):
    """Compile the reference's CUDA sources in place (only possible where /root/reference is mounted)."""
    if not os.path.isdir(REFERENCE):
        return LIB_PATH if os.path.exists(LIB_PATH) else None
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_ref/libws_refcuda.so"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH)


class RefCuda:
    def __init__(self, size, tau, max_weight, res):
        L = C.CDLL(LIB_PATH)
        vp, i32p, i64p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float)
        L.refcuda_create.restype = vp
        L.refcuda_create.argtypes = [C.c_int] * 6
        L.refcuda_destroy.argtypes = [vp]
        L.refcuda_set_points.argtypes = [vp, vp, C.c_int64]
        L.refcuda_update.argtypes = [vp, i32p, i32p]
        L.refcuda_reg_prepare.argtypes = [vp]
        L.refcuda_reg_step.argtypes = [vp, f32p, i64p, i64p, i32p, i32p]
        L.refcuda_download.argtypes = [vp, vp]
        self.L = L
        self.size = [s if s % 2 == 1 else s + 1 for s in size]
        self.h = L.refcuda_create(self.size[0], self.size[1], self.size[2], int(tau), int(max_weight), int(res))

    def set_points(self, pts):
        p = np.ascontiguousarray(pts, dtype=np.int32)
        self.L.refcuda_set_points(self.h, p.ctypes.data, len(p))

    def update(self, pos, up):
        a = np.ascontiguousarray(pos, dtype=np.int32)
        b = np.ascontiguousarray(up, dtype=np.int32)
        self.L.refcuda_update(self.h, a.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_int32)))

    def reg_prepare(self):
        self.L.refcuda_reg_prepare(self.h)

    def reg_step(self, T16):
        H = np.zeros(36, np.int64)
        g = np.zeros(6, np.int64)
        e, c = C.c_int32(), C.c_int32()
        T = np.ascontiguousarray(T16, dtype=np.float32)
        self.L.refcuda_reg_step(self.h, T.ctypes.data_as(C.POINTER(C.c_float)), H.ctypes.data_as(C.POINTER(C.c_int64)),
                                g.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(e), C.byref(c))
        return H, g, e.value, c.value

    def download(self):
        out = np.zeros(int(np.prod(np.array(self.size, np.int64))), np.uint32)
        self.L.refcuda_download(self.h, out.ctypes.data)
        return out

    def close(self):
        if self.h:
            self.L.refcuda_destroy(self.h)
            self.h = None
