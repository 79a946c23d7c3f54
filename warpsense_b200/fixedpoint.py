"""Host-side fixed-point helpers of the hot path (numpy mirrors of include/util/util.h).

All citations are file:line under /root/reference.  These are used to prepare inputs for the
CUDA path (pose -> sensor voxel + up vector, scan points into the map frame); the per-point
work itself runs on the device.
"""
import numpy as np

MATRIX_RESOLUTION = 32768  # include/warpsense/consts.h:12-13
WEIGHT_RESOLUTION = 64     # include/warpsense/consts.h:9-10


def _trunc_f2i(a):
    """C++ float -> int cast (truncation toward zero)."""
    return np.trunc(np.asarray(a, dtype=np.float32)).astype(np.int64).astype(np.int32)


def to_int_mat(mat):
    """include/util/util.h:8-11 : (mat * MATRIX_RESOLUTION).cast<int>() in float32."""
    m = np.asarray(mat, dtype=np.float32).reshape(4, 4)
    return _trunc_f2i(m * np.float32(MATRIX_RESOLUTION))


def transform_points(points, int_mat):
    """include/util/util.h:13-18 on an [n,3] int32 array: int32 multiply-add (wrapping),
    then a truncating divide by MATRIX_RESOLUTION."""
    p = np.asarray(points, dtype=np.int32).reshape(-1, 3).astype(np.int64)
    m = np.asarray(int_mat, dtype=np.int64).reshape(4, 4)
    acc = p @ m[:3, :3].T + m[:3, 3][None, :]
    acc = ((acc + 2 ** 31) % 2 ** 32) - 2 ** 31            # int32 wrap-around
    q = np.abs(acc) // MATRIX_RESOLUTION                   # trunc toward zero
    return (np.sign(acc) * q).astype(np.int32)


def to_map(pose, map_resolution):
    """include/util/util.h:52-56 : floor(t / res) per axis, in float32."""
    t = np.asarray(pose, dtype=np.float32).reshape(4, 4)[:3, 3]
    return np.floor(t / np.float32(map_resolution)).astype(np.int32)


def convert_pose_to_gpu(pose, map_resolution):
    """src/warpsense/tsdf_mapping.cpp:77-85 -> (scanner voxel pos, up vector)."""
    rot = np.eye(4, dtype=np.int32)
    rot[:3, :3] = to_int_mat(pose)[:3, :3]
    up = transform_points(np.array([[0, 0, MATRIX_RESOLUTION]], np.int32), rot)[0]
    return to_map(pose, map_resolution), up


def colmajor16(mat):
    """4x4 matrix -> the column-major float32[16] Eigen::Matrix4f stores (the C-ABI layout)."""
    return np.ascontiguousarray(np.asarray(mat, dtype=np.float32).reshape(4, 4).T).reshape(16)


def from_colmajor16(a):
    return np.asarray(a, dtype=np.float32).reshape(4, 4).T.copy()
