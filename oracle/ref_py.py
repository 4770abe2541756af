"""ctypes wrapper of oracle/_ref/libstvl_ref.so — the REFERENCE'S OWN translation units (spatio_temporal_voxel_grid.cpp and
the two frustum models, compiled unmodified from /root/reference against the stand-in third-party headers in oracle/ref_shim/,
see oracle/ref_driver.cpp).  TEST INFRASTRUCTURE: used by tests/test_ref_pin.py and tests/golden/make_ref_golden.py to pin the
restated oracle to the reference's code.  It only exists where /root/reference does (the authoring container)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libstvl_ref.so")
REFERENCE = "/root/reference/spatio_temporal_voxel_layer"


def available():
    return os.path.isdir(REFERENCE)


def build(force=False):
    if not available():
        raise RuntimeError("the reference sources are not on this machine")
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref"] + (["-B"] if force else []))
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp, d, u64 = C.c_void_p, C.c_double, C.c_uint64
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_float, d, C.c_int, d, C.c_int]
        L.ref_destroy.argtypes = [vp]
        L.ref_mark.argtypes = [vp, vp, u64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, d, d, d]
        L.ref_clear_frustums.argtypes = [vp, d, C.c_uint32, vp, vp]
        L.ref_active_count.restype = u64
        L.ref_active_count.argtypes = [vp]
        L.ref_get_voxels.argtypes = [vp, vp, vp, vp, vp, u64]
        L.ref_costmap_size.restype = u64
        L.ref_costmap_size.argtypes = [vp]
        L.ref_get_costmap.argtypes = [vp, vp, vp, vp, u64]
        L.ref_get_cleared.restype = u64
        L.ref_get_cleared.argtypes = [vp, vp, vp, u64]
        L.ref_get_points.restype = u64
        L.ref_get_points.argtypes = [vp, vp, u64]
        L.ref_reset.argtypes = [vp]
        L.ref_reset_area.argtypes = [vp, d, d, d, d, C.c_int]
        L.ref_index_to_world.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp]
        L.ref_decay_duration.restype = d
        L.ref_decay_duration.argtypes = [vp, d]
        L.ref_frustum_acceleration.restype = d
        L.ref_frustum_acceleration.argtypes = [vp, d, d]
        L.ref_frustum_is_inside.argtypes = [C.c_int, vp, vp, u64, vp]
        L.ref_depth_frustum_planes.argtypes = [vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def reading_params(vfov, vpad, hfov, min_z, max_z, origin, quat_xyzw, accel):
    return np.array([vfov, vpad, hfov, min_z, max_z, *origin, *quat_xyzw, accel], np.float64)


def frustum_is_inside(model_type, params, xyz):
    xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    out = np.zeros(len(xyz), np.uint8)
    p = np.ascontiguousarray(params, np.float64)
    lib().ref_frustum_is_inside(int(model_type), _ptr(p), _ptr(xyz), len(xyz), _ptr(out))
    return out.astype(bool)


def depth_frustum_planes(params):
    out = np.zeros(36, np.float64)
    p = np.ascontiguousarray(params, np.float64)
    valid = lib().ref_depth_frustum_planes(_ptr(p), _ptr(out))
    return bool(valid), out


class RefGrid:
    """volume_grid::SpatioTemporalVoxelGrid of the reference, clock injected per call."""

    def __init__(self, voxel_size=0.05, background=0.0, decay_model=0, voxel_decay=15.0, pub_voxels=False):
        self.L = lib()
        self.h = self.L.ref_create(np.float32(voxel_size), background, decay_model, voxel_decay, int(pub_voxels))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    def mark(self, cloud, origin, obstacle_range, now, marking=1.0, point_step=None, offsets=(0, 4, 8)):
        if point_step is None:
            a = np.ascontiguousarray(cloud, np.float32)
            raw, n, point_step = a.view(np.uint8).reshape(-1), a.shape[0], a.shape[1] * 4
        else:
            raw = np.ascontiguousarray(cloud, np.uint8).reshape(-1)
            n = len(raw) // point_step
        o = np.ascontiguousarray(origin, np.float64)
        self.L.ref_mark(self.h, _ptr(raw), n, point_step, offsets[0], offsets[1], offsets[2], _ptr(o), obstacle_range, float(marking), now)

    def clear_frustums(self, now, readings=()):
        """readings: [(model_type, reading_params(...))] in reading order."""
        n = len(readings)
        mt = np.array([r[0] for r in readings], np.int32)
        p = np.concatenate([r[1] for r in readings]) if n else np.zeros(0)
        self.L.ref_clear_frustums(self.h, now, n, _ptr(mt) if n else None, _ptr(p) if n else None)

    def active_count(self):
        return self.L.ref_active_count(self.h)

    def voxels(self):
        n = self.active_count()
        ix = np.zeros(n, np.int32); iy = np.zeros(n, np.int32); iz = np.zeros(n, np.int32); t = np.zeros(n, np.float64)
        self.L.ref_get_voxels(self.h, _ptr(ix), _ptr(iy), _ptr(iz), _ptr(t), n)
        return ix, iy, iz, t

    def costmap(self):
        n = self.L.ref_costmap_size(self.h)
        x = np.zeros(n, np.float64); y = np.zeros(n, np.float64); c = np.zeros(n, np.uint32)
        self.L.ref_get_costmap(self.h, _ptr(x), _ptr(y), _ptr(c), n)
        return x, y, c

    def cleared(self):
        n = self.L.ref_get_cleared(self.h, None, None, 0)
        x = np.zeros(n, np.float64); y = np.zeros(n, np.float64)
        self.L.ref_get_cleared(self.h, _ptr(x), _ptr(y), n)
        return x, y

    def points(self):
        n = self.L.ref_get_points(self.h, None, 0)
        xyz = np.zeros((n, 3), np.float32)
        self.L.ref_get_points(self.h, _ptr(xyz), n)
        return xyz

    def reset(self):
        return bool(self.L.ref_reset(self.h))

    def reset_area(self, sx, sy, ex, ey, invert=False):
        self.L.ref_reset_area(self.h, sx, sy, ex, ey, int(invert))

    def index_to_world(self, ix, iy, iz):
        out = np.zeros(3, np.float64)
        self.L.ref_index_to_world(self.h, ix, iy, iz, _ptr(out))
        return out

    def decay_duration(self, t):
        return self.L.ref_decay_duration(self.h, t)

    def frustum_acceleration(self, t, factor):
        return self.L.ref_frustum_acceleration(self.h, t, factor)
