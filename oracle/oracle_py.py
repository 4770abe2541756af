"""ctypes wrapper around oracle/libstvl_oracle.so (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstvl_oracle.so")

FRUSTUM_DTYPE = np.dtype([("type", "<i4"), ("reserved", "<i4"), ("accel_factor", "<f8"), ("p", "<f8", (36,))])
assert FRUSTUM_DTYPE.itemsize == 304


def build(force=False):
    """Compile the oracle with its Makefile (g++ only)."""
    src = os.path.join(_HERE, "stvl_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, d, i32, u32, u64, f32 = C.c_void_p, C.c_double, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
    L.oracle_create.restype = vp
    L.oracle_create.argtypes = [f32, d, C.c_int, d, C.c_int]
    L.oracle_destroy.argtypes = [vp]
    L.oracle_voxel_size.restype = d
    L.oracle_voxel_size.argtypes = [vp]
    L.oracle_build_depth_frustum.restype = C.c_int
    L.oracle_build_depth_frustum.argtypes = [d, d, d, d, vp, vp, d, vp]
    L.oracle_build_lidar_frustum.restype = C.c_int
    L.oracle_build_lidar_frustum.argtypes = [d, d, d, d, d, vp, vp, d, vp]
    L.oracle_is_inside.argtypes = [vp, vp, u64, vp]
    L.oracle_clear_frustums.argtypes = [vp, d, vp, C.c_int]
    L.oracle_mark.argtypes = [vp, vp, u64, u32, u32, u32, u32, vp, d, d, d, vp]
    L.oracle_active_count.restype = u64
    L.oracle_active_count.argtypes = [vp]
    L.oracle_get_voxels.restype = u64
    L.oracle_get_voxels.argtypes = [vp, vp, vp, vp, vp, u64]
    L.oracle_load_voxels.argtypes = [vp, vp, vp, vp, vp, u64]
    L.oracle_costmap_size.restype = u64
    L.oracle_costmap_size.argtypes = [vp]
    L.oracle_get_costmap.restype = u64
    L.oracle_get_costmap.argtypes = [vp, vp, vp, vp, u64]
    L.oracle_get_cleared.restype = u64
    L.oracle_get_cleared.argtypes = [vp, vp, vp, u64]
    L.oracle_get_points.restype = u64
    L.oracle_get_points.argtypes = [vp, vp, u64]
    L.oracle_reset.restype = C.c_int
    L.oracle_reset.argtypes = [vp]
    L.oracle_reset_area.argtypes = [vp, d, d, d, d, C.c_int]
    L.oracle_update_ros_costmap.restype = u64
    L.oracle_update_ros_costmap.argtypes = [vp, d, d, d, u32, u32, C.c_int, C.c_uint8, C.c_uint8, vp, vp]
    L.oracle_quantise.argtypes = [vp, f32, f32, f32, vp]
    L.oracle_index_to_world.argtypes = [vp, i32, i32, i32, vp]
    L.oracle_decay_duration.restype = d
    L.oracle_decay_duration.argtypes = [vp, d]
    L.oracle_frustum_acceleration.restype = d
    L.oracle_frustum_acceleration.argtypes = [d, d]
    L.oracle_filter_passthrough.restype = u64
    L.oracle_filter_passthrough.argtypes = [vp, u64, u32, u32, u32, u32, d, d, vp]
    L.oracle_transform_from_quat.restype = None
    L.oracle_transform_from_quat.argtypes = [vp, vp, vp]
    L.oracle_transform_cloud.restype = None
    L.oracle_transform_cloud.argtypes = [vp, u64, u32, u32, u32, u32, vp, vp]
    L.oracle_set_convex_polygon_cost.restype = C.c_int
    L.oracle_set_convex_polygon_cost.argtypes = [vp, d, d, d, u32, u32, vp, u32, C.c_uint8]
    L.oracle_update_costs.restype = None
    L.oracle_update_costs.argtypes = [vp, vp, u32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_filter_voxel.restype = u64
    L.oracle_filter_voxel.argtypes = [vp, u64, u32, u32, u32, u32, f32, d, d, C.c_int, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def depth_frustum(vfov, hfov, min_d, max_d, origin, quat_xyzw, accel=0.0):
    out = np.zeros(1, FRUSTUM_DTYPE)
    o = np.ascontiguousarray(origin, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    lib().oracle_build_depth_frustum(vfov, hfov, min_d, max_d, _ptr(o), _ptr(q), accel, _ptr(out))
    return out


def lidar_frustum(vfov, vpad, hfov, min_d, max_d, origin, quat_xyzw, accel=0.0):
    out = np.zeros(1, FRUSTUM_DTYPE)
    o = np.ascontiguousarray(origin, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    lib().oracle_build_lidar_frustum(vfov, vpad, hfov, min_d, max_d, _ptr(o), _ptr(q), accel, _ptr(out))
    return out


def is_inside(block, xyz):
    xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    out = np.zeros(len(xyz), np.uint8)
    b = np.ascontiguousarray(block)
    lib().oracle_is_inside(_ptr(b), _ptr(xyz), len(xyz), _ptr(out))
    return out.astype(bool)


class OracleGrid:
    """Mirror of volume_grid::SpatioTemporalVoxelGrid (reference
    include/spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp:122-189) with `now` injected."""

    def __init__(self, voxel_size=0.05, background=0.0, decay_model=0, voxel_decay=15.0, pub_voxels=False):
        self.L = lib()
        self.h = self.L.oracle_create(np.float32(voxel_size), background, decay_model, voxel_decay, int(pub_voxels))
        self.voxel_size = self.L.oracle_voxel_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    def clear_frustums(self, now, frustums=None):
        if frustums is None or len(frustums) == 0:
            self.L.oracle_clear_frustums(self.h, now, None, 0)
        else:
            f = np.ascontiguousarray(frustums, FRUSTUM_DTYPE)
            self.L.oracle_clear_frustums(self.h, now, _ptr(f), len(f))

    def mark(self, cloud, origin, obstacle_range, now, marking=1.0, point_step=None, offsets=(0, 4, 8)):
        """cloud: float32 array (N, k>=3) (x,y,z first) or a raw uint8 buffer with point_step/offsets."""
        if cloud.dtype == np.uint8:
            raw = np.ascontiguousarray(cloud)
            n = raw.size // point_step
        else:
            a = np.ascontiguousarray(cloud, np.float32)
            n = a.shape[0]
            point_step = a.shape[1] * 4
            raw = a
        o = np.ascontiguousarray(origin, np.float64)
        nf = C.c_uint64(0)
        self.L.oracle_mark(self.h, _ptr(raw), n, point_step, offsets[0], offsets[1], offsets[2], _ptr(o),
                           obstacle_range, float(marking), now, C.byref(nf))
        return nf.value

    def active_count(self):
        return self.L.oracle_active_count(self.h)

    def voxels(self):
        n = self.active_count()
        ix = np.zeros(n, np.int32); iy = np.zeros(n, np.int32); iz = np.zeros(n, np.int32); t = np.zeros(n, np.float64)
        self.L.oracle_get_voxels(self.h, _ptr(ix), _ptr(iy), _ptr(iz), _ptr(t), n)
        return ix, iy, iz, t

    def load_voxels(self, ix, iy, iz, t):
        ix = np.ascontiguousarray(ix, np.int32); iy = np.ascontiguousarray(iy, np.int32)
        iz = np.ascontiguousarray(iz, np.int32); t = np.ascontiguousarray(t, np.float64)
        self.L.oracle_load_voxels(self.h, _ptr(ix), _ptr(iy), _ptr(iz), _ptr(t), len(ix))

    def costmap(self):
        """GetFlattenedCostmap: (x, y, count) of the last sweep."""
        n = self.L.oracle_costmap_size(self.h)
        x = np.zeros(n, np.float64); y = np.zeros(n, np.float64); c = np.zeros(n, np.uint32)
        self.L.oracle_get_costmap(self.h, _ptr(x), _ptr(y), _ptr(c), n)
        return x, y, c

    def cleared(self):
        n = self.L.oracle_get_cleared(self.h, None, None, 0)
        x = np.zeros(n, np.float64); y = np.zeros(n, np.float64)
        self.L.oracle_get_cleared(self.h, _ptr(x), _ptr(y), n)
        return x, y

    def points(self):
        n = self.L.oracle_get_points(self.h, None, 0)
        xyz = np.zeros((n, 3), np.float32)
        self.L.oracle_get_points(self.h, _ptr(xyz), n)
        return xyz

    def reset(self):
        return bool(self.L.oracle_reset(self.h))

    def reset_area(self, sx, sy, ex, ey, invert=False):
        self.L.oracle_reset_area(self.h, sx, sy, ex, ey, int(invert))

    def update_ros_costmap(self, origin_x, origin_y, resolution, size_x, size_y, mark_threshold=0,
                           default_value=0, lethal=254, bounds=None):
        cm = np.zeros((size_y, size_x), np.uint8)
        b = np.array([1e308, 1e308, -1e308, -1e308] if bounds is None else bounds, np.float64)
        n = self.L.oracle_update_ros_costmap(self.h, origin_x, origin_y, resolution, size_x, size_y,
                                             mark_threshold, default_value, lethal, _ptr(cm), _ptr(b))
        return cm, b, n

    def quantise(self, x, y, z):
        out = np.zeros(3, np.int32)
        self.L.oracle_quantise(self.h, np.float32(x), np.float32(y), np.float32(z), _ptr(out))
        return tuple(int(v) for v in out)

    def index_to_world(self, ix, iy, iz):
        out = np.zeros(3, np.float64)
        self.L.oracle_index_to_world(self.h, ix, iy, iz, _ptr(out))
        return out

    def decay_duration(self, t):
        return self.L.oracle_decay_duration(self.h, t)


def frustum_acceleration(t, factor):
    return lib().oracle_frustum_acceleration(t, factor)


def set_convex_polygon_cost(costmap, origin_x, origin_y, resolution, polygon_xy, cost):
    """nav2 Costmap2D::setConvexPolygonCost on a (size_y, size_x) uint8 array, in place; False (nothing changed) if a
    vertex is off the map."""
    poly = np.ascontiguousarray(polygon_xy, np.float64).reshape(-1, 2)
    assert costmap.dtype == np.uint8 and costmap.flags["C_CONTIGUOUS"]
    return bool(lib().oracle_set_convex_polygon_cost(_ptr(costmap), origin_x, origin_y, resolution, costmap.shape[1], costmap.shape[0],
                                                     _ptr(poly), len(poly), cost))


def update_costs(layer, master, min_i, min_j, max_i, max_j, method):
    """nav2 CostmapLayer::updateWithOverwrite (0) / updateWithMax (1) of the window into `master`, in place."""
    assert layer.dtype == np.uint8 and master.dtype == np.uint8 and layer.shape == master.shape
    lib().oracle_update_costs(_ptr(layer), _ptr(master), layer.shape[1], min_i, min_j, max_i, max_j, method)


def transform_from_quat(translation, quat_xyzw):
    """tf2 TransformStamped -> row-major 3x4 float affine as tf2_sensor_msgs builds it (Eigen, float)."""
    t = np.ascontiguousarray(translation, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    m = np.zeros(12, np.float32)
    lib().oracle_transform_from_quat(_ptr(t), _ptr(q), _ptr(m))
    return m


def transform_cloud(cloud, m):
    """tf2::doTransform of every x,y,z (reference measurement_buffer.cpp:141-146); other columns are carried over."""
    a = np.ascontiguousarray(cloud, np.float32)
    m = np.ascontiguousarray(m, np.float32)
    out = np.zeros((a.shape[0], 3), np.float32)
    lib().oracle_transform_cloud(_ptr(a), a.shape[0], a.shape[1] * 4, 0, 4, 8, _ptr(m), _ptr(out))
    res = a.copy()
    res[:, :3] = out
    return res


def filter_passthrough(cloud, lo, hi):
    a = np.ascontiguousarray(cloud, np.float32)
    out = np.zeros((a.shape[0], 3), np.float32)
    n = lib().oracle_filter_passthrough(_ptr(a), a.shape[0], a.shape[1] * 4, 0, 4, 8, lo, hi, _ptr(out))
    return out[:n]


def filter_voxel(cloud, leaf, lo, hi, min_points=0):
    a = np.ascontiguousarray(cloud, np.float32)
    out = np.zeros((a.shape[0], 3), np.float32)
    n = lib().oracle_filter_voxel(_ptr(a), a.shape[0], a.shape[1] * 4, 0, 4, 8, np.float32(leaf), lo, hi,
                                  int(min_points), _ptr(out))
    return out[:n]
