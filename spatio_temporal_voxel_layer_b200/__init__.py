"""spatio_temporal_voxel_layer_b200 — B200-native per-cycle voxel update of STVL.

The product is the CUDA library `libstvl_b200.so` (C-ABI in include/stvl_b200.h) and the C++ façade in
host/.  This module is the thin ctypes binding used by the tests and bench.py; it mirrors the reference's
`volume_grid::SpatioTemporalVoxelGrid` interface (reference
spatio_temporal_voxel_layer/include/spatio_temporal_voxel_layer/spatio_temporal_voxel_grid.hpp:122-189)
with the clock injected.  There is no CPU fallback: importing works without a GPU (so the symbols can be
checked), but creating a grid raises if the library or a B200 is missing.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

__all__ = ["load_library", "SpatioTemporalVoxelGrid", "MeasurementReading", "build_depth_frustum",
           "build_lidar_frustum", "FRUSTUM_DTYPE", "StvlError", "device_count", "EXPORTED_SYMBOLS"]

# STVL_B200_LIB selects another build of the same library (the profiling-only trace variant); never a different backend
LIB_PATH = os.environ.get("STVL_B200_LIB") or _build.LIB_PATH

STVL_OK = 0
DECAY_LINEAR, DECAY_EXPONENTIAL, DECAY_PERSISTENT = 0, 1, 2
MODEL_DEPTH_CAMERA, MODEL_3D_LIDAR = 0, 1
FILTER_NONE, FILTER_VOXEL, FILTER_PASSTHROUGH = 0, 1, 2
OUT_CLEARED, OUT_AMBIGUOUS = 1, 2
TIME_SWEEP, TIME_MARK, TIME_COSTMAP = 1, 2, 4
LETHAL_OBSTACLE, NO_INFORMATION, FREE_SPACE = 254, 255, 0

FRUSTUM_DTYPE = np.dtype([("type", "<i4"), ("reserved", "<i4"), ("accel_factor", "<f8"), ("p", "<f8", (36,))])
assert FRUSTUM_DTYPE.itemsize == 304


class StvlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("stvl error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("voxel_size", C.c_float), ("decay_model", C.c_int32), ("background_value", C.c_double),
                ("voxel_decay", C.c_double), ("pub_voxels", C.c_int32), ("device", C.c_int32),
                ("leaf_capacity", C.c_uint32), ("max_points", C.c_uint32), ("shard_index", C.c_int32),
                ("shard_count", C.c_int32), ("shard_width", C.c_int32), ("reserved", C.c_int32)]


class Reading(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_points", C.c_uint64), ("point_step", C.c_uint32), ("off_x", C.c_uint32),
                ("off_y", C.c_uint32), ("off_z", C.c_uint32), ("origin", C.c_double * 3),
                ("obstacle_range", C.c_double), ("now", C.c_double), ("marking", C.c_int32), ("filter", C.c_int32),
                ("min_obstacle_height", C.c_double), ("max_obstacle_height", C.c_double),
                ("voxel_min_points", C.c_int32), ("has_transform", C.c_int32), ("transform", C.c_float * 12)]


class SweepStats(C.Structure):
    _fields_ = [("n_swept", C.c_uint64), ("n_survived", C.c_uint64), ("n_cleared", C.c_uint64),
                ("n_rewritten", C.c_uint64), ("n_ambiguous", C.c_uint64), ("cleared_bbox_index", C.c_int32 * 4),
                ("cleared_bbox_world", C.c_double * 4)]


class CostmapParams(C.Structure):
    _fields_ = [("origin_x", C.c_double), ("origin_y", C.c_double), ("resolution", C.c_double),
                ("size_x", C.c_uint32), ("size_y", C.c_uint32), ("mark_threshold", C.c_int32),
                ("default_value", C.c_uint8), ("lethal_value", C.c_uint8), ("pad", C.c_uint8 * 2)]


class CostmapResult(C.Structure):
    _fields_ = [("n_marked_columns", C.c_uint64), ("has_bounds", C.c_int32), ("reserved", C.c_int32),
                ("touched", C.c_double * 4)]


class Timing(C.Structure):
    _fields_ = [("sweep_ms", C.c_double), ("mark_ms", C.c_double), ("costmap_ms", C.c_double),
                ("n_sweep", C.c_uint64), ("n_mark", C.c_uint64), ("n_costmap", C.c_uint64)]


class PoolStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "leaf_capacity", "leaves_in_use", "leaves_free", "hash_capacity", "hash_tombstones", "tiles_in_use",
        "tile_capacity", "points_dropped_out_of_range", "points_dropped_nonfinite", "kernel_launches", "device_bytes")]


# every symbol include/stvl_b200.h declares
EXPORTED_SYMBOLS = [
    "stvl_abi_version", "stvl_last_error", "stvl_device_count", "stvl_create", "stvl_destroy",
    "stvl_build_depth_frustum", "stvl_build_lidar_frustum", "stvl_build_transform", "stvl_clear_frustums", "stvl_mark", "stvl_mark_device",
    "stvl_costmap", "stvl_costmap_device", "stvl_set_outputs", "stvl_get_sweep_stats", "stvl_get_columns",
    "stvl_get_cleared", "stvl_get_points", "stvl_get_ambiguous", "stvl_reset", "stvl_reset_area", "stvl_active_count",
    "stvl_get_voxels", "stvl_load_voxels", "stvl_index_to_world", "stvl_voxel_size", "stvl_columns_to_dense",
    "stvl_costmap_from_dense", "stvl_get_pool_stats", "stvl_get_stream", "stvl_synchronize",
    "stvl_update_cycle_device", "stvl_update_cycle_sharded_device", "stvl_update_cycle_async", "stvl_update_cycle_wait", "stvl_shard_word_intervals", "stvl_update_costs_device", "stvl_get_costmap_result", "stvl_enable_timing", "stvl_get_timing",
    "stvl_nccl_unique_id", "stvl_comm_init", "stvl_comm_destroy", "stvl_costmap_sharded_device", "stvl_costmap_sharded_flush",
]

_lib = None


def load_library(build_if_missing=True):
    """dlopen libstvl_b200.so (building it with nvcc first if it is missing or stale). No GPU needed."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        try:
            _build.build_library()
        except Exception:
            if not os.path.exists(LIB_PATH):
                raise
    if not os.path.exists(LIB_PATH):
        raise ImportError("libstvl_b200.so is missing: run `python -m spatio_temporal_voxel_layer_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, d, i32, u32, u64 = C.c_void_p, C.c_double, C.c_int32, C.c_uint32, C.c_uint64
    L.stvl_abi_version.restype = C.c_int
    L.stvl_last_error.restype = C.c_char_p
    L.stvl_device_count.argtypes = [C.POINTER(C.c_int)]
    L.stvl_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.stvl_destroy.argtypes = [vp]
    L.stvl_build_depth_frustum.argtypes = [d, d, d, d, vp, vp, d, vp]
    L.stvl_build_lidar_frustum.argtypes = [d, d, d, d, d, vp, vp, d, vp]
    L.stvl_build_transform.argtypes = [vp, vp, vp]
    L.stvl_clear_frustums.argtypes = [vp, d, vp, u32]
    L.stvl_mark.argtypes = [vp, C.POINTER(Reading), u32]
    L.stvl_mark_device.argtypes = [vp, C.POINTER(Reading), u32]
    L.stvl_costmap.argtypes = [vp, C.POINTER(CostmapParams), vp, C.POINTER(CostmapResult)]
    L.stvl_costmap_device.argtypes = [vp, C.POINTER(CostmapParams), vp, C.POINTER(CostmapResult)]
    L.stvl_set_outputs.argtypes = [vp, u32]
    L.stvl_get_sweep_stats.argtypes = [vp, C.POINTER(SweepStats)]
    L.stvl_get_columns.argtypes = [vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.stvl_get_cleared.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.stvl_get_points.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.stvl_get_ambiguous.argtypes = [vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.stvl_reset.argtypes = [vp]
    L.stvl_reset_area.argtypes = [vp, d, d, d, d, C.c_int]
    L.stvl_active_count.argtypes = [vp, C.POINTER(u64)]
    L.stvl_get_voxels.argtypes = [vp, vp, vp, vp, vp, u64, C.POINTER(u64)]
    L.stvl_load_voxels.argtypes = [vp, vp, vp, vp, vp, u64]
    L.stvl_index_to_world.restype = d
    L.stvl_index_to_world.argtypes = [vp, i32]
    L.stvl_voxel_size.restype = d
    L.stvl_voxel_size.argtypes = [vp]
    L.stvl_columns_to_dense.argtypes = [vp, i32, i32, u32, u32, vp]
    L.stvl_costmap_from_dense.argtypes = [vp, vp, i32, i32, u32, u32, C.POINTER(CostmapParams), vp,
                                          C.POINTER(CostmapResult)]
    L.stvl_get_pool_stats.argtypes = [vp, C.POINTER(PoolStats)]
    L.stvl_get_stream.argtypes = [vp, C.POINTER(vp)]
    L.stvl_synchronize.argtypes = [vp]
    L.stvl_update_cycle_device.argtypes = [vp, d, vp, u32, C.POINTER(Reading), u32, C.POINTER(CostmapParams), vp]
    L.stvl_update_cycle_async.argtypes = [vp, d, vp, u32, C.POINTER(Reading), u32, C.POINTER(CostmapParams), vp, C.POINTER(u64)]
    L.stvl_update_cycle_wait.argtypes = [vp, u64, C.POINTER(CostmapResult)]
    L.stvl_shard_word_intervals.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.POINTER(CostmapParams), vp, vp]
    L.stvl_get_costmap_result.argtypes = [vp, C.POINTER(CostmapResult)]
    L.stvl_update_costs_device.argtypes = [vp, C.POINTER(CostmapParams), vp, vp, vp, u32, C.c_int, i32, i32, i32, i32]
    L.stvl_update_cycle_sharded_device.argtypes = [vp, d, vp, u32, C.POINTER(Reading), u32, C.POINTER(CostmapParams), vp, C.c_int, C.c_int]
    L.stvl_enable_timing.argtypes = [vp, u32]
    L.stvl_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.stvl_nccl_unique_id.argtypes = [vp]
    L.stvl_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.stvl_comm_destroy.argtypes = [vp]
    L.stvl_costmap_sharded_device.argtypes = [vp, C.POINTER(CostmapParams), vp, C.c_int, C.c_int, C.c_int]
    L.stvl_costmap_sharded_flush.argtypes = [vp, C.POINTER(CostmapParams), vp, C.c_int]
    _lib = L
    return L


def _check(rc):
    if rc != STVL_OK:
        raise StvlError(rc, load_library().stvl_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    n = C.c_int(0)
    _check(load_library().stvl_device_count(C.byref(n)))
    return n.value


def build_depth_frustum(vfov, hfov, min_d, max_d, origin, quat_xyzw, accel=0.0):
    """geometry::DepthCameraFrustum ctor + SetPosition/SetOrientation/TransformModel (host only)."""
    out = np.zeros(1, FRUSTUM_DTYPE)
    o = np.ascontiguousarray(origin, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    _check(load_library().stvl_build_depth_frustum(vfov, hfov, min_d, max_d, _ptr(o), _ptr(q), accel, _ptr(out)))
    return out


def build_lidar_frustum(vfov, vpad, hfov, min_d, max_d, origin, quat_xyzw, accel=0.0):
    """geometry::ThreeDimensionalLidarFrustum ctor + SetPosition/SetOrientation/TransformModel (host only)."""
    out = np.zeros(1, FRUSTUM_DTYPE)
    o = np.ascontiguousarray(origin, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    _check(load_library().stvl_build_lidar_frustum(vfov, vpad, hfov, min_d, max_d, _ptr(o), _ptr(q), accel, _ptr(out)))
    return out


def shard_word_intervals(voxel_size, shard_count, shard_width, n_ranks, params):
    """stvl_shard_word_intervals (host only): per rank, the [w_lo, w_hi) interval of 32-cell words of a costmap row its x-slabs map into."""
    lo = np.zeros(n_ranks, np.int32)
    hi = np.zeros(n_ranks, np.int32)
    _check(load_library().stvl_shard_word_intervals(np.float32(voxel_size), shard_count, shard_width, n_ranks, C.byref(params), _ptr(lo), _ptr(hi)))
    return lo, hi


def build_transform(translation, quat_xyzw):
    """Row-major 3x4 float affine of a TransformStamped, as tf2::doTransform applies it to a cloud (host only)."""
    t = np.ascontiguousarray(translation, np.float64)
    q = np.ascontiguousarray(quat_xyzw, np.float64)
    m = np.zeros(12, np.float32)
    _check(load_library().stvl_build_transform(_ptr(t), _ptr(q), _ptr(m)))
    return m


class MeasurementReading:
    """observation::MeasurementReading (reference measurement_reading.h:60-125), numpy-backed.

    cloud: float32 (N, k>=3) array with x,y,z first, OR (data_ptr:int, n_points, point_step, (offx,offy,offz))
    for raw / device-resident buffers."""

    def __init__(self, cloud=None, origin=(0.0, 0.0, 0.0), orientation=(0.0, 0.0, 0.0, 1.0), obstacle_range=2.5,
                 min_z=0.0, max_z=10.0, vfov=0.7, vfov_padding=0.0, hfov=1.04, decay_acceleration=0.0,
                 marking=True, clearing=False, model_type=MODEL_DEPTH_CAMERA, now=None,
                 filter=FILTER_NONE, min_obstacle_height=0.0, max_obstacle_height=3.0, transform=None, voxel_min_points=0):
        self.cloud = cloud
        self.origin = tuple(float(v) for v in origin)
        self.orientation = tuple(float(v) for v in orientation)
        self.obstacle_range = float(obstacle_range)
        self.min_z, self.max_z = float(min_z), float(max_z)
        self.vfov, self.vfov_padding, self.hfov = float(vfov), float(vfov_padding), float(hfov)
        self.decay_acceleration = float(decay_acceleration)
        self.marking, self.clearing = marking, clearing
        self.model_type = model_type
        self.now = now
        self.filter = filter
        self.min_obstacle_height, self.max_obstacle_height = float(min_obstacle_height), float(max_obstacle_height)
        self.transform = None if transform is None else np.ascontiguousarray(transform, np.float32).reshape(12)
        self.voxel_min_points = int(voxel_min_points)

    def frustum(self):
        if self.model_type == MODEL_DEPTH_CAMERA:
            return build_depth_frustum(self.vfov, self.hfov, self.min_z, self.max_z, self.origin, self.orientation,
                                       self.decay_acceleration)
        if self.model_type == MODEL_3D_LIDAR:
            return build_lidar_frustum(self.vfov, self.vfov_padding, self.hfov, self.min_z, self.max_z, self.origin,
                                       self.orientation, self.decay_acceleration)
        return None   # unknown model types are skipped, reference spatio_temporal_voxel_grid.cpp:134-137

    def _fill(self, r, default_now, keep):
        c = self.cloud
        if isinstance(c, tuple):
            ptr, n, step, offs = c
        else:
            a = np.ascontiguousarray(c, np.float32)
            keep.append(a)
            ptr, n, step, offs = a.ctypes.data, a.shape[0], a.shape[1] * 4, (0, 4, 8)
        r.data = ptr
        r.n_points = n
        r.point_step = step
        r.off_x, r.off_y, r.off_z = offs
        r.origin[0], r.origin[1], r.origin[2] = self.origin
        r.obstacle_range = self.obstacle_range
        r.now = float(self.now if self.now is not None else default_now)
        r.marking = 1 if self.marking else 0
        r.filter = self.filter
        r.min_obstacle_height = self.min_obstacle_height
        r.max_obstacle_height = self.max_obstacle_height
        r.voxel_min_points = self.voxel_min_points
        r.has_transform = 0 if self.transform is None else 1
        if self.transform is not None:
            for i in range(12):
                r.transform[i] = float(self.transform[i])


class SpatioTemporalVoxelGrid:
    """volume_grid::SpatioTemporalVoxelGrid on the GPU.  Method names follow the reference; `now` is injected."""

    def __init__(self, voxel_size=0.05, background_value=0.0, decay_model=DECAY_LINEAR, voxel_decay=15.0,
                 pub_voxels=False, device=-1, leaf_capacity=0, max_points=0, shard_index=0, shard_count=1,
                 shard_width=1):
        self.L = load_library()
        cfg = Config(np.float32(voxel_size), decay_model, background_value, voxel_decay, int(pub_voxels), device,
                     leaf_capacity, max_points, shard_index, shard_count, shard_width, 0)
        h = C.c_void_p()
        _check(self.L.stvl_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.voxel_size = self.L.stvl_voxel_size(self.h)
        self.pub_voxels = bool(pub_voxels)

    def close(self):
        if getattr(self, "h", None):
            self.L.stvl_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- hot path ------------------------------------------------------------------
    def ClearFrustums(self, clearing_readings, now):
        """reference .cpp:96-146.  clearing_readings: MeasurementReading list or a FRUSTUM_DTYPE array."""
        if isinstance(clearing_readings, np.ndarray):
            f = np.ascontiguousarray(clearing_readings, FRUSTUM_DTYPE)
        else:
            blocks = [r.frustum() for r in clearing_readings]
            blocks = [b for b in blocks if b is not None]
            f = np.concatenate(blocks) if blocks else np.zeros(0, FRUSTUM_DTYPE)
        _check(self.L.stvl_clear_frustums(self.h, float(now), _ptr(f) if len(f) else None, len(f)))

    def Mark(self, marking_readings, now=None, device=False):
        """reference .cpp:254-309.  `now` is the default per-reading clock sample."""
        n = len(marking_readings)
        arr = (Reading * max(n, 1))()
        keep = []
        for i, m in enumerate(marking_readings):
            m._fill(arr[i], now, keep)
        fn = self.L.stvl_mark_device if device else self.L.stvl_mark
        _check(fn(self.h, arr, n))

    def UpdateROSCostmap(self, origin_x, origin_y, resolution, size_x, size_y, mark_threshold=0,
                         default_value=FREE_SPACE, out=None):
        """the grid loop of SpatioTemporalVoxelLayer::UpdateROSCostmap (reference
        spatio_temporal_voxel_layer.cpp:699-718): returns (costmap bytes (size_y,size_x), CostmapResult)."""
        p = CostmapParams(origin_x, origin_y, resolution, size_x, size_y, mark_threshold, default_value,
                          LETHAL_OBSTACLE)
        if out is None:
            out = np.empty((size_y, size_x), np.uint8)
        res = CostmapResult()
        _check(self.L.stvl_costmap(self.h, C.byref(p), _ptr(out), C.byref(res)))
        return out, res

    def update_cycle_device(self, now, clearing_readings, marking_readings, origin_x, origin_y, resolution, size_x, size_y,
                            costmap_dev_ptr, mark_threshold=0, default_value=FREE_SPACE):
        """stvl_update_cycle_device: the fused clear -> mark -> costmap cycle on device-resident clouds (every marking
        reading's cloud is a (data_ptr, n, step, offsets) tuple); enqueue only."""
        fr = [b for b in (r.frustum() for r in clearing_readings) if b is not None]
        blocks = np.concatenate(fr) if fr else np.zeros(0, FRUSTUM_DTYPE)
        n = len(marking_readings)
        arr = (Reading * max(n, 1))()
        keep = []
        for i, m in enumerate(marking_readings):
            m._fill(arr[i], now, keep)
        p = CostmapParams(origin_x, origin_y, resolution, size_x, size_y, mark_threshold, default_value, LETHAL_OBSTACLE)
        _check(self.L.stvl_update_cycle_device(self.h, float(now), _ptr(blocks) if len(blocks) else None, len(blocks), arr, n,
                                               C.byref(p), costmap_dev_ptr))

    def update_cycle_async(self, now, clearing_readings, marking_readings, params, costmap_host):
        """stvl_update_cycle_async: the pipelined host-buffer cycle.  marking_readings carry HOST clouds (numpy arrays or
        (ptr, n, step, offsets) tuples of pinned memory), costmap_host is a uint8 numpy array (size_y, size_x) that stays
        alive until update_cycle_wait(ticket).  Returns the ticket."""
        fr = [b for b in (r.frustum() for r in clearing_readings) if b is not None]
        blocks = np.concatenate(fr) if fr else np.zeros(0, FRUSTUM_DTYPE)
        n = len(marking_readings)
        arr = (Reading * max(n, 1))()
        keep = []
        for i, m in enumerate(marking_readings):
            m._fill(arr[i], now, keep)
        t = C.c_uint64(0)
        _check(self.L.stvl_update_cycle_async(self.h, float(now), _ptr(blocks) if len(blocks) else None, len(blocks), arr, n,
                                              C.byref(params), _ptr(costmap_host), C.byref(t)))
        self._async_keep = getattr(self, "_async_keep", {})
        self._async_keep[t.value] = (keep, costmap_host)
        return t.value

    def update_cycle_wait(self, ticket):
        res = CostmapResult()
        _check(self.L.stvl_update_cycle_wait(self.h, ticket, C.byref(res)))
        getattr(self, "_async_keep", {}).pop(ticket, None)
        return res

    def update_cycle_sharded_device(self, now, clearing_readings, marking_readings, params, costmap_dev_ptr, root=0, pipelined=False):
        """stvl_update_cycle_sharded_device: the fused cycle on a spatial shard + in-library NCCL merge (bytes mode)."""
        fr = [b for b in (r.frustum() for r in clearing_readings) if b is not None]
        blocks = np.concatenate(fr) if fr else np.zeros(0, FRUSTUM_DTYPE)
        n = len(marking_readings)
        arr = (Reading * max(n, 1))()
        keep = []
        for i, m in enumerate(marking_readings):
            m._fill(arr[i], now, keep)
        _check(self.L.stvl_update_cycle_sharded_device(self.h, float(now), _ptr(blocks) if len(blocks) else None, len(blocks), arr, n,
                                                       C.byref(params), costmap_dev_ptr, int(root), 1 if pipelined else 0))

    def update_costs_device(self, params, layer_costmap_dev_ptr, master_dev_ptr, footprint_xy=None, combination_method=1, window=None):
        """SpatioTemporalVoxelLayer::updateCosts on device byte grids (reference spatio_temporal_voxel_layer.cpp:666-696):
        footprint clearing in the layer's costmap, then updateWithOverwrite (0) / updateWithMax (1) of `window` =
        (min_i, min_j, max_i, max_j) into the master grid.  Enqueue only."""
        fp = None if footprint_xy is None else np.ascontiguousarray(footprint_xy, np.float64).reshape(-1, 2)
        if window is None:
            window = (0, 0, params.size_x, params.size_y)
        _check(self.L.stvl_update_costs_device(self.h, C.byref(params), layer_costmap_dev_ptr, master_dev_ptr,
                                               _ptr(fp) if fp is not None and len(fp) else None, 0 if fp is None else len(fp),
                                               int(combination_method), *[int(v) for v in window]))

    def costmap_result(self):
        res = CostmapResult()
        _check(self.L.stvl_get_costmap_result(self.h, C.byref(res)))
        return res

    def GetFlattenedCostmap(self):
        """reference .cpp:312-317: dict-like arrays (x, y, count) of the last sweep's occupied voxel columns."""
        ix, iy, cnt = self.columns()
        return self.index_to_world(ix), self.index_to_world(iy), cnt

    def GetOccupancyPointCloud(self):
        """reference .cpp:344-376: float32 (N,3) voxel centres of the last sweep's survivors."""
        n = C.c_uint64(0)
        _check(self.L.stvl_get_points(self.h, None, 0, C.byref(n)))
        xyz = np.zeros((n.value, 3), np.float32)
        if n.value:
            _check(self.L.stvl_get_points(self.h, _ptr(xyz), n.value, C.byref(n)))
        return xyz

    def ResetGrid(self):
        _check(self.L.stvl_reset(self.h))
        return self.active_count() == 0

    def ResetGridArea(self, start, end, invert_area=False):
        _check(self.L.stvl_reset_area(self.h, start[0], start[1], end[0], end[1], int(invert_area)))

    # ---- queries ---------------------------------------------------------------------
    def sweep_stats(self):
        s = SweepStats()
        _check(self.L.stvl_get_sweep_stats(self.h, C.byref(s)))
        return s

    def columns(self):
        n = C.c_uint64(0)
        _check(self.L.stvl_get_columns(self.h, None, None, None, 0, C.byref(n)))
        m = n.value
        ix = np.zeros(m, np.int32); iy = np.zeros(m, np.int32); cnt = np.zeros(m, np.uint32)
        if m:
            _check(self.L.stvl_get_columns(self.h, _ptr(ix), _ptr(iy), _ptr(cnt), m, C.byref(n)))
        return ix, iy, cnt

    def cleared(self):
        n = C.c_uint64(0)
        _check(self.L.stvl_get_cleared(self.h, None, None, 0, C.byref(n)))
        m = n.value
        ix = np.zeros(m, np.int32); iy = np.zeros(m, np.int32)
        if m:
            _check(self.L.stvl_get_cleared(self.h, _ptr(ix), _ptr(iy), m, C.byref(n)))
        return ix, iy

    def ambiguous(self):
        n = C.c_uint64(0)
        _check(self.L.stvl_get_ambiguous(self.h, None, None, None, 0, C.byref(n)))
        m = n.value
        ix = np.zeros(m, np.int32); iy = np.zeros(m, np.int32); iz = np.zeros(m, np.int32)
        if m:
            _check(self.L.stvl_get_ambiguous(self.h, _ptr(ix), _ptr(iy), _ptr(iz), m, C.byref(n)))
        return ix, iy, iz

    def active_count(self):
        n = C.c_uint64(0)
        _check(self.L.stvl_active_count(self.h, C.byref(n)))
        return n.value

    def voxels(self):
        m = self.active_count()
        ix = np.zeros(m, np.int32); iy = np.zeros(m, np.int32); iz = np.zeros(m, np.int32); t = np.zeros(m, np.float64)
        n = C.c_uint64(0)
        if m:
            _check(self.L.stvl_get_voxels(self.h, _ptr(ix), _ptr(iy), _ptr(iz), _ptr(t), m, C.byref(n)))
        return ix, iy, iz, t

    def load_voxels(self, ix, iy, iz, t):
        ix = np.ascontiguousarray(ix, np.int32); iy = np.ascontiguousarray(iy, np.int32)
        iz = np.ascontiguousarray(iz, np.int32); t = np.ascontiguousarray(t, np.float64)
        _check(self.L.stvl_load_voxels(self.h, _ptr(ix), _ptr(iy), _ptr(iz), _ptr(t), len(ix)))

    def index_to_world(self, idx):
        """IndexToWorld (reference .cpp:446-460): idx*vs then + vs/2, two roundings."""
        return np.asarray(idx, np.float64) * self.voxel_size + self.voxel_size / 2.0

    def set_outputs(self, flags):
        _check(self.L.stvl_set_outputs(self.h, flags))

    def pool_stats(self):
        s = PoolStats()
        _check(self.L.stvl_get_pool_stats(self.h, C.byref(s)))
        return s

    def stream(self):
        s = C.c_void_p()
        _check(self.L.stvl_get_stream(self.h, C.byref(s)))
        return s.value

    def synchronize(self):
        _check(self.L.stvl_synchronize(self.h))

    # ---- in-library NCCL merge of the sharded column counts -------------------------------------------------------
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_uint8 * 128)()
        _check(load_library().stvl_nccl_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self.L.stvl_comm_init(self.h, buf, rank, n_ranks))

    def comm_destroy(self):
        _check(self.L.stvl_comm_destroy(self.h))

    def costmap_sharded_device(self, params, costmap_dev_ptr, root=0, count_bytes=4, pipelined=False):
        _check(self.L.stvl_costmap_sharded_device(self.h, C.byref(params), costmap_dev_ptr, root, count_bytes, int(pipelined)))

    def costmap_sharded_flush(self, params, costmap_dev_ptr, root=0):
        _check(self.L.stvl_costmap_sharded_flush(self.h, C.byref(params), costmap_dev_ptr, root))

    def enable_timing(self, mask):
        _check(self.L.stvl_enable_timing(self.h, mask))

    def timing(self):
        t = Timing()
        _check(self.L.stvl_get_timing(self.h, C.byref(t)))
        return t

    # ---- device-resident variants (bench `value`, multi-GPU merge) ---------------------
    def costmap_device(self, params, costmap_dev_ptr):
        res = CostmapResult()
        _check(self.L.stvl_costmap_device(self.h, C.byref(params), costmap_dev_ptr, C.byref(res)))
        return res

    def columns_to_dense(self, ix0, iy0, nx, ny, dense_dev_ptr):
        _check(self.L.stvl_columns_to_dense(self.h, ix0, iy0, nx, ny, dense_dev_ptr))

    def costmap_from_dense(self, dense_dev_ptr, ix0, iy0, nx, ny, params, costmap_dev_ptr, sync=True):
        res = CostmapResult() if sync else None
        _check(self.L.stvl_costmap_from_dense(self.h, dense_dev_ptr, ix0, iy0, nx, ny, C.byref(params),
                                              costmap_dev_ptr, C.byref(res) if sync else None))
        return res
