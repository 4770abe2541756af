"""Spatial key-range sharding of the voxel grid across the GPUs of one box (BASELINE config 5).

Voxels are independent in Mark and in the clear sweep; the only cross-voxel step of the cycle is the per-(x,y) column
count.  Each rank owns the 8^3 leaves whose x index falls into its slabs; after the local sweep every rank scatters its
column counts into a dense window over the costmap, the windows are summed with one reduce (NCCL over NVLink on GPUs,
gloo in the CPU tests) and the destination rank thresholds the sum into costmap bytes.
"""
import math

import numpy as np


def floor_div(a, b):
    return a // b   # Python's // floors, like the device-side floor_div


def shard_of_leaf(bx, shard_width, shard_count):
    """Owner rank of leaf x-index bx — same rule as shard_owns() in csrc/stvl_device.cuh."""
    if shard_count <= 1:
        return 0
    return int(floor_div(int(bx), int(shard_width)) % int(shard_count))


def shard_of_voxel(ix, shard_width, shard_count):
    return shard_of_leaf(np.asarray(ix) >> 3 if isinstance(ix, np.ndarray) else int(ix) >> 3, shard_width, shard_count) \
        if not isinstance(ix, np.ndarray) else (np.floor_divide(np.asarray(ix) >> 3, shard_width) % max(shard_count, 1))


def dense_window(origin_x, origin_y, resolution, size_x, size_y, voxel_size):
    """Voxel-column index window [ix0, ix0+nx) x [iy0, iy0+ny) that covers every column whose centre can map into the
    costmap (two columns of slack on each side)."""
    vs = float(np.float32(voxel_size))
    ix0 = int(math.floor(origin_x / vs)) - 2
    iy0 = int(math.floor(origin_y / vs)) - 2
    nx = int(math.ceil(size_x * resolution / vs)) + 5
    ny = int(math.ceil(size_y * resolution / vs)) + 5
    return ix0, iy0, nx, ny


def merge_counts(dist, dense, dst=0):
    """Sum the per-rank dense column-count windows onto rank `dst` (one collective per cycle)."""
    dist.reduce(dense, dst=dst, op=dist.ReduceOp.SUM)
    return dense
