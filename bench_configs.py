"""Runners for the BASELINE.json configs other than the headline one, shared by bench.py (timed) and tests/ (full-size
parity): configs 1, 3, 4 (workloads.GrowingConfig: the store grows from empty while the robot drives) through the fused
device cycle and, cycle for cycle, through the CPU oracle.  Test / bench infrastructure: the oracle is imported here, never
by the product package."""
import ctypes as C
import time

import numpy as np


def _gpu_readings(stvl, sensors, now, dev_clouds, voxel_min_points):
    """C-ABI argument arrays of one cycle: frustum blocks (clearing) and marking readings on device-resident clouds."""
    filt = {"none": stvl.FILTER_NONE, "passthrough": stvl.FILTER_PASSTHROUGH, "voxel": stvl.FILTER_VOXEL}
    blocks, keep = [], []
    arr = (stvl.Reading * len(sensors))()
    for k, (s, t) in enumerate(zip(sensors, dev_clouds)):
        p = s.p
        if s.model_type == 0:
            blocks.append(stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], s.origin, s.quat, p["accel"]))
        else:
            blocks.append(stvl.build_lidar_frustum(p["vfov"], p["vpad"], p["hfov"], p["min_z"], p["max_z"], s.origin, s.quat, p["accel"]))
        r = stvl.MeasurementReading((t.data_ptr(), t.shape[0], 16, (0, 4, 8)), origin=s.origin, orientation=s.quat,
                                    obstacle_range=p["obstacle_range"], now=now + 1e-4 * k, filter=filt[s.filter],
                                    min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"],
                                    voxel_min_points=voxel_min_points)
        r._fill(arr[k], now, keep)
    return np.concatenate(blocks), arr, keep


def run_gpu(stvl, w, warmup=5, leaf_capacity=1 << 18, timed=True):
    """All cycles of `w` through stvl_update_cycle_device.  Returns dict(us_per_cycle, grid, costmap (last cycle), ...);
    the caller closes the grid."""
    import torch
    g = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, pub_voxels=False, leaf_capacity=leaf_capacity)
    g.set_outputs(0)
    cp = w.costmap_params()
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                                cp["mark_threshold"], cp["default_value"], 254)
    cm_dev = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    dev = [[torch.from_numpy(np.ascontiguousarray(s.cloud, np.float32)).cuda() for s in cyc] for cyc in w.cycles]
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(g.stream())
    prepared = []
    for j in range(w.n_cycles):
        now, sensors = w.cycle(j)
        blocks, arr, keep = _gpu_readings(stvl, sensors, now, dev[j % len(dev)], w.voxel_min_points)
        prepared.append((now, blocks, arr, keep))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = 0
    for j, (now, blocks, arr, keep) in enumerate(prepared):
        if j == warmup and timed:
            g.synchronize(); torch.cuda.synchronize()
            l0 = g.pool_stats().kernel_launches
            e0.record(stream)
        stvl._check(g.L.stvl_update_cycle_device(g.h, now, blocks.ctypes.data_as(C.c_void_p), len(blocks), arr, len(arr),
                                                 C.byref(params), cm_dev.data_ptr()))
    e1.record(stream)
    g.synchronize(); torch.cuda.synchronize()
    n_timed = w.n_cycles - warmup
    out = dict(grid=g, costmap=cm_dev.cpu().numpy(), params=cp, n_timed=n_timed)
    if timed and n_timed > 0:
        ms = e0.elapsed_time(e1)
        out.update(us_per_cycle=1e3 * ms / n_timed, cycles_per_s=n_timed / (ms / 1e3), launches=int(g.pool_stats().kernel_launches - l0))
    del dev
    return out


def run_oracle(w, warmup=5):
    """The same cycles on the CPU oracle (pre-filters applied as the reference's MeasurementBuffer would, outside the clock).
    Returns dict(grid, costmap, n_marked, seconds per timed cycle)."""
    from oracle import oracle_py as O
    og = O.OracleGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, False)
    cp = w.costmap_params()
    leaf = np.float32(w.voxel_size)
    times, cm, n_marked = [], None, 0
    for j in range(w.n_cycles):
        now, sensors = w.cycle(j)
        blocks, clouds = [], []
        for s in sensors:
            p = s.p
            if s.model_type == 0:
                blocks.append(O.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], s.origin, s.quat, p["accel"]))
            else:
                blocks.append(O.lidar_frustum(p["vfov"], p["vpad"], p["hfov"], p["min_z"], p["max_z"], s.origin, s.quat, p["accel"]))
            if s.filter == "voxel":
                clouds.append(O.filter_voxel(s.cloud, leaf, p["min_obstacle_height"], p["max_obstacle_height"], w.voxel_min_points))
            elif s.filter == "passthrough":
                clouds.append(np.ascontiguousarray(O.filter_passthrough(s.cloud, p["min_obstacle_height"], p["max_obstacle_height"])))
            else:
                clouds.append(np.ascontiguousarray(s.cloud[:, :3]))
        t0 = time.perf_counter()
        og.clear_frustums(now, np.concatenate(blocks))
        for k, (s, xyz) in enumerate(zip(sensors, clouds)):
            og.mark(xyz, s.origin, s.p["obstacle_range"], now + 1e-4 * k)
        cm, _, n_marked = og.update_ros_costmap(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                                                cp["mark_threshold"], cp["default_value"])
        if j >= warmup:
            times.append(time.perf_counter() - t0)
    return dict(grid=og, costmap=cm, n_marked=n_marked, times=times)


def compare(gpu, ora):
    """Bit-exact comparison of the final state: voxel keys + mark times, per-column counts, last costmap bytes."""
    g, og = gpu["grid"], ora["grid"]
    gv, ov = g.voxels(), og.voxels()
    gk, ok_ = np.lexsort((gv[2], gv[1], gv[0])), np.lexsort((ov[2], ov[1], ov[0]))
    same_vox = len(gv[0]) == len(ov[0]) and all(np.array_equal(np.asarray(a)[gk].view(np.uint8), np.asarray(b)[ok_].view(np.uint8)) for a, b in zip(gv, ov))
    ix, iy, cnt = g.columns()
    gx, gy = g.index_to_world(ix), g.index_to_world(iy)
    ox, oy, oc = og.costmap()
    a, b = np.lexsort((gy, gx)), np.lexsort((oy, ox))
    same_cols = (len(gx) == len(ox) and np.array_equal(gx[a].view(np.uint64), ox[b].view(np.uint64)) and
                 np.array_equal(gy[a].view(np.uint64), oy[b].view(np.uint64)) and np.array_equal(cnt[a], oc[b]))
    same_cm = bool(np.array_equal(gpu["costmap"], ora["costmap"]))
    return dict(voxels=bool(same_vox), columns=bool(same_cols), costmap=same_cm, n_voxels=int(len(gv[0])), n_columns=int(len(gx)),
                marked_cells=int((gpu["costmap"] == 254).sum()))
