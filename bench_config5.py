#!/usr/bin/env python
"""bench_config5.py — BASELINE config 5: warehouse map (~50 M active voxels at 0.02 m), spatial key-range shards across
the GPUs of one box, NCCL reduce of the column counts.  STRONG scaling: the map is fixed, every rank loads and sweeps
only its x-slab.  (bench.py is the driver's benchmark on config 2; this script produces the config-5 numbers quoted in
DESIGN.md / BASELINE.md.)

  python bench_config5.py [--voxels 50000000] [--steps 20] [--warmup 3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_config5.py ...

Checks (size-independent properties, the oracle is too slow at this size): every loaded voxel is swept, survivors +
cleared == swept, the number of marked costmap columns equals the number of distinct rack-face columns plus whatever the
cameras added, and the merged costmap at N ranks equals the bytes rank 0 would produce alone (checked at small sizes by
tests/test_gpu_parity.py::test_sharded_grids_union_equals_single).
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=int, default=50_000_000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import torch
    import spatio_temporal_voxel_layer_b200 as stvl
    from spatio_temporal_voxel_layer_b200 import sharding, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = workloads.Config5(target=args.voxels)
    p = w.p
    vs = w.vs
    # x-slabs of `width` leaves dealt round-robin: rank r owns the leaves with floor(bx / width) mod N == r (the device
    # rule, csrc/stvl_device.cuh shard_owns).  With the map straddling x = 0 a rank may own two slabs.
    bx_lo, bx_hi = w.x_lo >> 3, w.x_hi >> 3
    width = max(1, int(math.ceil((bx_hi - bx_lo + 1) / world)))
    slabs = [sl for sl in range(bx_lo // width, bx_hi // width + 1) if sl % world == rank]
    n_expected = sum(max(0, min(w.x_hi + 1, (sl + 1) * width * 8) - max(w.x_lo, sl * width * 8)) for sl in slabs) * len(w.faces) * (w.z_range[1] - w.z_range[0] + 1)
    g = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, pub_voxels=False, device=local_rank,
                                     leaf_capacity=max(1 << 16, int(n_expected / 48) + (1 << 15)),
                                     shard_index=rank if world > 1 else 0, shard_count=world, shard_width=width)
    g.set_outputs(0)
    t0 = time.perf_counter()
    n_local = 0
    step = 8_000_000
    for sl in slabs:
        ix, iy, iz, t = w.voxels(x_range=(sl * width * 8, (sl + 1) * width * 8))
        n_local += len(ix)
        for s in range(0, len(ix), step):
            g.load_voxels(ix[s:s + step], iy[s:s + step], iz[s:s + step], t[s:s + step])
        del ix, iy, iz, t
    n_active = g.active_count()
    load_s = time.perf_counter() - t0
    assert n_active == n_local == n_expected, (n_active, n_local, n_expected)

    ring = len(w.cycles)
    dev_clouds = [[torch.from_numpy(c).cuda() for c in cyc.clouds] for cyc in w.cycles]
    frustums = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])
                                for o, q in cyc.poses]) for cyc in w.cycles]
    args_dev = []
    for slot in range(ring):
        arr = (stvl.Reading * w.n_cams)()
        keep = []
        for k, ((o, q), c) in enumerate(zip(w.cycles[slot].poses, dev_clouds[slot])):
            stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q,
                                    obstacle_range=p["obstacle_range"], now=0.0, filter=stvl.FILTER_PASSTHROUGH,
                                    min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
        args_dev.append(arr)
    cp = w.costmap_params()
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0, 254)
    costmap_dev = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(g.stream())
    if world > 1:   # in-library NCCL merge: one communicator per handle, id from rank 0
        uid = [stvl.SpatioTemporalVoxelGrid.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
    L, h = g.L, g.h
    fr_ptrs = [f.ctypes.data_as(C.c_void_p) for f in frustums]

    def cycle(step_no):
        slot = step_no % ring
        now = w.now0 + w.dt * (step_no + 1)
        arr = args_dev[slot]
        for k in range(w.n_cams):
            arr[k].now = now + 1e-4 * k
        if world == 1:
            stvl._check(L.stvl_update_cycle_device(h, now, fr_ptrs[slot], w.n_cams, arr, w.n_cams, C.byref(params), costmap_dev.data_ptr()))
        else:
            # the fused cycle on this rank's slab; local {0, lethal} bytes + NCCL max into rank 0's costmap, pipelined one
            # cycle behind on the side stream
            stvl._check(L.stvl_update_cycle_sharded_device(h, now, fr_ptrs[slot], w.n_cams, arr, w.n_cams, C.byref(params),
                                                           costmap_dev.data_ptr(), 0, 1))

    def barrier():
        g.synchronize(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        g.synchronize(); torch.cuda.synchronize()

    n = 0
    for _ in range(args.warmup):
        cycle(n); n += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        cycle(n); n += 1
    if world > 1:
        stvl._check(L.stvl_costmap_sharded_flush(h, C.byref(params), costmap_dev.data_ptr(), 0))
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    # the sweep's own duration: the same cycles once more with the stages bracketed by CUDA events inside the library
    # (the events serialise the kernels, so this pass is not part of the cycle time above)
    g.enable_timing(stvl.TIME_SWEEP | stvl.TIME_MARK | stvl.TIME_COSTMAP)
    g.timing()
    for _ in range(args.steps):
        cycle(n); n += 1
    if world > 1:
        stvl._check(L.stvl_costmap_sharded_flush(h, C.byref(params), costmap_dev.data_ptr(), 0))
    barrier()
    tm = g.timing()
    g.enable_timing(0)
    st = g.sweep_stats()
    marked = int((costmap_dev == 254).sum().item()) if rank == 0 else 0
    sw = torch.tensor([float(st.n_swept), float(st.n_survived + st.n_cleared)], dtype=torch.float64, device="cuda")
    tmax = torch.tensor([ms, tm.sweep_ms / max(tm.n_sweep, 1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(sw)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        swept = int(sw[0].item())
        assert swept == int(sw[1].item())
        peak = 6452.5
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        sweep_ms = float(tmax[1])
        line = {"metric": "mark+clear+costmap cycles/s (config 5: warehouse %d voxels, 0.02 m)" % args.voxels,
                "value": args.steps / (float(tmax[0]) / 1e3), "unit": "cycles/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(tmax[0]) / args.steps, "scaling": "strong",
                "active_voxels_swept": swept, "voxels_per_s": swept / (float(tmax[0]) / args.steps / 1e3),
                "sweep_ms_max_rank": sweep_ms,
                # the sweep's first kernel also resets the costmap bytes of the fused cycle (Costmap2D::resetMaps): 16 B per
                # active voxel + 1 B per costmap cell
                "sweep_roofline": {"algorithmic_bytes": swept * 16 // world + cp["size_x"] * cp["size_y"],
                                   "voxel_bytes": swept * 16 // world, "costmap_reset_bytes": cp["size_x"] * cp["size_y"],
                                   "achieved_gbs": (swept * 16 / world + cp["size_x"] * cp["size_y"]) / (sweep_ms / 1e3) / 1e9,
                                   "frac_of_measured_peak": (swept * 16 / world + cp["size_x"] * cp["size_y"]) / (sweep_ms / 1e3) / 1e9 / peak},
                "costmap_cells": cp["size_x"] * cp["size_y"], "marked_cells": marked, "load_s": load_s,
                "leaf_pool": {k: getattr(g.pool_stats(), k) for k in ("leaves_in_use", "leaf_capacity", "tiles_in_use", "device_bytes")}}
        print(json.dumps(line))
    barrier()
    if world > 1:
        g.comm_destroy()
    del costmap_dev, dev_clouds
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
