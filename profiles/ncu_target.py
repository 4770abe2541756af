"""A short, deterministic kernel sequence for ncu captures: config-2 store (710,765 voxels), then staged cycles
(sweep, one batched Mark launch, costmap) at dt = 0.2 s.   ncu -k regex:mark_kernel -s 3 -c 1 ... python profiles/ncu_target.py
--config5 N: a config-5 slab of N voxels instead (the sweep of a large store)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatio_temporal_voxel_layer_b200 as stvl
from spatio_temporal_voxel_layer_b200 import workloads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=6)
    ap.add_argument("--config5", type=int, default=0)
    ap.add_argument("--fused", action="store_true")
    args = ap.parse_args()
    if args.config5:
        w = workloads.Config5(target=args.config5, dt=0.2)
        vox = w.voxels()
        g = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, 0, w.voxel_decay, leaf_capacity=1 << 21)
        g.load_voxels(vox[0], vox[1], vox[2], vox[3])
    else:
        w = workloads.Config2(n_poses=2, ring=4, dt=0.2)
        g = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, 0, w.voxel_decay, leaf_capacity=1 << 20)
        g.load_voxels(*w.seed_vox, w.seed_t)
    g.set_outputs(0)
    p = w.p
    L, h = g.L, g.h
    dev = [[torch.from_numpy(c).cuda() for c in cyc.clouds] for cyc in w.cycles]
    fr = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])
                          for o, q in cyc.poses]) for cyc in w.cycles]
    arrs = []
    keep = []
    for slot, cyc in enumerate(w.cycles):
        arr = (stvl.Reading * w.n_cams)()
        for k, ((o, q), c) in enumerate(zip(cyc.poses, dev[slot])):
            stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q,
                                    obstacle_range=p["obstacle_range"], now=0.0, filter=stvl.FILTER_PASSTHROUGH,
                                    min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
        arrs.append(arr)
    cp = w.costmap_params()
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0, 254)
    costmap = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    for step in range(args.cycles):
        slot = step % len(w.cycles)
        now = w.now0 + w.dt * (step + 1)
        for k in range(w.n_cams):
            arrs[slot][k].now = now + 1e-4 * k
        if args.fused:
            stvl._check(L.stvl_update_cycle_device(h, now, fr[slot].ctypes.data_as(C.c_void_p), w.n_cams, arrs[slot], w.n_cams,
                                                   C.byref(params), costmap.data_ptr()))
        else:
            stvl._check(L.stvl_clear_frustums(h, now, fr[slot].ctypes.data_as(C.c_void_p), w.n_cams))
            stvl._check(L.stvl_mark_device(h, arrs[slot], w.n_cams))
            stvl._check(L.stvl_costmap_device(h, C.byref(params), costmap.data_ptr(), None))
        g.synchronize()
    st = g.sweep_stats()
    print("last sweep: swept %d cleared %d rewritten %d; active %d" % (st.n_swept, st.n_cleared, st.n_rewritten, g.active_count()))
    g.close()


if __name__ == "__main__":
    main()
