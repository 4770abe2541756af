"""Micro-benchmarks of the individual kernels with CUDA events (inputs resident in HBM, > L2 working sets).
   python profiles/microbench.py
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spatio_temporal_voxel_layer_b200 as stvl
from spatio_temporal_voxel_layer_b200 import synth, workloads


def timed(stream, fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the GPU busy for a while first: an idle B200 sits at low clocks and the first launches after a pause are
    # 3x slower than in a running pipeline (which is what bench.py measures)
    t_end = time.perf_counter() + 0.4
    while time.perf_counter() < t_end:
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    w = workloads.Config2(n_poses=2, ring=8)
    p = w.p
    g = stvl.SpatioTemporalVoxelGrid(0.05, voxel_decay=15.0, leaf_capacity=1 << 20)
    g.set_outputs(0)
    g.load_voxels(*w.seed_vox, w.seed_t)
    stream = torch.cuda.ExternalStream(g.stream())
    L, h = g.L, g.h
    n_pts = w.points_per_cycle

    def mk_args(clouds_per_slot, filt=stvl.FILTER_PASSTHROUGH, rng=p["obstacle_range"]):
        out = []
        for slot, clouds in enumerate(clouds_per_slot):
            arr = (stvl.Reading * len(clouds))()
            keep = []
            for k, c in enumerate(clouds):
                o, q = w.cycles[slot % len(w.cycles)].poses[k]
                stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q, obstacle_range=rng,
                                        now=w.now0 + 1.0 + 1e-4 * k, filter=filt, min_obstacle_height=p["min_obstacle_height"],
                                        max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
            out.append(arr)
        return out

    real = [[torch.from_numpy(c).cuda() for c in cyc.clouds] for cyc in w.cycles]
    nan = [[torch.full_like(c, float("nan")) for c in cyc] for cyc in real]
    far = [[torch.full_like(c, 1000.0) for c in cyc] for cyc in real]   # finite but out of range
    variants = {"real clouds": mk_args(real), "all-NaN clouds (load + reject only)": mk_args(nan),
                "all out-of-range (load + distance test)": mk_args(far)}
    state = {"i": 0}
    g.enable_timing(stvl.TIME_MARK | stvl.TIME_SWEEP | stvl.TIME_COSTMAP)
    for name, args in variants.items():
        def run():
            a = args[state["i"] % len(args)]; state["i"] += 1
            for k in range(len(a)):     # clocks must keep advancing, or the library falls back to one launch per reading
                a[k].now = w.now0 + 1.0 + 1e-3 * state["i"] + 1e-5 * k
            stvl._check(L.stvl_mark_device(h, a, len(a)))
        g.timing()
        timed(stream, run, 40)
        t = g.timing()
        us = t.mark_ms / t.n_mark * 1e3
        print("mark  %-45s %8.1f us (kernel, library events) -> %7.1f GB/s algorithmic (16 B/pt)" % (name, us, n_pts * 16 / us / 1e3))
    # Mark when the fixed costs (launch, set-up, one flush per CTA, last-CTA fold: ~15 us) are amortised: the same seven
    # clouds replicated 8x per reading (16 readings of 1.23 M points = 19.7 M points = 315 MB per launch, > L2)
    big_src = [torch.cat([real[s % len(real)][k % 7] for s in range(8)]) for k in range(16)]
    arr = (stvl.Reading * 16)()
    keep = []
    for k, c in enumerate(big_src):
        o, q = w.cycles[0].poses[k % 7]
        stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q, obstacle_range=p["obstacle_range"],
                                now=w.now0 + 3.0 + 1e-4 * k, filter=stvl.FILTER_PASSTHROUGH, min_obstacle_height=p["min_obstacle_height"],
                                max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
    n_big = sum(c.shape[0] for c in big_src)
    def run_big():
        state["i"] += 1
        for k in range(16):
            arr[k].now = w.now0 + 1.0 + 1e-3 * state["i"] + 1e-5 * k
        stvl._check(L.stvl_mark_device(h, arr, 16))
    timed(stream, run_big, 3)
    g.timing()
    timed(stream, run_big, 10)
    t = g.timing()
    us = t.mark_ms / t.n_mark * 1e3
    print("mark  %-45s %8.1f us (kernel, library events) -> %7.1f GB/s algorithmic (16 B/pt)" % ("16 readings x 1.23 M points (315 MB)", us, n_big * 16 / us / 1e3))
    del big_src
    # plain device copy of the same bytes as a yardstick
    src = torch.empty(n_pts * 4 * 8, dtype=torch.float32, device="cuda")
    dst = torch.empty(n_pts * 4, dtype=torch.float32, device="cuda")
    st = {"i": 0}
    def cp():
        k = st["i"] % 8; st["i"] += 1
        dst.copy_(src[k * n_pts * 4:(k + 1) * n_pts * 4])
    with torch.cuda.stream(stream):
        us = timed(stream, cp, 40)
    print("torch copy of one cycle's clouds (read+write)      %8.1f us  -> %7.1f GB/s read" % (us, n_pts * 16 / us / 1e3))

    fr = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q in cyc.poses]) for cyc in w.cycles]
    now = {"t": w.now0 + 2.0}
    def sweep(with_frustums=True):
        now["t"] += 0.01
        f = fr[0]
        stvl._check(L.stvl_clear_frustums(h, now["t"], f.ctypes.data_as(C.c_void_p) if with_frustums else None, len(f) if with_frustums else 0))
    us = timed(stream, lambda: sweep(True), 40)
    st_ = g.sweep_stats()
    print("sweep 7 frustums (%d voxels)                    %8.1f us  -> %7.1f GB/s algorithmic (16 B/voxel)" % (st_.n_swept, us, st_.n_swept * 16 / us / 1e3))
    us = timed(stream, lambda: sweep(False), 40)
    print("sweep no frustums                                  %8.1f us" % us)
    cp_ = w.costmap_params()
    params = stvl.CostmapParams(cp_["origin_x"], cp_["origin_y"], cp_["resolution"], cp_["size_x"], cp_["size_y"], 0, 0, 254)
    cm = torch.zeros((cp_["size_y"], cp_["size_x"]), dtype=torch.uint8, device="cuda")
    us = timed(stream, lambda: stvl._check(L.stvl_costmap_device(h, C.byref(params), cm.data_ptr(), None)), 40)
    print("costmap %dx%d                                  %8.1f us" % (cp_["size_x"], cp_["size_y"], us))
    # updateCosts on device byte grids: footprint polygon + updateWithMax of the whole 1600 x 1600 window
    master = torch.zeros_like(cm)
    fp = np.array([[0.35, 0.25], [-0.35, 0.25], [-0.35, -0.25], [0.35, -0.25]]) + np.array([cp_["origin_x"] + 40.0, cp_["origin_y"] + 40.0])
    us = timed(stream, lambda: g.update_costs_device(params, cm.data_ptr(), master.data_ptr(), footprint_xy=fp, combination_method=1), 40)
    print("updateCosts (footprint + updateWithMax) %dx%d      %8.1f us  -> %7.1f GB/s (2 B read + 1 B written per cell)" % (
        cp_["size_x"], cp_["size_y"], us, cp_["size_x"] * cp_["size_y"] * 3 / us / 1e3))
    print("pool:", {k: getattr(g.pool_stats(), k) for k in ("leaves_in_use", "tiles_in_use", "kernel_launches")})


if __name__ == "__main__":
    main()
