"""Per-CTA phase timeline of the sweep and Mark kernels (profiling only).

Needs the trace variant of the library (never the product):
    python -m spatio_temporal_voxel_layer_b200.build --trace
    STVL_B200_LIB=spatio_temporal_voxel_layer_b200/libstvl_b200_trace.so python profiles/trace_phases.py [--cycles 60]

Thread 0 of every CTA stamps %globaltimer at phase boundaries (csrc/stvl_kernels.cu, TRACE(kernel, phase)); this script
runs config-2 cycles back to back on the 710,765-voxel seed (dt 0.2 s), reads the stamps of the LAST cycle and prints, per
kernel and phase, when the first / median / last CTA reached it, relative to the kernel's first CTA entry.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("STVL_B200_LIB", os.path.join(ROOT, "spatio_temporal_voxel_layer_b200", "libstvl_b200_trace.so"))

PHASES = {
    0: ("mark_kernel", ["entry", "setup done", "chunk0 consumed", "chunks done", "leaves resolved", "flushed", "exit"]),
    1: ("sweep_kernel", ["entry", "predecessor done", "re-arm issued", "leaves done", "published"]),
    3: ("costmap pass", ["entry", "tile_n read", "tiles done (thread 0)", "CTA done", "ticket taken"]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cycles", type=int, default=16)
    ap.add_argument("--dt", type=float, default=0.2)
    ap.add_argument("--staged", action="store_true", help="enqueue the cycle stage by stage (three C-ABI calls, no overlap) instead of the fused call")
    args = ap.parse_args()
    import torch
    import spatio_temporal_voxel_layer_b200 as stvl
    from spatio_temporal_voxel_layer_b200 import workloads
    w = workloads.Config2(n_poses=4, ring=8)
    p = w.p
    g = stvl.SpatioTemporalVoxelGrid(0.05, 0.0, 0, 15.0, leaf_capacity=1 << 20)
    g.set_outputs(0)
    g.load_voxels(*w.seed_vox, w.seed_t)
    L, h = g.L, g.h
    dev = [[torch.from_numpy(c).cuda() for c in cyc.clouds] for cyc in w.cycles]
    fr = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])
                          for o, q in cyc.poses]) for cyc in w.cycles]
    arrs = []
    for slot, cyc in enumerate(w.cycles):
        arr = (stvl.Reading * w.n_cams)()
        keep = []
        for k, ((o, q), c) in enumerate(zip(cyc.poses, dev[slot])):
            stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q,
                                    obstacle_range=p["obstacle_range"], now=0.0, filter=stvl.FILTER_PASSTHROUGH,
                                    min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
        arrs.append(arr)
    cp = w.costmap_params()
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0, 254)
    costmap = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    import time
    stream = torch.cuda.ExternalStream(g.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t = []
    for step in range(args.cycles):
        if step == 10:
            g.synchronize()
            e0.record(stream)
        th = time.perf_counter()
        slot = step % len(w.cycles)
        now = w.now0 + args.dt * (step + 1)
        for k in range(w.n_cams):
            arrs[slot][k].now = now + 1e-4 * k
        if args.staged:
            stvl._check(L.stvl_clear_frustums(h, now, fr[slot].ctypes.data_as(C.c_void_p), w.n_cams))
            stvl._check(L.stvl_mark_device(h, arrs[slot], w.n_cams))
            stvl._check(L.stvl_costmap_device(h, C.byref(params), costmap.data_ptr(), None))
        else:
            stvl._check(L.stvl_update_cycle_device(h, now, fr[slot].ctypes.data_as(C.c_void_p), w.n_cams, arrs[slot], w.n_cams,
                                                   C.byref(params), costmap.data_ptr()))
        host_t.append(time.perf_counter() - th)
    e1.record(stream)
    g.synchronize()
    print("cycle period (CUDA events, cycles 10..%d): %.2f us; host enqueue per cycle: median %.2f us (first 5 after the sync: %s)"
          % (args.cycles, e0.elapsed_time(e1) * 1e3 / (args.cycles - 10), np.median(host_t[10:]) * 1e6,
             ", ".join("%.1f" % (t * 1e6) for t in host_t[10:15])))
    tr = np.zeros((4, 1024, 8), np.uint64)
    L.stvl_debug_trace.argtypes = [C.c_void_p, C.c_size_t]
    rc = L.stvl_debug_trace(tr.ctypes.data_as(C.c_void_p), tr.nbytes)
    assert rc == 0, rc
    order = [1, 0, 3]   # stream order inside a cycle
    for k in order:
        name, phases = PHASES[k]
        used = tr[k, :, 0] > 0
        n = int(used.sum())
        if n == 0:
            continue
        t0 = int(tr[k, used, 0].min())
        sms = len(np.unique(tr[k, used, 7] & 0xFFFF))
        print("== %-15s %4d CTAs traced on %3d SMs; times in us relative to the kernel's first CTA entry: first / median / last CTA" % (name, n, sms))
        for ph, label in enumerate(phases):
            v = tr[k, used, ph].astype(np.int64)
            v = v[v > 0] - t0
            if len(v) == 0:
                continue
            print("   %-18s n=%4d  first %8.2f  median %8.2f  last %8.2f" % (label, len(v), v.min() / 1e3, np.median(v) / 1e3, v.max() / 1e3))
        for ph in range(1, len(phases)):
            a, b = tr[k, used, ph - 1].astype(np.int64), tr[k, used, ph].astype(np.int64)
            ok = (a > 0) & (b > 0)
            if ok.any():
                d = (b - a)[ok] / 1e3
                print("     d(%s -> %s): median %.2f  p90 %.2f  max %.2f us" % (phases[ph - 1], phases[ph], np.median(d), np.percentile(d, 90), d.max()))
    sw = tr[1]
    used = sw[:, 0] > 0
    if used.any():
        nl = (sw[used, 5] & 0xFFFFFFFF).astype(np.int64)
        nc = (sw[used, 5] >> 32).astype(np.int64)
        tb = (sw[used, 3].astype(np.int64) - sw[used, 6].astype(np.int64)) / 1e3
        ta = (sw[used, 6].astype(np.int64) - sw[used, 2].astype(np.int64)) / 1e3
        print("sweep pass 0 per CTA: leaves queued for the warp pass  min %d median %d max %d (sum %d); columns queued  min %d median %d max %d (sum %d)"
              % (nl.min(), np.median(nl), nl.max(), nl.sum(), nc.min(), np.median(nc), nc.max(), nc.sum()))
        print("   re-arm -> end of column pass (A + B'): median %.2f p90 %.2f max %.2f us;  warp pass (B) + C: median %.2f p90 %.2f max %.2f us"
              % (np.median(ta), np.percentile(ta, 90), ta.max(), np.median(tb), np.percentile(tb, 90), tb.max()))
        x = tr[2]
        if (x[:, 0] > 0).any():
            ok = used & (x[:, 0] > 0) & (x[:, 1] > 0) & (x[:, 2] > 0)
            d0 = (x[ok, 0].astype(np.int64) - sw[ok, 2].astype(np.int64)) / 1e3
            d1 = (x[ok, 1].astype(np.int64) - x[ok, 0].astype(np.int64)) / 1e3
            d2 = (x[ok, 2].astype(np.int64) - x[ok, 1].astype(np.int64)) / 1e3
            d3 = (sw[ok, 6].astype(np.int64) - x[ok, 2].astype(np.int64)) / 1e3
            for nm, d in (("re-arm -> headers loaded (warp 0)", d0), ("headers -> phase A done (warp 0)", d1), ("phase A done -> barrier passed", d2), ("barrier -> column pass done (warp 0)", d3)):
                print("   %-40s median %.2f p90 %.2f max %.2f us" % (nm, np.median(d), np.percentile(d, 90), d.max()))
        order = np.argsort(-tb)[:8]
        print("   slowest warp passes: " + ", ".join("%.1f us (%d leaves, %d columns)" % (tb[i], nl[i], nc[i]) for i in order))
    g.close()


if __name__ == "__main__":
    main()
