#!/usr/bin/env python
"""bench.py — mark+clear+costmap update cycles/s of the STVL hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      the CPU restatement (oracle) on the host cores

A step is one full update cycle in the reference's order (spatio_temporal_voxel_layer.cpp:772-795):
ClearFrustums(7 frustums) -> Mark(7 x 640x480 clouds) -> costmap bytes.  Workload at N=1 = BASELINE config 2.
  value : cycles/s with the clouds already resident in HBM (CUDA events on the library's stream)
  e2e   : cycles/s through the C-ABI with HOST buffers: pinned clouds copied H2D and the costmap bytes copied D2H
          inside every step (wall clock, synchronised on both sides)
  roofline : the dominant kernel's ALGORITHMIC bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline : the oracle (a port, 1 thread like the reference) on a bounded sample of the same cycles
N > 1 (torchrun): weak scaling — every rank owns one config-2 slab of an N-times larger store (spatial key-range
shards); per cycle the per-(x,y) column counts are merged with an NCCL reduce to rank 0, which writes the bytes.
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mark+clear+costmap cycles/s (7x640x480 RGBD, 0.05 m)"
UNIT = "cycles/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def oracle_frustums(O, p, poses):
    return np.concatenate([O.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q in poses])


def run_oracle_cycles(w, n_warm, n_steps, threshold_params):
    """The CPU restatement on the same cycles.  The PassThrough pre-filter runs upstream of the cycle in the reference
    (MeasurementBuffer, another thread), so it is applied before the clock starts; the GPU path fuses it into Mark."""
    from oracle import oracle_py as O
    p = w.p
    og = O.OracleGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, False)
    og.load_voxels(*w.seed_vox, w.seed_t)
    filtered = {}
    times = []
    cm = None
    for step in range(n_warm + n_steps):
        now, poses, clouds = w.cycle(step)
        key = step % len(w.cycles)
        if key not in filtered:
            filtered[key] = [np.ascontiguousarray(O.filter_passthrough(c, p["min_obstacle_height"], p["max_obstacle_height"])) for c in clouds]
        t0 = time.perf_counter()
        og.clear_frustums(now, oracle_frustums(O, p, poses))
        for k, ((o, q), xyz) in enumerate(zip(poses, filtered[key])):
            og.mark(xyz, o, p["obstacle_range"], now + 1e-4 * k)
        cm, _, _ = og.update_ros_costmap(threshold_params["origin_x"], threshold_params["origin_y"], threshold_params["resolution"],
                                         threshold_params["size_x"], threshold_params["size_y"], threshold_params["mark_threshold"],
                                         threshold_params["default_value"])
        t1 = time.perf_counter()
        if step >= n_warm:
            times.append(t1 - t0)
    return times, og, cm


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return model, os.cpu_count()


def bench_reference(args):
    """--impl reference: the reference's CPU path (restated: the real one needs OpenVDB/Eigen/PCL/ROS 2, absent here).
    Single-threaded like the reference (one mutex-guarded loop)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from spatio_temporal_voxel_layer_b200 import workloads
    w = workloads.Config2(n_poses=2, ring=4)
    cp = w.costmap_params()
    times, og, _ = run_oracle_cycles(w, args.warmup, args.steps, cp)
    total = sum(times)
    v = len(times) / total
    model, ncpu = host_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w.name, "points_per_cycle": w.points_per_cycle, "active_voxels": int(og.active_count()),
                   "host_cpu": model, "host_cpus": ncpu},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d full config-2 cycles after %d warm-up (oracle/stvl_oracle.cpp, 1 thread as upstream; PassThrough pre-filter excluded)" % (args.steps, args.warmup)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def bench_ours(args):
    import torch
    import spatio_temporal_voxel_layer_b200 as stvl
    from spatio_temporal_voxel_layer_b200 import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: rank r owns slab r (a config-2 store shifted by 80 m in x) --------------------------------------
    slab_pitch = 80.0
    # N=1: the store is centred on the origin (negative and positive coordinates); N>1: slab r spans x in [80r, 80r+80)
    slab_origin = (0.0, 0.0) if world == 1 else (slab_pitch * rank + slab_pitch / 2, 0.0)
    w = workloads.Config2(n_poses=4, ring=8, slab_origin=slab_origin, seed=1002 + rank)
    p = w.p
    vs = float(np.float32(w.voxel_size))
    leaf_w = int(round(slab_pitch / (8 * vs)))      # slab width in leaves (200)
    grid_kw = dict(voxel_size=w.voxel_size, decay_model=w.decay_model, voxel_decay=w.voxel_decay, pub_voxels=False,
                   device=local_rank, leaf_capacity=1 << 22)
    if world > 1:   # spatial key-range shard: this rank keeps leaves with floor(bx / 200) mod N == rank
        grid_kw.update(shard_index=rank, shard_count=world, shard_width=leaf_w)
    g = stvl.SpatioTemporalVoxelGrid(**grid_kw)
    # plugin flow: UpdateROSCostmap only touch()es the cleared cells, i.e. needs their bounding box (always produced)
    g.set_outputs(0)
    g.load_voxels(*w.seed_vox, w.seed_t)
    n_seed = g.active_count()

    ring = len(w.cycles)
    # clouds: pinned host copies (e2e leg) and device-resident copies (value leg)
    host_clouds = [[torch.from_numpy(c).pin_memory() for c in cyc.clouds] for cyc in w.cycles]
    dev_clouds = [[c.cuda(non_blocking=True) for c in cyc] for cyc in host_clouds]
    torch.cuda.synchronize()
    frustums = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])
                                for o, q in cyc.poses]) for cyc in w.cycles]
    cp = w.costmap_params()
    if world > 1:   # one costmap over all slabs
        cp["size_x"] = int(1600 + (world - 1) * slab_pitch / cp["resolution"])
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                                cp["mark_threshold"], cp["default_value"], 254)
    costmap_dev = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    costmap_host = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.ExternalStream(g.stream())
    bytes_h2d = sum(c.numel() * 4 for c in host_clouds[0]) + 7 * 304
    bytes_d2h = cp["size_x"] * cp["size_y"] + 256

    # per ring slot: the C-ABI argument arrays are built once; a step only refreshes the clock samples
    def build_args(src):
        out = []
        for slot in range(ring):
            arr = (stvl.Reading * w.n_cams)()
            keep = []
            for k, ((o, q), c) in enumerate(zip(w.cycles[slot].poses, src[slot])):
                stvl.MeasurementReading((c.data_ptr(), c.shape[0], 16, (0, 4, 8)), origin=o, orientation=q,
                                        obstacle_range=p["obstacle_range"], now=0.0, filter=stvl.FILTER_PASSTHROUGH,
                                        min_obstacle_height=p["min_obstacle_height"],
                                        max_obstacle_height=p["max_obstacle_height"])._fill(arr[k], 0.0, keep)
            out.append(arr)
        return out
    args_dev, args_host = build_args(dev_clouds), build_args(host_clouds)
    fr_ptrs = [f.ctypes.data_as(C.c_void_p) for f in frustums]
    L, h, check = g.L, g.h, stvl._check
    n_cams = w.n_cams

    def enqueue_clear_mark(step, device):
        slot = step % ring
        now = w.now0 + w.dt * (step + 1)
        arr = (args_dev if device else args_host)[slot]
        for k in range(n_cams):
            arr[k].now = now + 1e-4 * k
        check(L.stvl_clear_frustums(h, now, fr_ptrs[slot], n_cams))
        return arr

    params_ref = C.byref(params)
    cm_dev_ptr, cm_host_ptr = costmap_dev.data_ptr(), costmap_host.data_ptr()
    if world > 1:   # one NCCL communicator per handle, id created by rank 0
        uid = [stvl.SpatioTemporalVoxelGrid.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)

    def merge_pipelined():
        """Device-resident leg, N>1: every rank thresholds its own voxel columns (x-slab shards own whole columns) and the library reduces the
        {0, lethal} bytes to rank 0 with NCCL (max) on a side stream while its main stream already sweeps and marks the next cycle (column tiles are double-buffered),
        and rank 0 thresholds cycle k's merged counts with call k+1.  Enqueue only."""
        check(L.stvl_costmap_sharded_device(h, params_ref, cm_dev_ptr, 0, 0, 1))

    def merge_drain():
        check(L.stvl_costmap_sharded_flush(h, params_ref, cm_dev_ptr, 0))


    def step_device(step, fused=True):
        """One cycle with device-resident clouds.  N=1: ONE C-ABI call (stvl_update_cycle_device) enqueues the fused
        clear -> mark -> costmap cycle (three launches chained by programmatic dependent launch); nothing synchronises.
        fused=False enqueues the same cycle stage by stage (used for the per-stage / per-kernel event timings: in the
        fused cycle Mark overlaps the sweep, so a per-kernel duration is not defined there)."""
        if world == 1 and not fused:
            arr = enqueue_clear_mark(step, True)
            check(L.stvl_mark_device(h, arr, n_cams))
            check(L.stvl_costmap_device(h, params_ref, cm_dev_ptr, None))
        elif world == 1:
            slot = step % ring
            now = w.now0 + w.dt * (step + 1)
            arr = args_dev[slot]
            for k in range(n_cams):
                arr[k].now = now + 1e-4 * k
            check(L.stvl_update_cycle_device(h, now, fr_ptrs[slot], n_cams, arr, n_cams, params_ref, cm_dev_ptr))
        elif fused:
            # N>1: the same fused cycle on this rank's shard; its costmap pass writes the rank's {0, lethal} bytes into the
            # reduce window and the library merges the windows into rank 0's costmap with NCCL (max) on a side stream,
            # pipelined one cycle behind (stvl_update_cycle_sharded_device)
            slot = step % ring
            now = w.now0 + w.dt * (step + 1)
            arr = args_dev[slot]
            for k in range(n_cams):
                arr[k].now = now + 1e-4 * k
            check(L.stvl_update_cycle_sharded_device(h, now, fr_ptrs[slot], n_cams, arr, n_cams, params_ref, cm_dev_ptr, 0, 1))
        else:
            arr = enqueue_clear_mark(step, True)
            check(L.stvl_mark_device(h, arr, n_cams))
            merge_pipelined()

    def step_host(step):
        arr = enqueue_clear_mark(step, False)
        check(L.stvl_mark(h, arr, n_cams))
        if world == 1:
            res = stvl.CostmapResult()
            check(L.stvl_costmap(h, params_ref, cm_host_ptr, C.byref(res)))
            return res
        check(L.stvl_costmap_sharded_device(h, params_ref, cm_dev_ptr, 0, 0, 0))   # blocking merge: this cycle's bytes
        res = None
        if rank == 0:
            with torch.cuda.stream(stream):
                costmap_host.copy_(costmap_dev, non_blocking=True)
        g.synchronize()
        return res

    def barrier():
        g.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        g.synchronize()
        torch.cuda.synchronize()

    # ---- value leg: inputs resident in HBM ---------------------------------------------------------------------------
    step_no = 0
    for _ in range(args.warmup):
        step_device(step_no); step_no += 1
    barrier()
    launches0 = g.pool_stats().kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_device(step_no); step_no += 1
    if world > 1:
        merge_drain()                      # the last cycle's bytes are part of the timed region
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = e0.elapsed_time(e1)   # every rank's work (incl. the NCCL reduce) is ordered on the library's stream
    # The timed region lasts a few milliseconds — less than one nvidia-smi sampling period — so the sampler keeps running
    # while the very same cycle loop continues (untimed) for 8000 more cycles (~0.3 s): the clocks reported are those under this load.
    launches = g.pool_stats().kernel_launches - launches0
    for _ in range(8000):                  # (a fixed count: every rank must enqueue the same number of merges)
        step_device(step_no); step_no += 1
        if step_no % 100 == 0:
            g.synchronize()
    if world > 1:
        merge_drain()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "nvidia-smi -lms 100 from the start of the timed region through 8000 more (untimed) cycles of the same loop"
    # Per-kernel timing: K more cycles of the same ring, enqueued stage by stage with every stage bracketed by CUDA events
    # inside the library (on the library's stream).  Not part of `value`.  The Mark kernel's duration for the roofline
    # comes from here: standalone, serialised behind the sweep, inputs not in L2 (ring of 8 x 34.4 MB).
    g.enable_timing(stvl.TIME_SWEEP | stvl.TIME_MARK | stvl.TIME_COSTMAP)
    g.timing()
    for _ in range(args.steps):
        step_device(step_no, fused=False); step_no += 1
    t_all = g.timing()
    t_mark = t_all
    g.enable_timing(0)
    st = g.sweep_stats()
    active = g.active_count()

    if world > 1:
        merge_drain()
    barrier()
    # ---- e2e leg: host clouds in, host costmap bytes out ---------------------------------------------------------------
    for _ in range(args.warmup):
        step_host(step_no); step_no += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last_res = step_host(step_no); step_no += 1
    barrier()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    value = world * args.steps / (dev_ms / 1e3)
    e2e = world * args.steps / e2e_s

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.name, "points_per_cycle": w.points_per_cycle, "seeded_voxels": int(n_seed),
                       "active_voxels_end": int(active), "swept_last_cycle": int(st.n_swept),
                       "l2_policy": "inputs larger than L2: ring of %d distinct cycles x %.1f MB of clouds; leaf pool %.1f GB" % (
                           ring, w.points_per_cycle * 16 / 1e6, g.pool_stats().device_bytes / 1e9),
                       "costmap_cells": cp["size_x"] * cp["size_y"],
                       "parallelism": "1 GPU" if world == 1 else "%d spatial x-slab shards, fused cycle per shard + in-library NCCL merge of the per-shard costmap bytes (max), pipelined one cycle behind on a side stream" % world},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_h2d, "d2h_bytes_per_step": bytes_d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if t_mark.n_mark:
            mark_s = t_mark.mark_ms / t_mark.n_mark / 1e3
            algo_mark = w.points_per_cycle * 16          # 16 B read per input point
            algo_clear = st.n_swept * 16                 # 16 B per active voxel (key + time)
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "r1b_traffic.json")) as f:
                    traffic = json.load(f).get("mark_kernel_dram_bytes_per_launch")
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": "mark_kernel", "achieved": algo_mark / mark_s / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": algo_mark / mark_s / 1e9 / peak, "traffic": traffic,
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_mark,
                                "kernel_ms": t_mark.mark_ms / t_mark.n_mark, "launches_timed": int(t_mark.n_mark),
                                "timing": "CUDA events around the kernel on the library's stream, in a stage-by-stage pass of the same "
                                          "cycles right after the timed region (in the fused cycle Mark overlaps the sweep by "
                                          "programmatic dependent launch and has no duration of its own)"}
            if t_all.n_sweep and t_all.n_costmap and t_all.n_mark:
                sweep_s = t_all.sweep_ms / t_all.n_sweep / 1e3
                line["roofline"]["stage_ms"] = {"clear": t_all.sweep_ms / t_all.n_sweep, "mark": t_all.mark_ms / t_all.n_mark,
                                                "costmap": t_all.costmap_ms / t_all.n_costmap,
                                                "note": "stage-by-stage pass (stvl_clear_frustums, stvl_mark_device, stvl_costmap_device), serialised; the fused cycle of `value` overlaps them"}
                line["roofline"]["sweep_kernels"] = {"achieved": algo_clear / sweep_s / 1e9, "frac": algo_clear / sweep_s / 1e9 / peak,
                                                     "algorithmic_bytes_per_launch": int(algo_clear)}
        # cpu baseline: bounded sample of the same cycles on the host (rank 0, N=1 only)
        if world == 1 and not args.no_cpu_baseline:
            wb = w
            times, og, ocm = run_oracle_cycles(wb, 1, args.cpu_cycles, cp)
            model, ncpu = host_info()
            line["cpu_baseline"] = {"value": len(times) / sum(times), "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d full config-2 cycles after 1 warm-up, oracle/stvl_oracle.cpp single thread (the reference path is single-threaded); host %s x%s" % (len(times), model, ncpu)}
        print(json.dumps(line))
    # tear down in dependency order: tensors that were used on the library's stream are released while that stream is
    # still alive (the caching allocator records events on it), then the process group, then the grid
    barrier()
    if world > 1:
        merge_drain()
        torch.cuda.synchronize()
        g.comm_destroy()
    del costmap_dev, costmap_host, dev_clouds, host_clouds
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-cycles", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
