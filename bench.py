#!/usr/bin/env python
"""bench.py — mark+clear+costmap update cycles/s of the STVL hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      the CPU restatement (oracle) on the host cores

A step is one full update cycle in the reference's order (spatio_temporal_voxel_layer.cpp:772-795):
ClearFrustums(7 frustums) -> Mark(7 x 640x480 clouds) -> costmap bytes.  Workload at N=1 = BASELINE config 2.

The workload is NOT stationary (a 15 s linear decay empties a 710,765-voxel store that seven cameras on one robot cannot
re-mark), so every leg runs in EPISODES: the store is reset and re-seeded to the 710,765 voxels, one untimed settle
cycle runs, then EPISODE_CYCLES timed cycles at dt = 0.2 s — the store stays within 13 % of the seed in every
timed cycle of every leg.  Both arms (--impl ours / reference) build the same workloads.Config2 object and run the same
episodes, and the fused path's final state of an episode is compared bit for bit with the oracle's (`parity_checked`).

  value : cycles/s with the clouds already resident in HBM (CUDA events on the library's stream around each episode's
          timed cycles; the re-seed between episodes is not part of the path and is not timed)
  e2e   : cycles/s through the C-ABI with HOST buffers: pinned clouds copied H2D and the costmap bytes copied D2H inside
          every step (wall clock, synchronised on both sides of every episode)
  roofline : the dominant kernel's ALGORITHMIC bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline : the oracle (a port, 1 thread like the reference) on a bounded sample of the same episodes
N > 1 (torchrun): weak scaling — every rank owns one config-2 slab of an N-times larger store (spatial key-range
shards); per cycle the per-shard costmap cells are merged into rank 0's costmap inside the library.
"""
import argparse
import ctypes as C
import json
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
DT_HEADLINE = 0.2          # s between update cycles (the reference example runs its global costmap at 1 Hz)
EPISODE_CYCLES = 10        # timed cycles per episode (2 s of simulated time: the store decays from 710,765 to ~620 k)
SLAB_PITCH = 80.0          # m, x extent of one rank's slab at N > 1


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
        # the median over the samples taken under load (the sampler also sees the idle gaps of the re-seeds)
        hot = [v for v in sm if v >= 0.9 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(hot) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the workload both arms share
# ------------------------------------------------------------------------------------------------
def make_workload(rank, world):
    """Rank r's config-2 slab.  N=1: the store is centred on the origin (negative and positive coordinates);
    N>1: slab r spans x in [80 r, 80 r + 80)."""
    from spatio_temporal_voxel_layer_b200 import workloads
    origin = (0.0, 0.0) if world == 1 else (SLAB_PITCH * rank + SLAB_PITCH / 2, 0.0)
    return workloads.Config2(n_poses=4, ring=8, slab_origin=origin, seed=1002 + rank, dt=DT_HEADLINE)


def costmap_geometry(w, world):
    """Costmap parameters of the run.  N > 1: ONE costmap over all slabs, with slab 0's origin — identical on every rank (the
    ranks' cell intervals of the merge are derived from it, and the exchange sizes must agree)."""
    cp = w.costmap_params()
    if world > 1:
        from spatio_temporal_voxel_layer_b200 import synth
        slab0 = synth.retail_scene(size=75.0, seed=7, origin=(SLAB_PITCH / 2, 0.0))
        cp["origin_x"] = float(slab0.lo[:, 0].min() - 1.0)
        cp["origin_y"] = float(slab0.lo[:, 1].min() - 1.0)
        cp["size_x"] = int(1600 + (world - 1) * SLAB_PITCH / cp["resolution"])
    return cp


def shared_config(w, cp, world):
    """`config` of the JSON line — identical in both arms."""
    ring = len(w.cycles)
    return {"workload": w.name, "dt_s": w.dt, "episode_cycles": EPISODE_CYCLES,
            "episode": "store reset + re-seeded to %d voxels, 1 untimed settle cycle, then %d timed cycles; repeated until K steps are timed"
                       % (len(w.seed_vox[0]), EPISODE_CYCLES),
            "points_per_cycle": w.points_per_cycle, "seeded_voxels": len(w.seed_vox[0]),
            "costmap_cells": cp["size_x"] * cp["size_y"], "mark_threshold": cp["mark_threshold"],
            "l2_policy": "inputs larger than L2: ring of %d distinct cycles x %.1f MB of clouds" % (ring, w.points_per_cycle * 16 / 1e6),
            "n_slabs": world}


def episode_plan(n_steps):
    plan = [EPISODE_CYCLES] * (n_steps // EPISODE_CYCLES)
    if n_steps % EPISODE_CYCLES:
        plan.append(n_steps % EPISODE_CYCLES)
    return plan


def cycle_inputs(w, j, dt=None):
    """Cycle j of an episode (j = 0 is the settle cycle): (now, ring slot)."""
    return w.now0 + (w.dt if dt is None else dt) * (j + 1), j % len(w.cycles)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the same episodes
# ------------------------------------------------------------------------------------------------
class OracleRunner:
    def __init__(self, w, cp):
        from oracle import oracle_py as O
        self.O, self.w, self.cp = O, w, cp
        self.filtered = {}
        self.frustums = {}

    def _inputs(self, slot):
        """The PassThrough pre-filter runs upstream of the cycle in the reference (MeasurementBuffer, subscriber thread), so
        it is applied before the clock starts; the GPU path fuses it into Mark."""
        O, p = self.O, self.w.p
        if slot not in self.filtered:
            cyc = self.w.cycles[slot]
            self.filtered[slot] = [np.ascontiguousarray(O.filter_passthrough(c, p["min_obstacle_height"], p["max_obstacle_height"])) for c in cyc.clouds]
            self.frustums[slot] = np.concatenate([O.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q in cyc.poses])
        return self.filtered[slot], self.frustums[slot]

    def episode(self, n_cycles, dt=None):
        """reset + seed + settle + n_cycles.  Returns (per-cycle seconds of the timed cycles, grid, last costmap, marked cells)."""
        O, w, cp, p = self.O, self.w, self.cp, self.w.p
        og = O.OracleGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, False)
        og.load_voxels(*w.seed_vox, w.seed_t)
        times, cm, n_marked = [], None, 0
        for j in range(n_cycles + 1):
            now, slot = cycle_inputs(w, j, dt)
            clouds, fr = self._inputs(slot)
            t0 = time.perf_counter()
            og.clear_frustums(now, fr)
            for k, ((o, q), xyz) in enumerate(zip(w.cycles[slot].poses, clouds)):
                og.mark(xyz, o, p["obstacle_range"], now + 1e-4 * k)
            cm, _, n_marked = og.update_ros_costmap(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                                                    cp["mark_threshold"], cp["default_value"])
            t1 = time.perf_counter()
            if j >= 1:
                times.append(t1 - t0)
        return times, og, cm, n_marked


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
    Single-threaded like the reference (one mutex-guarded loop).  Same workload object, same episodes."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    w = make_workload(0, world)
    cp = costmap_geometry(w, world)
    runner = OracleRunner(w, cp)
    if args.warmup:
        runner.episode(min(args.warmup, EPISODE_CYCLES))
    times, og = [], None
    for n in episode_plan(args.steps):
        t, og, _, _ = runner.episode(n)
        times += t
    total = sum(times)
    v = len(times) / total
    model, ncpu = host_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(w, cp, world),
        "details": {"active_voxels_end_of_episode": int(og.active_count()), "host_cpu": model, "host_cpus": ncpu},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d config-2 cycles in episodes of %d after a warm-up episode (oracle/stvl_oracle.cpp, 1 thread as upstream; "
                                   "PassThrough pre-filter excluded)" % (args.steps, EPISODE_CYCLES)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, local_rank):
    """Best effort: run this rank (and therefore first-touch its pinned staging memory) on the CPUs of the NUMA node its GPU
    hangs off, so that N ranks' host->device streams do not all cross one node's memory controllers."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        if not hasattr(props, "pci_bus_id"):
            return "pci bus id unavailable in this torch build"
        dom = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, getattr(props, "pci_device_id", 0))
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % dom).read().strip())
        if node < 0:
            return "device reports no NUMA node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return "node %d CPUs are outside this process's cpuset (%d allowed CPUs)" % (node, len(allowed))
        os.sched_setaffinity(0, use)
        return "rank bound to NUMA node %d (%d CPUs)" % (node, len(use))
    except Exception as e:      # noqa: BLE001 — affinity is an optimisation, never a failure
        return "not bound (%s)" % type(e).__name__


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: warehouse map, ~50 M active voxels at 0.02 m, STRONG-scaled over the ranks (x-slab shards)
# ------------------------------------------------------------------------------------------------
def config5_record(args, torch, stvl, dist, rank, world, local_rank):
    """Every rank loads and sweeps only its x-slab of the fixed 50 M-voxel map; per cycle the ranks' lethal cells are merged
    into rank 0's costmap (packed cell bitmaps over ncclSend / ncclRecv, expanded on the root).  Size-independent checks (the
    oracle is too slow at this size; tests/ compare a 5 M-voxel instance with it): every loaded voxel is swept, survivors +
    cleared == swept, and the merged costmap holds exactly the cells the ranks marked."""
    import math
    from spatio_temporal_voxel_layer_b200 import workloads
    w = workloads.Config5(target=args.config5_voxels, dt=DT_HEADLINE)
    p = w.p
    bx_lo, bx_hi = w.x_lo >> 3, w.x_hi >> 3
    width = max(1, int(math.ceil((bx_hi - bx_lo + 1) / world)))
    slabs = [sl for sl in range(bx_lo // width, bx_hi // width + 1) if sl % world == rank]
    n_expected = sum(max(0, min(w.x_hi + 1, (sl + 1) * width * 8) - max(w.x_lo, sl * width * 8)) for sl in slabs) * len(w.faces) * (w.z_range[1] - w.z_range[0] + 1)
    g = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, pub_voxels=False, device=local_rank,
                                     leaf_capacity=max(1 << 16, int(n_expected / 44) + (1 << 15)),
                                     shard_index=rank if world > 1 else 0, shard_count=world, shard_width=width)
    g.set_outputs(0)
    t0 = time.perf_counter()
    n_local, step = 0, 8_000_000
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
    if world > 1:
        uid = [stvl.SpatioTemporalVoxelGrid.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)
    L, h, check = g.L, g.h, stvl._check
    fr_ptrs = [f.ctypes.data_as(C.c_void_p) for f in frustums]

    def cycle(step_no, fused=True):
        slot = step_no % ring
        now = w.now0 + w.dt * (step_no + 1)
        arr = args_dev[slot]
        for k in range(w.n_cams):
            arr[k].now = now + 1e-4 * k
        if not fused:
            check(L.stvl_clear_frustums(h, now, fr_ptrs[slot], w.n_cams))
            check(L.stvl_mark_device(h, arr, w.n_cams))
        elif world == 1:
            check(L.stvl_update_cycle_device(h, now, fr_ptrs[slot], w.n_cams, arr, w.n_cams, C.byref(params), costmap_dev.data_ptr()))
        else:
            check(L.stvl_update_cycle_sharded_device(h, now, fr_ptrs[slot], w.n_cams, arr, w.n_cams, C.byref(params),
                                                     costmap_dev.data_ptr(), 0, 1))

    def drain():
        if world > 1:
            check(L.stvl_costmap_sharded_flush(h, C.byref(params), costmap_dev.data_ptr(), 0))

    def barrier():
        g.synchronize(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        g.synchronize(); torch.cuda.synchronize()

    n, warm, steps = 0, 3, args.config5_steps
    for _ in range(warm):
        cycle(n); n += 1
    drain(); barrier()
    st0 = g.sweep_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        cycle(n); n += 1
    drain()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    st1 = g.sweep_stats()
    own_marked = int(g.costmap_result().n_marked_columns)
    merged_marked = int((costmap_dev == 254).sum().item()) if rank == 0 else 0
    # the sweep kernel's own duration: the next cycles stage by stage, the sweep bracketed by CUDA events inside the library
    g.enable_timing(stvl.TIME_SWEEP)
    g.timing()
    for _ in range(5):
        cycle(n, fused=False); n += 1
    barrier()
    tm = g.timing()
    g.enable_timing(0)
    st2 = g.sweep_stats()
    vals = torch.tensor([float(st1.n_swept), float(st1.n_survived + st1.n_cleared), float(own_marked), float(st0.n_swept), float(st2.n_swept)],
                        dtype=torch.float64, device="cuda")
    tmax = torch.tensor([ms, tm.sweep_ms / max(tm.n_sweep, 1)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    rec = None
    if rank == 0:
        swept = int(vals[0].item())
        peak, _ = measured_peak_gbs()
        sweep_ms = float(tmax[1])
        algo = int(vals[4].item()) * 16 // world          # 16 B per active voxel, per rank (max-rank time against the mean share)
        wpr = (cp["size_x"] + 31) // 32
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                traffic = json.load(f).get("config5_sweep_kernel")
        except Exception:
            pass
        rec = {"workload": w.name, "scaling": "strong", "n_gpus": world, "dt_s": w.dt, "steps": steps, "warmup": warm,
               "us_per_cycle": 1e3 * float(tmax[0]) / steps, "cycles_per_s": steps / (float(tmax[0]) / 1e3),
               "active_voxels_swept_first_timed_cycle": int(vals[3].item()), "active_voxels_swept_last_timed_cycle": swept,
               "voxels_per_s": swept / (float(tmax[0]) / steps / 1e3),
               "sweep_kernel": {"us_max_rank": 1e3 * sweep_ms, "algorithmic_bytes_per_rank": algo,
                                "algorithmic_gbs": algo / (sweep_ms / 1e3) / 1e9, "algorithmic_frac_of_measured_peak": algo / (sweep_ms / 1e3) / 1e9 / peak,
                                "note": "16 B per active voxel (SURVEY §8d) over the sweep kernel's CUDA-event time; leaves that cannot change are decided from their 84 B header, so the DRAM bytes actually moved are fewer — see dram_measured",
                                "dram_measured": traffic},
               "merge": {"what": "packed cell bitmaps of each rank's own x interval, ncclSend/ncclRecv to rank 0 on a side stream, expanded to bytes on the root; pipelined one cycle behind",
                         "nvlink_bytes_per_cycle": int((world - 1) * (wpr * 4 * cp["size_y"]) / max(world, 1)) if world > 1 else 0,
                         "dense_window_bytes_per_cycle_r1": int(world * cp["size_x"] * cp["size_y"]) if world > 1 else 0},
               "costmap_cells": cp["size_x"] * cp["size_y"], "marked_cells_merged": merged_marked, "marked_cells_sum_of_ranks": int(vals[2].item()),
               "checks": {"every_voxel_swept": swept == int(vals[1].item()), "merged_equals_union_of_ranks": merged_marked == int(vals[2].item())},
               "load_s": load_s}
    barrier()
    if world > 1:
        g.comm_destroy()
    del costmap_dev, dev_clouds
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    g.close()
    return rec


def other_configs_record(args, stvl):
    """Configs 1, 3 and 4 (N = 1): device-resident fused cycles per second at full size, the CPU oracle on the same cycles, and a
    bit-exact comparison of the final state (voxels + mark times, per-column counts, last costmap bytes)."""
    import bench_configs as BC
    from spatio_temporal_voxel_layer_b200 import workloads
    out = {}
    for name, cls, kw in (("config1", workloads.Config1, dict(n_cycles=20)), ("config3", workloads.Config3, dict(n_cycles=20)),
                          ("config4", workloads.Config4, dict(n_cycles=8))):
        w = cls(**kw)
        warm = 5 if w.n_cycles >= 15 else 3
        gpu = BC.run_gpu(stvl, w, warmup=warm)
        ora = BC.run_oracle(w, warmup=warm)
        cmpr = BC.compare(gpu, ora)
        st = gpu["grid"].sweep_stats()
        out[name] = {"workload": w.name, "points_per_cycle": w.points_per_cycle, "cycles_timed": gpu["n_timed"], "dt_s": w.dt,
                     "cycles_per_s": gpu.get("cycles_per_s"), "us_per_cycle": gpu.get("us_per_cycle"), "gpu_launches": gpu.get("launches"),
                     "mark_bytes_per_cycle": w.points_per_cycle * 16,
                     "mark_bytes_rate_gbs": (w.points_per_cycle * 16 / (gpu["us_per_cycle"] * 1e-6) / 1e9) if gpu.get("us_per_cycle") else None,
                     "active_voxels_end": cmpr["n_voxels"], "swept_last_cycle": int(st.n_swept),
                     "cpu_oracle_cycles_per_s": len(ora["times"]) / sum(ora["times"]) if ora["times"] else None,
                     "parity_checked": bool(cmpr["voxels"] and cmpr["columns"] and cmpr["costmap"]), "parity": cmpr}
        gpu["grid"].close()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def sorted_voxels(v):
    ix, iy, iz, t = (np.asarray(a) for a in v)
    k = np.lexsort((iz, iy, ix))
    return ix[k], iy[k], iz[k], t[k]


def bench_ours(args):
    import torch
    import spatio_temporal_voxel_layer_b200 as stvl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(torch, local_rank)   # before any pinned allocation: first touch decides the node
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The per-cycle exchange is a few hundred KB of cell bitmaps per peer: one NCCL point-to-point channel (one copy CTA per
        # peer) moves that in a fraction of a cycle and leaves the SMs to the cycle's persistent kernels.
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = make_workload(rank, world)
    p = w.p
    vs = float(np.float32(w.voxel_size))
    leaf_w = int(round(SLAB_PITCH / (8 * vs)))      # slab width in leaves (200)
    grid_kw = dict(voxel_size=w.voxel_size, decay_model=w.decay_model, voxel_decay=w.voxel_decay, pub_voxels=False,
                   device=local_rank, leaf_capacity=1 << 20)
    if world > 1:   # spatial key-range shard: this rank keeps leaves with floor(bx / 200) mod N == rank
        grid_kw.update(shard_index=rank, shard_count=world, shard_width=leaf_w)
    g = stvl.SpatioTemporalVoxelGrid(**grid_kw)
    # plugin flow: UpdateROSCostmap only touch()es the cleared cells, i.e. needs their bounding box (always produced)
    g.set_outputs(0)

    ring = len(w.cycles)
    # clouds: pinned host copies (e2e leg) and device-resident copies (value leg)
    host_clouds = [[torch.from_numpy(c).pin_memory() for c in cyc.clouds] for cyc in w.cycles]
    dev_clouds = [[c.cuda(non_blocking=True) for c in cyc] for cyc in host_clouds]
    torch.cuda.synchronize()
    frustums = [np.concatenate([stvl.build_depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])
                                for o, q in cyc.poses]) for cyc in w.cycles]
    cp = costmap_geometry(w, world)
    params = stvl.CostmapParams(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                                cp["mark_threshold"], cp["default_value"], 254)
    costmap_dev = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    costmap_host = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8).pin_memory()
    costmap_host2 = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8).pin_memory()   # (two cycles are in flight in the e2e leg)
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
    params_ref = C.byref(params)
    cm_dev_ptr, cm_host_ptr = costmap_dev.data_ptr(), costmap_host.data_ptr()
    if world > 1:   # one NCCL communicator per handle, id created by rank 0
        uid = [stvl.SpatioTemporalVoxelGrid.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0], rank, world)

    def stamp(arr, now):
        for k in range(n_cams):
            arr[k].now = now + 1e-4 * k

    def merge_drain():
        if world > 1:
            check(L.stvl_costmap_sharded_flush(h, params_ref, cm_dev_ptr, 0))

    def step_device(j, dt=None, fused=True):
        """Cycle j of an episode with device-resident clouds.  N=1: ONE C-ABI call (stvl_update_cycle_device) enqueues the
        fused clear -> mark -> costmap cycle; nothing synchronises.  fused=False enqueues the same cycle stage by stage (used
        for the per-stage / per-kernel event timings).  N>1: the same fused cycle on this rank's shard + the in-library merge
        into rank 0's costmap, pipelined one cycle behind (stvl_update_cycle_sharded_device)."""
        now, slot = cycle_inputs(w, j, dt)
        arr = args_dev[slot]
        stamp(arr, now)
        if world == 1 and fused:
            check(L.stvl_update_cycle_device(h, now, fr_ptrs[slot], n_cams, arr, n_cams, params_ref, cm_dev_ptr))
        elif world == 1:
            check(L.stvl_clear_frustums(h, now, fr_ptrs[slot], n_cams))
            check(L.stvl_mark_device(h, arr, n_cams))
            check(L.stvl_costmap_device(h, params_ref, cm_dev_ptr, None))
        else:
            check(L.stvl_update_cycle_sharded_device(h, now, fr_ptrs[slot], n_cams, arr, n_cams, params_ref, cm_dev_ptr, 0, 1))

    def step_host(j):
        """Cycle j through the host-buffer entry points: clouds H2D and costmap bytes D2H inside the step."""
        now, slot = cycle_inputs(w, j)
        arr = args_host[slot]
        stamp(arr, now)
        check(L.stvl_clear_frustums(h, now, fr_ptrs[slot], n_cams))
        check(L.stvl_mark(h, arr, n_cams))
        if world == 1:
            res = stvl.CostmapResult()
            check(L.stvl_costmap(h, params_ref, cm_host_ptr, C.byref(res)))
            return
        check(L.stvl_costmap_sharded_device(h, params_ref, cm_dev_ptr, 0, 0, 0))   # blocking merge: this cycle's bytes
        if rank == 0:
            with torch.cuda.stream(stream):
                costmap_host.copy_(costmap_dev, non_blocking=True)
        g.synchronize()

    def submit_host(j):
        """Cycle j through the pipelined host-buffer entry point (stvl_update_cycle_async): pinned clouds H2D, fused cycle,
        costmap bytes D2H into a pinned host buffer — enqueued on three streams, returns a ticket."""
        now, slot = cycle_inputs(w, j)
        arr = args_host[slot]
        stamp(arr, now)
        t = C.c_uint64(0)
        check(L.stvl_update_cycle_async(h, now, fr_ptrs[slot], n_cams, arr, n_cams, params_ref,
                                        (costmap_host if j % 2 == 0 else costmap_host2).data_ptr(), C.byref(t)))
        return t.value

    def host_episode(n_cycles):
        """e2e episode: reset + seed + settle, then n_cycles with host buffers; wall-clock seconds of the n_cycles."""
        merge_drain()
        reseed()
        step_host(0)
        barrier()
        t0 = time.perf_counter()
        # two cycles in flight: cycle j+1's H2D copy overlaps cycle j's kernels and read-back; every cycle's bytes and result
        # are waited for (stvl_update_cycle_wait) inside the timed region.  N > 1: every rank uploads its own slab's clouds, the
        # ranks' cells are merged into rank 0 inside the cycle and rank 0 reads the merged map back.
        prev = None
        res = stvl.CostmapResult()
        for j in range(1, n_cycles + 1):
            t = submit_host(j)
            if prev is not None:
                check(L.stvl_update_cycle_wait(h, prev, C.byref(res)))
            prev = t
        check(L.stvl_update_cycle_wait(h, prev, C.byref(res)))
        barrier()
        return time.perf_counter() - t0

    def barrier():
        g.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        g.synchronize()
        torch.cuda.synchronize()

    def reseed(grid=None):
        grid = grid or g
        grid.ResetGrid()
        grid.load_voxels(*w.seed_vox, w.seed_t)

    host_enqueue_s = []   # per episode: wall-clock seconds the host spent enqueuing one cycle (this rank)

    def device_episode(n_cycles, dt=None, fused=True, timed=True):
        """reset + seed + settle + n_cycles; returns (CUDA-event ms of the n_cycles, kernel launches inside them)."""
        merge_drain()
        reseed()
        step_device(0, dt, fused)
        merge_drain()
        barrier()
        l0 = g.pool_stats().kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        th0 = time.perf_counter()
        for j in range(1, n_cycles + 1):
            step_device(j, dt, fused)
        host_enqueue_s.append((time.perf_counter() - th0) / max(n_cycles, 1))   # the calls only enqueue: host time per cycle
        merge_drain()                      # N>1: the last cycle's merge is part of the timed region
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), g.pool_stats().kernel_launches - l0

    # ---- value leg: inputs resident in HBM ---------------------------------------------------------------------------
    device_episode(min(args.warmup, EPISODE_CYCLES))           # warm-up episode (W cycles after the settle cycle)
    host_enqueue_s.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, launches, episode_ms = 0.0, 0, []
    n_seed = len(w.seed_vox[0])
    for n in episode_plan(args.steps):
        ms, nl = device_episode(n)
        dev_ms += ms
        launches += nl
        episode_ms.append(ms / n)
    host_enqueue_us = 1e6 * float(np.median(host_enqueue_s)) if host_enqueue_s else None
    st_end = g.sweep_stats()               # the last timed cycle's sweep
    active_end = g.active_count()
    # The timed region lasts a few hundred microseconds per episode — less than one nvidia-smi sampling period — so the
    # sampler keeps running while the same cycle loop soaks for ~0.3 s (untimed, results discarded, store re-seeded after).
    reseed()
    for j in range(6000):                  # (a fixed count: every rank must enqueue the same number of merges)
        step_device(j % EPISODE_CYCLES)
        if j % 100 == 99:
            g.synchronize()
    merge_drain()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "nvidia-smi -lms 100 over the timed episodes and an untimed 6000-cycle soak of the same loop on a re-seeded store (discarded)"

    # ---- per-stage / per-kernel timing on a fresh episode, stage by stage ----------------------------------------------
    t_all = None
    st_stage = None
    if world == 1:
        reseed()
        step_device(0, fused=False)
        barrier()
        g.enable_timing(stvl.TIME_SWEEP | stvl.TIME_MARK | stvl.TIME_COSTMAP)
        g.timing()
        for j in range(1, EPISODE_CYCLES + 1):
            step_device(j, fused=False)
        t_all = g.timing()
        g.enable_timing(0)
        st_stage = g.sweep_stats()
        # the Mark kernel on its own: a burst of back-to-back launches over the ring's clouds (275 MB > L2), CUDA events on the
        # library's stream around the burst — the kernel's average launch duration without the event-to-kernel latency that a
        # bracket around a single launch includes
        n_burst = 24
        def mark_burst(j0):
            for j in range(j0, j0 + n_burst):
                now, slot = cycle_inputs(w, EPISODE_CYCLES + 1 + j)      # clocks keep advancing: one batched launch per call
                arr = args_dev[slot]
                stamp(arr, now)
                check(L.stvl_mark_device(h, arr, n_cams))
        mark_burst(0)
        barrier()
        # the host needs longer to enqueue a launch than the kernel runs: a 2 x 2 GB fill goes first, so that the whole burst
        # is already queued when the GPU reaches the first event
        scratch = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
        eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            scratch.zero_(); scratch.zero_()
        eb0.record(stream)
        mark_burst(n_burst)
        eb1.record(stream)
        barrier()
        del scratch
        mark_burst_ms = eb0.elapsed_time(eb1) / n_burst

    # ---- sensitivity of the device-resident cycle to the cycle period (2-3 s of simulated time each, fresh seed) ----------
    sensitivity = {}
    if world == 1 and not args.no_sensitivity:
        for dt, n in ((0.01, 200), (1.0, 3)):
            ms, _ = device_episode(n, dt=dt)
            sensitivity["dt_%g_s" % dt] = {"cycles": n, "us_per_cycle": 1e3 * ms / n, "cycles_per_s": n / (ms / 1e3),
                                           "swept_last_cycle": int(g.sweep_stats().n_swept)}

    # ---- parity: the fused path's state after one more episode vs the oracle's (full config-2 size) -----------------------
    parity = None
    oracle_times = []
    if not args.no_parity:
        device_episode(EPISODE_CYCLES)
        if world > 1:
            dist.broadcast(costmap_dev, src=0)      # the merged costmap: every rank checks its own slab's cells
        gcm = costmap_dev.cpu().numpy()
        gv = sorted_voxels(g.voxels())
        gcols = g.columns()
        runner = OracleRunner(w, cp)
        n_ep = 1 if (world > 1 or args.no_cpu_baseline) else max(1, args.cpu_cycles // EPISODE_CYCLES)
        og = ocm = None
        for e in range(n_ep):
            t, og_e, ocm_e, _ = runner.episode(EPISODE_CYCLES)
            oracle_times += t
            if e == 0:
                og, ocm = og_e, ocm_e
        ov = sorted_voxels(og.voxels())
        ok_vox = all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(gv, ov))
        ox_, oy_, oc_ = og.costmap()         # GetFlattenedCostmap of the last sweep: world (x, y) doubles + count
        okk = np.lexsort((oy_, ox_))
        gx, gy = g.index_to_world(gcols[0]), g.index_to_world(gcols[1])
        gk = np.lexsort((gy, gx))
        ok_cols = (len(gx) == len(ox_) and np.array_equal(gx[gk].view(np.uint64), ox_[okk].view(np.uint64)) and
                   np.array_equal(gy[gk].view(np.uint64), oy_[okk].view(np.uint64)) and np.array_equal(gcols[2][gk], oc_[okk]))
        if world == 1:
            ok_cm = bool(np.array_equal(gcm, ocm))
        else:
            # x cells of this rank's slab in the merged costmap (slab r spans [80 r, 80 r + 80) m from slab 0's origin - 37.5 - 1)
            own = ocm == 254
            ok_cm = bool(np.array_equal(gcm[own], ocm[own]))
            counts = torch.tensor([int(own.sum())], device="cuda")
            dist.all_reduce(counts)
            ok_cm = ok_cm and int(counts.item()) == int((gcm == 254).sum())     # nothing marked that no slab's oracle marks
        ok = bool(ok_vox and ok_cols and ok_cm)
        if dist is not None:
            f = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            ok = bool(f.item())
        parity = {"parity_checked": ok, "voxels_compared": int(len(gv[0])), "columns_compared": int(len(gcols[0])),
                  "marked_cells": int((gcm == 254).sum()),
                  "what": "after an episode of %d fused cycles on the 710,765-voxel seed: costmap bytes, per-column counts (keys as f64 bit patterns) "
                          "and the voxel set with its mark times, all bit-exact against oracle/stvl_oracle.cpp on the same inputs%s"
                          % (EPISODE_CYCLES, "" if world == 1 else "; every rank checks its own slab against the merged costmap")}
        if not ok:
            parity["detail"] = {"voxels": bool(ok_vox), "columns": bool(ok_cols), "costmap": bool(ok_cm)}

    # ---- e2e leg: host clouds in, host costmap bytes out ---------------------------------------------------------------
    e2e_s = 0.0
    host_episode(min(args.warmup, EPISODE_CYCLES))
    for n in episode_plan(args.steps):
        e2e_s += host_episode(n)
    e2e_cm_ok = None
    if world == 1 and not args.no_parity and parity is not None:
        # the last e2e episode ran the same EPISODE_CYCLES cycles as the parity episode when K is a multiple of it: its
        # last read-back must be the same bytes
        if args.steps % EPISODE_CYCLES == 0:
            last = costmap_host if EPISODE_CYCLES % 2 == 0 else costmap_host2
            e2e_cm_ok = bool(np.array_equal(last.numpy(), gcm))
    e2e_active_end = g.active_count()

    # max over ranks
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
    value = world * args.steps / (dev_ms / 1e3)
    e2e = world * args.steps / e2e_s

    # tear the headline handle's buffers down before the other configs allocate theirs
    others = None
    if world == 1 and not args.no_other_configs:
        others = other_configs_record(args, stvl)
    c5 = None
    if not args.no_config5:
        c5 = config5_record(args, torch, stvl, dist, rank, world, local_rank)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": shared_config(w, cp, world),
            "details": {"seeded_voxels": int(n_seed), "active_voxels_end_of_last_timed_episode": int(active_end),
                        "swept_last_timed_cycle": int(st_end.n_swept), "cleared_last_timed_cycle": int(st_end.n_cleared),
                        "rewritten_last_timed_cycle": int(st_end.n_rewritten),
                        "e2e_active_voxels_end": int(e2e_active_end), "us_per_cycle_by_episode": [round(1e3 * v, 2) for v in episode_ms],
                        "leaf_pool_gb": g.pool_stats().device_bytes / 1e9,
                        "host_enqueue_us_per_cycle_rank0": None if host_enqueue_us is None else round(host_enqueue_us, 2),
                        "host_enqueue_note": "wall clock of the enqueue-only C-ABI call(s) of one cycle from Python (median over the timed episodes); "
                                             "at or above ms_per_step the cycle rate is set by the host, not by the kernels",
                        "parallelism": "1 GPU" if world == 1 else "%d spatial x-slab shards, fused cycle per shard + in-library merge of the per-shard costmap cells into rank 0, pipelined one cycle behind" % world},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bytes_h2d, "d2h_bytes_per_step": bytes_d2h,
                    "how": "stvl_update_cycle_async / stvl_update_cycle_wait, two cycles in flight (H2D of cycle k+1 overlaps kernels and D2H of cycle k); wall clock over the episodes' timed cycles, every cycle's bytes waited for"
                           + ("" if world == 1 else "; every rank uploads its own slab's clouds, in-cycle merge into rank 0, rank 0 reads the merged map back"),
                    "numa": numa_note,
                    "pcie_bound_cycles_per_s": 53e9 / bytes_h2d, "readback_matches_device_leg": e2e_cm_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if parity is not None:
            line["parity_checked"] = parity["parity_checked"]
            line["parity"] = parity
        if sensitivity:
            line["sensitivity"] = sensitivity
        if others is not None:
            line["other_configs"] = others
        if c5 is not None:
            line["config5"] = c5
        if t_all is not None and t_all.n_mark:
            mark_s = mark_burst_ms / 1e3
            algo_mark = w.points_per_cycle * 16          # 16 B read per input point
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                    traffic = json.load(f).get("mark_kernel_dram_bytes_per_launch")
            except Exception:
                pass
            line["roofline"] = {"bound": "hbm", "kernel": "mark_kernel", "achieved": algo_mark / mark_s / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": algo_mark / mark_s / 1e9 / peak, "traffic": traffic,
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_mark,
                                "kernel_ms": mark_burst_ms, "launches_timed": n_burst,
                                "kernel_ms_single_bracketed": t_all.mark_ms / t_all.n_mark,
                                "timing": "CUDA events on the library's stream around a burst of %d back-to-back stvl_mark_device launches over the "
                                          "ring's clouds (average launch duration); kernel_ms_single_bracketed = events around single launches in the "
                                          "stage-by-stage pass, which adds the event-to-kernel latency (in the fused cycle Mark overlaps the sweep and "
                                          "has no duration of its own)" % n_burst}
            if t_all.n_sweep and t_all.n_costmap:
                sweep_s = t_all.sweep_ms / t_all.n_sweep / 1e3
                algo_clear = int(st_stage.n_swept) * 16      # 16 B per active voxel (key + time)
                line["roofline"]["stage_ms"] = {"clear": t_all.sweep_ms / t_all.n_sweep, "mark": t_all.mark_ms / t_all.n_mark,
                                                "costmap": t_all.costmap_ms / t_all.n_costmap,
                                                "note": "stage-by-stage pass (stvl_clear_frustums, stvl_mark_device, stvl_costmap_device) on a freshly seeded episode, serialised; the fused cycle of `value` overlaps them"}
                line["roofline"]["sweep_kernels"] = {"achieved": algo_clear / sweep_s / 1e9, "frac": algo_clear / sweep_s / 1e9 / peak,
                                                     "algorithmic_bytes_per_launch": int(algo_clear), "swept_voxels": int(st_stage.n_swept)}
        if world == 1 and not args.no_cpu_baseline and oracle_times:
            model, ncpu = host_info()
            line["cpu_baseline"] = {"value": len(oracle_times) / sum(oracle_times), "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d config-2 cycles (episodes of %d on the re-seeded store), oracle/stvl_oracle.cpp single thread (the reference path is single-threaded); host %s x%s"
                                              % (len(oracle_times), EPISODE_CYCLES, model, ncpu)}
        print(json.dumps(line))
    # tear down in dependency order: tensors that were used on the library's stream are released while that stream is
    # still alive (the caching allocator records events on it), then the process group, then the grid
    barrier()
    if world > 1:
        merge_drain()
        torch.cuda.synchronize()
        g.comm_destroy()
    del costmap_dev, costmap_host, costmap_host2, dev_clouds, host_clouds
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


def main():
    if os.environ.get("STVL_BENCH_WATCHDOG"):   # debugging aid: dump every thread's Python stack and exit if the run takes too long
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["STVL_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-cycles", type=int, default=80)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-sensitivity", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs 1, 3, 4 records (N = 1)")
    ap.add_argument("--no-config5", action="store_true", help="skip the strong-scaled config-5 record")
    ap.add_argument("--config5-voxels", type=int, default=50_000_000)
    ap.add_argument("--config5-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
