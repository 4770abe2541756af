"""Generates tests/golden/cycle_small.npz and cycle_small_c4.npz: small seeded multi-cycle scenarios (inputs) and what the
CPU oracle produces for them (outputs).  The reference ships no golden vectors, so this file pins OUR oracle against
regressions and gives the GPU tests a fixture that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O                       # noqa: E402
from spatio_temporal_voxel_layer_b200 import synth      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cycle_small.npz")
CM = dict(origin_x=-6.0, origin_y=-6.0, resolution=0.05, size_x=240, size_y=240, mark_threshold=1, default_value=255)


def scenario():
    p = synth.RGBD
    scene = synth.retail_scene(size=16.0, seed=4)
    cycles = []
    for c in range(4):
        now = 1000.0 + 0.7 * c
        centre = (-2.0 + 0.5 * c, 0.8, 0.0)
        readings = []
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.2 * c, n_cams=2)):
            readings.append((np.array(o), np.array(q), synth.depth_cloud(scene, o, q, width=80, height=60, seed=9 * c + k)))
        lo = np.array([centre[0], centre[1], 0.9])
        lq = synth.yaw_quat(0.1 * c)
        readings.append((lo, lq, synth.lidar_cloud(scene, lo, lq, rings=8, azimuths=360, seed=c)))
        cycles.append((now, readings))
    return p, cycles


def run_oracle(p, cycles):
    v = synth.VLP16
    og = O.OracleGrid(0.05, 255.0, 0, 2.0, True)
    outs = []
    for now, readings in cycles:
        blocks = [O.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q, _ in readings[:2]]
        lo, lq, _ = readings[2]
        blocks.append(O.lidar_frustum(v["vfov"], v["vpad"], v["hfov"], v["min_z"], v["max_z"], lo, lq, v["accel"]))
        og.clear_frustums(now, np.concatenate(blocks))
        for k, (o, q, cloud) in enumerate(readings):
            lim = (p if k < 2 else v)
            og.mark(O.filter_passthrough(cloud, lim["min_obstacle_height"], lim["max_obstacle_height"]), o, lim["obstacle_range"], now + 1e-3 * k)
        cm, b, n = og.update_ros_costmap(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"],
                                         CM["mark_threshold"], CM["default_value"])
        x, y, c = og.costmap()
        order = np.lexsort((y, x))
        outs.append(dict(costmap=cm, bounds=b, col_x=x[order], col_y=y[order], col_c=c[order]))
    ix, iy, iz, t = og.voxels()
    order = np.lexsort((iz, iy, ix))
    return outs, (ix[order], iy[order], iz[order], t[order])


# ---- second fixture: 0.02 m voxels, exponential decay, sensor-frame clouds + tf transform + VOXEL pre-filter ------------
OUT_C4 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cycle_small_c4.npz")
CM_C4 = dict(origin_x=-4.0, origin_y=-4.0, resolution=0.04, size_x=200, size_y=200, mark_threshold=2, default_value=0)
C4_GRID = dict(voxel_size=0.02, background=0.0, decay_model=1, voxel_decay=1.2)


def scenario_c4():
    """3 cycles; per cycle two depth cameras whose clouds arrive in the SENSOR frame (the reading carries the float affine
    tf2::doTransform would apply) and go through the VOXEL pre-filter (min 1 point per leaf), plus a lidar in the global
    frame with the PASSTHROUGH window.  The robot sits near the origin, so x and y change sign inside every cloud."""
    scene = synth.retail_scene(size=8.0, seed=12)
    cycles = []
    for c in range(3):
        now = 500.0 + 0.45 * c
        centre = (-0.3 + 0.25 * c, 0.2, 0.0)
        readings = []
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.3 * c + 0.1, n_cams=2)):
            glob = synth.depth_cloud(scene, o, q, width=64, height=48, seed=40 * c + k)
            m = O.transform_from_quat(o, q)
            rot = m.reshape(3, 4)[:, :3].astype(np.float64)
            sensor = glob.copy()
            sensor[:, :3] = ((glob[:, :3].astype(np.float64) - np.asarray(o)) @ rot).astype(np.float32)
            readings.append((np.array(o), np.array(q), sensor, m))
        lo = np.array([centre[0], centre[1], 0.7])
        lq = synth.yaw_quat(0.2 * c)
        readings.append((lo, lq, synth.lidar_cloud(scene, lo, lq, rings=8, azimuths=240, seed=70 + c), None))
        cycles.append((now, readings))
    return synth.RGBD, cycles


def run_oracle_c4(p, cycles):
    v = synth.VLP16
    og = O.OracleGrid(C4_GRID["voxel_size"], C4_GRID["background"], C4_GRID["decay_model"], C4_GRID["voxel_decay"], False)
    leaf = np.float32(C4_GRID["voxel_size"])
    outs = []
    for now, readings in cycles:
        blocks = [O.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q, _, _ in readings[:2]]
        lo, lq, _, _ = readings[2]
        blocks.append(O.lidar_frustum(v["vfov"], v["vpad"], v["hfov"], v["min_z"], v["max_z"], lo, lq, v["accel"]))
        og.clear_frustums(now, np.concatenate(blocks))
        for k, (o, q, cloud, m) in enumerate(readings):
            if m is not None:
                pts = O.filter_voxel(O.transform_cloud(cloud, m), leaf, p["min_obstacle_height"], p["max_obstacle_height"], 1)
                og.mark(pts, o, p["obstacle_range"], now + 1e-3 * k)
            else:
                og.mark(O.filter_passthrough(cloud, v["min_obstacle_height"], v["max_obstacle_height"]), o, v["obstacle_range"], now + 1e-3 * k)
        cm, b, n = og.update_ros_costmap(CM_C4["origin_x"], CM_C4["origin_y"], CM_C4["resolution"], CM_C4["size_x"], CM_C4["size_y"],
                                         CM_C4["mark_threshold"], CM_C4["default_value"])
        x, y, c = og.costmap()
        order = np.lexsort((y, x))
        outs.append(dict(costmap=cm, bounds=b, col_x=x[order], col_y=y[order], col_c=c[order]))
    ix, iy, iz, t = og.voxels()
    order = np.lexsort((iz, iy, ix))
    return outs, (ix[order], iy[order], iz[order], t[order])


def write(out_path, cycles, outs, vox):
    d = {"n_cycles": len(cycles)}
    for c, (now, readings) in enumerate(cycles):
        d["now_%d" % c] = now
        for k, r in enumerate(readings):
            d["origin_%d_%d" % (c, k)] = r[0]
            d["quat_%d_%d" % (c, k)] = r[1]
            d["cloud_%d_%d" % (c, k)] = r[2]
            if len(r) > 3 and r[3] is not None:
                d["tf_%d_%d" % (c, k)] = r[3]
        d["costmap_%d" % c] = np.packbits(outs[c]["costmap"] == 254)
        d["bounds_%d" % c] = outs[c]["bounds"]
        for key in ("col_x", "col_y", "col_c"):
            d["%s_%d" % (key, c)] = outs[c][key]
    d["vox_ix"], d["vox_iy"], d["vox_iz"], d["vox_t"] = vox
    np.savez_compressed(out_path, **d)
    print("wrote", out_path, os.path.getsize(out_path), "bytes;", len(vox[0]), "voxels")


def main():
    p4, cycles4 = scenario_c4()
    outs4, vox4 = run_oracle_c4(p4, cycles4)
    write(OUT_C4, cycles4, outs4, vox4)
    p, cycles = scenario()
    outs, vox = run_oracle(p, cycles)
    d = {"n_cycles": len(cycles)}
    for c, (now, readings) in enumerate(cycles):
        d["now_%d" % c] = now
        for k, (o, q, cloud) in enumerate(readings):
            d["origin_%d_%d" % (c, k)] = o
            d["quat_%d_%d" % (c, k)] = q
            d["cloud_%d_%d" % (c, k)] = cloud
        d["costmap_%d" % c] = np.packbits(outs[c]["costmap"] == 254)
        d["bounds_%d" % c] = outs[c]["bounds"]
        for key in ("col_x", "col_y", "col_c"):
            d["%s_%d" % (key, c)] = outs[c][key]
    d["vox_ix"], d["vox_iy"], d["vox_iz"], d["vox_t"] = vox
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(vox[0]), "voxels")


if __name__ == "__main__":
    main()
