"""Certifies the committed golden fixtures with the REFERENCE'S OWN CODE and writes tests/golden/ref_certified.json.

Runs only where /root/reference exists (the authoring container): oracle/_ref/libstvl_ref.so is the reference's
spatio_temporal_voxel_grid.cpp + frustum models compiled unmodified against stand-in third-party headers (oracle/ref_driver.cpp).
Every cycle of cycle_small.npz and cycle_small_c4.npz is replayed on it — ClearFrustums with the same readings, Mark with the
same (pre-filtered) clouds — and the per-column counts (GetFlattenedCostmap) of every cycle and the final voxel store with its
mark times must equal the fixture bit for bit.  The fixtures' costmap BYTES additionally involve nav2's worldToMap and the
fixtures' clouds of the second scenario PCL's VoxelGrid / tf2's transform, which stay restated (oracle/stvl_oracle.cpp).
The JSON records the sha256 of the fixture files that passed, so the GPU box (no reference there) can tell that the vectors
its GPU tests compare against are the certified ones.

    python tests/golden/make_ref_golden.py
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O, ref_py as R      # noqa: E402
from spatio_temporal_voxel_layer_b200 import synth  # noqa: E402


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def ref_readings(p, v, readings):
    out = []
    for k, r in enumerate(readings):
        o, q = r[0], r[1]
        if k < 2:
            out.append((0, R.reading_params(p["vfov"], 0.0, p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"])))
        else:
            out.append((1, R.reading_params(v["vfov"], v["vpad"], v["hfov"], v["min_z"], v["max_z"], o, q, v["accel"])))
    return out


def check_cycle(rg, g, c):
    x, y, cnt = rg.costmap()
    order = np.lexsort((y, x))
    ok = (np.array_equal(x[order].view(np.uint64), g["col_x_%d" % c].view(np.uint64)) and
          np.array_equal(y[order].view(np.uint64), g["col_y_%d" % c].view(np.uint64)) and
          np.array_equal(cnt[order], g["col_c_%d" % c]))
    return bool(ok), len(x)


def check_voxels(rg, g):
    ix, iy, iz, t = rg.voxels()
    order = np.lexsort((iz, iy, ix))
    ok = (np.array_equal(ix[order], g["vox_ix"]) and np.array_equal(iy[order], g["vox_iy"]) and
          np.array_equal(iz[order], g["vox_iz"]) and np.array_equal(t[order].view(np.uint64), g["vox_t"].view(np.uint64)))
    return bool(ok), len(ix)


def certify_small():
    mg = _mg()
    path = os.path.join(HERE, "cycle_small.npz")
    g = np.load(path)
    p, v = synth.RGBD, synth.VLP16
    rg = R.RefGrid(0.05, 255.0, 0, 2.0, True)
    n_cols = 0
    for c in range(int(g["n_cycles"])):
        now = float(g["now_%d" % c])
        readings = [(g["origin_%d_%d" % (c, k)], g["quat_%d_%d" % (c, k)], g["cloud_%d_%d" % (c, k)]) for k in range(3)]
        rg.clear_frustums(now, ref_readings(p, v, readings))
        ok, n = check_cycle(rg, g, c)
        assert ok, "cycle_small: column counts of cycle %d differ from the reference's" % c
        n_cols += n
        for k, (o, q, cloud) in enumerate(readings):
            lim = p if k < 2 else v
            rg.mark(O.filter_passthrough(cloud, lim["min_obstacle_height"], lim["max_obstacle_height"]), o, lim["obstacle_range"], now + 1e-3 * k)
    ok, n_vox = check_voxels(rg, g)
    assert ok, "cycle_small: final voxel store differs from the reference's"
    return path, dict(cycles=int(g["n_cycles"]), columns_compared=n_cols, voxels_compared=n_vox)


def certify_c4():
    mg = _mg()
    path = os.path.join(HERE, "cycle_small_c4.npz")
    g = np.load(path)
    p, v = synth.RGBD, synth.VLP16
    gp = mg.C4_GRID
    rg = R.RefGrid(gp["voxel_size"], gp["background"], gp["decay_model"], gp["voxel_decay"], False)
    leaf = np.float32(gp["voxel_size"])
    n_cols = 0
    for c in range(int(g["n_cycles"])):
        now = float(g["now_%d" % c])
        readings = []
        for k in range(3):
            key = "tf_%d_%d" % (c, k)
            readings.append((g["origin_%d_%d" % (c, k)], g["quat_%d_%d" % (c, k)], g["cloud_%d_%d" % (c, k)], g[key] if key in g.files else None))
        rg.clear_frustums(now, ref_readings(p, v, readings))
        ok, n = check_cycle(rg, g, c)
        assert ok, "cycle_small_c4: column counts of cycle %d differ from the reference's" % c
        n_cols += n
        for k, (o, q, cloud, m) in enumerate(readings):
            if m is not None:   # tf2 transform + PCL VoxelGrid stay restated (not part of the grid's translation units)
                pts = O.filter_voxel(O.transform_cloud(cloud, m), leaf, p["min_obstacle_height"], p["max_obstacle_height"], 1)
                rg.mark(pts, o, p["obstacle_range"], now + 1e-3 * k)
            else:
                rg.mark(O.filter_passthrough(cloud, v["min_obstacle_height"], v["max_obstacle_height"]), o, v["obstacle_range"], now + 1e-3 * k)
    ok, n_vox = check_voxels(rg, g)
    assert ok, "cycle_small_c4: final voxel store differs from the reference's"
    return path, dict(cycles=int(g["n_cycles"]), columns_compared=n_cols, voxels_compared=n_vox)


def main():
    R.build(force=True)
    cert = {"how": "python tests/golden/make_ref_golden.py — fixtures replayed on oracle/_ref/libstvl_ref.so (the reference's own "
                   "spatio_temporal_voxel_grid.cpp, depth_camera_frustum.cpp, three_dimensional_lidar_frustum.cpp, unmodified, "
                   "against stand-in OpenVDB / Eigen / ROS headers); per-cycle column counts and the final voxel store bit-exact",
            "fixtures": {}}
    for fn in (certify_small, certify_c4):
        path, info = fn()
        info["sha256"] = hashlib.sha256(open(path, "rb").read()).hexdigest()
        cert["fixtures"][os.path.basename(path)] = info
        print("certified", os.path.basename(path), info)
    with open(os.path.join(HERE, "ref_certified.json"), "w") as f:
        json.dump(cert, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
