"""The C++ host side (spatio_temporal_voxel_layer_b200/host): the drop-in SpatioTemporalVoxelGrid façade and the
SpatioTemporalVoxelLayer plugin flow, driven through a scenario file by stvl_host_demo.

CPU: the host code compiles with plain g++ against the shim headers, links libstvl_b200.so and refuses to run without
a GPU.  GPU: a multi-cycle scenario through updateBounds/updateCosts reproduces the oracle's costmap bytes, bounds and
flattened costmap — with the fused GPU costmap path and with the reference's literal host loops."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "spatio_temporal_voxel_layer_b200", "host")
DEMO = os.path.join(HOST, "stvl_host_demo")


def build_host():
    import spatio_temporal_voxel_layer_b200 as pkg
    pkg.load_library()
    subprocess.check_call(["make", "-C", HOST, "-s"])
    assert os.path.exists(DEMO)


def write_scenario(path, cfg, sources, cycles):
    with open(path, "wb") as f:
        f.write(b"STVLSCN1")
        f.write(struct.pack("<fidiii", cfg["voxel_size"], cfg["decay_model"], cfg["voxel_decay"], cfg["mark_threshold"],
                            cfg["track_unknown"], cfg["publish"]))
        f.write(struct.pack("<dddII", cfg["origin_x"], cfg["origin_y"], cfg["resolution"], cfg["size_x"], cfg["size_y"]))
        f.write(struct.pack("<I", len(sources)))
        for s in sources:
            f.write(struct.pack("<9d", s["obstacle_range"], s["min_z"], s["max_z"], s["vfov"], s["vpad"], s["hfov"], s["accel"],
                                s["min_h"], s["max_h"]))
            f.write(struct.pack("<4i", s["marking"], s["clearing"], s["model_type"], s["filter"]))
        f.write(struct.pack("<I", len(cycles)))
        for now, readings in cycles:
            f.write(struct.pack("<d", now))
            for (o, q, cloud) in readings:
                f.write(struct.pack("<7d", *o, *q))
                c = np.ascontiguousarray(cloud, np.float32)
                f.write(struct.pack("<I", len(c)))
                f.write(c.tobytes())


def read_output(path, n_cells):
    out = []
    with open(path, "rb") as f:
        assert f.read(8) == b"STVLOUT1"
        (n_cycles,) = struct.unpack("<I", f.read(4))
        for _ in range(n_cycles):
            (cells,) = struct.unpack("<I", f.read(4))
            cm = np.frombuffer(f.read(cells), np.uint8).copy()
            b = np.array(struct.unpack("<4d", f.read(32)))
            (ncol,) = struct.unpack("<Q", f.read(8))
            cols = np.frombuffer(f.read(ncol * 20), np.dtype([("x", "<f8"), ("y", "<f8"), ("c", "<u4")])).copy()
            (npts,) = struct.unpack("<Q", f.read(8))
            out.append((cm, b, cols, npts))
        master = np.frombuffer(f.read(n_cells), np.uint8).copy()
        (saved,) = struct.unpack("<I", f.read(4))
    return out, master, saved


def test_stvl2pc_converter(tmp_path):
    """stvl2pc (pure host code) on a hand-written .stvl file: lower corners like vdb2pc, centres and times on request."""
    subprocess.check_call(["make", "-C", HOST, "-s", "stvl2pc"])
    vs = float(np.float32(0.05))
    rec = np.array([(-3, 2, 7, 100.25), (0, 0, 0, -1.5), (123456, -654321, 40, 1.7e9)], np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("t", "<f8")]))
    path = tmp_path / "g.stvl"
    with open(path, "wb") as f:
        f.write(b"STVLGRD1" + struct.pack("<dQ", vs, len(rec)))
        for r in rec:
            f.write(struct.pack("<iiid", int(r["x"]), int(r["y"]), int(r["z"]), float(r["t"])))
    tool = os.path.join(HOST, "stvl2pc")
    out = tmp_path / "g.pcd"
    assert subprocess.run([tool, str(path), str(out)]).returncode == 0
    lines = open(out).read().splitlines()
    assert "FIELDS x y z" in lines and "POINTS 3" in lines
    pts = np.loadtxt(out, skiprows=11, dtype=np.float32).reshape(-1, 3)
    want = np.stack([(rec[k].astype(np.float64) * vs).astype(np.float32) for k in ("x", "y", "z")], 1)
    assert np.array_equal(pts, want)
    assert subprocess.run([tool, str(path), str(out), "--centres", "--times"]).returncode == 0
    full = np.loadtxt(out, skiprows=11).reshape(-1, 4)
    assert np.array_equal(full[:, :3].astype(np.float32), np.stack([(rec[k].astype(np.float64) * vs + vs / 2).astype(np.float32) for k in ("x", "y", "z")], 1))
    assert np.array_equal(full[:, 3], rec["t"])
    bad = tmp_path / "bad.stvl"
    bad.write_bytes(b"NOTAGRID" + b"\0" * 16)
    assert subprocess.run([tool, str(bad), str(out)], capture_output=True).returncode == 1


def test_host_code_builds_and_refuses_without_gpu(tmp_path, stvl):
    build_host()
    scen = tmp_path / "s.bin"
    cfg = dict(voxel_size=0.05, decay_model=0, voxel_decay=15.0, mark_threshold=0, track_unknown=0, publish=0,
               origin_x=-5.0, origin_y=-5.0, resolution=0.05, size_x=200, size_y=200)
    write_scenario(str(scen), cfg, [], [])
    r = subprocess.run([DEMO, str(scen), str(tmp_path / "o.bin")], capture_output=True, text=True)
    if stvl.device_count() == 0:
        assert r.returncode == 3, r.stderr
        assert "no CPU path" in r.stderr or "no CUDA device" in r.stderr
    else:
        assert r.returncode == 0, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("literal", [False, True])
def test_plugin_cycle_matches_oracle(tmp_path, stvl, oracle, literal):
    build_host()
    from spatio_temporal_voxel_layer_b200 import synth
    p = synth.RGBD
    scene = synth.retail_scene(size=24.0, seed=11)
    cfg = dict(voxel_size=0.05, decay_model=0, voxel_decay=1.5, mark_threshold=1, track_unknown=1, publish=1,
               origin_x=-9.0, origin_y=-9.0, resolution=0.05, size_x=360, size_y=360)
    n_cams = 3
    sources = [dict(obstacle_range=p["obstacle_range"], min_z=p["min_z"], max_z=p["max_z"], vfov=p["vfov"], vpad=0.0,
                    hfov=p["hfov"], accel=p["accel"], min_h=p["min_obstacle_height"], max_h=p["max_obstacle_height"],
                    marking=1, clearing=1, model_type=0, filter=2) for _ in range(n_cams)]
    cycles = []
    for c in range(8):
        now = 200.0 + 0.4 * c
        centre = (-4.0 + 0.6 * c, 0.8, 0.0)
        readings = []
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.15 * c, n_cams=n_cams)):
            readings.append((o, tuple(q), synth.depth_cloud(scene, o, q, width=160, height=120, seed=50 * c + k)))
        cycles.append((now, readings))
    scen, outp = tmp_path / "s.bin", tmp_path / "o.bin"
    write_scenario(str(scen), cfg, sources, cycles)
    r = subprocess.run([DEMO, str(scen), str(outp)] + (["--literal"] if literal else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got, master, saved = read_output(str(outp), cfg["size_x"] * cfg["size_y"])
    assert saved == 1 and os.path.exists(str(outp) + ".grid.stvl")

    og = oracle.OracleGrid(0.05, 255.0, 0, 1.5, True)
    master_ref = np.zeros(cfg["size_x"] * cfg["size_y"], np.uint8)   # the demo's master grid starts at 0
    footprints = []
    for (now, readings), (cm, b, cols, npts) in zip(cycles, got):
        fr = np.concatenate([oracle.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q, _ in readings])
        og.clear_frustums(now, fr)
        for (o, q, cloud) in readings:
            og.mark(oracle.filter_passthrough(cloud, p["min_obstacle_height"], p["max_obstacle_height"]), o, p["obstacle_range"], now)
        ocm, ob, on = og.update_ros_costmap(cfg["origin_x"], cfg["origin_y"], cfg["resolution"], cfg["size_x"], cfg["size_y"],
                                            cfg["mark_threshold"], 255)
        # updateFootprint (reference spatio_temporal_voxel_layer.cpp:529-548): the demo's 0.7 m x 0.5 m footprint at the
        # first source's origin, yaw 0.3 * cycle, touches the bounds; updateCosts (:682-684) frees the cells under it
        cyc = len(footprints)
        rx, ry, yaw = readings[0][0][0], readings[0][0][1], 0.3 * cyc
        fp = [(rx + (px * math.cos(yaw) - py * math.sin(yaw)), ry + (px * math.sin(yaw) + py * math.cos(yaw)))
              for px, py in ((0.35, 0.25), (-0.35, 0.25), (-0.35, -0.25), (0.35, -0.25))]
        footprints.append(fp)
        for x, y in fp:
            ob = np.array([min(ob[0], x), min(ob[1], y), max(ob[2], x), max(ob[3], y)])
        assert oracle.set_convex_polygon_cost(ocm, cfg["origin_x"], cfg["origin_y"], cfg["resolution"], fp, 0)
        assert np.array_equal(cm, ocm.reshape(-1)), "layer costmap bytes differ"
        np.testing.assert_array_equal(b, ob)
        ox, oy, oc = og.costmap()
        a = np.sort(cols, order=["x", "y"])
        ref = np.sort(np.rec.fromarrays([ox, oy, oc], names="x,y,c"), order=["x", "y"])
        assert len(a) == len(ref)
        assert np.array_equal(a["x"], ref["x"]) and np.array_equal(a["y"], ref["y"]) and np.array_equal(a["c"], ref["c"])
        assert npts == len(og.points())
        # updateCosts with combination_method 1 = updateWithMax, NO_INFORMATION cells of the layer are skipped
        layer = ocm.reshape(-1)
        upd = (layer != 255) & (master_ref < layer)
        master_ref[upd] = layer[upd]
    assert np.array_equal(master, master_ref)
    # the saved grid equals the oracle's voxels
    raw = open(str(outp) + ".grid.stvl", "rb").read()
    assert raw[:8] == b"STVLGRD1"
    vs, n = struct.unpack("<dQ", raw[8:24])
    rec = np.frombuffer(raw[24:], np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("t", "<f8")]))
    assert len(rec) == n == og.active_count()
    ix, iy, iz, t = og.voxels()
    a = np.sort(rec, order=["x", "y", "z"])
    o = np.sort(np.rec.fromarrays([ix, iy, iz, t], names="x,y,z,t"), order=["x", "y", "z"])
    assert np.array_equal(a["x"], o["x"]) and np.array_equal(a["z"], o["z"]) and np.array_equal(a["t"], o["t"])
    # SaveGrid -> LoadGrid -> SaveGrid round trip (the demo reloads the saved grid and saves it again): same voxels, same times
    raw2 = open(str(outp) + ".grid2.stvl", "rb").read()
    rec2 = np.frombuffer(raw2[24:], rec.dtype)
    assert raw2[:24] == raw[:24] and np.array_equal(np.sort(rec2, order=["x", "y", "z"]), a)
    # ... and the converter (the counterpart of the reference's vdb2pc) lists every voxel's lower corner, coord * voxel_size
    pcd = str(outp) + ".pcd"
    r2 = subprocess.run([os.path.join(HOST, "stvl2pc"), str(outp) + ".grid.stvl", pcd], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr
    pts = np.loadtxt(pcd, skiprows=11, dtype=np.float32).reshape(-1, 3)
    want = np.stack([(rec[k].astype(np.float64) * vs).astype(np.float32) for k in ("x", "y", "z")], 1)
    assert np.array_equal(pts, want)
