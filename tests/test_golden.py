"""Golden-fixture tests: tests/golden/cycle_small.npz and cycle_small_c4.npz (made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces it (regression pin).  GPU: the CUDA path reproduces it without consulting the oracle."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cycle_small.npz")


def load():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg, np.load(GOLD)


def cycles_from(g):
    out = []
    for c in range(int(g["n_cycles"])):
        readings = [(g["origin_%d_%d" % (c, k)], g["quat_%d_%d" % (c, k)], g["cloud_%d_%d" % (c, k)]) for k in range(3)]
        out.append((float(g["now_%d" % c]), readings))
    return out


def test_oracle_reproduces_golden(oracle):
    mg, g = load()
    from spatio_temporal_voxel_layer_b200 import synth
    cycles = cycles_from(g)
    outs, vox = mg.run_oracle(synth.RGBD, cycles)
    for c, o in enumerate(outs):
        assert np.array_equal(np.packbits(o["costmap"] == 254), g["costmap_%d" % c])
        np.testing.assert_array_equal(o["bounds"], g["bounds_%d" % c])
        for key in ("col_x", "col_y", "col_c"):
            assert np.array_equal(o[key], g["%s_%d" % (key, c)])
    for a, key in zip(vox, ("vox_ix", "vox_iy", "vox_iz", "vox_t")):
        assert np.array_equal(a, g[key])
    # generated inputs are reproducible too (the fixture can be rebuilt bit-identically)
    p, regen = mg.scenario()
    for (now, rs), (now2, rs2) in zip(cycles, regen):
        assert now == now2
        for (o, q, cl), (o2, q2, cl2) in zip(rs, rs2):
            assert np.array_equal(cl.view(np.uint32), cl2.view(np.uint32))


@pytest.mark.gpu
def test_gpu_reproduces_golden(stvl):
    mg, g = load()
    from spatio_temporal_voxel_layer_b200 import synth
    p, v = synth.RGBD, synth.VLP16
    cm = mg.CM
    gg = stvl.SpatioTemporalVoxelGrid(0.05, 255.0, 0, 2.0, pub_voxels=True)
    for c, (now, readings) in enumerate(cycles_from(g)):
        rs = []
        for k, (o, q, cloud) in enumerate(readings):
            lim = p if k < 2 else v
            rs.append(stvl.MeasurementReading(cloud, origin=o, orientation=q, obstacle_range=lim["obstacle_range"],
                                              min_z=lim["min_z"], max_z=lim["max_z"], vfov=lim["vfov"],
                                              vfov_padding=lim.get("vpad", 0.0), hfov=lim["hfov"], decay_acceleration=lim["accel"],
                                              model_type=0 if k < 2 else 1, now=now + 1e-3 * k, filter=stvl.FILTER_PASSTHROUGH,
                                              min_obstacle_height=lim["min_obstacle_height"], max_obstacle_height=lim["max_obstacle_height"]))
        gg.ClearFrustums(rs, now)
        gg.Mark(rs)
        out, res = gg.UpdateROSCostmap(cm["origin_x"], cm["origin_y"], cm["resolution"], cm["size_x"], cm["size_y"],
                                       cm["mark_threshold"], cm["default_value"])
        assert np.array_equal(np.packbits(out == 254), g["costmap_%d" % c]), "cycle %d" % c
        assert set(np.unique(out)) <= {254, 255}
        x, y, cnt = gg.GetFlattenedCostmap()
        order = np.lexsort((y, x))
        assert np.array_equal(x[order], g["col_x_%d" % c]) and np.array_equal(y[order], g["col_y_%d" % c])
        assert np.array_equal(cnt[order], g["col_c_%d" % c])
        st = gg.sweep_stats()
        b = [1e308, 1e308, -1e308, -1e308]
        for box, ok in ((res.touched, res.has_bounds), (st.cleared_bbox_world, st.n_cleared > 0)):
            if ok:
                b = [min(b[0], box[0]), min(b[1], box[1]), max(b[2], box[2]), max(b[3], box[3])]
        np.testing.assert_array_equal(np.array(b), g["bounds_%d" % c])
    ix, iy, iz, t = gg.voxels()
    order = np.lexsort((iz, iy, ix))
    assert np.array_equal(ix[order], g["vox_ix"]) and np.array_equal(iy[order], g["vox_iy"]) and np.array_equal(iz[order], g["vox_iz"])
    assert np.array_equal(t[order], g["vox_t"])
    gg.close()


# ---- second fixture: 0.02 m voxels, exponential decay, sensor-frame clouds + transform + VOXEL pre-filter ------------------
GOLD_C4 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cycle_small_c4.npz")


def cycles_c4_from(g):
    out = []
    for c in range(int(g["n_cycles"])):
        readings = []
        for k in range(3):
            key = "tf_%d_%d" % (c, k)
            readings.append((g["origin_%d_%d" % (c, k)], g["quat_%d_%d" % (c, k)], g["cloud_%d_%d" % (c, k)], g[key] if key in g.files else None))
        out.append((float(g["now_%d" % c]), readings))
    return out


def test_oracle_reproduces_golden_c4(oracle):
    mg, _ = load()
    g = np.load(GOLD_C4)
    from spatio_temporal_voxel_layer_b200 import synth
    outs, vox = mg.run_oracle_c4(synth.RGBD, cycles_c4_from(g))
    assert sum(int(np.unpackbits(g["costmap_%d" % c]).sum()) for c in range(int(g["n_cycles"]))) > 50
    for c, o in enumerate(outs):
        assert np.array_equal(np.packbits(o["costmap"] == 254), g["costmap_%d" % c])
        np.testing.assert_array_equal(o["bounds"], g["bounds_%d" % c])
        for key in ("col_x", "col_y", "col_c"):
            assert np.array_equal(o[key], g["%s_%d" % (key, c)])
    for a, key in zip(vox, ("vox_ix", "vox_iy", "vox_iz", "vox_t")):
        assert np.array_equal(a, g[key])


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_gpu_reproduces_golden_c4(stvl, fused):
    """The CUDA path on the second fixture, without the oracle: stage by stage with host clouds, and through the fused
    device cycle.  Transform, VOXEL filter, exponential decay, lidar hourglass, 0.02 m voxels, negative coordinates."""
    mg, _ = load()
    g = np.load(GOLD_C4)
    from spatio_temporal_voxel_layer_b200 import synth
    p, v = synth.RGBD, synth.VLP16
    cm, gp = mg.CM_C4, mg.C4_GRID
    gg = stvl.SpatioTemporalVoxelGrid(gp["voxel_size"], gp["background"], gp["decay_model"], gp["voxel_decay"])
    keep = []
    if fused:
        import torch
        cm_dev = torch.zeros((cm["size_y"], cm["size_x"]), dtype=torch.uint8, device="cuda")
    for c, (now, readings) in enumerate(cycles_c4_from(g)):
        rs = []
        for k, (o, q, cloud, m) in enumerate(readings):
            lim = p if k < 2 else v
            src = cloud
            if fused:
                t = torch.from_numpy(np.ascontiguousarray(cloud, np.float32)).cuda()
                keep.append(t)
                src = (t.data_ptr(), t.shape[0], 16, (0, 4, 8))
            rs.append(stvl.MeasurementReading(src, origin=o, orientation=q, obstacle_range=lim["obstacle_range"],
                                              min_z=lim["min_z"], max_z=lim["max_z"], vfov=lim["vfov"],
                                              vfov_padding=lim.get("vpad", 0.0), hfov=lim["hfov"], decay_acceleration=lim["accel"],
                                              model_type=0 if k < 2 else 1, now=now + 1e-3 * k,
                                              filter=stvl.FILTER_VOXEL if m is not None else stvl.FILTER_PASSTHROUGH,
                                              min_obstacle_height=lim["min_obstacle_height"], max_obstacle_height=lim["max_obstacle_height"],
                                              transform=m, voxel_min_points=1))
        if fused:
            gg.update_cycle_device(now, rs, rs, cm["origin_x"], cm["origin_y"], cm["resolution"], cm["size_x"], cm["size_y"],
                                   cm_dev.data_ptr(), mark_threshold=cm["mark_threshold"], default_value=cm["default_value"])
            res = gg.costmap_result()
            out = cm_dev.cpu().numpy()
        else:
            gg.ClearFrustums(rs, now)
            gg.Mark(rs)
            out, res = gg.UpdateROSCostmap(cm["origin_x"], cm["origin_y"], cm["resolution"], cm["size_x"], cm["size_y"],
                                           cm["mark_threshold"], cm["default_value"])
        assert np.array_equal(np.packbits(out == 254), g["costmap_%d" % c]), "cycle %d" % c
        x, y, cnt = gg.GetFlattenedCostmap()
        order = np.lexsort((y, x))
        assert np.array_equal(x[order], g["col_x_%d" % c]) and np.array_equal(y[order], g["col_y_%d" % c])
        assert np.array_equal(cnt[order], g["col_c_%d" % c])
        st = gg.sweep_stats()
        assert st.n_ambiguous == 0      # no exp()-dependent razor-edge decision in this fixture
        b = [1e308, 1e308, -1e308, -1e308]
        for box, ok in ((res.touched, res.has_bounds), (st.cleared_bbox_world, st.n_cleared > 0)):
            if ok:
                b = [min(b[0], box[0]), min(b[1], box[1]), max(b[2], box[2]), max(b[3], box[3])]
        np.testing.assert_array_equal(np.array(b), g["bounds_%d" % c])
    ix, iy, iz, t = gg.voxels()
    order = np.lexsort((iz, iy, ix))
    assert np.array_equal(ix[order], g["vox_ix"]) and np.array_equal(iy[order], g["vox_iy"]) and np.array_equal(iz[order], g["vox_iz"])
    assert np.array_equal(t[order], g["vox_t"])
    gg.close()
