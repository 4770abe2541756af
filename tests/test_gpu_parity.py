"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): voxel key sets, mark times, column counts and costmap bytes bit-exact;
voxels whose decision hinged on libm exp/atan must be listed (stvl_get_ambiguous)."""
import math

import ctypes

import numpy as np
import pytest

from helpers import assert_same_columns, assert_same_voxels, point_set

pytestmark = pytest.mark.gpu


def rgbd_reading(stvl, synth, scene, origin, quat, seed, w=160, h=120, now=None, **kw):
    cloud = synth.depth_cloud(scene, origin, quat, width=w, height=h, seed=seed)
    p = synth.RGBD
    return stvl.MeasurementReading(cloud, origin=origin, orientation=quat, obstacle_range=p["obstacle_range"],
                                   min_z=p["min_z"], max_z=p["max_z"], vfov=p["vfov"], hfov=p["hfov"],
                                   decay_acceleration=p["accel"], marking=True, clearing=True, now=now,
                                   filter=stvl.FILTER_PASSTHROUGH, min_obstacle_height=p["min_obstacle_height"],
                                   max_obstacle_height=p["max_obstacle_height"], **kw)


def oracle_mark(oracle, og, reading, now):
    """Mark one reading on the oracle, applying the same PassThrough pre-filter the GPU fuses into Mark."""
    cloud = np.ascontiguousarray(reading.cloud, np.float32)
    if reading.filter == 2:
        xyz = oracle.filter_passthrough(cloud, reading.min_obstacle_height, reading.max_obstacle_height)
    else:
        xyz = cloud[:, :3]
    og.mark(np.ascontiguousarray(xyz), reading.origin, reading.obstacle_range, reading.now if reading.now is not None else now,
            marking=1.0 if reading.marking else 0.0)


def oracle_frustums(oracle, readings):
    blocks = []
    for r in readings:
        if r.model_type == 0:
            blocks.append(oracle.depth_frustum(r.vfov, r.hfov, r.min_z, r.max_z, r.origin, r.orientation, r.decay_acceleration))
        elif r.model_type == 1:
            blocks.append(oracle.lidar_frustum(r.vfov, r.vfov_padding, r.hfov, r.min_z, r.max_z, r.origin, r.orientation, r.decay_acceleration))
    return np.concatenate(blocks) if blocks else None


def compare_cycle_outputs(gg, og, check_points=False, ambiguous_ok=False):
    gx, gy, gc = gg.GetFlattenedCostmap()
    assert_same_columns((gx, gy, gc), og.costmap())
    # cleared cells: the reference keeps a set of (x,y); the GPU lists one entry per cleared voxel
    cix, ciy = gg.cleared()
    gcl = np.unique(np.stack([gg.index_to_world(cix), gg.index_to_world(ciy)], 1), axis=0) if len(cix) else np.zeros((0, 2))
    ox, oy = og.cleared()
    ocl = np.unique(np.stack([ox, oy], 1), axis=0) if len(ox) else np.zeros((0, 2))
    assert np.array_equal(gcl.view(np.uint64), ocl.view(np.uint64)), "cleared cell sets differ"
    st = gg.sweep_stats()
    if len(ox):
        np.testing.assert_array_equal(np.array(st.cleared_bbox_world[:]), [ox.min(), oy.min(), ox.max(), oy.max()])
    if not ambiguous_ok:
        assert st.n_ambiguous == 0
    if check_points:
        assert np.array_equal(point_set(gg.GetOccupancyPointCloud()), point_set(og.points()))
    assert_same_voxels(gg.voxels(), og.voxels())
    return st


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_mark_quirks_bit_exact(stvl, oracle):
    """Quantisation quirks of reference .cpp:285-303 on adversarial points: negatives, exact voxel and range
    boundaries, the z-uses-y-sign bug, tiny distances, non-finite values."""
    rng = np.random.default_rng(1)
    n = 50000
    pts = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    pts[:5000] = (np.round(pts[:5000] / 0.05) * 0.05).astype(np.float32)          # on voxel boundaries
    pts[5000:6000, 0] = np.float32(3.0); pts[5000:6000, 1:] = 0                    # on the range boundary
    pts[6000:6100] = rng.uniform(-0.012, 0.012, (100, 3)).astype(np.float32)       # around distance^2 = 1e-4
    pts[6100] = [np.nan, 0, 0]; pts[6101] = [np.inf, 1, 1]; pts[6102] = [1, -np.inf, 1]
    cloud = np.zeros((n, 4), np.float32); cloud[:, :3] = pts
    for vs in (0.05, 0.02):
        gg = stvl.SpatioTemporalVoxelGrid(vs)
        og = oracle.OracleGrid(vs)
        r = stvl.MeasurementReading(cloud, origin=(0.001, -0.002, 0.003), obstacle_range=3.0, now=1000.25)
        gg.Mark([r])
        og.mark(cloud, r.origin, 3.0, 1000.25)
        assert_same_voxels(gg.voxels(), og.voxels())
        assert gg.pool_stats().points_dropped_nonfinite == 3
        gg.close()


def test_mark_strided_and_unaligned_layouts(stvl, oracle):
    """PointCloud2 layouts other than packed xyz-pad: 32 B XYZRGB-style, 22 B Velodyne-style (unaligned floats)."""
    rng = np.random.default_rng(2)
    n = 20000
    xyz = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
    for step, offs in ((32, (0, 4, 8)), (32, (16, 20, 4)), (22, (0, 4, 8)), (22, (2, 6, 14)), (12, (0, 4, 8))):
        raw = rng.integers(0, 255, (n, step), dtype=np.uint8)
        for a in range(3):
            raw[:, offs[a]:offs[a] + 4] = xyz[:, a:a + 1].copy().view(np.uint8)
        gg = stvl.SpatioTemporalVoxelGrid(0.05)
        og = oracle.OracleGrid(0.05)
        r = stvl.MeasurementReading((raw.ctypes.data, n, step, offs), origin=(0, 0, 0), obstacle_range=2.5, now=5.0)
        gg.Mark([r])
        og.mark(raw.reshape(-1), (0, 0, 0), 2.5, 5.0, point_step=step, offsets=offs)
        assert_same_voxels(gg.voxels(), og.voxels())
        gg.close()


def test_mark_overwrite_order(stvl, oracle):
    """Later readings overwrite earlier ones (reference .cpp:428) — batched atomicMax path for monotone clocks,
    per-reading plain-store path when clocks go backwards."""
    rng = np.random.default_rng(3)
    clouds = [np.concatenate([rng.uniform(-2, 2, (5000, 3)), np.zeros((5000, 1))], 1).astype(np.float32) for _ in range(4)]
    for nows in ([10.0, 10.1, 10.2, 10.3], [10.0, 9.0, 11.0, 8.0], [10.0, 10.0, 10.0, 10.0]):
        gg = stvl.SpatioTemporalVoxelGrid(0.05)
        og = oracle.OracleGrid(0.05)
        rs = [stvl.MeasurementReading(c, origin=(0, 0, 0), obstacle_range=3.0, now=t) for c, t in zip(clouds, nows)]
        gg.Mark(rs)
        for c, t in zip(clouds, nows):
            og.mark(c, (0, 0, 0), 3.0, t)
        assert_same_voxels(gg.voxels(), og.voxels())
        # a second Mark with an older clock must still overwrite
        gg.Mark([stvl.MeasurementReading(clouds[0], origin=(0, 0, 0), obstacle_range=3.0, now=5.0)])
        og.mark(clouds[0], (0, 0, 0), 3.0, 5.0)
        assert_same_voxels(gg.voxels(), og.voxels())
        gg.close()


@pytest.mark.parametrize("decay_model,voxel_decay", [(0, 15.0), (1, 15.0), (2, 15.0), (2, -1.0)])
def test_sweep_decay_models_no_frustum(stvl, oracle, decay_model, voxel_decay):
    rng = np.random.default_rng(4)
    n = 60000
    vox = np.unique(rng.integers(-300, 300, (n, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 40
    vox = np.unique(vox, axis=0)
    now = 1.7e9
    t = now - rng.uniform(0, 30, len(vox))
    t[:100] = now - 15.0                      # exactly at the linear decay boundary: survives (strict <)
    t[100:200] = np.nextafter(now - 15.0, 0)  # one ulp older: cleared
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=decay_model, voxel_decay=voxel_decay, pub_voxels=True)
    og = oracle.OracleGrid(0.05, decay_model=decay_model, voxel_decay=voxel_decay, pub_voxels=True)
    gg.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    gg.ClearFrustums([], now)
    og.clear_frustums(now)
    st = compare_cycle_outputs(gg, og, check_points=True, ambiguous_ok=(decay_model == 1))
    assert st.n_swept == len(vox)
    assert st.n_survived + st.n_cleared == st.n_swept
    gg.close()


def test_sweep_depth_frustums_bit_exact(stvl, oracle, ):
    """Config 1/2 shape: depth-camera plane sets with acceleration, voxels scattered densely through and around the
    frustums, including voxels sitting exactly ON frustum planes (camera placed on voxel-centre lattice planes)."""
    rng = np.random.default_rng(5)
    vsf = float(np.float32(0.05))
    vox = np.unique(rng.integers(-160, 160, (400000, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 48
    vox = np.unique(vox, axis=0)
    now = 2000.0
    t = now - rng.uniform(0, 14.9, len(vox))
    readings = []
    import spatio_temporal_voxel_layer_b200.synth as synth
    for k in range(7):
        yaw = k * 2 * math.pi / 7
        # origins ON the voxel-centre lattice so that near/far/side planes pass exactly through voxel centres
        o = ((rng.integers(-20, 20)) * vsf + vsf / 2, (rng.integers(-20, 20)) * vsf + vsf / 2, 12 * vsf + vsf / 2)
        readings.append(stvl.MeasurementReading(None, origin=o, orientation=synth.optical_quat(yaw, 0.05 * k),
                                                min_z=0.1, max_z=7.0, vfov=0.8745, hfov=1.048,
                                                decay_acceleration=[5.0, 0.0, 0.5, 50.0, 5.0, 1.0, 5.0][k]))
    # plus an axis-aligned camera whose far/near planes coincide with voxel-centre planes
    readings.append(stvl.MeasurementReading(None, origin=(vsf / 2, vsf / 2, vsf / 2), orientation=(0, 0, 0, 1), min_z=0.1,
                                            max_z=7.0, vfov=0.8745, hfov=1.048, decay_acceleration=3.0))
    readings.append(stvl.MeasurementReading(None, origin=(0, 0, 0), orientation=(0, 0, 0, 1), vfov=0.0, hfov=0.0))   # bogus
    gg = stvl.SpatioTemporalVoxelGrid(0.05)
    og = oracle.OracleGrid(0.05)
    gg.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    gg.ClearFrustums(readings, now)
    og.clear_frustums(now, oracle_frustums(oracle, readings))
    st = compare_cycle_outputs(gg, og)
    assert st.n_rewritten > 1000 and st.n_cleared > 1000
    # second sweep on the rewritten times
    gg.ClearFrustums(readings[:3], now + 1.0)
    og.clear_frustums(now + 1.0, oracle_frustums(oracle, readings[:3]))
    compare_cycle_outputs(gg, og)
    gg.close()


@pytest.mark.parametrize("hfov", [6.29, 2.0])
def test_sweep_lidar_hourglass(stvl, oracle, hfov):
    """Config 3: VLP-16 hourglass frustum, exponential decay.  hfov 6.29 = full circle (no atan);
    hfov 2.0 exercises the atan wedge — mismatches, if any, must be inside the listed ambiguous set."""
    rng = np.random.default_rng(6)
    vox = np.unique(rng.integers(-200, 200, (300000, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 60 - 10
    vox = np.unique(vox, axis=0)
    now = 500.0
    t = now - rng.uniform(0, 6.0, len(vox))
    q = np.array([0.02, -0.03, math.sin(0.4), math.cos(0.4)]); q /= np.linalg.norm(q)
    r = stvl.MeasurementReading(None, origin=(0.31, -0.12, 0.9), orientation=q, min_z=1.0, max_z=8.0, vfov=0.523,
                                vfov_padding=0.05, hfov=hfov, decay_acceleration=5.0, model_type=1)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=1, voxel_decay=15.0)
    og = oracle.OracleGrid(0.05, decay_model=1, voxel_decay=15.0)
    gg.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    gg.ClearFrustums([r], now)
    og.clear_frustums(now, oracle_frustums(oracle, [r]))
    st = gg.sweep_stats()
    assert st.n_cleared > 100 and st.n_rewritten > 100
    gv, ov = gg.voxels(), og.voxels()
    if st.n_ambiguous == 0:
        compare_cycle_outputs(gg, og, ambiguous_ok=True)
    else:
        # every disagreement must be a listed voxel
        ax, ay, az = gg.ambiguous()
        listed = set(zip(ax.tolist(), ay.tolist(), az.tolist()))
        gs = dict(zip(zip(gv[0].tolist(), gv[1].tolist(), gv[2].tolist()), gv[3].tolist()))
        os_ = dict(zip(zip(ov[0].tolist(), ov[1].tolist(), ov[2].tolist()), ov[3].tolist()))
        diff = {k for k in set(gs) | set(os_) if gs.get(k) != os_.get(k)}
        assert diff <= listed, "unlisted disagreement: %s" % sorted(diff - listed)[:5]
    gg.close()


def test_multi_cycle_scenario_with_costmap(stvl, oracle):
    """Config 1/2 end to end at reduced image size: robot drives down an aisle with 3 RGBD cameras for 12 cycles:
    clear -> mark -> costmap each cycle, order as the reference's updateBounds (spatio_temporal_voxel_layer.cpp:772-795)."""
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=30.0, seed=3)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=2.0, pub_voxels=True, leaf_capacity=256)  # tiny pool: forces growth
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=2.0, pub_voxels=True)
    size, res, ox, oy = 400, 0.05, -10.0, -10.0
    for cyc in range(12):
        now = 100.0 + 0.4 * cyc
        centre = (-6.0 + 0.5 * cyc, 2.3 - 13.5 + 12.0, 0.0)
        readings = []
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.1 * cyc, n_cams=3)):
            readings.append(rgbd_reading(stvl, synth, scene, o, q, seed=100 * cyc + k, now=now + 0.001 * k))
        gg.ClearFrustums(readings, now)
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        gg.Mark(readings)
        for r in readings:
            oracle_mark(oracle, og, r, now)
        cm, res_g = gg.UpdateROSCostmap(ox, oy, res, size, size, mark_threshold=(cyc % 3), default_value=255 if cyc % 2 else 0)
        ocm, ob, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=(cyc % 3), default_value=255 if cyc % 2 else 0)
        assert np.array_equal(cm, ocm), "costmap bytes differ in cycle %d" % cyc
        assert res_g.n_marked_columns == on
        st = compare_cycle_outputs(gg, og, check_points=True)
        # bounds: touch() over marked columns and cleared cells == oracle's
        b = [1e308, 1e308, -1e308, -1e308]
        if res_g.has_bounds:
            b = [min(b[0], res_g.touched[0]), min(b[1], res_g.touched[1]), max(b[2], res_g.touched[2]), max(b[3], res_g.touched[3])]
        if st.n_cleared:
            w = st.cleared_bbox_world
            b = [min(b[0], w[0]), min(b[1], w[1]), max(b[2], w[2]), max(b[3], w[3])]
        np.testing.assert_array_equal(np.array(b), ob)
    assert og.active_count() > 1000
    gg.close()


@pytest.mark.parametrize("mixed_layouts", [False, True])
def test_fused_update_cycle_matches_oracle(stvl, oracle, mixed_layouts):
    """stvl_update_cycle_device — the three-launch fused cycle (triage kernel resets the costmap, leaf pass and Mark
    chained by programmatic dependent launch, costmap pass carried by the Mark kernel) — against the oracle's
    clear -> mark -> costmap, over 10 cycles with device-resident clouds.  Costmap bytes, marked-column count, touched
    bounds, column counts, cleared cells and the voxel store must all be bit-identical; cycle 6 has nothing to mark
    (the plain costmap kernel runs instead), a dirty costmap buffer checks the in-kernel reset."""
    import torch
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=30.0, seed=3)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=2.0, leaf_capacity=1 << 14)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=2.0)
    size, res, ox, oy = 400, 0.05, -10.0, -10.0
    stream = torch.cuda.ExternalStream(gg.stream())
    cm_dev = torch.empty((size + 1) * size + 3, dtype=torch.uint8, device="cuda")[3:3 + size * size]   # deliberately unaligned
    for cyc in range(10):
        now = 100.0 + 0.4 * cyc
        centre = (-6.0 + 0.5 * cyc, 2.3 - 13.5 + 12.0, 0.0)
        readings = []
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.1 * cyc, n_cams=3)):
            readings.append(rgbd_reading(stvl, synth, scene, o, q, seed=100 * cyc + k, now=now + 0.001 * k))
        marking = [] if cyc == 6 else readings
        if cyc == 2:      # 18 marking readings with increasing clocks: two batched Mark launches, the costmap pass rides on the last
            marking = [stvl.MeasurementReading(r.cloud, origin=r.origin, orientation=r.orientation, obstacle_range=r.obstacle_range - 0.1 * (j // 3),
                                               now=now + 0.001 * j, filter=r.filter, min_obstacle_height=r.min_obstacle_height,
                                               max_obstacle_height=r.max_obstacle_height) for j, r in enumerate(readings * 6)]
        elif cyc == 3:    # the same reading list twice: clocks not monotone -> one Mark launch per reading with plain stores
            marking = readings * 2
        clearing = readings * 3 if cyc == 4 else readings     # 9 frustums: more than fit the kernel parameters -> upload path
        dev_readings, keep = [], []
        for k, r in enumerate(marking):
            host = np.ascontiguousarray(r.cloud, np.float32)
            if mixed_layouts and k == 1:      # one strided cloud switches the whole batch to the scalar-load kernel variant
                wide = np.zeros((len(host), 6), np.float32)
                wide[:, 2:5] = host[:, :3]
                t = torch.from_numpy(wide).cuda()
                layout = (t.data_ptr(), len(host), 24, (8, 12, 16))
            else:
                t = torch.from_numpy(host).cuda()
                layout = (t.data_ptr(), len(host), 16, (0, 4, 8))
            keep.append(t)
            # cycle 8: one reading goes through the VOXEL pre-filter (extra kernels between the sweep and Mark)
            filt = stvl.FILTER_VOXEL if (cyc == 8 and k == 2 and not (mixed_layouts and k == 1)) else r.filter
            dev_readings.append(stvl.MeasurementReading(layout, origin=r.origin, orientation=r.orientation, obstacle_range=r.obstacle_range,
                                                        now=r.now, filter=filt, min_obstacle_height=r.min_obstacle_height,
                                                        max_obstacle_height=r.max_obstacle_height, voxel_min_points=2))
        with torch.cuda.stream(stream):
            cm_dev.fill_(77)      # stale bytes: the cycle itself must reset them
        thr, dflt = cyc % 3, (255 if cyc % 2 else 0)
        gg.update_cycle_device(now, clearing, dev_readings, ox, oy, res, size, size, cm_dev.data_ptr(), mark_threshold=thr, default_value=dflt)
        og.clear_frustums(now, oracle_frustums(oracle, clearing))
        for k, r in enumerate(marking):
            if cyc == 8 and k == 2:
                cent = oracle.filter_voxel(np.ascontiguousarray(r.cloud, np.float32), np.float32(0.05), r.min_obstacle_height, r.max_obstacle_height, 2)
                og.mark(cent, r.origin, r.obstacle_range, r.now)
            else:
                oracle_mark(oracle, og, r, now)
        ocm, ob, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=thr, default_value=dflt)
        res_g = gg.costmap_result()
        cm = cm_dev.cpu().numpy().reshape(size, size)
        assert np.array_equal(cm, ocm), "costmap bytes differ in cycle %d" % cyc
        assert res_g.n_marked_columns == on
        st = compare_cycle_outputs(gg, og)
        b = [1e308, 1e308, -1e308, -1e308]
        if res_g.has_bounds:
            b = [min(b[0], res_g.touched[0]), min(b[1], res_g.touched[1]), max(b[2], res_g.touched[2]), max(b[3], res_g.touched[3])]
        if st.n_cleared:
            w = st.cleared_bbox_world
            b = [min(b[0], w[0]), min(b[1], w[1]), max(b[2], w[2]), max(b[3], w[3])]
        np.testing.assert_array_equal(np.array(b), ob)
        del keep
    assert og.active_count() > 1000
    gg.close()


def test_mark_large_scattered_clouds(stvl, oracle):
    """Clouds large and scattered enough that every Mark CTA gets many chunks, fills its shared-memory table and has to
    flush it early (and points that find the table full take the per-point path): two readings of 1.5 M points each,
    uniformly random in a 5 m cube, the second overwriting the first where they overlap (later clock sample wins)."""
    rng = np.random.default_rng(77)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, leaf_capacity=1 << 13)
    og = oracle.OracleGrid(0.05)
    readings = []
    for k, centre in enumerate(((0.3, -0.2, 0.4), (1.1, 0.6, 0.2))):
        cloud = np.zeros((1_500_000, 4), np.float32)
        cloud[:, :3] = rng.uniform(-2.5, 2.5, (len(cloud), 3)) + np.asarray(centre, np.float32)
        cloud[rng.integers(0, len(cloud), 5000), 1] = np.nan
        now = 50.0 + 0.25 * k
        readings.append(stvl.MeasurementReading(cloud, origin=centre, obstacle_range=4.0, now=now))
        og.mark(cloud[np.isfinite(cloud[:, 1])], centre, 4.0, now)
    gg.Mark(readings)
    assert og.active_count() > 500_000
    assert_same_voxels(gg.voxels(), og.voxels())
    assert gg.pool_stats().points_dropped_nonfinite >= 9000
    gg.close()


def test_fused_cycle_large_map_small_cloud(stvl, oracle):
    """Config-5 shape at test size: far more voxel columns than a small cloud's Mark grid could take along, so the fused
    cycle launches the costmap kernel between the leaf pass and Mark instead of carrying the pass on Mark."""
    import torch
    import spatio_temporal_voxel_layer_b200.synth as synth
    rng = np.random.default_rng(5)
    vox = np.unique(np.concatenate([rng.integers(-600, 600, (120000, 2)), rng.integers(2, 30, (120000, 1))], axis=1), axis=0)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=10.0)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=10.0)
    gg.set_outputs(0)
    t = 100.0 - rng.uniform(0, 12, len(vox))
    gg.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t); og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    size, res, ox, oy = 1280, 0.05, -32.0, -32.0
    scene = synth.retail_scene(size=20.0, seed=2)
    cm_dev = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    keep = []
    for cyc in range(3):
        now = 100.5 + 1.5 * cyc
        o, q = (0.5 * cyc, 0.3, 0.6), synth.optical_quat(0.4 * cyc)
        r = rgbd_reading(stvl, synth, scene, o, q, seed=cyc, now=now, w=64, h=48)
        tcl = torch.from_numpy(np.ascontiguousarray(r.cloud, np.float32)).cuda()
        keep.append(tcl)
        rd = stvl.MeasurementReading((tcl.data_ptr(), tcl.shape[0], 16, (0, 4, 8)), origin=r.origin, orientation=r.orientation,
                                     obstacle_range=r.obstacle_range, now=now, filter=r.filter, min_obstacle_height=r.min_obstacle_height,
                                     max_obstacle_height=r.max_obstacle_height)
        gg.update_cycle_device(now, [r], [rd], ox, oy, res, size, size, cm_dev.data_ptr(), mark_threshold=1, default_value=255)
        og.clear_frustums(now, oracle_frustums(oracle, [r]))
        oracle_mark(oracle, og, r, now)
        ocm, ob, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=1, default_value=255)
        res_g = gg.costmap_result()
        assert np.array_equal(cm_dev.cpu().numpy(), ocm), "costmap bytes differ in cycle %d" % cyc
        assert res_g.n_marked_columns == on and on > 50000
        assert_same_voxels(gg.voxels(), og.voxels())
    assert gg.pool_stats().tiles_in_use > 20000
    gg.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_mixed_api_soak(stvl, oracle, seed):
    """30 cycles on ONE handle with a random mix of everything the boundary offers — fused device cycles, stage-by-stage
    cycles with host or device clouds, cycles that are only inspected several cycles later (no synchronisation in
    between), stage timing switched on and off (it changes how kernels are chained), area resets, a full reset, a tiny
    leaf pool that has to grow — bit-compared with the oracle whenever the state is read."""
    import torch
    import spatio_temporal_voxel_layer_b200.synth as synth
    rng = np.random.default_rng(1000 + seed)
    scene = synth.retail_scene(size=30.0, seed=3 + seed)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=1.5, leaf_capacity=128)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=1.5)
    gg.set_outputs(0)
    size, res, ox, oy = 400, 0.05, -10.0, -10.0
    stream = torch.cuda.ExternalStream(gg.stream())
    cm_dev = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    keep_alive = []
    now = 200.0
    checked = 0
    for cyc in range(30):
        now += float(rng.uniform(0.05, 0.6))
        centre = (-6.0 + 0.4 * cyc, 2.3 - 13.5 + 12.0 + float(rng.uniform(-0.3, 0.3)), 0.0)
        n_cams = int(rng.integers(1, 4))
        readings = [rgbd_reading(stvl, synth, scene, o, q, seed=1000 * seed + 10 * cyc + k, now=now + 0.001 * k, w=120, h=90)
                    for k, (o, q) in enumerate(synth.camera_ring(centre, 0.2 * cyc, n_cams=n_cams))]
        mode = int(rng.integers(0, 3))          # 0 fused device cycle, 1 stage-by-stage with host clouds, 2 with device clouds
        gg.enable_timing(int(rng.integers(0, 8)) if rng.random() < 0.4 else 0)
        thr, dflt = int(rng.integers(0, 3)), int(rng.choice([0, 255]))
        dev_readings = []
        if mode != 1:
            for r in readings:
                t = torch.from_numpy(np.ascontiguousarray(r.cloud, np.float32)).cuda()
                keep_alive.append(t)
                dev_readings.append(stvl.MeasurementReading((t.data_ptr(), t.shape[0], 16, (0, 4, 8)), origin=r.origin, orientation=r.orientation,
                                                            obstacle_range=r.obstacle_range, now=r.now, filter=r.filter,
                                                            min_obstacle_height=r.min_obstacle_height, max_obstacle_height=r.max_obstacle_height))
        if mode == 0:
            gg.update_cycle_device(now, readings, dev_readings, ox, oy, res, size, size, cm_dev.data_ptr(), mark_threshold=thr, default_value=dflt)
        else:
            gg.ClearFrustums(readings, now)
            gg.Mark(readings if mode == 1 else dev_readings, device=(mode == 2))
            p = stvl.CostmapParams(ox, oy, res, size, size, thr, dflt, 254)
            stvl._check(gg.L.stvl_costmap_device(gg.h, ctypes.byref(p), cm_dev.data_ptr(), None))     # enqueue only
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        for r in readings:
            oracle_mark(oracle, og, r, now)
        ocm, ob, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=thr, default_value=dflt)
        if rng.random() < 0.5 or cyc == 29:      # read the state back only now and then: cycles pile up unsynchronised in between
            res_g = gg.costmap_result()
            assert np.array_equal(cm_dev.cpu().numpy(), ocm), "costmap bytes differ in cycle %d (mode %d)" % (cyc, mode)
            assert res_g.n_marked_columns == on
            assert_same_columns(gg.GetFlattenedCostmap(), og.costmap())
            assert_same_voxels(gg.voxels(), og.voxels())
            st = gg.sweep_stats()
            assert st.n_swept == st.n_survived + st.n_cleared
            keep_alive.clear()
            checked += 1
        if rng.random() < 0.15:
            a = rng.uniform(-8, 4, 2)
            inv = bool(rng.integers(0, 2))
            gg.ResetGridArea((a[0], a[1]), (a[0] + 4.0, a[1] + 3.0), invert_area=inv)
            og.reset_area(a[0], a[1], a[0] + 4.0, a[1] + 3.0, inv)
        if cyc == 17:
            assert gg.ResetGrid() and og.reset()
    assert checked >= 8 and og.active_count() > 200
    assert gg.pool_stats().leaf_capacity > 128      # the pool had to grow on the way
    gg.close()


@pytest.mark.parametrize("sx,sy", [(317, 251), (320, 256)])     # odd row pitch: byte path; multiple of 16: 128-bit path
def test_update_costs_on_device(stvl, oracle, sx, sy):
    """SURVEY §8f rank 3 — SpatioTemporalVoxelLayer::updateCosts on device byte grids (stvl_update_costs_device): footprint
    clearing (nav2 setConvexPolygonCost: Bresenham outline + column fill) and updateWithMax / updateWithOverwrite of a
    window into the master grid, against the literal CPU restatement.  Random convex polygons of every size and
    orientation, some with a vertex off the map (nothing may change), degenerate ones, unaligned windows."""
    import torch
    rng = np.random.default_rng(12)
    gg = stvl.SpatioTemporalVoxelGrid(0.05)
    res, ox, oy = 0.05, -7.3, -4.1
    params = stvl.CostmapParams(ox, oy, res, sx, sy, 0, 0, 254)
    n_filled = n_failed = 0
    for it in range(60):
        layer = rng.choice(np.array([0, 1, 99, 128, 253, 254, 255], np.uint8), (sy, sx), p=[0.5, 0.05, 0.05, 0.05, 0.05, 0.2, 0.1])
        master = rng.integers(0, 256, (sy, sx)).astype(np.uint8)
        # a convex polygon: points on an ellipse, random rotation / size / vertex count; every 6th one pokes out of the map
        nv = int(rng.integers(3, 13))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        if it % 7 == 3:
            ang = ang[::-1].copy()                      # clockwise
        a, b = rng.uniform(0.03, 3.0, 2)
        cx = rng.uniform(ox + 3.2, ox + sx * res - 3.2) if it % 6 else ox + sx * res - 0.5
        cy = rng.uniform(oy + 3.2, oy + sy * res - 3.2)
        th = rng.uniform(0, np.pi)
        px, py = a * np.cos(ang), b * np.sin(ang)
        poly = np.stack([cx + px * np.cos(th) - py * np.sin(th), cy + px * np.sin(th) + py * np.cos(th)], 1)
        if it == 10:
            poly = np.repeat(poly[:1], 4, axis=0)       # all vertices in one cell
        if it == 11:
            poly = poly[:2]                             # not a polygon: nothing happens
        method = int(rng.integers(0, 3))                # 2 = neither: master untouched
        i0, j0 = int(rng.integers(0, sx // 2)), int(rng.integers(0, sy // 2))
        window = (i0, j0, int(rng.integers(i0, sx + 1)), int(rng.integers(j0, sy + 1)))
        want_layer, want_master = layer.copy(), master.copy()
        ok = oracle.set_convex_polygon_cost(want_layer, ox, oy, res, poly, 0) if len(poly) >= 3 else False
        n_filled += int(ok and (want_layer != layer).any())
        n_failed += int(not ok)
        if method in (0, 1):
            oracle.update_costs(want_layer, want_master, *window, method)
        layer_dev, master_dev = torch.from_numpy(layer).cuda(), torch.from_numpy(master).cuda()
        torch.cuda.synchronize()
        gg.update_costs_device(params, layer_dev.data_ptr(), master_dev.data_ptr(), footprint_xy=poly, combination_method=method, window=window)
        gg.synchronize()
        assert np.array_equal(layer_dev.cpu().numpy(), want_layer), "layer costmap differs (iteration %d, %d vertices)" % (it, len(poly))
        assert np.array_equal(master_dev.cpu().numpy(), want_master), "master grid differs (iteration %d, method %d)" % (it, method)
    assert n_filled >= 35 and n_failed >= 8
    gg.close()


def test_leaf_pool_growth_replays_mark(stvl, oracle):
    """A Mark that exhausts the leaf pool is replayed after the pool (and its hash) grew: nothing is lost."""
    rng = np.random.default_rng(13)
    n = 30000
    cloud = np.zeros((n, 4), np.float32)
    cloud[:, :3] = rng.uniform(-40, 40, (n, 3))          # scattered: almost one leaf per point
    gg = stvl.SpatioTemporalVoxelGrid(0.05, leaf_capacity=64)
    og = oracle.OracleGrid(0.05)
    gg.Mark([stvl.MeasurementReading(cloud, origin=(0, 0, 0), obstacle_range=100.0, now=3.0)])
    og.mark(cloud, (0, 0, 0), 100.0, 3.0)
    assert_same_voxels(gg.voxels(), og.voxels())
    ps = gg.pool_stats()
    assert ps.leaf_capacity >= 16384 and ps.leaves_in_use > 10000
    # and the grown grid keeps working: sweep + second mark + costmap
    gg.ClearFrustums([], 4.0); og.clear_frustums(4.0)
    gg.Mark([stvl.MeasurementReading(cloud[:1000] + np.float32(0.3), origin=(0, 0, 0), obstacle_range=100.0, now=4.5)])
    og.mark(cloud[:1000] + np.float32(0.3), (0, 0, 0), 100.0, 4.5)
    compare_cycle_outputs(gg, og)
    gg.close()


def test_reset_and_reset_area(stvl, oracle):
    rng = np.random.default_rng(8)
    vox = np.unique(rng.integers(-100, 100, (50000, 3)), axis=0)
    t = rng.uniform(1, 2, len(vox))
    gg = stvl.SpatioTemporalVoxelGrid(0.05)
    og = oracle.OracleGrid(0.05)
    gg.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t); og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    vs = gg.voxel_size
    # rectangle edges exactly ON voxel centres: strict inequalities (reference .cpp:409-411)
    s, e = (-20 * vs + vs / 2, -10 * vs + vs / 2), (30 * vs + vs / 2, 40 * vs + vs / 2)
    gg.ResetGridArea(s, e, invert_area=True); og.reset_area(s[0], s[1], e[0], e[1], True)
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.ResetGridArea((-4.0, -4.0), (4.0, 4.0), invert_area=False); og.reset_area(-4.0, -4.0, 4.0, 4.0, False)
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.ClearFrustums([], 3.0); og.clear_frustums(3.0)
    compare_cycle_outputs(gg, og)
    assert gg.ResetGrid() and og.reset()
    assert gg.active_count() == 0
    gg.ClearFrustums([], 4.0)
    assert len(gg.columns()[0]) == 0
    # the grid is usable again after a reset
    gg.load_voxels([1, 2], [3, 4], [5, 6], [7.0, 8.0])
    assert gg.active_count() == 2
    gg.close()


def test_leaf_recycling_and_hash_tombstones(stvl, oracle):
    """Leaves die and are reused over many cycles; results stay identical to the oracle."""
    rng = np.random.default_rng(9)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=1.0, leaf_capacity=64)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=1.0)
    for cyc in range(30):
        now = 10.0 + cyc * 0.6
        gg.ClearFrustums([], now); og.clear_frustums(now)
        c = rng.uniform(-1, 1, 3) * [20, 20, 0] + [0, 0, 1.0]
        pts = (c + rng.normal(0, 0.6, (4000, 3))).astype(np.float32)
        cloud = np.zeros((4000, 4), np.float32); cloud[:, :3] = pts
        gg.Mark([stvl.MeasurementReading(cloud, origin=tuple(c), obstacle_range=5.0, now=now)])
        og.mark(cloud, tuple(c), 5.0, now)
        if cyc % 5 == 4:
            compare_cycle_outputs(gg, og)
    ps = gg.pool_stats()
    assert ps.leaves_free > 0 or ps.hash_tombstones >= 0
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.close()


def test_config4_small_voxels_mixed_sensors(stvl, oracle):
    """Config 4 shape at reduced size: 0.02 m voxels, one lidar + two RGBD, linear decay."""
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=20.0, seed=5)
    gg = stvl.SpatioTemporalVoxelGrid(0.02, decay_model=0, voxel_decay=3.0)
    og = oracle.OracleGrid(0.02, decay_model=0, voxel_decay=3.0)
    for cyc in range(4):
        now = 50.0 + cyc
        centre = (0.3 * cyc - 1.0, 0.8, 0.0)
        readings = []
        lq = synth.yaw_quat(0.2 * cyc)
        lo = (centre[0], centre[1], 0.9)
        lc = synth.lidar_cloud(scene, lo, lq, rings=32, azimuths=512, vfov=0.6, seed=cyc)
        readings.append(stvl.MeasurementReading(lc, origin=lo, orientation=lq, obstacle_range=8.0, min_z=0.5, max_z=8.0,
                                                vfov=0.6, vfov_padding=0.03, hfov=6.29, decay_acceleration=2.0,
                                                model_type=1, now=now, filter=stvl.FILTER_PASSTHROUGH,
                                                min_obstacle_height=0.1, max_obstacle_height=2.2))
        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.3, n_cams=2)):
            readings.append(rgbd_reading(stvl, synth, scene, o, q, seed=10 * cyc + k, now=now + 0.01 * (k + 1)))
        gg.ClearFrustums(readings, now)
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        gg.Mark(readings)
        for r in readings:
            oracle_mark(oracle, og, r, now)
        compare_cycle_outputs(gg, og)
    assert og.active_count() > 5000
    gg.close()


def test_sharded_grids_union_equals_single(stvl, oracle):
    """Config 5 mechanics on one GPU: N spatial shards, each with its own handle, produce column counts whose
    dense sum (what the NCCL reduce computes) thresholds to the same costmap bytes as the unsharded grid."""
    import torch
    rng = np.random.default_rng(10)
    vox = np.unique(rng.integers(-250, 250, (200000, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 30
    vox = np.unique(vox, axis=0)
    now = 100.0
    t = now - rng.uniform(0, 20, len(vox))
    og = oracle.OracleGrid(0.05)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.clear_frustums(now)
    size, res, ox, oy = 512, 0.05, -12.8, -12.8
    ocm, _, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=2, default_value=0)
    nshards = 4
    ix0, iy0, nx, ny = -260, -260, 520, 520
    dense = torch.zeros((ny, nx), dtype=torch.int32, device="cuda")
    total_active = 0
    shards = []
    for s in range(nshards):
        g = stvl.SpatioTemporalVoxelGrid(0.05, shard_index=s, shard_count=nshards, shard_width=4)
        g.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)      # every shard sees everything, keeps its slabs
        total_active += g.active_count()
        g.ClearFrustums([], now)
        g.columns_to_dense(ix0, iy0, nx, ny, dense.data_ptr())
        g.synchronize()      # the scatter is asynchronous on the handle's stream
        shards.append(g)
    assert total_active == len(vox)
    cm = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    p = stvl.CostmapParams(ox, oy, res, size, size, 2, 0, 254)
    r = shards[0].costmap_from_dense(dense.data_ptr(), ix0, iy0, nx, ny, p, cm.data_ptr())
    assert np.array_equal(cm.cpu().numpy(), ocm)
    assert r.n_marked_columns == on
    for g in shards:
        g.close()


@pytest.mark.parametrize("vs,min_points", [(0.05, 0), (0.05, 3), (0.02, 1)])
def test_voxel_filter_then_mark(stvl, oracle, vs, min_points):
    """Config 4: the VOXEL pre-filter of BufferROSCloud (PCL VoxelGrid, reference measurement_buffer.cpp:153-164) on the
    device, then Mark.  Centroids are float sums in ascending point order (the oracle's definition), so the resulting
    voxel sets and mark times are bit-exact."""
    import spatio_temporal_voxel_layer_b200.synth as synth
    rng = np.random.default_rng(31)
    scene = synth.retail_scene(size=20.0, seed=9)
    o, q = (0.3, 0.8, 0.7), synth.optical_quat(0.9)
    cloud = synth.depth_cloud(scene, o, q, width=320, height=240, seed=5)           # organised, with NaN misses
    lid = synth.lidar_cloud(scene, (0.3, 0.8, 0.9), synth.yaw_quat(0.2), rings=32, azimuths=1024, vfov=0.6, seed=6)
    extra = np.zeros((4000, 4), np.float32); extra[:, :3] = rng.uniform(-2.5, 2.5, (4000, 3))
    gg = stvl.SpatioTemporalVoxelGrid(vs)
    og = oracle.OracleGrid(vs)
    leaf = np.float32(vs)
    for k, (c, origin, zl) in enumerate(((cloud, o, (0.3, 2.0)), (lid, (0.3, 0.8, 0.9), (0.1, 2.2)), (extra, (0.0, 0.0, 0.5), (-1.0, 1.5)))):
        now = 77.0 + k
        r = stvl.MeasurementReading(c, origin=origin, obstacle_range=4.0, now=now, filter=stvl.FILTER_VOXEL,
                                    min_obstacle_height=zl[0], max_obstacle_height=zl[1])
        arr = (stvl.Reading * 1)()
        keep = []
        r._fill(arr[0], now, keep)
        arr[0].voxel_min_points = min_points
        stvl._check(gg.L.stvl_mark(gg.h, arr, 1))
        cent = oracle.filter_voxel(c, leaf, zl[0], zl[1], min_points)
        og.mark(cent, origin, 4.0, now)
    assert og.active_count() > 500
    assert_same_voxels(gg.voxels(), og.voxels())
    assert gg.pool_stats().points_dropped_nonfinite == 0     # filtered slots carry the sensor origin, not NaN
    gg.close()


@pytest.mark.parametrize("filt", ["none", "passthrough", "voxel"])
def test_sensor_frame_cloud_transformed_on_device(stvl, oracle, filt):
    """BufferROSCloud's tf2::doTransform (reference measurement_buffer.cpp:141-146) folded into Mark: the reading carries
    the sensor-frame cloud and the float affine; the device applies it before the filter and the quantisation.  Oracle:
    transform_cloud -> filter -> Mark.  Bit-exact (float products summed left to right, no FMA)."""
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=20.0, seed=11)
    o, q = (0.4, -0.7, 0.8), synth.optical_quat(-2.1)
    glob = synth.depth_cloud(scene, o, q, width=320, height=240, seed=8)            # global frame, NaN misses
    m = stvl.build_transform(o, q)
    assert np.array_equal(m.view(np.uint32), oracle.transform_from_quat(o, q).view(np.uint32))
    rot = m.reshape(3, 4)[:, :3].astype(np.float64)
    sensor = glob.copy()
    sensor[:, :3] = ((glob[:, :3].astype(np.float64) - np.asarray(o)) @ rot).astype(np.float32)   # R^T (p - t)
    fcode = {"none": stvl.FILTER_NONE, "passthrough": stvl.FILTER_PASSTHROUGH, "voxel": stvl.FILTER_VOXEL}[filt]
    zl = (0.25, 1.9)
    want = oracle.transform_cloud(sensor, m)
    assert np.nanmax(np.abs(want[:, :3] - glob[:, :3])) < 1e-4
    if filt == "passthrough":
        want = oracle.filter_passthrough(want, *zl)
    elif filt == "voxel":
        want = oracle.filter_voxel(want, np.float32(0.05), zl[0], zl[1], 2)
    og = oracle.OracleGrid(0.05)
    og.mark(want, o, 3.5, 12.5)
    assert og.active_count() > 300
    # host-pointer call, packed 16-byte points
    gg = stvl.SpatioTemporalVoxelGrid(0.05)
    r = stvl.MeasurementReading(sensor, origin=o, orientation=q, obstacle_range=3.5, now=12.5, filter=fcode,
                                min_obstacle_height=zl[0], max_obstacle_height=zl[1], transform=m, voxel_min_points=2)
    gg.Mark([r], now=12.5)
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.close()
    # strided 20-byte points with x,y,z at 4,8,12 (the scalar-load variant of the kernel)
    wide = np.zeros((len(sensor), 5), np.float32)
    wide[:, 1:4] = sensor[:, :3]
    gg = stvl.SpatioTemporalVoxelGrid(0.05)
    r = stvl.MeasurementReading((wide.ctypes.data, len(wide), 20, (4, 8, 12)), origin=o, orientation=q, obstacle_range=3.5,
                                now=12.5, filter=fcode, min_obstacle_height=zl[0], max_obstacle_height=zl[1], transform=m,
                                voxel_min_points=2)
    gg.Mark([r], now=12.5)
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.close()


def test_config2_full_size_parity_and_properties(stvl, oracle):
    """BASELINE config 2 at FULL size (7 x 640x480 clouds = 2,150,400 points per cycle, 710,765 seeded voxels): two complete
    cycles bit-compared with the oracle, plus size-independent properties: sweep conservation, Mark idempotence."""
    from spatio_temporal_voxel_layer_b200 import workloads
    w = workloads.Config2(n_poses=1, ring=2)
    p = w.p
    assert w.points_per_cycle == 7 * 640 * 480 and len(w.seed_vox[0]) == workloads.RETAIL_VOXELS
    gg = stvl.SpatioTemporalVoxelGrid(0.05, 0.0, 0, 15.0, leaf_capacity=1 << 16)
    og = oracle.OracleGrid(0.05, 0.0, 0, 15.0, False)
    gg.load_voxels(*w.seed_vox, w.seed_t)
    og.load_voxels(*w.seed_vox, w.seed_t)
    cp = w.costmap_params()
    for step in range(2):
        now, poses, clouds = w.cycle(step * 40)      # 0.4 s apart so that voxels expire between the cycles
        rs = [stvl.MeasurementReading(c, origin=o, orientation=q, obstacle_range=p["obstacle_range"], min_z=p["min_z"],
                                      max_z=p["max_z"], vfov=p["vfov"], hfov=p["hfov"], decay_acceleration=p["accel"],
                                      now=now + 1e-4 * k, filter=stvl.FILTER_PASSTHROUGH,
                                      min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"])
              for k, ((o, q), c) in enumerate(zip(poses, clouds))]
        gg.ClearFrustums(rs, now)
        og.clear_frustums(now, oracle_frustums(oracle, rs))
        st = gg.sweep_stats()
        assert st.n_swept == st.n_survived + st.n_cleared and st.n_ambiguous == 0
        gg.Mark(rs)
        for r in rs:
            oracle_mark(oracle, og, r, now)
        cm, res = gg.UpdateROSCostmap(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0)
        ocm, ob, on = og.update_ros_costmap(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0)
        assert np.array_equal(cm, ocm) and res.n_marked_columns == on
    assert_same_voxels(gg.voxels(), og.voxels())
    # Mark is idempotent: the same readings again change nothing
    before = gg.voxels()
    gg.Mark(rs)
    assert_same_voxels(gg.voxels(), before)
    assert st.n_swept > 600000 and st.n_cleared > 100 and st.n_rewritten > 100
    gg.close()


def test_async_host_cycle_matches_oracle(stvl, oracle):
    """stvl_update_cycle_async / stvl_update_cycle_wait — the pipelined host-buffer cycle (H2D copy, fused cycle and D2H
    read-back on three streams, two cycles in flight) — against the oracle: every cycle's costmap bytes and marked-cell
    count, and the final voxel store."""
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=30.0, seed=5)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=3.0, leaf_capacity=1 << 14)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=3.0)
    size, res, ox, oy = 400, 0.05, -10.0, -10.0
    params = stvl.CostmapParams(ox, oy, res, size, size, 1, 0, 254)
    outs = [np.full((size, size), 9, np.uint8) for _ in range(3)]
    expected, tickets = [], []
    n_cycles = 9
    for cyc in range(n_cycles):
        now = 50.0 + 0.5 * cyc
        centre = (-6.0 + 0.6 * cyc, 2.3 - 13.5 + 12.0, 0.0)
        readings = [rgbd_reading(stvl, synth, scene, o, q, seed=10 * cyc + k, now=now + 0.001 * k)
                    for k, (o, q) in enumerate(synth.camera_ring(centre, 0.2 * cyc, n_cams=3))]
        tickets.append(gg.update_cycle_async(now, readings, readings, params, outs[cyc % 3]))
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        for r in readings:
            oracle_mark(oracle, og, r, now)
        ocm, _, on = og.update_ros_costmap(ox, oy, res, size, size, mark_threshold=1, default_value=0)
        expected.append((ocm, on))
        if cyc >= 1:      # wait for the previous cycle while this one is in flight
            r = gg.update_cycle_wait(tickets[cyc - 1])
            assert np.array_equal(outs[(cyc - 1) % 3], expected[cyc - 1][0]), "costmap bytes differ in cycle %d" % (cyc - 1)
            assert r.n_marked_columns == expected[cyc - 1][1]
    r = gg.update_cycle_wait(tickets[-1])
    assert np.array_equal(outs[(n_cycles - 1) % 3], expected[-1][0])
    assert r.n_marked_columns == expected[-1][1]
    with pytest.raises(stvl.StvlError):
        gg.update_cycle_wait(tickets[-1])      # already consumed
    assert_same_voxels(gg.voxels(), og.voxels())
    assert og.active_count() > 1000
    gg.close()


def test_fused_cycles_tiny_clouds_back_to_back(stvl, oracle):
    """Many fused cycles with tiny clouds and a tiny store, enqueued back to back without any synchronisation: every
    kernel's grid is a single CTA, so under programmatic dependent launch consecutive cycles' kernels are resident together
    (the chunk cursors of consecutive Mark launches alternate and are never reset).  Marked counts and the final store must
    match the oracle exactly."""
    import torch
    rng = np.random.default_rng(9)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=50.0, leaf_capacity=1 << 12)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=50.0)
    size, res, ox, oy = 128, 0.05, -3.2, -3.2
    cm_dev = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    clouds = [np.concatenate([rng.uniform(-2.5, 2.5, (n, 3)), np.zeros((n, 1))], 1).astype(np.float32) for n in (37, 1500, 1024, 1, 2049, 300)]
    dev = [torch.from_numpy(c).cuda() for c in clouds]
    torch.cuda.synchronize()
    n_cycles = 240
    for cyc in range(n_cycles):
        now = 10.0 + 0.01 * cyc
        k = cyc % len(clouds)
        rd = stvl.MeasurementReading((dev[k].data_ptr(), len(clouds[k]), 16, (0, 4, 8)), origin=(0.1, -0.2, 0.3), obstacle_range=2.4, now=now)
        gg.update_cycle_device(now, [], [rd], ox, oy, res, size, size, cm_dev.data_ptr())
        og.clear_frustums(now, None)
        og.mark(np.ascontiguousarray(clouds[k][:, :3]), (0.1, -0.2, 0.3), 2.4, now)
    ocm, _, on = og.update_ros_costmap(ox, oy, res, size, size)
    assert gg.costmap_result().n_marked_columns == on
    assert np.array_equal(cm_dev.cpu().numpy(), ocm)
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.close()


def test_remark_over_negative_stored_times(stvl, oracle):
    """Clocks near zero (sim time): an accelerated sweep rewrites mark times to value - accel < 0; a later batched Mark must
    still overwrite them (signed 64-bit max on the bit pattern — a negative double is a huge unsigned integer)."""
    rng = np.random.default_rng(4)
    pts = np.concatenate([rng.uniform(0.5, 2.5, (4000, 3)), np.zeros((4000, 1))], 1).astype(np.float32)
    pts[:, 1] = rng.uniform(-0.8, 0.8, 4000).astype(np.float32)
    pts[:, 2] = rng.uniform(-0.5, 0.5, 4000).astype(np.float32)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=2, voxel_decay=1e6)
    og = oracle.OracleGrid(0.05, decay_model=2, voxel_decay=1e6)
    quat = (0.5, -0.5, 0.5, -0.5)      # optical frame looking along +x
    def reading(now):
        return stvl.MeasurementReading(pts, origin=(0.0, 0.0, 0.0), orientation=quat, obstacle_range=4.0, min_z=0.05, max_z=6.0,
                                       vfov=1.2, hfov=1.4, decay_acceleration=15.0, now=now)
    for now in (1.5, 3.0, 3.2, 6.0):
        r = reading(now)
        gg.ClearFrustums([r], now)
        og.clear_frustums(now, oracle_frustums(oracle, [r]))
        gg.Mark([r])
        oracle_mark(oracle, og, r, now)
        assert_same_voxels(gg.voxels(), og.voxels())
    gg.ClearFrustums([reading(6.5)], 6.5)
    og.clear_frustums(6.5, oracle_frustums(oracle, [reading(6.5)]))
    compare_cycle_outputs(gg, og)
    assert og.active_count() > 500
    gg.close()


@pytest.mark.parametrize("config", ["config1", "config3", "config4"])
def test_full_size_configs_fused_cycle(stvl, oracle, config):
    """BASELINE configs 1, 3 and 4 at FULL size (640x480 depth images, 16 x 1800 VLP-16 scans, 128 x 1024 Ouster scans + 4
    RGBD with the VOXEL pre-filter at 0.02 m) through the fused device cycle, cycle for cycle against the oracle: final voxel
    store with mark times, per-column counts and the last costmap bytes, bit-exact."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench_configs as BC
    from spatio_temporal_voxel_layer_b200 import workloads
    w = {"config1": lambda: workloads.Config1(n_cycles=8), "config3": lambda: workloads.Config3(n_cycles=8),
         "config4": lambda: workloads.Config4(n_cycles=3)}[config]()
    gpu = BC.run_gpu(stvl, w, warmup=0, timed=False)
    ora = BC.run_oracle(w, warmup=0)
    c = BC.compare(gpu, ora)
    gpu["grid"].close()
    assert c["voxels"] and c["columns"] and c["costmap"], c
    assert c["n_voxels"] > 1000 and c["marked_cells"] > 50, c


def test_config5_five_million_voxels_against_oracle(stvl, oracle):
    """Config 5 mechanics at a size the oracle still handles (5 M voxels at 0.02 m, ~1 s per oracle cycle): two fused cycles on
    one unsharded grid — costmap bytes of a 10000 x 6000 map, per-column counts, the whole voxel store with its mark times."""
    import torch
    from spatio_temporal_voxel_layer_b200 import synth, workloads
    w = workloads.Config5(target=5_000_000, dt=0.2)
    p = w.p
    ix, iy, iz, t = w.voxels()
    assert 4_000_000 < len(ix) < 7_000_000
    gg = stvl.SpatioTemporalVoxelGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, leaf_capacity=len(ix) // 40 + (1 << 15))
    gg.set_outputs(0)
    og = oracle.OracleGrid(w.voxel_size, 0.0, w.decay_model, w.voxel_decay, False)
    gg.load_voxels(ix, iy, iz, t)
    og.load_voxels(ix, iy, iz, t)
    cp = w.costmap_params()
    cm_dev = torch.zeros((cp["size_y"], cp["size_x"]), dtype=torch.uint8, device="cuda")
    keep = []
    for cyc in range(2):
        now, poses, clouds = w.cycle(cyc)
        readings, dev_readings = [], []
        for k, ((o, q), cl) in enumerate(zip(poses, clouds)):
            r = stvl.MeasurementReading(cl, origin=o, orientation=q, obstacle_range=p["obstacle_range"], min_z=p["min_z"], max_z=p["max_z"],
                                        vfov=p["vfov"], hfov=p["hfov"], decay_acceleration=p["accel"], now=now + 1e-4 * k,
                                        filter=stvl.FILTER_PASSTHROUGH, min_obstacle_height=p["min_obstacle_height"],
                                        max_obstacle_height=p["max_obstacle_height"])
            readings.append(r)
            td = torch.from_numpy(np.ascontiguousarray(cl, np.float32)).cuda()
            keep.append(td)
            dev_readings.append(stvl.MeasurementReading((td.data_ptr(), len(cl), 16, (0, 4, 8)), origin=o, orientation=q,
                                                        obstacle_range=p["obstacle_range"], now=now + 1e-4 * k, filter=stvl.FILTER_PASSTHROUGH,
                                                        min_obstacle_height=p["min_obstacle_height"], max_obstacle_height=p["max_obstacle_height"]))
        gg.update_cycle_device(now, readings, dev_readings, cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"],
                               cm_dev.data_ptr(), mark_threshold=0, default_value=0)
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        for r in readings:
            oracle_mark(oracle, og, r, now)
        ocm, _, on = og.update_ros_costmap(cp["origin_x"], cp["origin_y"], cp["resolution"], cp["size_x"], cp["size_y"], 0, 0)
        assert gg.costmap_result().n_marked_columns == on
        assert np.array_equal(cm_dev.cpu().numpy(), ocm), "costmap bytes differ in cycle %d" % cyc
        gx, gy, gc = gg.GetFlattenedCostmap()
        assert_same_columns((gx, gy, gc), og.costmap())
        st = gg.sweep_stats()
        assert st.n_swept == st.n_survived + st.n_cleared and st.n_swept > 4_000_000
    assert_same_voxels(gg.voxels(), og.voxels())
    gg.close()


def test_costmap_survives_pool_growth_after_mark(stvl, oracle):
    """The per-column counts of the last sweep (the reference's _cost_map, which reflects the pre-Mark survivors) must survive
    a leaf / tile pool growth triggered by the Mark that follows the sweep: costmap and column queries made before the next
    sweep still see them."""
    rng = np.random.default_rng(23)
    a = np.zeros((400, 4), np.float32)
    a[:, :3] = rng.uniform(-1.5, 1.5, (400, 3))
    b = np.zeros((30000, 4), np.float32)
    b[:, :3] = rng.uniform(-40, 40, (30000, 3))          # scattered: almost one leaf (and many tiles) per point
    gg = stvl.SpatioTemporalVoxelGrid(0.05, leaf_capacity=64)
    og = oracle.OracleGrid(0.05)
    gg.Mark([stvl.MeasurementReading(a, origin=(0, 0, 0), obstacle_range=100.0, now=3.0)])
    og.mark(a, (0, 0, 0), 100.0, 3.0)
    gg.ClearFrustums([], 4.0); og.clear_frustums(4.0)
    gg.Mark([stvl.MeasurementReading(b, origin=(0, 0, 0), obstacle_range=100.0, now=4.5)])     # exhausts both pools
    og.mark(b, (0, 0, 0), 100.0, 4.5)
    cm, res = gg.UpdateROSCostmap(-4.0, -4.0, 0.05, 160, 160, mark_threshold=1, default_value=0)   # first call after the Mark: grows + replays
    ocm, _, on = og.update_ros_costmap(-4.0, -4.0, 0.05, 160, 160, mark_threshold=1, default_value=0)
    assert on > 100 and res.n_marked_columns == on
    assert np.array_equal(cm, ocm)
    gx, gy, gc = gg.GetFlattenedCostmap()
    assert_same_columns((gx, gy, gc), og.costmap())
    assert_same_voxels(gg.voxels(), og.voxels())
    assert gg.pool_stats().leaf_capacity >= 16384
    gg.close()


@pytest.mark.parametrize("pipelined", [False, True])
def test_sharded_cycle_single_rank_bitmap_path(stvl, oracle, pipelined):
    """stvl_update_cycle_sharded_device on a ONE-rank communicator (runs on a 1-GPU box): the fused cycle's costmap pass
    writes the packed cell bitmap, the (empty) exchange runs, and the root's expand kernel turns the bits into costmap bytes —
    blocking, and pipelined one cycle behind with the flush at the end — against the oracle, with an odd-sized map, a
    non-zero default value and a mark threshold."""
    import torch
    import spatio_temporal_voxel_layer_b200.synth as synth
    scene = synth.retail_scene(size=30.0, seed=3)
    gg = stvl.SpatioTemporalVoxelGrid(0.05, decay_model=0, voxel_decay=2.0, leaf_capacity=1 << 14)
    gg.comm_init(stvl.SpatioTemporalVoxelGrid.nccl_unique_id(), 0, 1)
    og = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=2.0)
    sx, sy, res, ox, oy = 397, 411, 0.05, -10.0, -10.0
    params = stvl.CostmapParams(ox, oy, res, sx, sy, 2, 255, 254)
    cm_dev = torch.full((sy, sx), 77, dtype=torch.uint8, device="cuda")
    expected, keep = [], []
    n_cycles = 6
    for cyc in range(n_cycles):
        now = 100.0 + 0.4 * cyc
        centre = (-6.0 + 0.5 * cyc, 2.3 - 13.5 + 12.0, 0.0)
        readings = [rgbd_reading(stvl, synth, scene, o, q, seed=100 * cyc + k, now=now + 0.001 * k)
                    for k, (o, q) in enumerate(synth.camera_ring(centre, 0.1 * cyc, n_cams=3))]
        dev_readings = []
        for r in readings:
            t = torch.from_numpy(np.ascontiguousarray(r.cloud, np.float32)).cuda()
            keep.append(t)
            dev_readings.append(stvl.MeasurementReading((t.data_ptr(), t.shape[0], 16, (0, 4, 8)), origin=r.origin, orientation=r.orientation,
                                                        obstacle_range=r.obstacle_range, now=r.now, filter=r.filter,
                                                        min_obstacle_height=r.min_obstacle_height, max_obstacle_height=r.max_obstacle_height))
        gg.update_cycle_sharded_device(now, readings, dev_readings, params, cm_dev.data_ptr(), root=0, pipelined=pipelined)
        og.clear_frustums(now, oracle_frustums(oracle, readings))
        for r in readings:
            oracle_mark(oracle, og, r, now)
        ocm, _, on = og.update_ros_costmap(ox, oy, res, sx, sy, mark_threshold=2, default_value=255)
        expected.append(ocm)
        gg.synchronize()
        got = cm_dev.cpu().numpy()
        if not pipelined:
            assert np.array_equal(got, expected[cyc]), "cycle %d" % cyc
            assert gg.costmap_result().n_marked_columns == on
        elif cyc >= 1:
            assert np.array_equal(got, expected[cyc - 1]), "bytes of cycle %d must be complete after call %d" % (cyc - 1, cyc)
    if pipelined:
        gg.costmap_sharded_flush(params, cm_dev.data_ptr(), root=0)
        gg.synchronize()
        assert np.array_equal(cm_dev.cpu().numpy(), expected[-1])
    assert_same_voxels(gg.voxels(), og.voxels())
    assert int((expected[-1] == 254).sum()) > 50
    gg.comm_destroy()
    gg.close()
