"""Known-answer tests that pin the CPU oracle to the reference SOURCE (the reference ships no tests or
golden vectors: SURVEY.md §4/§8c — "parity unpinned").  Every expected value below was derived by hand
from the cited reference lines, independently of oracle/stvl_oracle.cpp.
Paths: GRID = spatio_temporal_voxel_layer/src/spatio_temporal_voxel_grid.cpp,
DEPTH = .../src/frustum_models/depth_camera_frustum.cpp, LIDAR = .../three_dimensional_lidar_frustum.cpp,
LAYER = .../src/spatio_temporal_voxel_layer.cpp."""
import math

import numpy as np
import pytest


def test_voxel_size_is_float_rounded(oracle):
    # GRID:51,54 — ctor takes `const float &`, member is double
    g = oracle.OracleGrid(0.05)
    assert g.voxel_size == 0.05000000074505806
    assert 1.0 / g.voxel_size == 19.99999970197678
    g2 = oracle.OracleGrid(0.02)
    assert g2.voxel_size == 0.019999999552965164


@pytest.mark.parametrize("pt,expected", [
    ((1.0, 0.0, 0.0), (19, 0, 0)),        # 1.0 * 19.99999970197678 = 19.9999997 -> trunc 19 (not floor(1/0.05)=20)
    ((-0.05, 0.0, 0.0), (-2, 0, 0)),      # negative: shifted by -vs first, then truncated toward zero (GRID:293)
    ((0.3, -0.2, 0.7), (6, -5, 12)),      # z is shifted because y < 0 (sic, GRID:295): (0.7-0.05)*20 = 12.99 -> 12
    ((0.3, 0.2, -0.7), (6, 4, -13)),      # z negative but y >= 0: NOT shifted: -0.7*20 = -13.99 -> -13
    ((-1e-3, 1e-3, 0.0), (-1, 0, 0)),     # (-0.001-0.05)*20 = -1.02 -> -1
    ((0.0, 0.0, 0.0), (0, 0, 0)),
])
def test_quantisation_quirks(oracle, pt, expected):
    g = oracle.OracleGrid(0.05)
    # independent restatement in numpy float64 of GRID:293-303
    vs = float(np.float32(0.05))
    f = [float(np.float32(v)) for v in pt]
    x = f[0] - vs if f[0] < 0 else f[0]
    y = f[1] - vs if f[1] < 0 else f[1]
    z = f[2] - vs if f[1] < 0 else f[2]
    manual = tuple(int(math.trunc(v * (1.0 / vs))) for v in (x, y, z))
    assert manual == expected
    assert g.quantise(*pt) == expected


def test_index_to_world_two_roundings(oracle):
    # GRID:446-460: coord*vs, then + vs/2
    g = oracle.OracleGrid(0.05)
    vs = g.voxel_size
    for i in (-754, -1, 0, 7, 19, 100003):
        w = g.index_to_world(i, i, i)
        assert w[0] == i * vs + vs / 2.0


def test_decay_models(oracle):
    # GRID:320-331: 0 linear, 1 exponential, anything else persistent
    assert oracle.OracleGrid(0.05, decay_model=0, voxel_decay=15.0).decay_duration(4.0) == 11.0
    assert oracle.OracleGrid(0.05, decay_model=1, voxel_decay=15.0).decay_duration(4.0) == 15.0 * math.exp(-4.0)
    assert oracle.OracleGrid(0.05, decay_model=2, voxel_decay=15.0).decay_duration(4.0) == 15.0
    assert oracle.OracleGrid(0.05, decay_model=-1, voxel_decay=15.0).decay_duration(4.0) == 15.0
    # GRID:334-341: 1/6 * factor * t^3 in that association
    assert oracle.frustum_acceleration(3.0, 5.0) == (1. / 6. * 5.0) * ((3.0 * 3.0) * 3.0)


def test_depth_frustum_membership(oracle):
    # camera at the origin, identity orientation: +Z forward (DEPTH:76-95).  Side planes contain pairs of the corner
    # rays R_x(+-v/2) R_y(+-h/2) Z = (+-sin h2, -+sin v2 cos h2, cos v2 cos h2), hence lateral bounds
    # |x|/z <= tan(h2)/cos(v2) and |y|/z <= tan(v2); near/far planes are z = d * cos(v2) * cos(h2).
    v, h, dmin, dmax = 0.8745, 1.048, 0.1, 7.0
    f = oracle.depth_frustum(v, h, dmin, dmax, [0, 0, 0], [0, 0, 0, 1], 5.0)
    assert int(f["type"][0]) == 0
    bx = math.tan(h / 2) / math.cos(v / 2)
    by = math.tan(v / 2)
    zf = dmax * math.cos(v / 2) * math.cos(h / 2)
    zn = dmin * math.cos(v / 2) * math.cos(h / 2)
    pts, exp = [], []
    for z in (0.5, 1.0, 3.0):
        for s in (1, -1):
            pts += [[s * (bx - 0.01) * z, 0, z], [s * (bx + 0.01) * z, 0, z], [0, s * (by - 0.01) * z, z], [0, s * (by + 0.01) * z, z]]
            exp += [True, False, True, False]
    pts += [[0, 0, zf - 0.01], [0, 0, zf + 0.01], [0, 0, zn + 0.005], [0, 0, zn - 0.005], [0, 0, -1.0]]
    exp += [True, False, True, False, False]
    assert oracle.is_inside(f, pts).tolist() == exp
    # the plane block: far normal +Z, near normal -Z (DEPTH:142-149)
    p = f["p"][0].reshape(6, 6)
    np.testing.assert_allclose(p[4, :3], [0, 0, 1], atol=1e-15)
    np.testing.assert_allclose(p[5, :3], [0, 0, -1], atol=1e-15)
    np.testing.assert_allclose(p[4, 5], zf, rtol=1e-15)


def test_depth_frustum_pose(oracle):
    # rotate the camera 90 deg about Y (+Z forward becomes +X) and move it to (10, -3, 1) (DEPTH:160-174)
    q = [0, math.sin(math.pi / 4), 0, math.cos(math.pi / 4)]
    f = oracle.depth_frustum(0.8745, 1.048, 0.1, 7.0, [10, -3, 1], q, 0.0)
    assert oracle.is_inside(f, [[12, -3, 1], [8, -3, 1], [10, -3, 3], [10.5, -3.0, 1.05]]).tolist() == [True, False, False, True]


def test_bogus_frustum_is_never_inside(oracle):
    # DEPTH:71-74, 289-291
    f = oracle.depth_frustum(0.0, 0.0, 0.1, 7.0, [0, 0, 0], [0, 0, 0, 1], 0.0)
    assert int(f["type"][0]) == -1
    assert not oracle.is_inside(f, [[0, 0, 1]]).any()


def test_lidar_hourglass_membership(oracle):
    # LIDAR:77-116: min/max are RADIAL XY distances; vertical test (|z|+pad)^2 / r^2 <= tan^2(vFOV/2)
    v, pad = 0.523, 0.05
    f = oracle.lidar_frustum(v, pad, 6.29, 1.0, 8.0, [0, 0, 0], [0, 0, 0, 1], 5.0)
    t = math.tan(v / 2)
    pts = [[2, 0, 0], [2, 0, 2 * t - pad - 0.01], [2, 0, 2 * t - pad + 0.01], [0.5, 0, 0], [9, 0, 0], [-3, 4, 0.5],
           [0, -7.9, 0], [0, 8.1, 0], [2, 0, -(2 * t - pad - 0.01)]]
    exp = [True, True, False, False, False, True, True, False, True]
    assert oracle.is_inside(f, pts).tolist() == exp
    # a 90 degree wedge (hFOV = pi/2): only |azimuth| <= pi/4 in front (+X) passes (LIDAR:104-113)
    w = oracle.lidar_frustum(v, pad, math.pi / 2, 1.0, 8.0, [0, 0, 0], [0, 0, 0, 1], 5.0)
    pts = [[3, 0, 0], [3, 2.9, 0], [3, 3.1, 0], [-3, 0, 0], [0, 3, 0]]
    assert oracle.is_inside(w, pts).tolist() == [True, True, False, False, False]


def test_mark_range_and_overwrite(oracle):
    # GRID:274-291: skip if distance^2 (float) > range^2 (float) or < 0.0001; GRID:428: later marks overwrite
    g = oracle.OracleGrid(0.05)
    cloud = np.array([[1.0, 0, 0.5], [2.9, 0, 0.5], [3.2, 0, 0.5], [0.005, 0.0, 0.5], [1.0, 0, 0.5]], np.float32)
    g.mark(cloud, origin=(0, 0, 0.5), obstacle_range=3.0, now=100.0)
    ix, iy, iz, t = g.voxels()
    assert sorted(zip(ix.tolist(), iy.tolist(), iz.tolist())) == [(19, 0, 9), (58, 0, 9)]   # 0.5*19.9999997 -> 9 ; 2.9f*19.9999997 = 58.000001 -> 58
    assert (t == 100.0).all()
    g.mark(cloud[:1], origin=(0, 0, 0.5), obstacle_range=3.0, now=101.5)
    ix, iy, iz, t = g.voxels()
    assert dict(zip(ix.tolist(), t.tolist())) == {19: 101.5, 58: 100.0}
    # marking == 0 -> reading ignored (GRID:273)
    g.mark(np.array([[2.0, 0, 0.5]], np.float32), origin=(0, 0, 0.5), obstacle_range=3.0, now=102.0, marking=0.0)
    assert g.active_count() == 2


def test_sweep_semantics_and_costmap_order(oracle):
    # GRID:149-224 + SURVEY §3.1: the costmap of a cycle reflects the survivors of THAT cycle's sweep, before Mark
    g = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=15.0, pub_voxels=True)
    g.clear_frustums(10.0)                     # empty grid: early return (GRID:104-108)
    assert len(g.costmap()[0]) == 0
    cloud = np.array([[1.0, 0.02, 0.5], [1.0, 0.02, 0.56], [2.0, 0.02, 0.5]], np.float32)
    g.mark(cloud, origin=(0, 0, 0.5), obstacle_range=3.0, now=10.0)
    assert len(g.costmap()[0]) == 0            # Mark never touches _cost_map
    g.clear_frustums(12.0)
    x, y, c = g.costmap()
    vs = g.voxel_size
    got = {(round(a, 9), round(b, 9)): int(n) for a, b, n in zip(x, y, c)}
    assert got == {(round(19 * vs + vs / 2, 9), round(vs / 2, 9)): 2, (round(39 * vs + vs / 2, 9), round(vs / 2, 9)): 1}
    assert len(g.points()) == 3
    # strict `<`: a voxel exactly at its decay time survives (GRID:203), one ulp later it is cleared
    g.clear_frustums(25.0)
    assert g.active_count() == 3
    g.clear_frustums(math.nextafter(25.0, 26.0))
    assert g.active_count() == 0
    assert len(g.cleared()[0]) == 2            # two distinct (x,y) columns were cleared


def test_frustum_acceleration_first_frustum_wins(oracle):
    # GRID:171-198: the first containing frustum decides; survivor's mark time becomes value - accel
    g = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=15.0)
    g.load_voxels([20], [0], [10], [100.0])    # centre (1.025, 0.025, 0.525)
    q = [0, math.sin(math.pi / 4), 0, math.cos(math.pi / 4)]   # looking along +X
    f_slow = oracle.depth_frustum(0.8745, 1.048, 0.1, 7.0, [0, 0, 0.5], q, 1.0)
    f_fast = oracle.depth_frustum(0.8745, 1.048, 0.1, 7.0, [0, 0, 0.5], q, 100.0)
    g.clear_frustums(102.0, np.concatenate([f_slow, f_fast]))
    t = g.voxels()[3]
    assert t[0] == 100.0 - (1. / 6. * 1.0) * ((2.0 * 2.0) * 2.0)
    g2 = oracle.OracleGrid(0.05, decay_model=0, voxel_decay=15.0)
    g2.load_voxels([20], [0], [10], [100.0])
    g2.clear_frustums(102.0, np.concatenate([f_fast, f_slow]))   # 15 - 2 - 133.3 < 0 -> cleared
    assert g2.active_count() == 0


def test_update_ros_costmap(oracle):
    # LAYER:699-725 with nav2 worldToMap: threshold per voxel column, bounds touch marked and cleared cells
    g = oracle.OracleGrid(0.05)
    g.load_voxels([0, 0, 0, 5, -3, 400], [0, 0, 0, 5, -3, 0], [1, 2, 3, 1, 1, 1], [10.0] * 6)
    g.clear_frustums(11.0)
    cm, b, n = g.update_ros_costmap(-1.0, -1.0, 0.05, 100, 100, mark_threshold=0, default_value=255)
    vs = g.voxel_size
    assert n == 3 and cm.sum() == 254 * 3 + 255 * (100 * 100 - 3)      # column x=400 is outside the 5 m map
    assert cm[int((0 * vs + vs / 2 + 1.0) / 0.05), int((0 * vs + vs / 2 + 1.0) / 0.05)] == 254
    assert cm[int((-3 * vs + vs / 2 + 1.0) / 0.05), int((-3 * vs + vs / 2 + 1.0) / 0.05)] == 254
    np.testing.assert_array_equal(b, [-3 * vs + vs / 2, -3 * vs + vs / 2, 5 * vs + vs / 2, 5 * vs + vs / 2])
    cm, b, n = g.update_ros_costmap(-1.0, -1.0, 0.05, 100, 100, mark_threshold=2, default_value=0)
    assert n == 1 and cm.sum() == 254


def test_reset_area(oracle):
    # GRID:397-418: strict inequalities on voxel centres; invert flips
    g = oracle.OracleGrid(0.05)
    g.load_voxels([0, 10, 20, 30], [0, 0, 0, 0], [1, 1, 1, 1], [1.0] * 4)
    # with invert_area == false the test `in_range == invert_area` clears everything OUTSIDE the rectangle
    g.reset_area(0.4, -1.0, 1.2, 1.0, invert=False)
    assert sorted(g.voxels()[0].tolist()) == [10, 20]
    g.reset_area(0.4, -1.0, 0.8, 1.0, invert=True)    # invert: clear INSIDE
    assert sorted(g.voxels()[0].tolist()) == [20]
    assert g.reset() and g.active_count() == 0


def test_passthrough_and_voxel_filters(oracle):
    # measurement_buffer.cpp:153-175
    cloud = np.array([[0.01, 0.01, 0.5], [0.02, 0.02, 0.52], [0.30, 0.0, 0.5], [0.0, 0.0, 0.1], [0.0, 0.0, 2.5],
                      [np.nan, 0, 1.0]], np.float32)
    p = oracle.filter_passthrough(cloud, 0.3, 2.0)
    assert p.shape == (3, 3) and np.array_equal(p, cloud[:3, :3])
    v = oracle.filter_voxel(cloud, 0.05, 0.3, 2.0, 0)
    assert v.shape == (2, 3)
    c = (cloud[0, :3] + cloud[1, :3]) / np.float32(2)
    assert any(np.array_equal(r, c) for r in v) and any(np.array_equal(r, cloud[2, :3]) for r in v)
    assert oracle.filter_voxel(cloud, 0.05, 0.3, 2.0, 2).shape == (1, 3)


def test_update_costs_known_answers(oracle):
    """updateCosts building blocks, hand-checked: setConvexPolygonCost fills the cells between the
    Bresenham outlines column by column, fails without side effects when a vertex is off the map; updateWithMax never
    lowers a cost and treats 255 as "no information" on both sides; updateWithOverwrite copies everything but 255."""
    cm = np.full((12, 16), 254, np.uint8)
    # axis-aligned rectangle, resolution 0.5, origin (-1, -2): vertices map to cells (3..9, 2..7)
    assert oracle.set_convex_polygon_cost(cm, -1.0, -2.0, 0.5, [(0.6, -0.9), (3.9, -0.9), (3.9, 1.9), (0.6, 1.9)], 0)
    want = np.full((12, 16), 254, np.uint8)
    want[2:8, 3:10] = 0
    assert np.array_equal(cm, want)
    # a triangle: (2,2) -> (10,2) -> (2,10) in cells: column x holds rows 2 .. 12 - x (the hypotenuse is an exact diagonal)
    cm = np.full((16, 16), 9, np.uint8)
    assert oracle.set_convex_polygon_cost(cm, 0.0, 0.0, 1.0, [(2.5, 2.5), (10.5, 2.5), (2.5, 10.5)], 0)
    for x in range(16):
        col = np.nonzero(cm[:, x] == 0)[0]
        if 2 <= x <= 10:
            assert col.min() == 2 and col.max() == 12 - x and len(col) == 11 - x, (x, col)
        else:
            assert len(col) == 0
    # a vertex off the map: nothing changes
    before = cm.copy()
    assert not oracle.set_convex_polygon_cost(cm, 0.0, 0.0, 1.0, [(2.5, 2.5), (20.5, 2.5), (2.5, 10.5)], 0)
    assert np.array_equal(cm, before)
    layer = np.array([[0, 100, 255, 254], [254, 255, 0, 50]], np.uint8)
    master = np.array([[255, 200, 7, 10], [253, 1, 255, 60]], np.uint8)
    m = master.copy()
    oracle.update_costs(layer, m, 0, 0, 4, 2, 1)      # max
    assert m.tolist() == [[0, 200, 7, 254], [254, 1, 0, 60]]
    m = master.copy()
    oracle.update_costs(layer, m, 0, 0, 4, 2, 0)      # overwrite
    assert m.tolist() == [[0, 100, 7, 254], [254, 1, 0, 50]]
    m = master.copy()
    oracle.update_costs(layer, m, 1, 0, 3, 1, 0)      # window [1,3) x [0,1)
    assert m.tolist() == [[255, 100, 7, 10], [253, 1, 255, 60]]


def test_polygon_fill_equals_column_extent_of_the_outline(oracle):
    """The device kernel fills every column between the lowest and the highest Bresenham outline cell instead of
    sorting the outline and scanning it like nav2's convexFillCells.  An independent numpy model of that rule must agree
    with the literal restatement in the oracle on random convex polygons (both windings, slivers, single-cell ones)."""
    rng = np.random.default_rng(3)

    def outline(x0, y0, x1, y1):
        dx, dy = x1 - x0, y1 - y0
        adx, ady = abs(dx), abs(dy)
        sx, sy = (1 if dx > 0 else -1), (1 if dy > 0 else -1)
        cells, x, y = [], x0, y0
        if adx >= ady:
            err = adx // 2
            for _ in range(adx):
                cells.append((x, y)); x += sx; err += ady
                if err >= adx:
                    y += sy; err -= adx
        else:
            err = ady // 2
            for _ in range(ady):
                cells.append((x, y)); y += sy; err += adx
                if err >= ady:
                    x += sx; err -= ady
        cells.append((x, y))
        return cells

    sx_, sy_, res, ox, oy = 97, 83, 0.1, -2.0, -3.0
    for it in range(200):
        nv = int(rng.integers(3, 10))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        if it % 3 == 0:
            ang = ang[::-1]
        a, b = rng.uniform(0.02, 3.5, 2)
        th = rng.uniform(0, np.pi)
        cx, cy = rng.uniform(ox + 3.6, ox + sx_ * res - 3.6), rng.uniform(oy + 3.6, oy + sy_ * res - 3.6)
        px, py = a * np.cos(ang), b * np.sin(ang)
        poly = np.stack([cx + px * np.cos(th) - py * np.sin(th), cy + px * np.sin(th) + py * np.cos(th)], 1)
        got = np.full((sy_, sx_), 200, np.uint8)
        assert oracle.set_convex_polygon_cost(got, ox, oy, res, poly, 0)
        m = [(int((x - ox) / res), int((y - oy) / res)) for x, y in poly]
        lo, hi = {}, {}
        for k in range(nv):
            for (x, y) in outline(*m[k], *m[(k + 1) % nv]):
                lo[x] = min(lo.get(x, y), y); hi[x] = max(hi.get(x, y), y)
        want = np.full((sy_, sx_), 200, np.uint8)
        for x in lo:
            want[lo[x]:hi[x] + 1, x] = 0
        assert np.array_equal(got, want), it


def test_cloud_transform_against_numpy_float32(oracle):
    """tf2::doTransform restatement: quaternion -> float matrix as Eigen's toRotationMatrix, then
    ((m0*x + m1*y) + m2*z) + t per row with every operation rounded to float32 — checked against numpy float32 scalars
    evaluated in that order, and against the double-precision rotation within float accuracy."""
    rng = np.random.default_rng(21)
    f = np.float32
    for _ in range(20):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.uniform(-30, 30, 3)
        m = oracle.transform_from_quat(t, q)
        x, y, z, w = (f(v) for v in q)
        tx, ty, tz = f(2) * x, f(2) * y, f(2) * z
        twx, twy, twz = tx * w, ty * w, tz * w
        txx, txy, txz, tyy, tyz, tzz = tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
        want = np.array([[f(1) - (tyy + tzz), txy - twz, txz + twy, f(t[0])],
                         [txy + twz, f(1) - (txx + tzz), tyz - twx, f(t[1])],
                         [txz - twy, tyz + twx, f(1) - (txx + tyy), f(t[2])]], np.float32)
        assert np.array_equal(m.reshape(3, 4).view(np.uint32), want.view(np.uint32))
        pts = np.zeros((500, 4), np.float32)
        pts[:, :3] = rng.uniform(-8, 8, (500, 3))
        out = oracle.transform_cloud(pts, m)
        for i in range(0, 500, 37):
            px, py, pz = pts[i, :3]
            for r in range(3):
                ref = ((want[r, 0] * px + want[r, 1] * py) + want[r, 2] * pz) + want[r, 3]
                assert out[i, r].view(np.uint32) == np.float32(ref).view(np.uint32)
        # sanity against the exact rotation
        xq, yq, zq, wq = q
        R = np.array([[1 - 2 * (yq * yq + zq * zq), 2 * (xq * yq - zq * wq), 2 * (xq * zq + yq * wq)],
                      [2 * (xq * yq + zq * wq), 1 - 2 * (xq * xq + zq * zq), 2 * (yq * zq - xq * wq)],
                      [2 * (xq * zq - yq * wq), 2 * (yq * zq + xq * wq), 1 - 2 * (xq * xq + yq * yq)]])
        assert np.abs(out[:, :3] - (pts[:, :3].astype(np.float64) @ R.T + t)).max() < 2e-4
