"""A second, independent restatement of the cycle in plain Python floats (IEEE doubles, one correctly rounded operation
per expression node, no FMA), written from the reference SOURCE lines cited below, run against the C++ oracle on a small
seeded multi-cycle scene.  The reference ships no tests: two restatements that agree bit for bit do not
make the oracle "pinned", but a transcription slip would have to be made twice, the same way, to go unnoticed.
GRID = spatio_temporal_voxel_layer/src/spatio_temporal_voxel_grid.cpp, DEPTH = .../frustum_models/depth_camera_frustum.cpp,
LAYER = .../src/spatio_temporal_voxel_layer.cpp."""
import math

import numpy as np


class ModelGrid:
    def __init__(self, voxel_size, decay_model, voxel_decay):
        self.vs = float(np.float32(voxel_size))      # GRID:51,54: float parameter, double member
        self.inv = 1.0 / self.vs                     # OpenVDB UniformScaleMap caches the inverse
        self.decay_model, self.voxel_decay = decay_model, voxel_decay
        self.vox = {}                                # (ix, iy, iz) -> mark time
        self.columns = {}                            # (x, y) world doubles -> count   (_cost_map)
        self.cleared = set()

    def mark(self, cloud, origin, obstacle_range, now):
        # GRID:269-309
        mark_range_2 = float(np.float32(obstacle_range * obstacle_range))
        ox, oy, oz = (float(v) for v in origin)
        for fx, fy, fz in np.asarray(cloud, np.float32)[:, :3]:
            x, y, z = float(fx), float(fy), float(fz)
            d2 = float(np.float32(((x - ox) * (x - ox) + (y - oy) * (y - oy)) + (z - oz) * (z - oz)))
            if d2 > mark_range_2 or d2 < 0.0001:
                continue
            X = x - self.vs if x < 0 else x
            Y = y - self.vs if y < 0 else y
            Z = z - self.vs if y < 0 else z          # sic: z is shifted on y's sign (GRID:295)
            key = (int(X * self.inv), int(Y * self.inv), int(Z * self.inv))   # int(): truncation toward zero
            self.vox[key] = now                      # setValueOn: overwrite (GRID:428)

    def centre(self, i):
        return i * self.vs + self.vs / 2.0           # GRID:446-460: two roundings

    @staticmethod
    def inside_depth(planes, px, py, pz):
        # DEPTH:286-304 with Eigen's normalize(): divide by sqrt(squaredNorm) when squaredNorm > 0
        for k in range(6):
            nx, ny, nz, qx, qy, qz = planes[6 * k:6 * k + 6]
            dx, dy, dz = px - qx, py - qy, pz - qz
            n2 = (dx * dx + dy * dy) + dz * dz
            if n2 > 0.0:
                n = math.sqrt(n2)
                dx, dy, dz = dx / n, dy / n, dz / n
            if (nx * dx + ny * dy) + nz * dz > 0.0:
                return False
        return True

    @staticmethod
    def lidar_model(vfov, vpad, hfov, min_d, max_d, position, quat_xyzw):
        """ThreeDimensionalLidarFrustum (LIDAR = .../frustum_models/three_dimensional_lidar_frustum.cpp:44-116) as a
        predicate, built from the raw parameters — nothing of the oracle's frustum block is used."""
        h_half, tan2 = hfov / 2.0, math.tan(vfov / 2.0) * math.tan(vfov / 2.0)
        min2, max2, full = min_d * min_d, max_d * max_d, hfov > 6.27
        qx, qy, qz, qw = -quat_xyzw[0], -quat_xyzw[1], -quat_xyzw[2], quat_xyzw[3]      # conjugate (LIDAR:72)
        px, py, pz = position

        def cross(a, b):
            return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])

        def inside(x, y, z):
            v = (x - px, y - py, z - pz)
            uv = cross((qx, qy, qz), v)                  # Eigen: uv = 2 * (q.vec x v); v + w * uv + q.vec x uv
            uv = (uv[0] + uv[0], uv[1] + uv[1], uv[2] + uv[2])
            c2 = cross((qx, qy, qz), uv)
            t = tuple((v[i] + qw * uv[i]) + c2[i] for i in range(3))
            r2 = (t[0] * t[0]) + (t[1] * t[1])
            if r2 > max2 or r2 < min2:
                return False
            vp = math.fabs(t[2]) + vpad
            if r2 == 0.0:
                if vp * vp > 0.0:                        # x / 0 = +inf > tan^2
                    return False
            elif (vp * vp / r2) > tan2:
                return False
            if not full:
                half_pi = math.pi / 2
                if t[0] > 0:
                    if math.fabs(math.atan(t[1] / t[0])) > h_half:
                        return False
                else:
                    ratio = (t[0] / t[1]) if t[1] != 0.0 else (math.copysign(math.inf, t[0]) if t[0] != 0.0 else math.nan)
                    if math.fabs(math.atan(ratio)) + half_pi > h_half:
                        return False
            return True
        return inside

    def clear_frustums(self, now, frustums):
        # GRID:96-224; frustums: list of (36 plane doubles OR a predicate, accel_factor), in reading order
        self.columns, self.cleared = {}, set()
        if not self.vox:
            return
        for key in list(self.vox):
            v = self.vox[key]
            t = now - v
            if self.decay_model == 0:
                base = self.voxel_decay - t
            elif self.decay_model == 1:
                base = self.voxel_decay * math.exp(-t)
            else:
                base = self.voxel_decay
            cx, cy, cz = self.centre(key[0]), self.centre(key[1]), self.centre(key[2])
            gone = False
            for planes, factor in frustums:
                if (planes(cx, cy, cz) if callable(planes) else self.inside_depth(planes, cx, cy, cz)):
                    accel = ((1. / 6.) * factor) * ((t * t) * t)     # GRID:334-341
                    if base - accel < 0.:
                        gone = True
                    else:
                        self.vox[key] = v - accel
                    break                                            # first containing frustum decides
            else:
                gone = base < 0.
            if gone:
                del self.vox[key]
                self.cleared.add((cx, cy))
            else:
                self.columns[(cx, cy)] = self.columns.get((cx, cy), 0) + 1

    def costmap(self, origin_x, origin_y, resolution, size_x, size_y, mark_threshold, default_value):
        # LAYER:699-725 with nav2's worldToMap
        cm = np.full((size_y, size_x), default_value, np.uint8)
        b = [1e308, 1e308, -1e308, -1e308]

        def touch(x, y):
            b[0], b[1], b[2], b[3] = min(b[0], x), min(b[1], y), max(b[2], x), max(b[3], y)

        for (x, y), count in self.columns.items():
            if count < mark_threshold or x < origin_x or y < origin_y:
                continue
            mx, my = int((x - origin_x) / resolution), int((y - origin_y) / resolution)
            if mx < size_x and my < size_y:
                cm[my, mx] = 254
                touch(x, y)
        for x, y in self.cleared:
            touch(x, y)
        return cm, b


def test_python_model_agrees_with_the_oracle(oracle):
    from spatio_temporal_voxel_layer_b200 import synth
    p = synth.RGBD
    scene = synth.retail_scene(size=12.0, seed=21)
    for decay_model, voxel_decay in ((0, 1.1), (1, 1.1), (2, 5.0)):
        og = oracle.OracleGrid(0.05, 0.0, decay_model, voxel_decay, False)
        mg = ModelGrid(0.05, decay_model, voxel_decay)
        n_cleared_total = 0
        for cyc in range(5):
            now = 50.0 + 0.45 * cyc
            centre = (-1.5 + 0.4 * cyc, 0.6, 0.0)
            readings = [(o, q, synth.depth_cloud(scene, o, q, width=48, height=36, seed=7 * cyc + k))
                        for k, (o, q) in enumerate(synth.camera_ring(centre, 0.25 * cyc, n_cams=2))]
            blocks = [oracle.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], o, q, p["accel"]) for o, q, _ in readings]
            og.clear_frustums(now, np.concatenate(blocks))
            mg.clear_frustums(now, [(b["p"][0].tolist(), float(b["accel_factor"][0])) for b in blocks])
            n_cleared_total += len(mg.cleared)
            for k, (o, q, cloud) in enumerate(readings):
                pts = oracle.filter_passthrough(cloud, p["min_obstacle_height"], p["max_obstacle_height"])
                og.mark(pts, o, p["obstacle_range"], now + 1e-3 * k)
                mg.mark(pts, o, p["obstacle_range"], now + 1e-3 * k)
            # voxel store: keys and mark times, bit for bit
            ix, iy, iz, t = og.voxels()
            got = {(int(a), int(b), int(c)): float(v) for a, b, c, v in zip(ix, iy, iz, t)}
            assert got.keys() == mg.vox.keys(), "cycle %d: key sets differ" % cyc
            assert all(np.float64(got[k]).view(np.uint64) == np.float64(mg.vox[k]).view(np.uint64) for k in got), "cycle %d: mark times differ" % cyc
            # column counts and cleared cells of this cycle's sweep
            x, y, c = og.costmap()
            assert {(float(a), float(b)): int(n) for a, b, n in zip(x, y, c)} == mg.columns
            cxs, cys = og.cleared()
            assert {(float(a), float(b)) for a, b in zip(cxs, cys)} == mg.cleared
            # costmap bytes and bounds
            ocm, ob, on = og.update_ros_costmap(-5.0, -5.0, 0.05, 200, 200, mark_threshold=cyc % 3, default_value=255)
            mcm, mb = mg.costmap(-5.0, -5.0, 0.05, 200, 200, cyc % 3, 255)
            assert np.array_equal(ocm, mcm)
            np.testing.assert_array_equal(ob, np.array(mb))
        assert len(mg.vox) > 150
        if decay_model != 2:
            assert n_cleared_total > 0      # something actually expired / was accelerated away


def test_python_model_agrees_with_the_oracle_lidar(oracle):
    """The same with 3-D lidar frustums (hourglass model, full circle and a 2 rad wedge that goes through atan) next to a
    depth camera, exponential decay: the lidar predicate is modelled from the raw sensor parameters."""
    from spatio_temporal_voxel_layer_b200 import synth
    p, v = synth.RGBD, synth.VLP16
    scene = synth.retail_scene(size=12.0, seed=22)
    og = oracle.OracleGrid(0.05, 0.0, 1, 0.9, False)
    mg = ModelGrid(0.05, 1, 0.9)
    n_cleared_total = 0
    for cyc in range(5):
        now = 80.0 + 0.5 * cyc
        lo, lq = (-1.0 + 0.5 * cyc, 0.7, 0.9), synth.yaw_quat(0.7 * cyc)
        co, cq = (-1.0 + 0.5 * cyc, 0.7, 0.6), synth.optical_quat(0.3 - 0.5 * cyc)
        hfov = 6.29 if cyc % 2 == 0 else 2.0
        lidar_cloud = synth.lidar_cloud(scene, lo, lq, rings=8, azimuths=180, seed=cyc)
        depth = synth.depth_cloud(scene, co, cq, width=48, height=36, seed=30 + cyc)
        lb = oracle.lidar_frustum(v["vfov"], v["vpad"], hfov, v["min_z"], v["max_z"], lo, lq, v["accel"])
        db = oracle.depth_frustum(p["vfov"], p["hfov"], p["min_z"], p["max_z"], co, cq, p["accel"])
        og.clear_frustums(now, np.concatenate([lb, db]))
        mg.clear_frustums(now, [(ModelGrid.lidar_model(v["vfov"], v["vpad"], hfov, v["min_z"], v["max_z"], lo, lq), v["accel"]),
                                (db["p"][0].tolist(), float(db["accel_factor"][0]))])
        n_cleared_total += len(mg.cleared)
        for k, (o, cloud, lim) in enumerate(((lo, lidar_cloud, v), (co, depth, p))):
            pts = oracle.filter_passthrough(cloud, lim["min_obstacle_height"], lim["max_obstacle_height"])
            og.mark(pts, o, lim["obstacle_range"], now + 1e-3 * k)
            mg.mark(pts, o, lim["obstacle_range"], now + 1e-3 * k)
        ix, iy, iz, t = og.voxels()
        got = {(int(a), int(b), int(c)): float(tv) for a, b, c, tv in zip(ix, iy, iz, t)}
        assert got.keys() == mg.vox.keys(), "cycle %d: key sets differ" % cyc
        assert all(np.float64(got[k]).view(np.uint64) == np.float64(mg.vox[k]).view(np.uint64) for k in got)
        x, y, c = og.costmap()
        assert {(float(a), float(b)): int(n) for a, b, n in zip(x, y, c)} == mg.columns
    assert len(mg.vox) > 150 and n_cleared_total > 0
