"""Synthetic sensor data for the BASELINE.json configs (numpy only, fixed seeds).

Scenes are sets of axis-aligned boxes (walls, shelving, pallets); depth cameras and spinning lidars
are ray-cast against them, so clouds have the structure real ones have (dense surface patches, many
points per voxel, vertical sheets) rather than uniform noise.  Used by tests/ and bench.py.
"""
import math

import numpy as np

# parameters of the reference's example configs (reference spatio_temporal_voxel_layer/example/
# standard_indoor_environment_config.yaml:7-9,25-27,39-43 and vlp16_config.yaml:39-45)
RGBD = dict(vfov=0.8745, hfov=1.048, min_z=0.1, max_z=7.0, accel=5.0, obstacle_range=3.0,
            min_obstacle_height=0.3, max_obstacle_height=2.0)
VLP16 = dict(vfov=0.523, vpad=0.05, hfov=6.29, min_z=1.0, max_z=8.0, accel=5.0, obstacle_range=8.0,
             min_obstacle_height=0.2, max_obstacle_height=2.5)


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_to_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def matrix_to_quat(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = math.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2
    y = math.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2
    z = math.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2
    x = math.copysign(x, R[2, 1] - R[1, 2])
    y = math.copysign(y, R[0, 2] - R[2, 0])
    z = math.copysign(z, R[1, 0] - R[0, 1])
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


def optical_quat(yaw, pitch=0.0):
    """Orientation (x,y,z,w) of an optical frame (+Z forward, +X right, +Y down) looking along `yaw`
    in the world XY plane, tilted down by `pitch` — the frame DepthCameraFrustum assumes
    (reference depth_camera_frustum.cpp:76-95: rays are deflections of +Z)."""
    cy, sy, cp, sp = math.cos(yaw), math.sin(yaw), math.cos(pitch), math.sin(pitch)
    fwd = np.array([cy * cp, sy * cp, -sp])
    right = np.array([sy, -cy, 0.0])
    down = np.cross(fwd, right)
    return matrix_to_quat(np.stack([right, down, fwd], axis=1))


def yaw_quat(yaw):
    return np.array([0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)])


class Scene:
    """Axis-aligned boxes; raycast() returns the nearest hit distance per ray (inf = miss)."""

    def __init__(self, boxes):
        self.lo = np.asarray([b[0] for b in boxes], np.float64)
        self.hi = np.asarray([b[1] for b in boxes], np.float64)

    def raycast(self, origin, dirs, max_range=None, chunk=1 << 13):
        origin = np.asarray(origin, np.float64)
        out = np.full(len(dirs), np.inf)
        lo, hi = self.lo, self.hi
        if max_range is not None:   # only boxes that can be hit within max_range
            gap = np.maximum(np.maximum(lo - origin, origin - hi), 0.0)
            near = (gap * gap).sum(axis=1) <= max_range * max_range
            lo, hi = lo[near], hi[near]
        if len(lo) == 0:
            return out
        for s in range(0, len(dirs), chunk):
            d = dirs[s:s + chunk]
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = 1.0 / d
                t0 = (lo[None, :, :] - origin) * inv[:, None, :]
                t1 = (hi[None, :, :] - origin) * inv[:, None, :]
            tn = np.nanmax(np.minimum(t0, t1), axis=2)
            tf = np.nanmin(np.maximum(t0, t1), axis=2)
            hit = (tf >= np.maximum(tn, 0.0))
            t = np.where(hit, np.where(tn > 0, tn, tf), np.inf)
            out[s:s + chunk] = t.min(axis=1)
        return out


def retail_scene(size=75.0, aisle_pitch=3.0, shelf_depth=0.6, shelf_height=2.2, seed=7, origin=(0.0, 0.0)):
    """A store floor: perimeter walls plus rows of shelving split by cross aisles. Spans negative and
    positive coordinates (exercises the reference's negative-coordinate quantisation quirks)."""
    rng = np.random.default_rng(seed)
    h = size / 2
    ox, oy = origin
    boxes = [((ox - h - 0.2, oy - h, 0), (ox - h, oy + h, 3.0)), ((ox + h, oy - h, 0), (ox + h + 0.2, oy + h, 3.0)),
             ((ox - h, oy - h - 0.2, 0), (ox + h, oy - h, 3.0)), ((ox - h, oy + h, 0), (ox + h, oy + h + 0.2, 3.0))]
    y = -h + 2.0
    while y + shelf_depth < h - 2.0:
        x = -h + 2.0
        while x < h - 4.0:
            length = float(rng.uniform(6.0, 12.0))
            x1 = min(x + length, h - 2.0)
            boxes.append(((ox + x, oy + y, 0.0), (ox + x1, oy + y + shelf_depth, shelf_height)))
            x = x1 + float(rng.uniform(1.8, 2.6))
        y += aisle_pitch
    for _ in range(40):   # pallets / displays in the aisles
        cx, cy = rng.uniform(-h + 3, h - 3, 2)
        sx, sy, sz = rng.uniform(0.3, 1.0), rng.uniform(0.3, 1.0), rng.uniform(0.5, 1.6)
        boxes.append(((ox + cx, oy + cy, 0.0), (ox + cx + sx, oy + cy + sy, sz)))
    return Scene(boxes)


def depth_cloud(scene, origin, quat, width=640, height=480, hfov=RGBD["hfov"], vfov=RGBD["vfov"],
                max_range=10.0, noise=0.004, seed=0, nan_misses=True):
    """A width x height organised XYZ cloud in the GLOBAL frame, float32 (N,4) (x,y,z,pad): point_step 16."""
    rng = np.random.default_rng(seed)
    u = (np.arange(width) + 0.5) / width * 2 - 1
    v = (np.arange(height) + 0.5) / height * 2 - 1
    uu, vv = np.meshgrid(u * math.tan(hfov / 2), v * math.tan(vfov / 2))
    d = np.stack([uu.ravel(), vv.ravel(), np.ones(uu.size)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dw = d @ quat_to_matrix(quat).T
    t = scene.raycast(origin, dw, max_range)
    t = t + rng.normal(0.0, noise, t.shape) * np.minimum(t, max_range)
    miss = ~np.isfinite(t) | (t > max_range)
    pts = np.asarray(origin)[None, :] + dw * np.where(miss, 0.0, t)[:, None]
    out = np.zeros((len(pts), 4), np.float32)
    out[:, :3] = pts
    if nan_misses:
        out[miss, :3] = np.nan
    else:
        out = out[~miss]
    return out


def lidar_cloud(scene, origin, quat, rings=16, azimuths=1800, vfov=VLP16["vfov"], max_range=30.0, noise=0.01,
                seed=0):
    """A spinning 3-D lidar scan in the GLOBAL frame, float32 (N,4); misses are dropped (unorganised)."""
    rng = np.random.default_rng(seed)
    el = np.linspace(-vfov / 2, vfov / 2, rings)
    az = np.arange(azimuths) * (2 * math.pi / azimuths)
    ee, aa = np.meshgrid(el, az)
    d = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], axis=2).reshape(-1, 3)
    dw = d @ quat_to_matrix(quat).T
    t = scene.raycast(origin, dw, max_range)
    ok = np.isfinite(t) & (t < max_range)
    t = t[ok] + rng.normal(0.0, noise, ok.sum())
    out = np.zeros((ok.sum(), 4), np.float32)
    out[:, :3] = np.asarray(origin)[None, :] + dw[ok] * t[:, None]
    return out


def surface_voxels(scene, vs, zmin, zmax, target=None, seed=0):
    """Unique voxel indices (ix,iy,iz) on the vertical faces of the scene's boxes between zmin and zmax —
    the map a robot would have built after driving every aisle.  Thin vertical sheets: no solid 8^3 blocks."""
    rng = np.random.default_rng(seed)
    cols = []
    for lo, hi in zip(scene.lo, scene.hi):
        x0, x1 = int(math.floor(lo[0] / vs)), int(math.floor(hi[0] / vs))
        y0, y1 = int(math.floor(lo[1] / vs)), int(math.floor(hi[1] / vs))
        xs = np.arange(x0, x1 + 1)
        ys = np.arange(y0, y1 + 1)
        top = min(hi[2], zmax)
        z0, z1 = int(math.ceil(max(lo[2], zmin) / vs)), int(math.floor(top / vs))
        if z1 < z0:
            continue
        zs = np.arange(z0, z1 + 1)
        for face_x in (x0, x1):
            yy, zz = np.meshgrid(ys, zs)
            cols.append(np.stack([np.full(yy.size, face_x), yy.ravel(), zz.ravel()], 1))
        for face_y in (y0, y1):
            xx, zz = np.meshgrid(xs, zs)
            cols.append(np.stack([xx.ravel(), np.full(xx.size, face_y), zz.ravel()], 1))
    vox = np.unique(np.concatenate(cols).astype(np.int64), axis=0)
    if target is not None:
        if len(vox) < target:
            # thicken: add the voxel layer behind each face until the target is reached
            extra = [vox]
            k = 1
            while sum(len(e) for e in extra) < target * 1.05 and k < 6:
                for dx, dy in ((k, 0), (-k, 0), (0, k), (0, -k)):
                    extra.append(vox + np.array([dx, dy, 0]))
                k += 1
            vox = np.unique(np.concatenate(extra), axis=0)
        if len(vox) > target:
            vox = vox[np.sort(rng.choice(len(vox), target, replace=False))]
    return vox[:, 0].astype(np.int32), vox[:, 1].astype(np.int32), vox[:, 2].astype(np.int32)


def camera_ring(center, yaw0, n_cams=7, radius=0.25, height=0.6, pitch=0.0):
    """n_cams RGBD cameras on a ring around the robot centre, yaw k*2*pi/n (BASELINE config 2)."""
    poses = []
    for k in range(n_cams):
        yaw = yaw0 + k * 2 * math.pi / n_cams
        o = (center[0] + radius * math.cos(yaw), center[1] + radius * math.sin(yaw), height)
        poses.append((o, optical_quat(yaw, pitch)))
    return poses
