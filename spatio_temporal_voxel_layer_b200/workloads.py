"""The BASELINE.json configs as concrete, seeded workloads (numpy only).  bench.py and the tests feed the SAME
arrays to the GPU path and to the CPU oracle.

config2: 7 RGBD cameras 640x480 with frustum acceleration, 0.05 m voxels, retail-scale map pre-seeded to
         710,765 active voxels (reference README.md:53), linear decay 15 s.
"""
import math

import numpy as np

from . import synth

RETAIL_VOXELS = 710765          # reference README.md:53
IMG_W, IMG_H = 640, 480


class Cycle:
    """One update cycle: the clearing/marking readings of every sensor at one robot pose."""

    def __init__(self, now, poses, clouds):
        self.now = now
        self.poses = poses          # [(origin, quat)]
        self.clouds = clouds        # [float32 (N,4)]


class Config2:
    name = "config2: 7xRGBD 640x480, 0.05 m, retail map 710,765 voxels, linear decay 15 s, frustum accel 5.0"
    voxel_size = 0.05
    decay_model = 0
    voxel_decay = 15.0
    n_cams = 7

    def __init__(self, n_poses=4, ring=8, seed=1002, width=IMG_W, height=IMG_H, preseed=RETAIL_VOXELS, now0=1.7e9,
                 dt=0.01, slab_origin=(0.0, 0.0), scene_size=75.0):
        """n_poses distinct robot poses are ray-cast; `ring` cycle inputs are derived from them with fresh
        sensor noise, so consecutive cycles never re-read the same cloud (ring*34.4 MB > 126 MB L2)."""
        self.p = synth.RGBD
        self.now0 = now0
        self.dt = dt
        self.width, self.height = width, height
        self.scene = synth.retail_scene(size=scene_size, seed=7, origin=slab_origin)
        vs = float(np.float32(self.voxel_size))
        rng = np.random.default_rng(seed)
        ix, iy, iz = synth.surface_voxels(self.scene, vs, self.p["min_obstacle_height"], self.p["max_obstacle_height"],
                                          target=preseed, seed=seed)
        self.seed_vox = (ix, iy, iz)
        self.seed_t = now0 - rng.uniform(0.0, self.voxel_decay, len(ix))
        # robot drives along the aisle between the shelf rows at y = 0.5 .. 1.1 and 3.5 .. 4.1
        ox, oy = slab_origin
        base = []
        for k in range(n_poses):
            centre = (ox - 6.0 + 1.5 * k, oy + 2.3, 0.0)
            poses = synth.camera_ring(centre, 0.05 + 0.11 * k, n_cams=self.n_cams)
            hits = [synth.depth_cloud(self.scene, o, q, width=width, height=height, noise=0.0, seed=0) for o, q in poses]
            base.append((poses, hits))
        self.cycles = []
        for c in range(ring):
            poses, hits = base[c % n_poses]
            clouds = []
            for k, ((o, q), h) in enumerate(zip(poses, hits)):
                r = np.random.default_rng(seed * 1000 + c * 16 + k)
                cl = h.copy()
                d = cl[:, :3] - np.asarray(o, np.float32)
                cl[:, :3] += (d * r.normal(0.0, 0.004, (len(cl), 1))).astype(np.float32)   # range noise along the ray
                clouds.append(np.ascontiguousarray(cl))
            self.cycles.append(Cycle(0.0, poses, clouds))

    @property
    def points_per_cycle(self):
        return sum(len(c) for c in self.cycles[0].clouds)

    def cycle(self, step):
        """(now, poses, clouds) of step `step` — `now` advances by dt per step."""
        c = self.cycles[step % len(self.cycles)]
        return self.now0 + self.dt * (step + 1), c.poses, c.clouds

    def costmap_params(self):
        """A costmap covering the store: 75 m + margin at 0.05 m."""
        size = 1600
        res = 0.05
        ox, oy = self.scene.lo[:, 0].min() - 1.0, self.scene.lo[:, 1].min() - 1.0
        return dict(origin_x=float(ox), origin_y=float(oy), resolution=res, size_x=size, size_y=size,
                    mark_threshold=0, default_value=0)


class Sensor:
    """One observation source of a cycle: pose, cloud (float32 (N,4), global frame) and the reference's per-source parameters."""

    def __init__(self, model_type, origin, quat, cloud, params, filter_name):
        self.model_type = model_type      # 0 depth camera, 1 3-D lidar (measurement_reading.h:49-53)
        self.origin, self.quat, self.cloud = origin, quat, cloud
        self.p = params                   # synth.RGBD / synth.VLP16 style dict
        self.filter = filter_name         # "passthrough" | "voxel" | "none"


class GrowingConfig:
    """Configs 1, 3 and 4 of BASELINE.json: the store starts EMPTY and grows while the robot drives (`n_cycles` distinct
    cycles, dt seconds apart).  cycle(j) -> (now, [Sensor...]); clearing and marking use the same sensors, as in the
    reference's example configs."""
    voxel_size = 0.05
    decay_model = 0
    voxel_decay = 15.0
    voxel_min_points = 0

    def __init__(self, n_cycles, dt=0.2, now0=1.7e9):
        self.n_cycles, self.dt, self.now0 = n_cycles, dt, now0
        self.cycles = []

    def cycle(self, j):
        return self.now0 + self.dt * (j + 1), self.cycles[j % len(self.cycles)]

    @property
    def points_per_cycle(self):
        return sum(len(s.cloud) for s in self.cycles[0])

    def costmap_params(self):
        return dict(origin_x=-16.0, origin_y=-16.0, resolution=self.voxel_size, size_x=int(round(32.0 / self.voxel_size)),
                    size_y=int(round(32.0 / self.voxel_size)), mark_threshold=0, default_value=255)   # track_unknown_space: true


class Config1(GrowingConfig):
    name = "config1: single depth camera 640x480, 0.05 m, linear decay 15 s, store grows from empty"

    def __init__(self, n_cycles=20, width=IMG_W, height=IMG_H, seed=1001, **kw):
        super().__init__(n_cycles, **kw)
        scene = synth.retail_scene(size=30.0, seed=3)
        for c in range(n_cycles):
            yaw = 0.05 + 0.04 * c
            o = (-6.0 + 0.15 * c, 2.3 - 13.5 + 12.0, 0.6)
            q = synth.optical_quat(yaw)
            self.cycles.append([Sensor(0, o, q, synth.depth_cloud(scene, o, q, width=width, height=height, seed=seed * 100 + c), synth.RGBD, "passthrough")])


class Config3(GrowingConfig):
    name = "config3: VLP-16 hourglass lidar (16 x 1800), 360 deg hFOV, exponential decay, VOXEL pre-filter, 0.05 m"
    decay_model = 1

    def __init__(self, n_cycles=20, rings=16, azimuths=1800, seed=1003, **kw):
        super().__init__(n_cycles, **kw)
        scene = synth.retail_scene(size=30.0, seed=3)
        for c in range(n_cycles):
            o = (-6.0 + 0.15 * c, 2.3 - 13.5 + 12.0, 0.9)
            q = synth.yaw_quat(0.03 * c)
            # (the reference example marks with obstacle_range 3.0 and a 0.3 .. 2.0 m window, vlp16_config.yaml:26-28)
            prm = dict(synth.VLP16, obstacle_range=3.0, min_obstacle_height=0.3, max_obstacle_height=2.0)
            self.cycles.append([Sensor(1, o, q, synth.lidar_cloud(scene, o, q, rings=rings, azimuths=azimuths, seed=seed * 100 + c), prm, "voxel")])


class Config4(GrowingConfig):
    name = "config4: Ouster-128 (128 x 1024) + 4 RGBD 640x480, VOXEL pre-filter on, 0.02 m, linear decay 15 s"
    voxel_size = 0.02
    voxel_min_points = 1

    def __init__(self, n_cycles=12, rings=128, azimuths=1024, width=IMG_W, height=IMG_H, seed=1004, **kw):
        super().__init__(n_cycles, **kw)
        scene = synth.retail_scene(size=30.0, seed=3)
        for c in range(n_cycles):
            centre = (-6.0 + 0.15 * c, 2.3 - 13.5 + 12.0, 0.0)
            lo, lq = (centre[0], centre[1], 1.0), synth.yaw_quat(0.03 * c)
            lidar = dict(vfov=0.785, vpad=0.03, hfov=6.29, min_z=0.5, max_z=8.0, accel=5.0, obstacle_range=6.0,
                         min_obstacle_height=0.1, max_obstacle_height=2.2)
            sensors = [Sensor(1, lo, lq, synth.lidar_cloud(scene, lo, lq, rings=rings, azimuths=azimuths, vfov=0.785, seed=seed * 100 + c), lidar, "voxel")]
            for k, (o, q) in enumerate(synth.camera_ring(centre, 0.1 + 0.03 * c, n_cams=4)):
                sensors.append(Sensor(0, o, q, synth.depth_cloud(scene, o, q, width=width, height=height, seed=seed * 100 + 10 * c + k), synth.RGBD, "voxel"))
            self.cycles.append(sensors)

    def costmap_params(self):
        return dict(origin_x=-16.0, origin_y=-16.0, resolution=0.05, size_x=640, size_y=640, mark_threshold=0, default_value=255)


class Config5:
    """BASELINE config 5: warehouse map, ~50 M active voxels at 0.02 m (thin vertical rack faces: no solid 8^3 blocks),
    7 RGBD cameras on the robot (small mark load), linear decay.  The map is built procedurally as disjoint sheets, so the
    voxel arrays are unique by construction.  `x_range` restricts generation to a spatial slab (multi-GPU shards)."""
    name = "config5: warehouse ~50M voxels, 0.02 m, 7xRGBD 160x120 mark load, linear decay 15 s"
    voxel_size = 0.02
    decay_model = 0
    voxel_decay = 15.0
    n_cams = 7

    def __init__(self, target=50_000_000, now0=1.7e9, dt=0.01, seed=1005, ring=4, width=160, height=120,
                 size_x=200.0, size_y=120.0):
        self.p = synth.RGBD
        self.now0, self.dt = now0, dt
        self.size_x, self.size_y = size_x, size_y
        self.vs = float(np.float32(self.voxel_size))
        self.seed = seed
        self.target = target
        zlo, zhi = int(math.ceil(0.3 / self.vs)), int(math.floor(2.0 / self.vs))
        self.z_range = (zlo, zhi)
        nz = zhi - zlo + 1
        # rack rows along x, every 3 m in y, two faces per row (0.6 m apart)
        n_faces = max(2, int(math.ceil(target / ((size_x - 4.0) / self.vs * nz))))
        self.faces = []
        y = -size_y / 2 + 2.0
        while len(self.faces) < n_faces and y < size_y / 2 - 2.0:
            for dy in (0.0, 0.6):
                if len(self.faces) < n_faces:
                    self.faces.append(int(math.floor((y + dy) / self.vs)))
            y += 3.0
        self.x_lo, self.x_hi = int(math.floor((-size_x / 2 + 2.0) / self.vs)), int(math.floor((size_x / 2 - 2.0) / self.vs))
        # robot in the aisle next to a central rack row; clouds from a local box scene
        yc = self.faces[len(self.faces) // 2] * self.vs
        boxes = [((-20.0, yc - 0.6, 0.0), (20.0, yc, 2.2)), ((-20.0, yc + 2.4, 0.0), (20.0, yc + 3.0, 2.2))]
        self.scene = synth.Scene(boxes)
        self.cycles = []
        for c in range(ring):
            centre = (-3.0 + 1.5 * c, yc + 1.2, 0.0)
            poses = synth.camera_ring(centre, 0.05 + 0.11 * c, n_cams=self.n_cams)
            clouds = [synth.depth_cloud(self.scene, o, q, width=width, height=height, seed=seed + 16 * c + k)
                      for k, (o, q) in enumerate(poses)]
            self.cycles.append(Cycle(0.0, poses, clouds))

    def voxels(self, x_range=None):
        """(ix, iy, iz, t) of the seeded map, optionally restricted to voxel x indices in [x_range[0], x_range[1])."""
        xlo, xhi = self.x_lo, self.x_hi + 1
        if x_range is not None:
            xlo, xhi = max(xlo, x_range[0]), min(xhi, x_range[1])
        if xhi <= xlo:
            z = np.zeros(0, np.int32)
            return z, z, z, np.zeros(0, np.float64)
        xs = np.arange(xlo, xhi, dtype=np.int32)
        zs = np.arange(self.z_range[0], self.z_range[1] + 1, dtype=np.int32)
        nx, nz, nf = len(xs), len(zs), len(self.faces)
        ix = np.broadcast_to(xs[None, :, None], (nf, nx, nz)).reshape(-1)
        iz = np.broadcast_to(zs[None, None, :], (nf, nx, nz)).reshape(-1)
        iy = np.broadcast_to(np.asarray(self.faces, np.int32)[:, None, None], (nf, nx, nz)).reshape(-1)
        rng = np.random.default_rng(self.seed)
        t = self.now0 - rng.uniform(0.0, self.voxel_decay, ix.shape[0])
        return np.ascontiguousarray(ix), np.ascontiguousarray(iy), np.ascontiguousarray(iz), t

    @property
    def points_per_cycle(self):
        return sum(len(c) for c in self.cycles[0].clouds)

    def cycle(self, step):
        c = self.cycles[step % len(self.cycles)]
        return self.now0 + self.dt * (step + 1), c.poses, c.clouds

    def costmap_params(self):
        res = 0.02
        return dict(origin_x=-self.size_x / 2, origin_y=-self.size_y / 2, resolution=res,
                    size_x=int(round(self.size_x / res)), size_y=int(round(self.size_y / res)), mark_threshold=0, default_value=0)
