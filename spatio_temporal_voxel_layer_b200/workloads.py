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
