"""CPU-only checks of the product library: it loads, exports every symbol include/stvl_b200.h declares, its
host-side frustum geometry is bit-identical to the oracle's restatement, and it refuses to run without a GPU
(no CPU fallback).  No kernels are launched here."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(stvl):
    L = stvl.load_library()
    header = open(os.path.join(ROOT, "include", "stvl_b200.h")).read()
    declared = set(re.findall(r"\b(stvl_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(stvl.EXPORTED_SYMBOLS), declared ^ set(stvl.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), "missing export " + name
    assert L.stvl_abi_version() == 2


def test_struct_layouts_match_header(stvl):
    assert C.sizeof(stvl.Config) == 56
    assert C.sizeof(stvl.Reading) == 152
    assert C.sizeof(stvl.SweepStats) == 88
    assert C.sizeof(stvl.CostmapParams) == 40
    assert C.sizeof(stvl.CostmapResult) == 48
    assert stvl.FRUSTUM_DTYPE.itemsize == 304


def test_depth_frustum_blocks_match_oracle_bitwise(stvl, oracle):
    rng = np.random.default_rng(11)
    for _ in range(200):
        v, h = rng.uniform(0.2, 1.5, 2)
        dmin = rng.uniform(0.0, 1.0)
        dmax = dmin + rng.uniform(0.5, 10.0)
        o = rng.uniform(-50, 50, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        a = float(rng.uniform(0, 10))
        mine = stvl.build_depth_frustum(v, h, dmin, dmax, o, q, a)
        ref = oracle.depth_frustum(v, h, dmin, dmax, o, q, a)
        assert mine["type"][0] == ref["type"][0] == 0
        assert mine["accel_factor"][0] == ref["accel_factor"][0]
        assert np.array_equal(mine["p"].view(np.uint64), ref["p"].view(np.uint64))
    bogus = stvl.build_depth_frustum(0.0, 0.0, 0.1, 7.0, [0, 0, 0], [0, 0, 0, 1], 0.0)
    assert bogus["type"][0] == -1


def test_lidar_frustum_blocks_match_oracle_bitwise(stvl, oracle):
    rng = np.random.default_rng(12)
    for _ in range(200):
        v = rng.uniform(0.1, 1.5)
        pad = rng.uniform(0, 0.2)
        h = rng.choice([6.29, 3.0, 1.5])
        dmin = rng.uniform(0.0, 1.5)
        dmax = dmin + rng.uniform(0.5, 30.0)
        o = rng.uniform(-50, 50, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        mine = stvl.build_lidar_frustum(v, pad, h, dmin, dmax, o, q, 5.0)
        ref = oracle.lidar_frustum(v, pad, h, dmin, dmax, o, q, 5.0)
        assert mine["type"][0] == ref["type"][0] == 1
        assert np.array_equal(mine["p"].view(np.uint64), ref["p"].view(np.uint64))


def test_measurement_reading_builds_frustums(stvl):
    r = stvl.MeasurementReading(origin=(1, 2, 0.5), orientation=(0, 0, 0, 1), vfov=0.8745, hfov=1.048, min_z=0.1,
                                max_z=7.0, decay_acceleration=5.0, model_type=stvl.MODEL_DEPTH_CAMERA)
    assert r.frustum()["type"][0] == 0 and r.frustum()["accel_factor"][0] == 5.0
    r.model_type = stvl.MODEL_3D_LIDAR
    assert r.frustum()["type"][0] == 1
    r.model_type = 7
    assert r.frustum() is None   # unknown model types are skipped (reference spatio_temporal_voxel_grid.cpp:134-137)


def test_no_cpu_fallback(stvl):
    """Without a B200 the library must refuse to create a grid instead of computing on the CPU."""
    if stvl.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(stvl.StvlError) as e:
        stvl.SpatioTemporalVoxelGrid(0.05)
    assert e.value.code == 3   # STVL_ERR_NO_DEVICE


def test_product_does_not_reference_oracle():
    """The product tree must never import / link / execute anything under oracle/."""
    pkg = os.path.join(ROOT, "spatio_temporal_voxel_layer_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile", ".txt")):
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle_py|libstvl_oracle|stvl_oracle|from oracle|import oracle", text):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders


def test_build_transform_matches_oracle_and_known_answers(stvl, oracle):
    """stvl_build_transform (host only) against the oracle's restatement of tf2_sensor_msgs' float affine, plus
    hand-checked answers: identity rotation is a pure float translation, a +90 deg yaw maps x -> y."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.uniform(-50, 50, 3)
        assert np.array_equal(stvl.build_transform(t, q).view(np.uint32), oracle.transform_from_quat(t, q).view(np.uint32))
    m = stvl.build_transform((1.5, -2.25, 0.125), (0, 0, 0, 1))
    assert np.array_equal(m.reshape(3, 4), np.array([[1, 0, 0, 1.5], [0, 1, 0, -2.25], [0, 0, 1, 0.125]], np.float32))
    pts = np.array([[0.1, 0.2, 0.3, 0.0], [np.nan, 1.0, 1.0, 0.0]], np.float32)
    out = oracle.transform_cloud(pts, m)
    assert out[0, 0] == np.float32(0.1) + np.float32(1.5) and out[0, 1] == np.float32(0.2) + np.float32(-2.25)
    assert np.isnan(out[1, :3]).all()     # 0 * NaN = NaN: a NaN coordinate poisons the whole point, as in Eigen's product
    s = math.sqrt(0.5)
    m = stvl.build_transform((0, 0, 0), (0, 0, s, s)).reshape(3, 4)
    out = oracle.transform_cloud(np.array([[2.0, 0.0, 1.0, 0.0]], np.float32), m)
    assert abs(out[0, 0]) < 1e-6 and abs(out[0, 1] - 2.0) < 1e-6 and out[0, 2] == 1.0
