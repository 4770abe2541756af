"""Shared comparison helpers for the parity tests (GPU path vs CPU oracle)."""
import numpy as np


def voxel_dict(ix, iy, iz, t):
    order = np.lexsort((iz, iy, ix))
    return np.stack([ix[order], iy[order], iz[order]], 1).astype(np.int64), np.asarray(t)[order]


def assert_same_voxels(gpu_vox, ora_vox, exact_times=True, rtol=1e-6, allowed_missing=None):
    gk, gt = voxel_dict(*gpu_vox)
    ok, ot = voxel_dict(*ora_vox)
    assert gk.shape == ok.shape, "active voxel count differs: gpu %d vs oracle %d" % (len(gk), len(ok))
    assert np.array_equal(gk, ok), "voxel key sets differ"
    if exact_times:
        assert np.array_equal(gt.view(np.uint64), ot.view(np.uint64)), "mark times differ in bits (max abs diff %g)" % np.abs(gt - ot).max()
    else:
        np.testing.assert_allclose(gt, ot, rtol=rtol, atol=0)


def column_dict(x, y, c):
    order = np.lexsort((y, x))
    return np.stack([np.asarray(x)[order], np.asarray(y)[order]], 1), np.asarray(c)[order]


def assert_same_columns(gpu_cols, ora_cols):
    gk, gc = column_dict(*gpu_cols)
    ok, oc = column_dict(*ora_cols)
    assert gk.shape == ok.shape, "occupied column count differs: gpu %d vs oracle %d" % (len(gk), len(ok))
    assert np.array_equal(gk.view(np.uint64), ok.view(np.uint64)), "column keys (world x,y doubles) differ"
    assert np.array_equal(gc, oc), "column counts differ"


def point_set(xyz):
    a = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    return np.unique(a.view(np.uint32).reshape(-1, 3), axis=0)
