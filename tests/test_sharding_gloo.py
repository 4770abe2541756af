"""World-size-2 (and 3) gloo test of the multi-GPU merge protocol on CPU: every rank sweeps only the leaves it owns
(oracle restricted to its slabs), scatters its column counts into the dense window, the windows are reduced to rank 0
and thresholded — the result must equal the unsharded oracle's costmap bytes.  The CUDA side of the same protocol
(stvl_columns_to_dense / stvl_costmap_from_dense and the device ownership rule) is covered by
test_gpu_parity.py::test_sharded_grids_union_equals_single."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scene():
    rng = np.random.default_rng(21)
    vox = np.unique(rng.integers(-200, 200, (60000, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 24
    vox = np.unique(vox, axis=0)
    now = 50.0
    t = now - rng.uniform(0, 20, len(vox))
    return vox, t, now


CM = dict(origin_x=-9.0, origin_y=-9.0, resolution=0.05, size_x=360, size_y=360, mark_threshold=2, default_value=0)


def _threshold(dense, ix0, iy0, vs, cm):
    """numpy restatement of UpdateROSCostmap's loop on a dense count window (test-side reference)."""
    out = np.full((cm["size_y"], cm["size_x"]), cm["default_value"], np.uint8)
    iy, ix = np.nonzero(dense)
    cnt = dense[iy, ix]
    keep = cnt.astype(np.int64) >= cm["mark_threshold"]
    wx = (ix[keep] + ix0).astype(np.float64) * vs + vs / 2.0
    wy = (iy[keep] + iy0).astype(np.float64) * vs + vs / 2.0
    ok = (wx >= cm["origin_x"]) & (wy >= cm["origin_y"])
    mx = ((wx[ok] - cm["origin_x"]) / cm["resolution"]).astype(np.uint64)
    my = ((wy[ok] - cm["origin_y"]) / cm["resolution"]).astype(np.uint64)
    inb = (mx < cm["size_x"]) & (my < cm["size_y"])
    out[my[inb], mx[inb]] = 254
    return out


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as O
    from spatio_temporal_voxel_layer_b200 import sharding
    vox, t, now = _scene()
    width = 5
    mine = sharding.shard_of_voxel(vox[:, 0].astype(np.int64), width, world) == rank
    og = O.OracleGrid(0.05)
    og.load_voxels(vox[mine, 0], vox[mine, 1], vox[mine, 2], t[mine])
    og.clear_frustums(now)
    vs = og.voxel_size
    ix0, iy0, nx, ny = sharding.dense_window(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"], 0.05)
    dense = torch.zeros((ny, nx), dtype=torch.int32)
    x, y, c = og.costmap()
    ix = np.rint((x - vs / 2.0) / vs).astype(np.int64) - ix0
    iy = np.rint((y - vs / 2.0) / vs).astype(np.int64) - iy0
    inw = (ix >= 0) & (iy >= 0) & (ix < nx) & (iy < ny)
    dense.numpy()[iy[inw], ix[inw]] += c[inw].astype(np.int32)
    n_local = int(mine.sum())
    tot = torch.tensor([n_local], dtype=torch.int64)
    dist.all_reduce(tot)
    sharding.merge_counts(dist, dense, dst=0)
    if rank == 0:
        ret["costmap"] = _threshold(dense.numpy(), ix0, iy0, vs, CM)
        ret["total"] = int(tot[0])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_equals_unsharded(world, oracle):
    vox, t, now = _scene()
    og = oracle.OracleGrid(0.05)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.clear_frustums(now)
    ref, _, _ = og.update_ros_costmap(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"],
                                      CM["mark_threshold"], CM["default_value"])
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret["total"] == len(vox)           # the shards partition the voxels
    assert np.array_equal(ret["costmap"], ref)
    assert ref.sum() > 0


def test_ownership_rule_matches_device_convention():
    from spatio_temporal_voxel_layer_b200 import sharding
    # floor semantics for negative leaf indices, slabs of `width` leaves dealt round-robin
    assert [sharding.shard_of_leaf(b, 4, 3) for b in (-9, -8, -5, -4, -1, 0, 3, 4, 11, 12)] == [0, 1, 1, 2, 2, 0, 0, 1, 2, 0]
    assert sharding.shard_of_leaf(123, 7, 1) == 0
    ix = np.array([-72, -64, -33, -1, 0, 31, 32, 95, 96])
    assert sharding.shard_of_voxel(ix, 4, 3).tolist() == [sharding.shard_of_leaf(int(v) >> 3, 4, 3) for v in ix]


# ---- the bitmap merge of the sharded fused cycle (stvl_update_cycle_sharded_device) on CPU -----------------------------------
def _bitmap_worker(rank, world, port, ret, width):
    """Rank r: sweep its own slabs on the oracle, threshold its own columns into costmap cells, pack them into the bitmap of its
    own word interval (stvl_shard_word_intervals — the library's host arithmetic), ship it to rank 0 (gloo send / recv standing
    in for the grouped ncclSend / ncclRecv); rank 0 ORs the bitmaps into bytes like expand_bitmaps_kernel."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as O
    import spatio_temporal_voxel_layer_b200 as stvl
    from spatio_temporal_voxel_layer_b200 import sharding
    vox, t, now = _scene()
    mine = sharding.shard_of_voxel(vox[:, 0].astype(np.int64), width, world) == rank
    og = O.OracleGrid(0.05)
    og.load_voxels(vox[mine, 0], vox[mine, 1], vox[mine, 2], t[mine])
    og.clear_frustums(now)
    own, _, _ = og.update_ros_costmap(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"], CM["mark_threshold"], 0)
    params = stvl.CostmapParams(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"], CM["mark_threshold"], 0, 254)
    lo, hi = stvl.shard_word_intervals(0.05, world, width, world, params)
    my, mx = np.nonzero(own == 254)
    inside = bool(len(mx) == 0 or ((mx >> 5).min() >= lo[rank] and (mx >> 5).max() < hi[rank]))   # every cell of the rank lies in its interval
    wpr = int(hi[rank] - lo[rank])
    bits = np.zeros((CM["size_y"], max(wpr, 1)), np.uint32)
    np.bitwise_or.at(bits, (my, (mx >> 5) - lo[rank]), (np.uint32(1) << (mx & 31).astype(np.uint32)))
    flags = torch.tensor([1 if inside else 0], dtype=torch.int64)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank != 0:
        dist.send(torch.from_numpy(bits.view(np.int32).reshape(-1).copy()), dst=0)
    else:
        out = np.full((CM["size_y"], CM["size_x"]), CM["default_value"], np.uint8)
        allbits = {0: bits}
        for r in range(1, world):
            buf = torch.zeros(CM["size_y"] * max(int(hi[r] - lo[r]), 1), dtype=torch.int32)
            dist.recv(buf, src=r)
            allbits[r] = buf.numpy().view(np.uint32).reshape(CM["size_y"], -1)
        xs = np.arange(CM["size_x"])
        for r, bm in allbits.items():
            w = xs >> 5
            sel = (w >= lo[r]) & (w < hi[r])
            cols = xs[sel]
            lethal = ((bm[:, (cols >> 5) - lo[r]] >> (cols & 31).astype(np.uint32)) & 1).astype(bool)
            sub = out[:, cols]
            sub[lethal] = 254
            out[:, cols] = sub
        ret["costmap"] = out
        ret["inside"] = int(flags[0])
        ret["words"] = int((hi - lo).sum())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,width", [(2, 5), (3, 5), (2, 23)])
def test_bitmap_merge_equals_unsharded(world, width, oracle, stvl):
    vox, t, now = _scene()
    og = oracle.OracleGrid(0.05)
    og.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
    og.clear_frustums(now)
    ref, _, _ = og.update_ros_costmap(CM["origin_x"], CM["origin_y"], CM["resolution"], CM["size_x"], CM["size_y"],
                                      CM["mark_threshold"], CM["default_value"])
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 2000) + world * 7 + width
    mp.spawn(_bitmap_worker, args=(world, port, ret, width), nprocs=world, join=True)
    assert ret["inside"] == 1
    assert np.array_equal(ret["costmap"], ref)
    assert int((ref == 254).sum()) > 1000
    if width == 23:     # wide slabs: the ranks' intervals hardly overlap, the exchange moves about one bit per cell in total
        assert ret["words"] <= ((CM["size_x"] + 31) // 32) + 2 * world


def test_shard_word_intervals_cover_every_column(stvl):
    """Property of the interval arithmetic on its own: for random geometries every voxel column a rank owns maps (if it maps into
    the costmap at all) into that rank's word interval."""
    from spatio_temporal_voxel_layer_b200 import sharding
    rng = np.random.default_rng(5)
    for _ in range(200):
        vs_param = float(rng.choice([0.02, 0.05, 0.1]))
        vs = float(np.float32(vs_param))
        n = int(rng.integers(1, 9))
        width = int(rng.integers(1, 300))
        res = float(rng.choice([vs_param, 0.05, 0.1, 0.013]))
        sx, sy = int(rng.integers(1, 3000)), 7
        ox = float(rng.uniform(-100, 100))
        params = stvl.CostmapParams(ox, 0.0, res, sx, sy, 0, 0, 254)
        lo, hi = stvl.shard_word_intervals(vs_param, n, width, n, params)
        ix = rng.integers(int((ox - 5) / vs), int((ox + sx * res + 5) / vs) + 1, 4000)
        wx = ix.astype(np.float64) * vs + vs / 2.0
        ok = wx >= ox
        mx = ((wx[ok] - ox) / res).astype(np.int64)
        inb = mx < sx
        owner = sharding.shard_of_voxel(ix[ok][inb].astype(np.int64), width, n)
        w = mx[inb] >> 5
        assert np.all(w >= lo[owner]) and np.all(w < hi[owner])
