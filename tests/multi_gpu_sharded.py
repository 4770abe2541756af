"""Run under torchrun on >= 2 GPUs (tests/test_multi_gpu.py launches it): the in-library NCCL merge of spatially sharded
grids must reproduce the costmap bytes of one unsharded grid — u32 and u8 count windows, blocking and pipelined."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spatio_temporal_voxel_layer_b200 as stvl  # noqa: E402


FAILED = []


def check(label, cond):
    """Record a failed comparison (reported by rank 0 at the end) and return whether it held."""
    if not cond:
        FAILED.append(label)
    return bool(cond)


def voxel_totals(g, ref, rank, label):
    """Active voxels summed over the shards against the unsharded grid's (a failed check names the stage)."""
    n = torch.tensor([g.active_count()], device="cuda")
    dist.all_reduce(n)
    if rank == 0:
        return check("voxel total after %s: shards %d, unsharded %d" % (label, int(n.item()), ref.active_count()), int(n.item()) == ref.active_count())
    return True


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rng = np.random.default_rng(77)
    vox = np.unique(rng.integers(-300, 300, (250000, 3)), axis=0)
    vox[:, 2] = vox[:, 2] % 20
    vox = np.unique(vox, axis=0)
    size, res, ox, oy = 640, 0.05, -16.0, -16.0
    params = stvl.CostmapParams(ox, oy, res, size, size, 2, 0, 254)

    g = stvl.SpatioTemporalVoxelGrid(0.05, voxel_decay=15.0, shard_index=rank, shard_count=world, shard_width=6)
    uid = [stvl.SpatioTemporalVoxelGrid.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(uid[0], rank, world)
    ref = stvl.SpatioTemporalVoxelGrid(0.05, voxel_decay=15.0) if rank == 0 else None
    cm = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    cm_ref = torch.zeros((size, size), dtype=torch.uint8, device="cuda")
    now = 100.0
    ok = True
    expected = []
    n_cycles = 8
    # cycles 0-2: blocking merge with u32 / u8 count windows and bytes mode; 3-4: pipelined u8 counts; 5-7: pipelined bytes mode
    for cyc in range(n_cycles):
        t = now - rng.uniform(0, 20, len(vox))            # same stream of random numbers on every rank
        g.ResetGrid(); g.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
        g.ClearFrustums([], now)
        # cycle 2 (bytes mode, blocking) tracks unknown space: default value 255 where nothing is marked
        pc = stvl.CostmapParams(ox, oy, res, size, size, 2, 255, 254) if cyc == 2 else params
        if rank == 0:
            ref.ResetGrid(); ref.load_voxels(vox[:, 0], vox[:, 1], vox[:, 2], t)
            ref.ClearFrustums([], now)
            ref.costmap_device(pc, cm_ref.data_ptr())
            ref.synchronize()
            expected.append(cm_ref.clone())
        if cyc < 3:
            g.costmap_sharded_device(pc, cm.data_ptr(), root=0, count_bytes=(4, 1, 0)[cyc], pipelined=False)
            g.synchronize()
            if rank == 0:
                ok = check("window merge, blocking, cycle %d" % cyc, torch.equal(cm, expected[cyc]) and int((cm == 254).sum()) > 1000) and ok
        else:
            if cyc == 5:   # mode switch: drain the count-window pipeline first
                g.costmap_sharded_flush(params, cm.data_ptr(), root=0)
                g.synchronize()
                if rank == 0:
                    ok = check("window pipeline drain", torch.equal(cm, expected[4])) and ok
            g.costmap_sharded_device(params, cm.data_ptr(), root=0, count_bytes=1 if cyc < 5 else 0, pipelined=True)
            g.synchronize()
            if rank == 0 and cyc in (4, 6, 7):   # the bytes of cycle k arrive with call k+1
                ok = check("window / bytes merge, pipelined, cycle %d" % cyc, torch.equal(cm, expected[cyc - 1])) and ok
        now += 1.0
    g.costmap_sharded_flush(params, cm.data_ptr(), root=0)
    g.synchronize()
    if rank == 0:
        ok = check("final flush of the merge", torch.equal(cm, expected[n_cycles - 1])) and ok
    ok = voxel_totals(g, ref, rank, "the window merges") and ok
    # fused sharded cycles (stvl_update_cycle_sharded_device): sweep + Mark of a device-resident cloud + costmap pass into
    # the reduce window + NCCL merge, blocking then pipelined, against the unsharded grid's fused cycle
    cloud = np.zeros((200000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-6.0, 6.0, (len(cloud), 3)).astype(np.float32)
    cloud[:, 2] = np.abs(cloud[:, 2]) * 0.15
    cloud_dev = torch.from_numpy(cloud).cuda()
    fused_expected = []
    for cyc in range(4):
        now += 1.0
        reading = stvl.MeasurementReading((cloud_dev.data_ptr(), len(cloud), 16, (0, 4, 8)), origin=(0.5 * cyc, 0.0, 0.5), obstacle_range=7.0,
                                          now=now + 0.001)
        if rank == 0:
            ref.update_cycle_device(now, [], [reading], ox, oy, res, size, size, cm_ref.data_ptr(), mark_threshold=2, default_value=0)
            ref.synchronize()
            fused_expected.append(cm_ref.clone())
        g.update_cycle_sharded_device(now, [], [reading], params, cm.data_ptr(), root=0, pipelined=cyc >= 2)
        g.synchronize()
        if rank == 0:
            if cyc < 2:
                ok = check("fused sharded cycle, blocking, %d" % cyc, torch.equal(cm, fused_expected[cyc]) and int((cm == 254).sum()) > 1000) and ok
            elif cyc == 3:
                ok = check("fused sharded cycle, pipelined", torch.equal(cm, fused_expected[2])) and ok
    g.costmap_sharded_flush(params, cm.data_ptr(), root=0)
    g.synchronize()
    if rank == 0:
        ok = check("fused sharded cycle, flush", torch.equal(cm, fused_expected[3])) and ok
    ok = voxel_totals(g, ref, rank, "the fused sharded cycles") and ok
    # the pipelined host-buffer cycle on sharded handles (stvl_update_cycle_async): every rank uploads the same host cloud, keeps
    # its own slabs, the cells are merged into rank 0 inside the cycle and rank 0 reads the merged bytes back
    host_cloud = torch.from_numpy(cloud).pin_memory()
    host_out = [np.zeros((size, size), np.uint8) for _ in range(2)]
    async_expected, tickets = [], []
    for cyc in range(3):
        now += 1.0
        origin = (0.3 * cyc - 0.4, 0.2, 0.5)
        hr = stvl.MeasurementReading((host_cloud.data_ptr(), len(cloud), 16, (0, 4, 8)), origin=origin, obstacle_range=7.0, now=now + 0.001)
        if rank == 0:
            dr = stvl.MeasurementReading((cloud_dev.data_ptr(), len(cloud), 16, (0, 4, 8)), origin=origin, obstacle_range=7.0, now=now + 0.001)
            ref.update_cycle_device(now, [], [dr], ox, oy, res, size, size, cm_ref.data_ptr(), mark_threshold=2, default_value=0)
            ref.synchronize()
            async_expected.append(cm_ref.cpu().numpy().copy())
        tickets.append(g.update_cycle_async(now, [], [hr], params, host_out[cyc % 2]))
        if cyc >= 1:
            g.update_cycle_wait(tickets[cyc - 1])
            if rank == 0:
                ok = check("async host cycle %d" % (cyc - 1), np.array_equal(host_out[(cyc - 1) % 2], async_expected[cyc - 1])) and ok
    g.update_cycle_wait(tickets[-1])
    if rank == 0:
        ok = check("async host cycle, last", np.array_equal(host_out[0], async_expected[2]) and int((async_expected[2] == 254).sum()) > 1000) and ok
    ok = voxel_totals(g, ref, rank, "the pipelined host-buffer cycles") and ok   # the shards together hold exactly the unsharded grid's voxels
    n_local = torch.tensor([g.active_count()], device="cuda")
    dist.all_reduce(n_local)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    if rank == 0:
        print("MULTI_GPU_OK" if bool(flag.item()) else "MULTI_GPU_MISMATCH", "ranks", world, "voxels", int(n_local.item()),
              ("failed checks: " + ", ".join(FAILED)) if FAILED else "")
    g.comm_destroy()
    g.close()
    if ref is not None:
        ref.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
