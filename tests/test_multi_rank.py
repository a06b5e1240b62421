"""CPU test of the N>1 host logic with world_size 2 over gloo: contiguous sharding with halo and the
single all-gather of pose records, reassembled in frame order on every rank."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from plslam_b200 import sharding


def test_shard_frames_partitions_the_sequence():
    for n in (1, 2, 7, 64, 513):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_frames(n, world, r) for r in range(world)]
            assert sum(c for _, c, _ in blocks) == n
            pos = 0
            for first, count, halo in blocks:
                assert first == pos and halo == (1 if first > 0 and count > 0 else 0)
                pos += count


def _worker(rank, world, n_frames, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count, halo = sharding.shard_frames(n_frames, world, rank)
    rng = np.random.default_rng(100)                        # same stream on every rank: pose of frame k is a function of k
    all_poses = rng.normal(size=(n_frames, 16)).astype(np.float32)
    rec = sharding.make_records(first, all_poses[first:first + count], np.arange(first, first + count) % 7,
                                1000 + np.arange(first, first + count), 200 + np.zeros(count))
    out = sharding.gather_records(rec, n_frames, world, dist)
    ok = (out.shape == (n_frames, sharding.RECORD_FLOATS) and np.array_equal(out[:, 16], np.arange(n_frames)) and
          np.array_equal(out[:, :16], all_poses) and np.array_equal(out[:, 17], np.arange(n_frames) % 7))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_of_pose_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 13, 2, 29611            # uneven blocks (7 + 6) exercise the padding
    procs = [ctx.Process(target=_worker, args=(r, world, n_frames, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(120)
    assert res == [(0, True), (1, True)]
