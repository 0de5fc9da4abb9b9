"""Multi-process host logic on CPU: sequence sharding and the trajectory gather over
torch.distributed (gloo, world_size 2) -- the same helper bench.py uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_sequences(pkg):
    from slam_rgbd_b200.dist import owner_of, shard_sequences

    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = shard_sequences(64, world, r)
            assert len(mine) == 64 // world  # 64 sequences -> 8 per GPU at 8 GPUs (configs[3])
            assert all(owner_of(s, world) == r for s in mine)
            seen += mine
        assert sorted(seen) == list(range(64))
    assert shard_sequences(3, 2, 0) == [0, 2] and shard_sequences(3, 2, 1) == [1]
    with pytest.raises(ValueError):
        shard_sequences(4, 2, 2)


def fake_traj(seq, frames):
    rng = np.random.default_rng(1000 + seq)
    return rng.normal(size=(frames, 12)).astype(np.float32)


def _worker(rank, world, port, n_seq, frames, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import youth_pkg

    youth_pkg.load()
    from slam_rgbd_b200.dist import gather_trajectories, shard_sequences

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_sequences(n_seq, world, rank)
    local = np.stack([fake_traj(s, frames) for s in mine]) if mine else np.zeros((0, frames, 12), np.float32)
    out = gather_trajectories(local, n_seq, world, rank)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_seq", [4, 3])
def test_gather_trajectories_world2(pkg, n_seq):
    frames = 7
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_seq, frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([fake_traj(s, frames) for s in range(n_seq)])
    for rank, out in results:
        assert out.shape == (n_seq, frames, 12)
        assert np.array_equal(out, want), f"rank {rank}"  # bitwise: a gather must not change a trajectory


def test_gather_single_process(pkg):
    from slam_rgbd_b200.dist import gather_trajectories

    local = np.stack([fake_traj(s, 5) for s in range(3)])
    assert np.array_equal(gather_trajectories(local, 3, 1, 0), local)
