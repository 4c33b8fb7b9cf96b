"""Host-side multi-GPU logic on CPU: world_size-2 gloo group (scene-sharded, no data-path collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from styl3r_b200.sharding import gather_images, shard_range, shard_scene_views


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            parts = [list(shard_range(n, w, r)) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_scene_views_splits_views_when_scenes_are_scarce():
    # cfg4: 32 scenes x 6 views over 8 ranks -> 4 whole scenes per rank
    for r in range(8):
        mine = shard_scene_views(32, 6, 8, r)
        assert [s for s, _ in mine] == list(range(4 * r, 4 * r + 4)) and all(v == list(range(6)) for _, v in mine)
    # pose-align: 1 scene x 6 views over 4 ranks -> views split
    got = [shard_scene_views(1, 6, 4, r) for r in range(4)]
    views = sum([v for part in got for _, v in part], [])
    assert sorted(views) == list(range(6))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_scenes, V = 5, 2
        mine = shard_scene_views(n_scenes, V, world, rank)
        # "render": image k of scene s is filled with s + k/10
        imgs = torch.stack([torch.full((3, 4, 4), s + k / 10.0) for s, vs in mine for k in vs]) if mine else torch.zeros(0, 3, 4, 4)
        counts = [sum(len(v) for _, v in shard_scene_views(n_scenes, V, world, r)) for r in range(world)]
        out = gather_images(imgs, counts)
        # timing protocol used by bench.py: max over ranks
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put((out[:, 0, 0, 0].tolist(), float(t)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_and_max_reduce():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    vals, tmax = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    expect = [s + k / 10.0 for s in range(5) for k in range(2)]
    assert vals == pytest.approx(expect) and tmax == 2.0
