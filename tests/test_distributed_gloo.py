"""N>1 host logic on CPU: env shards and the episode-statistics reduction (the only
collective on the path) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from gym_auv_b200.sharding import reduce_stats, shard_range, summarize_stats  # noqa: E402


def test_shard_range_partitions_exactly():
    for n, w in ((1_000_000, 8), (65536, 2), (10, 4), (7, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = torch.zeros(16, dtype=torch.float64)
    # rank r finished (r+1)*10 episodes with reward r+1 each
    stats[0] = (rank + 1) * 10
    stats[1] = (rank + 1) * 10 * (rank + 1.0)
    stats[2] = (rank + 1) * 10 * (rank + 1.0) ** 2
    stats[4] = rank * 3
    stats[6] = (rank + 1) * 100
    stats[9] = 1000 * (rank + 1)
    red = reduce_stats(stats)
    lo, hi = shard_range(101, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    if rank == 0:
        torch.save(dict(red=red, spans=spans), out)
    dist.destroy_process_group()


def test_stats_allreduce_world2(tmp_path):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    red = res["red"].numpy()
    assert red[0] == 30 and red[1] == 10 * 1 + 20 * 2 and red[9] == 3000
    s = summarize_stats(red, t_step_size=0.5)
    assert s["episodes"] == 30 and s["reward"] == pytest.approx(50 / 30)
    assert s["collision"] == pytest.approx(3 / 30) and s["timesteps"] == pytest.approx(300 / 30)
    assert s["duration"] == pytest.approx(0.5 * 300 / 30)
    assert res["spans"] == [(0, 51), (51, 101)]
