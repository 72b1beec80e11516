"""Multi-process (gloo, world_size 2, CPU) test of the host-side sharding + score aggregation used at N > 1."""
import os
import socket
from statistics import mean, stdev

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cage_challenge_4_b200.evaluation import aggregate_scores, shard
    lo, hi = shard(n_total, rank, world)
    local = torch.tensor([float(-(i * 37 % 101)) for i in range(lo, hi)], dtype=torch.float64)
    out = aggregate_scores(local, n_total, rank, world)
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 100])
def test_score_aggregation_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vals = [float(-(i * 37 % 101)) for i in range(n_total)]
    assert out["rewards"] == vals
    assert out["mean"] == mean(vals) and out["stdev"] == stdev(vals) and out["episodes"] == n_total


def test_shards_partition_the_episode_range():
    from cage_challenge_4_b200.evaluation import shard
    for n in (1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
