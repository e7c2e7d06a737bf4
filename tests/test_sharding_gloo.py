"""world_size-2 gloo test of the multi-GPU host logic (CPU): environment sharding + the episode-end all_gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_envs, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from av_aloha_b200 import sharding
    lo, hi = sharding.shard_range(num_envs, rank, world)
    env = torch.arange(lo, hi)
    # deterministic per-env "results" so every rank can check the gathered vector
    succ, mx, sm = (env % 3 == 0), (env % 5).to(torch.int32), env.to(torch.float32) * 0.5
    g = sharding.gather_episode_stats(succ, mx, sm, num_envs)
    full = torch.arange(num_envs)
    ok = (torch.equal(g[0], (full % 3 == 0).to(torch.uint8)) and torch.equal(g[1], (full % 5).to(torch.int32))
          and torch.allclose(g[2], full.float() * 0.5))
    q.put((rank, lo, hi, bool(ok), sharding.aggregate(*g)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_envs", [8, 11])
def test_gather_episode_stats_world2(num_envs):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == num_envs          # contiguous cover
    assert all(r[3] for r in res)
    assert res[0][4] == res[1][4]                                                        # same aggregate on every rank


def test_shard_range_covers_everything():
    from av_aloha_b200 import sharding
    for n in (1, 7, 4096, 51200):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)
