"""The N>1 path on CPU: world_size-2 gloo processes each own a shard of the instance space
(pogema_b200.sharding); the union of their results equals the unsharded run and the counter
aggregation sums correctly.  The stepping itself is done by the C oracle here (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pogema_b200.sharding import aggregate_counters, shard_range, shard_seeds
from tests.helpers import make_actions
from tests.oracle_c import COracle

GC = dict(size=10, density=0.2, num_agents=12, obs_radius=3, max_episode_steps=16, collision_system="priority",
          on_target="finish")
N, T, BASE = 7, 20, 50


def run_shard(seeds, actions):
    co = COracle.from_python_oracle(GC, seeds)
    out = co.run(actions, auto_reset=True)
    return out["rewards_sum"], co.pos.copy(), out["obs"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard_range(N, rank, world)
    seeds = shard_seeds(BASE, N, rank, world)
    actions = make_actions(T, N, GC["num_agents"], seed=9)[:, first:last]
    rsum, pos, obs = run_shard([int(s) for s in seeds], actions)
    agg = aggregate_counters({"agent_steps": (last - first) * T * GC["num_agents"], "reward": float(rsum.sum())})
    gathered = [None] * world
    dist.all_gather_object(gathered, (first, last, rsum, pos, obs))
    if rank == 0:
        ret["agg"] = agg
        ret["parts"] = gathered
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert shard_seeds(10, 7, 1, 2).tolist() == [14, 15, 16]


def test_two_rank_gloo_union_equals_unsharded():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    actions = make_actions(T, N, GC["num_agents"], seed=9)
    full_rsum, full_pos, full_obs = run_shard(list(range(BASE, BASE + N)), actions)
    for first, last, rsum, pos, obs in ret["parts"]:
        assert np.array_equal(rsum, full_rsum[first:last])
        assert np.array_equal(pos, full_pos[first:last])
        assert np.array_equal(obs, full_obs[first:last])
    assert ret["agg"]["agent_steps"] == N * T * GC["num_agents"]
    assert ret["agg"]["reward"] == float(full_rsum.sum())
