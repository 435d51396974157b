"""The register-resident step kernel (pgm_fast.cuh) against the generic one (PGM_FAST=0) and the C oracle:
every (team, agents per thread) pair the planner can choose, all three collision systems, radii 2..7, multi-step
launches, partial last batches, bit-packed observations, the packed host transport."""
import itertools

import numpy as np
import pytest

from tests.helpers import make_actions
from tests.oracle_c import COracle

pytestmark = pytest.mark.gpu


def build(gc, n, seeds, monkeypatch, fast, team=0, fast_team=None, fmt="u8", auto_reset=True):
    from pogema_b200 import BatchedPogema, GridConfig
    monkeypatch.setenv("PGM_FAST", "1" if fast else "0")
    if fast_team:
        monkeypatch.setenv("PGM_FAST_TEAM", str(fast_team))
    else:
        monkeypatch.delenv("PGM_FAST_TEAM", raising=False)
    env = BatchedPogema(GridConfig(**gc), num_envs=n, seeds=seeds, auto_reset=auto_reset, team_threads=team, obs_format=fmt)
    assert env.engine.plan()["fast_step_kernel"] == fast
    return env


@pytest.mark.parametrize("coll", ["priority", "block_both", "soft"])
@pytest.mark.parametrize("ot", ["finish", "nothing", "restart"])
@pytest.mark.parametrize("fast_team,A", [(32, 32), (32, 64), (32, 128), (64, 64), (64, 208), (128, 256), (256, 1008), (128, 48)])
def test_fast_equals_generic_and_oracle(coll, ot, fast_team, A, monkeypatch):
    import torch
    size = 20 if A <= 64 else (40 if A <= 256 else 72)
    gc = dict(size=size, density=0.15, num_agents=A, obs_radius=3 + (A % 3), max_episode_steps=9,
              collision_system=coll, on_target=ot)
    n = 5
    seeds = list(range(40, 40 + n))
    T = 14
    fast = build(gc, n, seeds, monkeypatch, True, fast_team=fast_team)
    p = fast.engine.plan()["fast"]
    assert p["team_threads"] == fast_team and p["agents_per_thread"] in (1, 2, 4)
    slow = build(gc, n, seeds, monkeypatch, False)
    of, os_ = fast.reset(), slow.reset()
    assert torch.equal(of, os_)
    co = COracle.from_python_oracle(gc, seeds) if A <= 256 else None
    acts = make_actions(T, n, A, seed=3)
    for t in range(T):
        a = torch.from_numpy(acts[t]).cuda()
        rf, rs = fast.step(a), slow.step(a)
        for x, y in zip(rf, rs):
            assert torch.equal(x, y), (coll, ot, fast_team, A, t)
        assert torch.equal(fast._state(), slow._state())
        assert torch.equal(fast.was_on_goal, slow.was_on_goal) and torch.equal(fast.elapsed_steps, slow.elapsed_steps)
        if co is not None:
            out = co.run(acts[t:t + 1], auto_reset=True)
            assert np.array_equal(rf[0].cpu().numpy(), out["obs"]) and np.array_equal(rf[1].cpu().numpy(), out["rewards"])
            assert np.array_equal(rf[2].cpu().numpy(), out["terminated"]) and np.array_equal(rf[3].cpu().numpy(), out["truncated"])
    # multi-step launch == single steps (fresh engines), metrics and checkpoints agree
    f2 = build(gc, n, seeds, monkeypatch, True, fast_team=fast_team)
    f2.reset()
    ring = torch.stack([f2.new_obs_buffer() for _ in range(3)])
    obs, rew, term, trunc = f2.rollout(torch.from_numpy(acts).cuda(), obs_out=ring)
    assert torch.equal(obs[(T - 1) % 3], rf[0]) and torch.equal(rew[-1], rf[1])
    assert np.array_equal(f2.engine.checkpoint(), fast.engine.checkpoint())
    assert np.array_equal(f2.engine.checkpoint(), slow.engine.checkpoint())
    for k, v in fast.metrics().items():
        assert np.array_equal(v, slow.metrics()[k])
    fast.check_errors(), slow.check_errors(), f2.check_errors()


@pytest.mark.parametrize("r", [2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("size", [12, 32, 33, 70])
def test_fast_radii_and_map_widths(r, size, monkeypatch):
    """narrow maps (<= 32 wide, two bitmap words per row: 64-bit row loads) and wide ones, every static radius,
    u8 and bits formats."""
    import torch
    A = 48
    gc = dict(size=size, density=0.2, num_agents=A, obs_radius=r, max_episode_steps=7, collision_system="priority",
              on_target="finish")
    seeds = [3, 4, 5]
    co = COracle.from_python_oracle(gc, seeds)
    envs = [build(gc, 3, seeds, monkeypatch, True, fmt=f) for f in ("u8", "bits")]
    for e in envs:
        e.reset()
    acts = make_actions(10, 3, A, seed=r)
    D = 2 * r + 1
    for t in range(10):
        a = torch.from_numpy(acts[t]).cuda()
        out = co.run(acts[t:t + 1], auto_reset=True)
        o8 = envs[0].step(a)[0].cpu().numpy()
        ob = envs[1].step(a)[0].cpu().numpy()
        assert np.array_equal(o8, out["obs"]), (r, size, t)
        bits = np.unpackbits(ob.view(np.uint8), bitorder="little").reshape(3, A, -1)[:, :, :3 * D * D].reshape(3, A, 3, D, D)
        assert np.array_equal(bits, out["obs"]), (r, size, t)


def test_fast_kernel_behind_the_packed_host_transport(monkeypatch):
    """pgm_step_host with the packed transport: the fast kernel writes the raw stream in batches of its team size,
    the reseeding observe pass (generic kernel) must use the same geometry."""
    gc = dict(size=24, density=0.2, num_agents=80, obs_radius=4, max_episode_steps=5, collision_system="soft",
              on_target="restart")
    n = 7
    for ar in (True, "reseed"):
        a = build(gc, n, list(range(n)), monkeypatch, True, fast_team=32, auto_reset=ar)
        b = build(gc, n, list(range(n)), monkeypatch, False, auto_reset=ar)
        a.reset(), b.reset()
        a.engine.set_host_transport("packed", 3)
        b.engine.set_host_transport("plain")
        acts = make_actions(17, n, 80, seed=11)
        for t in range(17):
            ra, rb = a.step_host(acts[t]), b.step_host(acts[t])
            for x, y in zip(ra, rb):
                assert np.array_equal(x, y), (ar, t)
        assert np.array_equal(a.current_seeds(), b.current_seeds())


def test_fast_kernel_falls_back_on_misaligned_observation_pointers(monkeypatch):
    import torch
    gc = dict(size=16, density=0.2, num_agents=32, obs_radius=3, max_episode_steps=8)
    a = build(gc, 4, [1, 2, 3, 4], monkeypatch, True)
    b = build(gc, 4, [1, 2, 3, 4], monkeypatch, True)
    a.reset(), b.reset()
    raw = torch.empty(a.engine.obs_bytes + 64, dtype=torch.uint8, device="cuda")
    odd = raw[3:3 + a.engine.obs_bytes].view(a.engine.obs_shape())      # 3 bytes off: generic kernel
    acts = make_actions(6, 4, 32, seed=2)
    for t in range(6):
        act = torch.from_numpy(acts[t]).cuda()
        assert torch.equal(a.step(act, out=odd)[0], b.step(act)[0])


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("K,ring", [(40, 3), (25, 4), (33, 33), (64, 1)])
def test_long_rollouts(fast, K, ring, monkeypatch):
    """pgm_step_many with many steps per launch and observation rings that do not divide K: same results as K single
    steps, every step's observation in slot k % ring, outputs at [k]."""
    import torch
    gc = dict(size=14, density=0.2, num_agents=32, obs_radius=3, max_episode_steps=11, collision_system="soft",
              on_target="restart")
    n = 5
    a = build(gc, n, list(range(n)), monkeypatch, fast)
    b = build(gc, n, list(range(n)), monkeypatch, fast)
    a.reset(), b.reset()
    acts = torch.from_numpy(make_actions(K, n, 32, seed=K)).cuda()
    obs_ring = torch.zeros((ring,) + tuple(a.engine.obs_shape()), dtype=torch.uint8, device="cuda")
    obs, rew, term, trunc = a.rollout(acts, obs_out=obs_ring)
    last = {}
    for k in range(K):
        o, r, te, tr = b.step(acts[k])
        assert torch.equal(rew[k], r) and torch.equal(term[k], te) and torch.equal(trunc[k], tr), k
        last[k % ring] = o.clone()
    for slot, o in last.items():
        assert torch.equal(obs[slot], o), slot
    assert np.array_equal(a.engine.checkpoint(), b.engine.checkpoint())
    a.check_errors(), b.check_errors()
