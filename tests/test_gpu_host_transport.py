"""Packed host transport of pgm_step_host / pgm_observe_host (the GPU writes the observation bit stream, the
copy engine moves it, host threads widen it): obs_host must receive exactly the bytes of the plain DMA path
and of the device-pointer path - for ragged shapes, several observation batches per instance, float32,
reseeding auto-reset, and against the oracle."""
import numpy as np
import pytest

from tests.helpers import make_actions, run_oracle

pytestmark = pytest.mark.gpu


def _host_bufs(e):
    n, a = e.num_envs, e.num_agents
    # +3: a deliberately misaligned destination (the caller owns the buffer; any alignment must work)
    raw = np.zeros(int(np.prod(e.obs_shape())) * np.dtype(e.obs_dtype()).itemsize + 64, np.uint8)
    off = {"f32": 4, "f16": 2}.get(e.obs_format, 3)
    obs = raw[off:off + raw.size - 64].view(e.obs_dtype()).reshape(e.obs_shape())
    return obs, np.empty((n, a), np.float32), np.empty((n, a), np.uint8), np.empty((n, a), np.uint8)


CASES = [
    # size, agents, r, envs, collision, on_target, fmt, auto_reset, threads
    (8, 3, 1, 5, "priority", "finish", "u8", True, 3),          # 27 bytes per agent: nothing is aligned
    (16, 12, 3, 7, "soft", "restart", "u8", True, 0),
    (32, 64, 5, 64, "priority", "finish", "u8", True, 0),        # configs[1] shape
    (32, 64, 5, 33, "block_both", "nothing", "f32", True, 2),
    (10, 5, 2, 4, "priority", "finish", "f32", False, 1),
    (32, 64, 5, 48, "soft", "finish", "f16", True, 0),
    (9, 5, 3, 6, "priority", "restart", "f16", True, 2),
    (24, 40, 60, 3, "priority", "finish", "u8", True, 4),        # r=60: several observation batches per instance
    (12, 9, 4, 11, "soft", "finish", "u8", "reseed", 0),         # rebuilt tasks: masked observe pass writes the stream
]


@pytest.mark.parametrize("size,agents,r,envs,coll,ot,fmt,auto_reset,threads", CASES)
def test_packed_transport_equals_plain_and_device(size, agents, r, envs, coll, ot, fmt, auto_reset, threads):
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = GridConfig(size=size, density=0.2, num_agents=agents, obs_radius=r, max_episode_steps=9,
                    collision_system=coll, on_target=ot, seed=3)
    mk = lambda: BatchedPogema(gc, num_envs=envs, auto_reset=auto_reset, obs_format=fmt)
    dev, plain, packed = mk(), mk(), mk()
    plain.engine.set_host_transport("plain")
    packed.engine.set_host_transport("packed", threads)
    assert packed.engine.host_transport_info()["packed"] and not plain.engine.host_transport_info()["packed"]
    o_dev = dev.reset()
    plain.reset(), packed.reset()
    assert np.array_equal(packed.engine.observe_host(), o_dev.cpu().numpy())
    assert np.array_equal(plain.engine.observe_host(), o_dev.cpu().numpy())
    acts = make_actions(25, envs, agents, seed=11)
    bp, bk = _host_bufs(plain.engine), _host_bufs(packed.engine)
    for t in range(acts.shape[0]):
        o, rew, te, tr = dev.step(torch.from_numpy(acts[t]).cuda())
        plain.engine.step_host(acts[t], *bp)
        packed.engine.step_host(acts[t], *bk)
        assert np.array_equal(bk[0], o.cpu().numpy()), f"packed obs differ at step {t}"
        assert np.array_equal(bp[0], bk[0])
        for x, y in zip(bp[1:], bk[1:]):
            assert np.array_equal(x, y)
        assert np.array_equal(bk[1], rew.cpu().numpy())
    info = packed.engine.host_transport_info()
    assert info["threads"] >= 1 and info["d2h_bytes"] < plain.engine.host_transport_info()["d2h_bytes"]
    # skipping observations still works on the packed engine
    packed.engine.step_host(acts[0], None, *bk[1:])
    dev.step(torch.from_numpy(acts[0]).cuda())
    assert np.array_equal(packed.engine.observe_host(), dev.observe().cpu().numpy())


def test_packed_transport_against_oracle():
    from pogema_b200 import BatchedPogema, GridConfig
    kw = dict(size=12, density=0.3, num_agents=8, obs_radius=5, max_episode_steps=16, collision_system="priority",
              on_target="finish")
    seeds = list(range(6))
    env = BatchedPogema(GridConfig(**kw), num_envs=len(seeds), seeds=seeds, auto_reset=True)
    env.engine.set_host_transport("packed", 2)
    env.reset()
    acts = make_actions(40, len(seeds), kw["num_agents"], seed=5)
    refs = [run_oracle(kw, s, acts[:, k], auto_reset=True) for k, s in enumerate(seeds)]
    assert np.array_equal(env.engine.observe_host(), np.stack([r["obs"][0] for r in refs]))
    for t in range(acts.shape[0]):
        o, rew, te, tr = env.step_host(acts[t])
        assert np.array_equal(o, np.stack([r["obs"][t + 1] for r in refs])), t
        assert np.array_equal(rew, np.stack([r["rewards"][t] for r in refs]))
        assert np.array_equal(te, np.stack([r["terminated"][t] for r in refs]))
        assert np.array_equal(tr, np.stack([r["truncated"][t] for r in refs]))


def test_packed_transport_rejected_for_bits_and_auto_threshold():
    from pogema_b200 import BatchedPogema, GridConfig
    from pogema_b200._native import PgmError
    gc = GridConfig(size=8, density=0.2, num_agents=4, obs_radius=2, seed=1)
    env = BatchedPogema(gc, num_envs=2, obs_format="bits")
    with pytest.raises(PgmError):
        env.engine.set_host_transport("packed")
    small = BatchedPogema(gc, num_envs=2)
    assert not small.engine.host_transport_info()["packed"]            # auto: tiny tensors use the plain DMA
    big = BatchedPogema(GridConfig(size=16, density=0.2, num_agents=32, obs_radius=5, seed=1), num_envs=512)
    assert big.engine.host_transport_info()["packed"]                  # 5.9 MB of observations


@pytest.mark.parametrize("envs,mode", [(1, "auto"), (3, "auto"), (700, "plain"), (700, "packed")])
def test_step_host_ex_flags_match_state(envs, mode):
    """pgm_step_host_ex: is_active / was_on_goal arrive with the step results (single small copy for tiny engines,
    separate copies for large ones) and equal what pgm_get_state reports afterwards."""
    from pogema_b200 import BatchedPogema, GridConfig
    from pogema_b200 import _native as nat
    gc = GridConfig(size=10, density=0.2, num_agents=7, obs_radius=5, max_episode_steps=40, collision_system="soft",
                    on_target="finish", seed=2)
    a = BatchedPogema(gc, num_envs=envs, auto_reset=False)
    b = BatchedPogema(gc, num_envs=envs, auto_reset=False)
    a.engine.set_host_transport(mode)
    a.reset(), b.reset()
    acts = make_actions(30, envs, 7, seed=9)
    obs, rew, te, tr = _host_bufs(a.engine)
    active = np.full((envs, 7), 9, np.uint8)
    was = np.full((envs, 7), 9, np.uint8)
    import torch
    for t in range(acts.shape[0]):
        a.engine.step_host(acts[t], obs, rew, te, tr, active=active, was_on_goal=was)
        o, r, term, trunc = b.step(torch.from_numpy(acts[t]).cuda())
        assert np.array_equal(obs, o.cpu().numpy()) and np.array_equal(rew, r.cpu().numpy())
        assert np.array_equal(te.astype(bool), term.cpu().numpy()) and np.array_equal(tr.astype(bool), trunc.cpu().numpy())
        assert np.array_equal(active, b.engine.get_state(nat.STATE_ACTIVE))
        assert np.array_equal(was, b.engine.get_state(nat.STATE_WAS_ON_GOAL))
    assert active.min() == 0          # some agents finished and disappeared
    # flags are optional, individually
    a.engine.step_host(acts[0], None, rew, te, tr, active=active)
    a.engine.step_host(acts[0], obs, rew, te, tr, was_on_goal=was)


def test_failed_packed_call_leaves_the_pool_usable():
    """A packed host call that fails after the widening threads were woken (step before any task exists) must
    return an error, not hang, and the next call must work."""
    from pogema_b200 import GridConfig
    from pogema_b200._native import PgmError
    from pogema_b200.engine import Engine
    gc = GridConfig(size=8, density=0.2, num_agents=4, obs_radius=2, seed=1)
    e = Engine(gc, 6)
    e.set_host_transport("packed", 3)
    obs, rew, te, tr = _host_bufs(e)
    acts = np.zeros((6, 4), np.uint8)
    for _ in range(2):
        with pytest.raises(PgmError):
            e.step_host(acts, obs, rew, te, tr)
    e.generate(list(range(6)))
    e.reset()
    ref = e.observe_host()
    e.step_host(acts, obs, rew, te, tr)           # everybody stays: the observation equals the reset one
    assert np.array_equal(obs, ref)
    e.close()
