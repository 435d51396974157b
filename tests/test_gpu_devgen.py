"""Device-side task generation (pgm_generate_device) must build exactly what the host generator
builds (which the CPU tests pin against the Python oracle / numpy): obstacles, starts, goals,
lifelong generators and component tables - including the instances that fall back to the host
because upstream's retry loop is needed."""
import numpy as np
import pytest

from pogema_b200 import _native as nat

pytestmark = pytest.mark.gpu


def build(gc_kwargs, seeds, on_device, auto_reset=True):
    from pogema_b200 import BatchedPogema, GridConfig
    return BatchedPogema(GridConfig(**gc_kwargs), num_envs=len(seeds), seeds=seeds, auto_reset=auto_reset,
                         generate_on_device=on_device)


def assert_same_tasks(a, b):
    assert np.array_equal(a.get_obstacles(), b.get_obstacles())
    assert np.array_equal(a.engine.get_state(nat.STATE_POSITIONS), b.engine.get_state(nat.STATE_POSITIONS))
    assert np.array_equal(a.engine.get_state(nat.STATE_TARGETS), b.engine.get_state(nat.STATE_TARGETS))
    assert np.array_equal(a.engine.checkpoint(), b.engine.checkpoint())      # includes the lifelong PCG64 states


@pytest.mark.parametrize("gc", [
    dict(size=8, density=0.3, num_agents=4, obs_radius=5),
    dict(size=32, density=0.3, num_agents=64, obs_radius=5),
    dict(size=17, density=0.0, num_agents=40, obs_radius=2),
    dict(size=20, density=0.45, num_agents=30, obs_radius=3),
    dict(size=24, density=0.7, num_agents=6, obs_radius=3),
    dict(size=64, density=0.3, num_agents=256, obs_radius=5, on_target="restart", collision_system="soft"),
    dict(size=16, density=0.35, num_agents=12, obs_radius=4, on_target="restart"),
])
def test_device_generation_equals_host_generation(gc):
    import torch
    seeds = list(range(1000, 1000 + 96))
    gc = dict(max_episode_steps=16, **gc)
    try:
        host = build(gc, seeds, on_device=False)
    except OverflowError:
        with pytest.raises(OverflowError):
            build(gc, seeds, on_device=True)
        return
    dev = build(gc, seeds, on_device=True)
    assert_same_tasks(host, dev)
    # stepping consumes the lifelong component tables and generators
    oh, od = host.reset(), dev.reset()
    assert torch.equal(oh, od)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(40):
        a = host.sample_actions(g)
        rh, rd = host.step(a), dev.step(a)
        for x, y in zip(rh, rd):
            assert torch.equal(x, y), t
    assert_same_tasks(host, dev)


def test_fallback_instances_and_overflow():
    from pogema_b200 import BatchedPogema, GridConfig
    # 5x5 maps at density 0.6: some seeds need upstream's retry loop, some cannot be placed at all
    gc = dict(size=5, density=0.6, num_agents=6, obs_radius=2)
    good = []
    for s in range(80):
        try:
            BatchedPogema(GridConfig(**gc), num_envs=1, seeds=[s], generate_on_device=False)
            good.append(s)
        except OverflowError:
            pass
    assert 5 < len(good) < 80
    host = build(gc, good, on_device=False)
    dev = build(gc, good, on_device=True)
    assert dev.engine.last_host_fallbacks > 0          # the retry loop ran on the host for some of them
    assert_same_tasks(host, dev)
    with pytest.raises(OverflowError):
        build(gc, list(range(80)), on_device=True)


def test_fixed_map_and_reset_with_new_seeds():
    import torch
    from pogema_b200.maps import maze_map
    m = maze_map(32, 1)
    gc = dict(map=m.tolist(), num_agents=40, obs_radius=3, on_target="restart", max_episode_steps=8)
    a = build(gc, list(range(10)), on_device=True)
    b = build(gc, list(range(50, 60)), on_device=False)
    a.reset(seeds=list(range(50, 60)))                   # rebuild on the device for the new seeds
    b.reset()
    assert_same_tasks(a, b)
    assert torch.equal(a.observe(), b.observe())


def test_device_generation_is_fast():
    import time
    import torch
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=5)
    env = build(gc, list(range(4096)), on_device=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.reset(seeds=list(range(4096, 8192)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"device regeneration of 4096 instances: {dt * 1e3:.2f} ms")
    assert dt < 0.25


@pytest.mark.parametrize("ot,coll", [("finish", "priority"), ("restart", "soft"), ("nothing", "block_both")])
def test_auto_reset_with_new_seeds_matches_oracle(ot, coll):
    """auto_reset='reseed': every finished episode is followed by a task built from seed + stride,
    i.e. a fresh oracle env with that seed."""
    import torch
    from oracle import pogema_oracle as orc
    from pogema_b200 import BatchedPogema, GridConfig
    kw = dict(size=8, density=0.25, num_agents=3, obs_radius=3, max_episode_steps=6, collision_system=coll,
              on_target=ot)
    seeds = [3, 4, 5, 6, 7]
    stride = 100
    env = BatchedPogema(GridConfig(**kw), num_envs=len(seeds), seeds=seeds, auto_reset="reseed", reseed_stride=stride)
    obs = env.reset().cpu().numpy()
    cur = list(seeds)
    refs = []
    for k, s in enumerate(seeds):
        r = orc.pogema_v0(orc.GridConfig(seed=s, **kw))
        o, _ = r.reset()
        assert np.array_equal(obs[k], np.stack(o).astype(np.uint8))
        refs.append(r)
    rng = np.random.default_rng(0)
    episodes = 0
    for t in range(40):
        acts = rng.integers(0, 5, size=(len(seeds), kw["num_agents"])).astype(np.uint8)
        o, r, te, tr = env.step(torch.from_numpy(acts).cuda())
        o, r, te, tr = o.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()
        for k in range(len(seeds)):
            ro, rr, rte, rtr, _ = refs[k].step(list(acts[k]))
            assert np.array_equal(r[k], np.array(rr, dtype=np.float32)) and np.array_equal(te[k], np.array(rte))
            assert np.array_equal(tr[k], np.array(rtr))
            if all(rte) or all(rtr):
                cur[k] += stride
                refs[k] = orc.pogema_v0(orc.GridConfig(seed=cur[k], **kw))
                ro, _ = refs[k].reset()
                episodes += 1
            assert np.array_equal(o[k], np.stack(ro).astype(np.uint8)), (t, k)
    assert episodes >= 25
    assert env.current_seeds().tolist() == cur
    env.check_errors()
    with pytest.raises(Exception):
        env.rollout(torch.zeros((4, len(seeds), kw["num_agents"]), dtype=torch.uint8, device="cuda"))
