"""GPU parity: the CUDA step engine (through the C-ABI, via BatchedPogema) against
the CPU oracle on identical seeds and action streams - bit-exact on positions,
targets, active flags, rewards, terminated, truncated and every observation byte."""
import itertools

import numpy as np
import pytest

from tests.helpers import make_actions, run_oracle

pytestmark = pytest.mark.gpu

COLLISIONS = ("priority", "block_both", "soft")
ON_TARGETS = ("finish", "nothing", "restart")


def run_gpu(gc_kwargs, seeds, actions, auto_reset=False, obs_format="u8", team_threads=0):
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    env = BatchedPogema(GridConfig(**gc_kwargs), num_envs=len(seeds), seeds=seeds, auto_reset=auto_reset,
                        obs_format=obs_format, team_threads=team_threads)
    out = {k: [] for k in ("obs", "pos", "tgt", "active", "rewards", "terminated", "truncated")}

    def snap(o):
        out["obs"].append(o.cpu().numpy().copy())
        out["pos"].append(env.get_agents_xy().cpu().numpy())
        out["tgt"].append(env.get_targets_xy().cpu().numpy())
        out["active"].append(env.is_active.cpu().numpy().astype(np.uint8))

    snap(env.reset())
    for t in range(actions.shape[0]):
        o, r, te, tr = env.step(torch.from_numpy(actions[t]).cuda())
        snap(o)
        out["rewards"].append(r.cpu().numpy().copy())
        out["terminated"].append(te.cpu().numpy().copy())
        out["truncated"].append(tr.cpu().numpy().copy())
    env.check_errors()
    res = {k: np.stack(v) for k, v in out.items()}
    res["env"] = env
    return res


def compare(gc_kwargs, seeds, T, auto_reset=False, team_threads=0, action_seed=7):
    A = gc_kwargs["num_agents"]
    actions = make_actions(T, len(seeds), A, seed=action_seed)
    gpu = run_gpu(gc_kwargs, seeds, actions, auto_reset=auto_reset, team_threads=team_threads)
    for k, seed in enumerate(seeds):
        ref = run_oracle(gc_kwargs, seed, actions[:, k], auto_reset=auto_reset)
        for key in ("pos", "tgt", "active", "obs"):
            g = gpu[key][:, k]
            assert g.shape == ref[key].shape, (key, g.shape, ref[key].shape)
            if not np.array_equal(g, ref[key]):
                t = int(np.argmax([not np.array_equal(g[i], ref[key][i]) for i in range(g.shape[0])]))
                raise AssertionError(f"{key} differs: instance {k} seed {seed} first at t={t}\n"
                                     f"gpu={g[t].tolist() if key != 'obs' else 'obs'}\nref={ref[key][t].tolist() if key != 'obs' else 'obs'}\n"
                                     f"cfg={gc_kwargs}")
        for key in ("rewards", "terminated", "truncated"):
            g = gpu[key][:, k]
            assert np.array_equal(g, ref[key]), (key, k, seed, gc_kwargs)
    return gpu


@pytest.mark.parametrize("coll,ot", list(itertools.product(COLLISIONS, ON_TARGETS)))
def test_all_modes_small(coll, ot):
    """BASELINE.json configs[0] shape (8x8, 4 agents, r=5) for all 9 mode combinations."""
    gc = dict(size=8, density=0.3, num_agents=4, obs_radius=5, max_episode_steps=64, collision_system=coll,
              on_target=ot)
    compare(gc, seeds=list(range(16)), T=70)


@pytest.mark.parametrize("coll,ot", list(itertools.product(COLLISIONS, ON_TARGETS)))
def test_all_modes_crowded(coll, ot):
    """Crowded 10x10 maps: many conflicts, chains and rotations."""
    gc = dict(size=10, density=0.1, num_agents=40, obs_radius=3, max_episode_steps=32, collision_system=coll,
              on_target=ot)
    compare(gc, seeds=list(range(100, 124)), T=40)


@pytest.mark.parametrize("coll,ot", list(itertools.product(COLLISIONS, ON_TARGETS)))
def test_all_modes_autoreset(coll, ot):
    gc = dict(size=8, density=0.2, num_agents=6, obs_radius=2, max_episode_steps=10, collision_system=coll,
              on_target=ot)
    compare(gc, seeds=list(range(8)), T=35, auto_reset=True)


@pytest.mark.parametrize("team", [32, 64, 128, 256])
def test_config2_shape_team_sizes(team):
    """BASELINE.json configs[1] instance shape (32x32, 64 agents, r=5, priority/finish), few instances."""
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
              collision_system="priority", on_target="finish")
    compare(gc, seeds=list(range(6)), T=66, team_threads=team)


@pytest.mark.parametrize("r", [1, 2, 3, 5, 7, 10, 16, 20])
def test_obs_radius_sweep(r):
    gc = dict(size=12, density=0.25, num_agents=9, obs_radius=r, max_episode_steps=16,
              collision_system="soft", on_target="restart")
    compare(gc, seeds=list(range(5)), T=20)


def test_lifelong_64x64():
    """BASELINE.json configs[2] shape, scaled down in instance count (random maps; maze maps in test_maps)."""
    gc = dict(size=64, density=0.3, num_agents=256, obs_radius=5, max_episode_steps=64,
              collision_system="soft", on_target="restart")
    compare(gc, seeds=[0, 1], T=40)


def test_block_both_many_agents():
    gc = dict(size=48, density=0.2, num_agents=600, obs_radius=5, max_episode_steps=32,
              collision_system="block_both", on_target="finish")
    compare(gc, seeds=[3], T=24)


@pytest.mark.parametrize("coll,ot", list(itertools.product(COLLISIONS, ON_TARGETS)))
def test_multi_step_launch_equals_single_steps(coll, ot):
    """pgm_step_many (K steps in one launch) == K pgm_step launches == the oracle."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = dict(size=10, density=0.15, num_agents=24, obs_radius=3, max_episode_steps=9, collision_system=coll,
              on_target=ot)
    seeds = list(range(30, 40))
    K = 25
    actions = make_actions(K, len(seeds), gc["num_agents"], seed=11)
    a = BatchedPogema(GridConfig(**gc), num_envs=len(seeds), seeds=seeds, auto_reset=True)
    b = BatchedPogema(GridConfig(**gc), num_envs=len(seeds), seeds=seeds, auto_reset=True)
    a.reset(), b.reset()
    act = torch.from_numpy(actions).cuda()
    obs, rew, term, trunc = a.rollout(act)
    for k in range(K):
        o, r, te, tr = b.step(act[k])
        assert torch.equal(obs[k], o), k
        assert torch.equal(rew[k], r) and torch.equal(term[k], te) and torch.equal(trunc[k], tr), k
    assert torch.equal(a._state(), b._state())
    assert torch.equal(a.elapsed_steps, b.elapsed_steps)
    assert np.array_equal(a.engine.get_state(7), b.engine.get_state(7))       # metric counters
    for i, seed in enumerate(seeds[:4]):
        ref = run_oracle(gc, seed, actions[:, i], auto_reset=True)
        assert np.array_equal(obs[:, i].cpu().numpy(), ref["obs"][1:])
        assert np.array_equal(rew[:, i].cpu().numpy(), ref["rewards"])
    # observation ring smaller than K: slot k % R holds the last step written to it
    c = BatchedPogema(GridConfig(**gc), num_envs=len(seeds), seeds=seeds, auto_reset=True)
    c.reset()
    ring = torch.empty((4,) + tuple(obs.shape[1:]), dtype=obs.dtype, device="cuda")
    c.rollout(act, obs_out=ring)
    for k in range(K - 4, K):
        assert torch.equal(ring[k % 4], obs[k])


@pytest.mark.parametrize("coll,ot", list(itertools.product(COLLISIONS, ON_TARGETS)))
def test_bucket_occupancy_variant(coll, ot, monkeypatch):
    """The tile-bucket occupancy structure (chosen automatically for large maps) forced on crowded small maps."""
    monkeypatch.setenv("PGM_OCC", "1")
    gc = dict(size=10, density=0.1, num_agents=40, obs_radius=3, max_episode_steps=32, collision_system=coll,
              on_target=ot)
    gpu = compare(gc, seeds=list(range(200, 216)), T=40, auto_reset=True)
    assert gpu["env"].engine.plan()["occupancy_buckets"] == 1


@pytest.mark.parametrize("coll", COLLISIONS)
def test_large_map_few_agents(coll):
    """A 400x400 map does not fit a dense cell grid in shared memory: the planner must pick the tile buckets."""
    gc = dict(size=400, density=0.2, num_agents=24, obs_radius=4, max_episode_steps=16, collision_system=coll,
              on_target="finish")
    gpu = compare(gc, seeds=[1, 2], T=12)
    assert gpu["env"].engine.plan()["occupancy_buckets"] == 1


@pytest.mark.parametrize("coll", COLLISIONS)
def test_maximum_map_size(coll):
    """GridConfig's maximum size (1024): the bitmaps no longer fit in shared memory next to each other, the
    obstacle bitmap is read from global memory.  Compared against the C oracle built from the engine's state."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    from tests.test_gpu_fullsize import c_oracle_from_engine
    gc = dict(size=1024, density=0.3, num_agents=300, obs_radius=5, max_episode_steps=6, collision_system=coll,
              on_target="finish")
    env = BatchedPogema(GridConfig(**gc), num_envs=2, seeds=[5, 6], auto_reset=True)
    obs = env.reset()
    co = c_oracle_from_engine(env)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = []
    for t in range(8):
        a = env.sample_actions(g)
        acts.append(a.cpu().numpy())
        obs, rew, term, trunc = env.step(a)
    out = co.run(np.stack(acts), auto_reset=True)
    assert np.array_equal(env.get_agents_xy().cpu().numpy() + 5, co.pos)
    assert np.array_equal(obs.cpu().numpy(), out["obs"])
    assert np.array_equal(rew.cpu().numpy(), out["rewards"])


def test_unsupported_shape_fails_loudly():
    """More agents than one SM's shared memory can hold scratch for: a clear error, never a silent fallback."""
    from pogema_b200 import BatchedPogema, GridConfig
    from pogema_b200._native import PgmError
    with pytest.raises(PgmError, match="does not fit"):
        BatchedPogema(GridConfig(size=400, density=0.1, num_agents=20000, seed=0), num_envs=1, generate_on_device=False)
