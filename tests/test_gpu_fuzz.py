"""Randomised differential test: random shapes / radii / modes / team sizes / occupancy structures,
every output of every step against the C oracle (which tests/test_oracle_c.py pins to the Python oracle)."""
import os
import zlib

import numpy as np
import pytest

from tests.helpers import make_actions
from tests.oracle_c import COracle

pytestmark = pytest.mark.gpu

COLLS = ("priority", "block_both", "soft")
ONTS = ("finish", "nothing", "restart")


def run_both(gc, seeds, T, auto_reset, team, fmt="u8", path="device"):
    """path: 'device' = pgm_step with device pointers; 'plain' / 'packed' = pgm_step_host_ex with host buffers
    through the plain DMA or the packed transport (GPU-written bit stream, host threads widen it)."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    env = BatchedPogema(GridConfig(**gc), num_envs=len(seeds), seeds=seeds, auto_reset=auto_reset, team_threads=team,
                        obs_format=fmt)
    co = COracle.from_python_oracle(gc, seeds)
    A = env.num_agents
    r = gc["obs_radius"]
    actions = make_actions(T, len(seeds), A, seed=zlib.crc32(repr(sorted(gc.items(), key=str)).encode()) % 1000)  # reproducible across processes
    obs = env.reset()
    if path != "device":
        env.engine.set_host_transport(path, 1 + len(seeds) % 3)
        n = len(seeds)
        hb = (np.empty(env.engine.obs_shape(), env.engine.obs_dtype()), np.empty((n, A), np.float32),
              np.empty((n, A), np.uint8), np.empty((n, A), np.uint8), np.empty((n, A), np.uint8), np.empty((n, A), np.uint8))
    for t in range(T):
        if path == "device":
            o, rew, te, tr = env.step(torch.from_numpy(actions[t]).cuda())
            o, rew, te, tr = o.cpu().numpy(), rew.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()
        else:
            env.engine.step_host(actions[t], hb[0], hb[1], hb[2], hb[3], active=hb[4], was_on_goal=hb[5])
            o, rew, te, tr = hb[0], hb[1], hb[2].astype(bool), hb[3].astype(bool)
        out = co.run(actions[t:t + 1], auto_reset=auto_reset)
        assert np.array_equal(env.get_agents_xy().cpu().numpy() + r, co.pos), (gc, t)
        assert np.array_equal(env.get_targets_xy().cpu().numpy() + r, co.tgt), (gc, t)
        assert np.array_equal(env.is_active.cpu().numpy().astype(np.uint8), co.active), (gc, t)
        if path != "device":
            assert np.array_equal(hb[4], co.active), (gc, t)
        og = o
        if fmt == "bits":
            D = 2 * r + 1
            og = np.unpackbits(og.view(np.uint8), bitorder="little").reshape(len(seeds), A, -1)[:, :, :3 * D * D]
            og = og.reshape(len(seeds), A, 3, D, D)
        assert np.array_equal(og, out["obs"]), (gc, t)
        assert np.array_equal(rew, out["rewards"]), (gc, t)
        assert np.array_equal(te, out["terminated"]) and np.array_equal(tr, out["truncated"])
    env.check_errors()
    env.close()


@pytest.mark.parametrize("case", range(int(os.environ.get("PGM_FUZZ_CASES", "36"))))
def test_random_configurations(case, monkeypatch):
    rng = np.random.default_rng(1000 + case)
    size = int(rng.integers(4, 36))
    r = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 12, 17]))
    density = float(rng.choice([0.0, 0.1, 0.2, 0.3, 0.4]))
    free = int(size * size * (1 - density))
    A = int(min(rng.choice([1, 2, 3, 5, 8, 17, 33, 64, 100]), max(1, free // 4)))
    gc = dict(size=size, density=density, num_agents=A, obs_radius=r, max_episode_steps=int(rng.integers(3, 20)),
              collision_system=COLLS[case % 3], on_target=ONTS[(case // 3) % 3])
    team = int(rng.choice([0, 0, 32, 64, 128]))
    if rng.random() < 0.4:
        monkeypatch.setenv("PGM_OCC", "1")
    seeds = [int(s) for s in rng.integers(0, 10_000, size=int(rng.integers(2, 6)))]
    from oracle import pogema_oracle as orc
    ok_seeds = []
    for s in seeds:
        try:
            orc.Grid(orc.GridConfig(seed=s, **gc))
            ok_seeds.append(s)
        except OverflowError:
            pass
    if not ok_seeds:
        pytest.skip("no placeable seed")
    fmt = "bits" if case % 5 == 4 else "u8"
    path = "device" if fmt == "bits" else ("device", "packed", "plain")[(case // 2) % 3]
    run_both(gc, ok_seeds, T=24, auto_reset=bool(case % 2), team=team, fmt=fmt, path=path)


def test_rectangular_map_and_single_agent():
    m = (np.random.default_rng(3).random((7, 19)) < 0.2).astype(np.uint8)
    for coll in COLLS:
        gc = dict(map=m.tolist(), num_agents=9, obs_radius=3, max_episode_steps=10, collision_system=coll,
                  on_target="restart")
        run_both(gc, [1, 2, 3], T=25, auto_reset=True, team=0)
    run_both(dict(size=2, density=0.0, num_agents=1, obs_radius=1, max_episode_steps=5), [0, 1, 2, 3], T=12,
             auto_reset=True, team=0)
    run_both(dict(size=6, density=0.2, num_agents=1, obs_radius=5, max_episode_steps=7, on_target="restart"),
             [0, 1, 2], T=20, auto_reset=True, team=0)


def test_observation_batches_when_the_stage_does_not_fit():
    """r=60 (D=121): 43923 bits per agent; 48 agents do not fit the stage buffer at once."""
    gc = dict(size=12, density=0.2, num_agents=48, obs_radius=60, max_episode_steps=6, collision_system="soft",
              on_target="finish")
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    env = BatchedPogema(GridConfig(**gc), num_envs=2, seeds=[4, 5])
    assert env.engine.plan()["agents_per_obs_batch"] < 48
    env.close()
    run_both(gc, [4, 5], T=8, auto_reset=True, team=0)


def test_maximum_observation_radius():
    """GridConfig's maximum obs_radius (128): D = 257, 198147 values per agent."""
    gc = dict(size=6, density=0.1, num_agents=3, obs_radius=128, max_episode_steps=4, collision_system="priority",
              on_target="restart")
    run_both(gc, [1, 2], T=6, auto_reset=True, team=0)
