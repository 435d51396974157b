"""The oracle against hand-derived outcomes (tests/scenarios.py) and against hand-computed
observation / bookkeeping facts.  No GPU needed."""
import numpy as np
import pytest

from oracle import pogema_oracle as orc
from tests.scenarios import SCENARIOS, clean_map


@pytest.mark.parametrize("sc", SCENARIOS, ids=[s["name"] for s in SCENARIOS])
@pytest.mark.parametrize("system", ["priority", "block_both", "soft"])
def test_collision_scenarios(sc, system):
    gc = orc.GridConfig(map=clean_map(sc["map"]), obs_radius=2, collision_system=system, on_target="nothing", seed=0)
    env = orc.pogema_v0(gc)
    env.reset()
    env.step(sc["actions"])
    got = [tuple(p) for p in env.unwrapped.grid.get_agents_xy(ignore_borders=True)]
    assert got == sc["expect"][system]


def test_finished_agent_vanishes_and_frees_its_cell():
    gc = orc.GridConfig(map="aAb..\n.....\n....B", obs_radius=2, on_target="finish", seed=0)
    for system in ("priority", "block_both", "soft"):
        gc.collision_system = system
        env = orc.pogema_v0(gc)
        env.reset()
        obs, rew, term, trunc, info = env.step([4, 0])          # a steps onto its target
        assert rew == [1.0, 0.0] and term == [True, False] and trunc == [False, False]
        assert [i["is_active"] for i in info] == [False, True]
        # b (at (0,2)) looks left: the finished agent is gone from the agents channel, b itself is at the centre
        assert obs[1][1, 2, 1] == 0.0 and obs[1][1, 2, 2] == 1.0
        # the finished agent no longer sees itself
        assert obs[0][1, 2, 2] == 0.0
        obs, rew, term, trunc, info = env.step([0, 3])          # b walks onto the freed cell
        assert env.unwrapped.grid.get_agents_xy(ignore_borders=True)[1] == [0, 1]
        assert rew == [0.0, 0.0] and term == [True, False]       # terminated stays True, reward only once


def test_border_ring_and_target_projection():
    gc = orc.GridConfig(map="a..\n...\n..A", obs_radius=2, seed=0)
    env = orc.pogema_v0(gc)
    obs, _ = env.reset()
    exp_obst = np.array([[0, 0, 0, 0, 0],
                         [0, 1, 1, 1, 1],
                         [0, 1, 0, 0, 0],
                         [0, 1, 0, 0, 0],
                         [0, 1, 0, 0, 0]], dtype=np.float32)
    assert obs[0].shape == (3, 5, 5) and obs[0].dtype == np.float32
    assert np.array_equal(obs[0][0], exp_obst)
    exp_agents = np.zeros((5, 5), np.float32)
    exp_agents[2, 2] = 1
    assert np.array_equal(obs[0][1], exp_agents)
    exp_tgt = np.zeros((5, 5), np.float32)
    exp_tgt[4, 4] = 1                                            # target 2 cells down-right: inside the window
    assert np.array_equal(obs[0][2], exp_tgt)
    # far target: clamped to the window edge/corner
    gc = orc.GridConfig(map="a......\n.......\n.......\n.......\n.......\n......A", obs_radius=2, seed=0)
    env = orc.pogema_v0(gc)
    obs, _ = env.reset()
    assert obs[0][2].sum() == 1 and obs[0][2][4, 4] == 1
    gc = orc.GridConfig(map="A......\n.......\n.......\n...a...", obs_radius=2, seed=0)
    obs, _ = orc.pogema_v0(gc).reset()
    assert obs[0][2][0, 0] == 1                                  # 3 up, 3 left -> clamped to (-2,-2) -> corner


def test_time_limit_and_metrics():
    gc = orc.GridConfig(map="a....\n.....\n....A", obs_radius=2, max_episode_steps=3, seed=0)
    env = orc.pogema_v0(gc)
    env.reset()
    for t in range(3):
        obs, rew, term, trunc, info = env.step([0])
        assert trunc == [t == 2] and term == [False]
    assert info[0]["metrics"] == {"ISR": 0.0, "CSR": 0.0, "ep_length": 3.0}
    # cooperative finish: reward only when everybody stands on its goal at once
    gc = orc.GridConfig(map="aA.\nbB.", obs_radius=2, on_target="nothing", max_episode_steps=8, seed=0)
    env = orc.pogema_v0(gc)
    env.reset()
    _, rew, term, _, _ = env.step([4, 0])
    assert rew == [0.0, 0.0] and term == [False, False]
    _, rew, term, _, info = env.step([0, 4])
    assert rew == [1.0, 1.0] and term == [True, True]
    # SumOfCostsAndMakespanMetric: a has stood on its goal since step 0, b since step 1 -> SoC = 0 + 1 + 2 agents
    assert info[0]["metrics"] == {"ISR": 1.0, "CSR": 1.0, "ep_length": 2, "SoC": 3, "makespan": 2}


def test_sum_of_costs_counts_the_final_stay():
    """a arrives, overshoots, returns at step 2; b stands on its goal from step 0 and steps off on the finishing step
    (upstream keeps the start of that stay); c never arrives and costs the last step."""
    gc = orc.GridConfig(map="aA..\nbB..\nc..C", obs_radius=2, on_target="nothing", max_episode_steps=5, seed=0)
    env = orc.pogema_v0(gc)
    env.reset()
    for acts in ([4, 4, 0], [4, 0, 0], [3, 0, 0], [0, 0, 0], [0, 4, 0]):
        _, _, term, trunc, info = env.step(acts)
    assert trunc == [True] * 3 and term == [False] * 3
    assert info[0]["metrics"]["SoC"] == (2 + 0 + 4) + 3 and info[0]["metrics"]["makespan"] == 5
    assert info[0]["metrics"]["ISR"] == 1 / 3


def test_lifelong_new_target_comes_from_the_agents_generator():
    gc = orc.GridConfig(map="aA...\n.....\n.....", obs_radius=2, on_target="restart", max_episode_steps=8, seed=5)
    env = orc.pogema_v0(gc)
    env.reset()
    seeds = np.random.default_rng(5).integers(np.iinfo(np.int32).max, size=1)
    g = np.random.default_rng(int(seeds[0]))
    comp = [(x + 2, y + 2) for x in range(3) for y in range(5)]   # one component, row-major, padded coords
    exp = tuple(int(v) for v in tuple(*g.choice(comp, 1)))
    _, rew, term, _, _ = env.step([4])
    assert rew == [1.0] and term == [False]
    assert env.unwrapped.grid.finishes_xy[0] == exp


def test_density_and_runtime_wrappers_of_the_package_on_the_oracle_env():
    """pogema_b200's host-side metric wrappers are plain Python over the env API: wrapped around the ORACLE env they must
    report what the oracle's own restatement of upstream's wrappers reports (no GPU, no engine involved)."""
    from pogema_b200.wrappers import AgentsDensityWrapper, RuntimeMetricWrapper
    kw = dict(size=8, density=0.2, num_agents=6, obs_radius=2, max_episode_steps=7, seed=3, observation_type="POMAPF")
    a = RuntimeMetricWrapper(AgentsDensityWrapper(orc.pogema_v0(orc.GridConfig(**kw))))
    b = orc.RuntimeMetricWrapper(orc.AgentsDensityWrapper(orc.pogema_v0(orc.GridConfig(**kw))))
    a.reset(), b.reset()
    rng = np.random.default_rng(0)
    seen = 0
    for t in range(30):
        acts = [int(v) for v in rng.integers(0, 5, size=6)]
        ra, rb = a.step(acts), b.step(acts)
        ma, mb = ra[4][0].get("metrics"), rb[4][0].get("metrics")
        assert (ma is None) == (mb is None)
        if ma is not None:
            ma, mb = dict(ma), dict(mb)
            assert 0.0 <= ma.pop("runtime") < 5.0 and 0.0 <= mb.pop("runtime") < 5.0
            assert ma == mb and 0.0 < ma["avg_agents_density"] <= 1.0
            seen += 1
            a.reset(), b.reset()
    assert seen >= 3
    with pytest.raises(TypeError):
        AgentsDensityWrapper(orc.pogema_v0(orc.GridConfig(size=8, num_agents=2, seed=0))).reset()


def test_two_recollections_of_the_soft_rule_agree():
    """SURVEY.md section 9 item 1 (LOW confidence upstream): a second, independently recalled form of `soft` /
    `_revert_action` (tools/prototypes/soft_variants.py: obstacle test folded into the vertex pass, only the first
    follower reverted per recursion, stays taking part in the swap test) gives the oracle's result on random crowded
    scenarios (93 000 scenarios when run as a script)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "prototypes", "soft_variants.py")
    spec = importlib.util.spec_from_file_location("soft_variants", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    moved, cancelled = mod.fuzz(600, seed=11)
    assert moved > 2000 and cancelled > 2000
