"""SoC / makespan (upstream wrappers/metrics.py :: SumOfCostsAndMakespanMetric, on_target='nothing'): the per-agent
costs the step kernels latch at the end of an episode against the oracle's wrapper, on both kernels, with policies that
actually reach (and sometimes leave) their goals; RuntimeMetricWrapper."""
import numpy as np
import pytest

import oracle.pogema_oracle as orc

pytestmark = pytest.mark.gpu


def greedy_actions(ref, rng, p_greedy):
    """towards the target with probability p_greedy (agents arrive, stay, get pushed off), else uniform"""
    g = ref.unwrapped.grid
    out = []
    for (x, y), (tx, ty) in zip(g.positions_xy, g.finishes_xy):
        if rng.random() < p_greedy:
            if (x, y) == (tx, ty):
                out.append(0)
            elif abs(tx - x) >= abs(ty - y):
                out.append(2 if tx > x else 1)
            else:
                out.append(4 if ty > y else 3)
        else:
            out.append(int(rng.integers(0, 5)))
    return out


@pytest.mark.parametrize("coll", ["priority", "block_both", "soft"])
@pytest.mark.parametrize("fast,A,r,size,density,max_steps", [(False, 5, 2, 8, 0.1, 14), (True, 16, 3, 8, 0.1, 14),
                                                             (True, 32, 5, 12, 0.1, 14), (True, 16, 2, 16, 0.0, 30)])
def test_soc_and_makespan_match_the_oracle_wrapper(coll, fast, A, r, size, density, max_steps, monkeypatch):
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    monkeypatch.setenv("PGM_FAST", "1" if fast else "0")
    kw = dict(size=size, density=density, num_agents=A, obs_radius=r, max_episode_steps=max_steps,
              collision_system=coll, on_target="nothing")
    n = 6
    seeds = list(range(10, 10 + n))
    env = BatchedPogema(GridConfig(**kw), num_envs=n, seeds=seeds, auto_reset=True)
    assert env.engine.plan()["fast_step_kernel"] == fast
    env.reset()
    refs = [orc.pogema_v0(orc.GridConfig(seed=s, **kw)) for s in seeds]
    for ref in refs:
        ref.reset()
    rng = np.random.default_rng(1)
    ended, early = 0, 0   # episodes ended | of them, with an agent that reached its goal before the last step and stayed
    for t in range(90):
        p = 0.95 if (t // 14) % 2 == 0 else 0.6
        acts = np.array([greedy_actions(ref, rng, p) for ref in refs], dtype=np.uint8)
        env.step(torch.from_numpy(acts).cuda())
        done = env.episode_done.cpu().numpy()
        m = None
        for k, ref in enumerate(refs):
            _, _, te, tr, info = ref.step(list(acts[k]))
            fin = all(te) or all(tr)
            assert bool(done[k]) == fin
            if fin:
                m = m or env.metrics()
                want = info[0]["metrics"]
                assert set(want) == {"ISR", "CSR", "ep_length", "SoC", "makespan"}
                for key, val in want.items():
                    assert m[key][k] == val, (coll, fast, t, k, key, m[key][k], val)
                ended += 1
                early += int(want["SoC"] < A * want["ep_length"])
                ref.reset()
    assert ended >= 12 and early >= 10
    env.check_errors()


def test_list_api_reports_soc_makespan_and_runtime():
    from pogema_b200 import GridConfig, RuntimeMetricWrapper, pogema_v0
    # a reaches A at step 0 and stays; b reaches B at step 1: solve times [0, 1] -> SoC 1 + 2, makespan 2
    kw = dict(map="aA.\nbB.", obs_radius=2, on_target="nothing", max_episode_steps=8, seed=0)
    env = RuntimeMetricWrapper(pogema_v0(GridConfig(**kw)))
    ref = orc.RuntimeMetricWrapper(orc.pogema_v0(orc.GridConfig(**kw)))
    env.reset(), ref.reset()
    for acts in ([4, 0], [0, 4]):
        out, rout = env.step(acts), ref.step(acts)
    m, rm = dict(out[4][0]["metrics"]), dict(rout[4][0]["metrics"])
    assert 0.0 <= m.pop("runtime") < 5.0 and 0.0 <= rm.pop("runtime") < 5.0
    assert m == rm == {"ISR": 1.0, "CSR": 1.0, "ep_length": 2, "SoC": 3, "makespan": 2}
    # an agent that stands on its goal, leaves, and comes back counts from its return; one that steps off on the
    # finishing step keeps the start of its stay (upstream's rule); one that never arrives costs the last step
    kw = dict(map="aA..\nbB..\nc..C", obs_radius=2, on_target="nothing", max_episode_steps=5, seed=0)
    env, ref = pogema_v0(GridConfig(**kw)), orc.pogema_v0(orc.GridConfig(**kw))
    env.reset(), ref.reset()
    for acts in ([4, 4, 0], [4, 0, 0], [3, 0, 0], [0, 0, 0], [0, 4, 0]):
        out, rout = env.step(acts), ref.step(acts)
    assert out[3] == [True] * 3
    assert out[4][0]["metrics"] == rout[4][0]["metrics"]
    assert out[4][0]["metrics"]["SoC"] == (2 + 0 + 4) + 3 and out[4][0]["metrics"]["makespan"] == 5


@pytest.mark.parametrize("kind", ["POMAPF", "MAPF"])
def test_agents_density_wrapper_matches_the_oracle(kind):
    from pogema_b200 import AgentsDensityWrapper, GridConfig, pogema_v0
    kw = dict(size=8, density=0.2, num_agents=6, obs_radius=2, max_episode_steps=7, seed=3, observation_type=kind)
    env = AgentsDensityWrapper(pogema_v0(GridConfig(**kw)))
    ref = orc.AgentsDensityWrapper(orc.pogema_v0(orc.GridConfig(**kw)))
    env.reset(), ref.reset()
    rng = np.random.default_rng(0)
    seen = 0
    for t in range(24):
        acts = [int(v) for v in rng.integers(0, 5, size=6)]
        out, rout = env.step(acts), ref.step(acts)
        assert out[4][0].get("metrics") == rout[4][0].get("metrics"), t
        if "metrics" in out[4][0]:
            assert "avg_agents_density" in out[4][0]["metrics"]
            seen += 1
            env.reset(), ref.reset()
    assert seen >= 3
