"""Host-side wrappers: PersistentWrapper history / step_back and the SVG
AnimationMonitor on top of the CUDA list env - recorded states must equal the oracle's trajectory, and a
step undone and redone must reproduce it exactly (device state incl. lifelong generators restored)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from oracle import pogema_oracle as orc
from tests.helpers import grid_snapshot

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ot,coll", [("finish", "priority"), ("restart", "soft"), ("nothing", "block_both")])
def test_persistent_history_and_step_back(ot, coll):
    from pogema_b200 import GridConfig, PersistentWrapper, pogema_v0
    kw = dict(size=10, density=0.25, num_agents=6, obs_radius=3, max_episode_steps=30, collision_system=coll,
              on_target=ot, seed=7)
    env = pogema_v0(GridConfig(persistent=True, **kw))
    assert isinstance(env, PersistentWrapper)
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    obs, _ = env.reset()
    robs, _ = ref.reset()
    rng = np.random.default_rng(3)
    acts = [rng.integers(0, 5, size=kw["num_agents"]) for _ in range(20)]
    traj = [grid_snapshot(ref)]
    outs = []
    for a in acts:
        out = ref.step(list(a))
        outs.append(out)
        traj.append(grid_snapshot(ref))
        if all(out[2]) or all(out[3]):
            break
    n = len(outs)
    for t in range(n):
        o, r, te, tr, info = env.step(list(acts[t]))
        assert np.array_equal(np.stack(o), np.stack(outs[t][0])) and r == list(outs[t][1])
    hist = env.get_history()
    assert len(hist) == kw["num_agents"] and all(len(h) == n + 1 for h in hist)
    for t in range(n + 1):
        pos, tgt, act = traj[t]
        for i, h in enumerate(hist):
            s = h[t]
            assert (s.x, s.y) == tuple(pos[i]) and (s.tx, s.ty) == tuple(tgt[i]) and s.active == bool(act[i]) and s.step == t
    # undo three steps, redo them: identical observations / rewards / flags (lifelong generators restored too)
    back = min(3, n)
    for _ in range(back):
        assert env.step_back()
    assert all(len(h) == n + 1 - back for h in env.get_history())
    for t in range(n - back, n):
        o, r, te, tr, info = env.step(list(acts[t]))
        assert np.array_equal(np.stack(o), np.stack(outs[t][0]))
        assert r == list(outs[t][1]) and te == list(outs[t][2]) and tr == list(outs[t][3])
    for _ in range(n):
        assert env.step_back()
    assert not env.step_back()
    assert all(len(h) == 1 for h in env.get_history())


def test_animation_monitor_writes_svg(tmp_path):
    from pogema_b200 import AnimationConfig, AnimationMonitor, GridConfig, pogema_v0
    gc = GridConfig(size=8, density=0.2, num_agents=4, obs_radius=2, max_episode_steps=12, seed=5)
    env = AnimationMonitor(pogema_v0(gc), AnimationConfig(directory=str(tmp_path) + "/", save_every_idx_episode=1))
    steps = 0
    env.reset()
    while True:
        _, _, term, trunc, _ = env.step(env.sample_actions())
        steps += 1
        if all(term) or all(trunc):
            break
    path = tmp_path / "pogema-ep00000-seed5.svg"
    assert path.exists()
    root = ET.parse(path).getroot()
    ns = "{http://www.w3.org/2000/svg}"
    agents = [c for c in root.iter(ns + "circle") if c.get("class") == "a"]
    targets = [c for c in root.iter(ns + "circle") if c.get("class") == "t"]
    assert len(agents) == 4 and len(targets) == 4
    for c in agents:
        anims = {a.get("attributeName"): a.get("values").split(";") for a in c.iter(ns + "animate")}
        assert len(anims["cx"]) == steps + 1 and len(anims["cy"]) == steps + 1
    obstacles = env.unwrapped.grid.get_obstacles(ignore_borders=True)
    rects = list(root.iter(ns + "rect"))
    assert len(rects) == int(obstacles.sum()) + 4 * 8 + 4      # map obstacles + the wall ring
    # on demand, static and egocentric variants
    env.reset()
    env.step(env.sample_actions())
    p2 = env.save_animation(str(tmp_path / "static.svg"), AnimationConfig(static=True, show_lines=True))
    assert not list(ET.parse(p2).getroot().iter(ns + "animate"))
    p3 = env.save_animation(str(tmp_path / "ego.svg"), AnimationConfig(egocentric_idx=0, show_border=False))
    assert len(list(ET.parse(p3).getroot().iter(ns + "circle"))) == 8


def test_auto_reset_wrapper_and_persistent_auto_reset():
    from pogema_b200 import AutoResetWrapper, GridConfig, PersistentWrapper, pogema_v0
    kw = dict(size=8, density=0.2, num_agents=3, obs_radius=2, max_episode_steps=5, seed=4)
    a = AutoResetWrapper(pogema_v0(GridConfig(**kw)))
    b = PersistentWrapper(pogema_v0(GridConfig(auto_reset=True, **kw)))
    a.reset(), b.reset()
    for t in range(12):
        act = a.sample_actions()
        oa, ra, ta, tra, _ = a.step(act)
        ob, rb, tb, trb, _ = b.step(act)
        assert np.array_equal(np.stack(oa), np.stack(ob)) and ra == rb and ta == tb and tra == trb
        assert len(b.get_history()[0]) == (t + 1) % 5 + 1


def test_single_agent_gym_matches_oracle():
    from pogema_b200 import GridConfig, make_single_agent_gym
    kw = dict(size=8, density=0.2, num_agents=1, obs_radius=3, max_episode_steps=20, seed=11)
    env = make_single_agent_gym(GridConfig(**kw))
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    o, info = env.reset()
    ro, rinfo = ref.reset()
    assert o.shape == (3, 7, 7) and np.array_equal(o, ro[0]) and info == rinfo[0]
    assert env.reset(return_info=False).shape == (3, 7, 7)
    rng = np.random.default_rng(0)
    for t in range(20):
        a = int(rng.integers(0, 5))
        o, r, te, tr, info = env.step(a)
        ro, rr, rte, rtr, rinfo = ref.step([a])
        assert np.array_equal(o, ro[0]) and r == rr[0] and te == rte[0] and tr == rtr[0]
        assert info["is_active"] == rinfo[0]["is_active"]
        if te or tr:
            assert info["metrics"] == rinfo[0]["metrics"]
            break
    multi = make_single_agent_gym(GridConfig(size=8, density=0.1, num_agents=3, obs_radius=2, seed=1))
    o, _ = multi.reset()
    o, r, te, tr, info = multi.step(0)               # the other two agents act at random
    assert o.shape == (3, 5, 5) and isinstance(r, float)


def test_sample_factory_wrapper_stack():
    from pogema_b200 import AutoResetWrapper, GridConfig, pogema_v0
    kw = dict(size=8, density=0.2, num_agents=3, obs_radius=2, max_episode_steps=6, seed=4)
    env = pogema_v0(GridConfig(integration="SampleFactory", **kw))
    assert isinstance(env, AutoResetWrapper) and env.is_multiagent and env.num_agents == 3
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    env.reset(), ref.reset()
    ends = 0
    for t in range(20):
        act = ref.sample_actions()
        o, r, te, tr, infos = env.step(act)
        ro, rr, rte, rtr, rinfos = ref.step(act)
        assert r == list(rr) and te == list(rte) and tr == list(rtr)
        if all(rte) or all(rtr):
            ends += 1
            assert infos[0]["episode_extra_stats"] == rinfos[0]["metrics"] == infos[0]["metrics"]
            ro, _ = ref.reset()                      # the stack auto-resets: the observation is the reset one
        assert np.array_equal(np.stack(o), np.stack(ro))
    assert ends >= 3
