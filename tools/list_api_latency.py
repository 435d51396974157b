"""Latency of one step through the upstream-style list API (pogema_v0, configs[0]: one 8x8 instance, 4 agents)
next to the numpy oracle on the same host, and the parts of it (development aid)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pogema_b200 import GridConfig, pogema_v0
from oracle import pogema_oracle as orc
kw = dict(size=8, density=0.3, num_agents=4, obs_radius=5, max_episode_steps=64, seed=0)
for name, env in (("cuda list env", pogema_v0(GridConfig(**kw))), ("oracle (numpy)", orc.pogema_v0(orc.GridConfig(**kw)))):
    env.reset()
    for i in range(50): env.step(env.sample_actions())
    env.reset()
    t0 = time.perf_counter(); n = 2000
    for i in range(n):
        o, r, te, tr, inf = env.step(env.sample_actions())
        if all(te) or all(tr): env.reset()
    print("%s: %.1f us per step" % (name, (time.perf_counter() - t0) / n * 1e6))
env = pogema_v0(GridConfig(**kw)); env.reset()
e = env._engine
import ctypes as C
act = np.zeros((1, 4), np.uint8)
t0 = time.perf_counter()
for i in range(2000): e.step_host(act, env._h_obs, env._h_rew, env._h_term, env._h_trunc)
print("step_host alone: %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
from pogema_b200 import _native as nat
t0 = time.perf_counter()
for i in range(2000): e.get_state(nat.STATE_ACTIVE)
print("get_state(ACTIVE): %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
t0 = time.perf_counter()
for i in range(2000): e.get_state(nat.STATE_WAS_ON_GOAL)
print("get_state(WAS_ON_GOAL): %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
t0 = time.perf_counter()
for i in range(2000): env._obs_list(env._h_obs); env._get_infos(env._h_active[0])
print("python list building + infos: %.1f us" % ((time.perf_counter() - t0) / 2000 * 1e6))
