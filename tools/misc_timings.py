"""Miscellaneous timings for the docs: reseeding auto-reset overhead, single-instance list API latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig, pogema_v0

gc = GridConfig(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64)
for mode in (True, "reseed"):
    env = BatchedPogema(gc, num_envs=4096, auto_reset=mode)
    env.reset()
    acts = [env.sample_actions() for _ in range(16)]
    bufs = [env.new_obs_buffer() for _ in range(4)]
    for i in range(70): env.step(acts[i % 16], out=bufs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 640
    for i in range(n): env.step(acts[i % 16], out=bufs[i % 4])
    e1.record(); torch.cuda.synchronize()
    print("auto_reset=%r: %.2f us/step (python-launched, 10 episode ends of all 4096 instances inside)" % (mode, e0.elapsed_time(e1) / n * 1e3))
    env.check_errors()
env = pogema_v0(GridConfig(size=8, num_agents=4, seed=0))
env.reset()
t0 = time.perf_counter()
for i in range(300):
    o, r, te, tr, inf = env.step(env.sample_actions())
    if all(te) or all(tr): env.reset()
print("pogema_v0 (8x8, 4 agents) list API: %.1f us per step" % ((time.perf_counter() - t0) / 300 * 1e6))
