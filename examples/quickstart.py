"""Quick start: the three ways to drive the engine (needs a CUDA device and a built libpgm_b200.so).

    python examples/quickstart.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pogema_b200 import BatchedPogema, GridConfig, pogema_v0

# 1. upstream-style single instance, Python lists in and out -------------------------------------------
env = pogema_v0(GridConfig(size=8, density=0.3, num_agents=4, obs_radius=5, seed=0))
obs, infos = env.reset()
while True:
    obs, reward, terminated, truncated, infos = env.step(env.sample_actions())
    if all(terminated) or all(truncated):
        break
print("single instance metrics:", infos[0]["metrics"])

# 2. closed loop on 4096 instances: a (toy) policy consumes the observation tensor on the GPU -----------
cfg = GridConfig(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64)
venv = BatchedPogema(cfg, num_envs=4096, auto_reset="reseed")        # a fresh random map for every episode
policy = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 11 * 11, 5)).cuda().half()
obs = venv.reset()                                                    # uint8 [4096, 64, 3, 11, 11]
total = 0.0
for t in range(128):
    with torch.no_grad():
        logits = policy(obs.view(-1, 3, 11, 11).half())
    actions = logits.argmax(-1).view(4096, 64).to(torch.uint8)
    obs, rewards, terminated, truncated = venv.step(actions)           # one kernel launch
    total += float(rewards.sum())
print("closed loop: %d agent-steps, total reward %.0f, seeds now %s..." % (128 * 4096 * 64, total, venv.current_seeds()[:3]))

# 2b. the same closed loop, double-buffered: two groups of 2048 instances on their own streams - while the policy works on
# one group's observations the other group steps, so one group's dependent front (state loads, move resolution, first
# bit assembly) overlaps the other's observation stores: 0.85 -> 0.96 of the HBM roofline for the step kernels
groups = BatchedPogema.groups(cfg, 4096, groups=2, auto_reset=True)
obs_g, acc = [], []
for genv, gstream in groups:
    with torch.cuda.stream(gstream):
        obs_g.append(genv.reset())
        acc.append(torch.zeros((), device="cuda"))                    # per-group reward sum, accumulated on the group's stream
for t in range(64):
    for k, (genv, gstream) in enumerate(groups):
        with torch.cuda.stream(gstream), torch.no_grad():             # group k's policy pass and step, in stream order
            actions = policy(obs_g[k].view(-1, 3, 11, 11).half()).argmax(-1).view(2048, 64).to(torch.uint8)
            obs_g[k], rewards, terminated, truncated = genv.step(actions)
            acc[k] += rewards.sum()                                   # (no host synchronisation inside the loop)
torch.cuda.synchronize()
print("closed loop in two groups: %d agent-steps, total reward %.0f" % (64 * 4096 * 64, float(acc[0] + acc[1])))

# 3. open loop: K steps per launch for actions known in advance ---------------------------------------------
venv = BatchedPogema(cfg, num_envs=4096)                                # same-task auto reset
venv.reset()
ring = torch.empty((2,) + tuple(venv.engine.obs_shape()), dtype=torch.uint8, device="cuda")   # keep only 2 observations
for launch in range(4):                                                 # 64 steps = one full episode
    actions = torch.randint(0, 5, (16, 4096, 64), dtype=torch.uint8, device="cuda")
    obs_k, rew_k, term_k, trunc_k = venv.rollout(actions, obs_out=ring)
print("rollout:", tuple(rew_k.shape), "rewards in the last 16 steps", float(rew_k.sum()))
print("metrics of the finished episodes:", {k: float(v.mean()) for k, v in venv.metrics().items()})

# 4. host buffers (numpy in, numpy out): the packed transport brings uint8 observations over PCIe as bits ----
import numpy as np
acts = np.random.default_rng(0).integers(0, 5, size=(4096, 64)).astype(np.uint8)
obs_h, rew_h, term_h, trunc_h = venv.step_host(acts)                     # pgm_step_host, 'auto' = packed for big tensors
print("host step:", obs_h.shape, obs_h.dtype, venv.engine.host_transport_info())

# 5. history, undo and an SVG animation of one episode ---------------------------------------------------------
from pogema_b200 import AnimationConfig, AnimationMonitor
env = AnimationMonitor(pogema_v0(GridConfig(size=8, density=0.3, num_agents=4, obs_radius=5, seed=0, max_episode_steps=32)),
                       AnimationConfig(directory="renders/", save_every_idx_episode=None))
env.reset()
for t in range(10):
    env.step(env.sample_actions())
env.env.step_back()                                                     # PersistentWrapper: device state restored
print("history length after 10 steps and one undo:", len(env.env.get_history()[0]), "->", env.save_animation("renders/quickstart.svg"))
