"""Closed loop with G groups of instances (the double-buffered sampling of RL frameworks: while the policy works on the
observations of one group, the other group steps).  Every group is its own engine on its own stream and advances with
ONE LAUNCH PER STEP (a CUDA graph of 16 single-step launches per group, replayed); a group's step t+1 starts only
after its step t has completed, but the groups overlap each other: one group's dependent front (state loads, move
resolution, first bit assembly) runs while the other group's observation stores keep HBM busy.
    python tools/two_groups.py [--groups 2] [--n 4096 ...]   -> one JSON line (whole-job rate over all groups)"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig

ap = argparse.ArgumentParser()
ap.add_argument("--groups", type=int, default=2)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--size", type=int, default=32)
ap.add_argument("--agents", type=int, default=64)
ap.add_argument("--r", type=int, default=5)
ap.add_argument("--coll", default="priority")
ap.add_argument("--ot", default="finish")
ap.add_argument("--steps", type=int, default=2048)
ap.add_argument("--map", default="random", choices=["random", "maze", "warehouse"])
ap.add_argument("--ring", type=int, default=4)
a = ap.parse_args()
from pogema_b200.maps import maze_map, warehouse_map
mp = None if a.map == "random" else (maze_map(a.size, 3) if a.map == "maze" else warehouse_map(a.size)).tolist()
gc = GridConfig(size=a.size, density=0.3, num_agents=a.agents, obs_radius=a.r, max_episode_steps=64,
                collision_system=a.coll, on_target=a.ot, map=mp)
G = a.groups
n_g = a.n // G
groups = []
for k in range(G):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        env = BatchedPogema(gc, num_envs=n_g, seeds=list(range(k * n_g, (k + 1) * n_g)), auto_reset=True)
        env.reset()
        acts = [env.sample_actions() for _ in range(16)]
        bufs = [env.new_obs_buffer() for _ in range(a.ring)]
        for i in range(20):
            env.step(acts[i % 16], out=bufs[i % a.ring])
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for i in range(16):
                env.step(acts[i], out=bufs[i % a.ring])
    groups.append((s, env, g, acts, bufs))
torch.cuda.synchronize()
for s, env, g, _, _ in groups:
    with torch.cuda.stream(s):
        g.replay()
torch.cuda.synchronize()
reps = max(1, a.steps // 16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = torch.cuda.current_stream()
# keep the GPU busy while the CPU enqueues, so that the events bracket device time
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
flush.fill_(1)
e0.record(main)
for s, *_ in groups:
    s.wait_event(e0)
for _ in range(reps):
    for s, env, g, _, _ in groups:
        with torch.cuda.stream(s):
            g.replay()
for s, *_ in groups:
    main.wait_stream(s)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (reps * 16)
for _, env, *_ in groups:
    env.check_errors()
D, P = 2 * a.r + 1, a.size + 2 * a.r
bytes_per = 3 * D * D + 21 + ((P * P + 7) // 8) / a.agents
rate = n_g * G * a.agents / (ms * 1e-3)
print(json.dumps({"groups": G, "instances_per_group": n_g, "steps_per_group": reps * 16, "ms_per_step_of_all_groups": round(ms, 5),
                  "agent_steps_per_s": rate, "frac_6541": rate * bytes_per / 6541.5e9,
                  "plan": groups[0][1].engine.plan().get("fast")}))
