"""Quick device-resident timing of the step kernel (development aid, not the bench contract)."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--size", type=int, default=32)
ap.add_argument("--agents", type=int, default=64)
ap.add_argument("--r", type=int, default=5)
ap.add_argument("--coll", default="priority")
ap.add_argument("--ot", default="finish")
ap.add_argument("--team", type=int, default=0)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--fmt", default="u8")
ap.add_argument("--noobs", action="store_true")
ap.add_argument("--map", default="random", choices=["random", "maze", "warehouse"])
ap.add_argument("--max-steps", type=int, default=64)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--many", type=int, default=0, help="steps per launch (pgm_step_many)")
ap.add_argument("--sets", type=int, default=0, help="with --many: rotate over this many sets of action / reward / flag tensors (the memory footprint of a longer rollout at the same steps per launch)")
ap.add_argument("--ring", type=int, default=4, help="observation buffers written in turn (1: the same buffer every step - it may stay in L2)")
a = ap.parse_args()
from pogema_b200.maps import maze_map, warehouse_map
mp = None if a.map == "random" else (maze_map(a.size, 3) if a.map == "maze" else warehouse_map(a.size)).tolist()
gc = GridConfig(size=a.size, density=0.3, num_agents=a.agents, obs_radius=a.r, max_episode_steps=a.max_steps,
                collision_system=a.coll, on_target=a.ot, map=mp)
t0 = time.time()
env = BatchedPogema(gc, num_envs=a.n, auto_reset=True, team_threads=a.team, obs_format=a.fmt)
t_gen = time.time() - t0
env.reset()
acts = [env.sample_actions() for _ in range(16)]
bufs = [env.new_obs_buffer() for _ in range(a.ring)]
for i in range(20):
    env.step(acts[i % 16], out=bufs[i % a.ring], compute_obs=not a.noobs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if a.many:
    K = a.many
    act_k = torch.stack([acts[i % 16] for i in range(K)])
    ring = torch.stack(bufs)
    for _ in range(3):
        env.rollout(act_k, obs_out=ring, compute_obs=not a.noobs)
    torch.cuda.synchronize()
    reps = max(1, a.steps // K)
    a.steps = reps * K
    if a.sets:
        N, A = env.num_envs, env.num_agents
        sets = [(act_k.clone(), torch.empty((K, N, A), dtype=torch.float32, device="cuda"),
                 torch.empty((K, N, A), dtype=torch.bool, device="cuda"), torch.empty((K, N, A), dtype=torch.bool, device="cuda"))
                for _ in range(a.sets)]
        sp = int(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            ac, rw, te, tr = sets[i % a.sets]
            env.engine.step_many(K, ac.data_ptr(), 1, ring.data_ptr(), a.ring, rw.data_ptr(), te.data_ptr(), tr.data_ptr(), sp)
        e1.record()
    else:
        e0.record()
        for _ in range(reps):
            env.rollout(act_k, obs_out=ring, compute_obs=not a.noobs)
        e1.record()
elif a.graph:
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(a.graph):
                env.step(acts[i % 16], out=bufs[i % a.ring], compute_obs=not a.noobs)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    reps = max(1, a.steps // a.graph)
    a.steps = reps * a.graph
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
else:
    e0.record()
    for i in range(a.steps):
        env.step(acts[i % 16], out=bufs[i % a.ring], compute_obs=not a.noobs)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
env.check_errors()
D = 2 * a.r + 1
P = a.size + 2 * a.r
bytes_per = ({"u8": 3 * D * D, "f16": 6 * D * D, "f32": 12 * D * D}.get(a.fmt, 4 * ((3 * D * D + 31) // 32))) + 21 + ((P * P + 7) // 8) / a.agents
if a.noobs:
    bytes_per = 21 + ((P * P + 7) // 8) / a.agents
rate = a.n * a.agents / (ms * 1e-3)
print(json.dumps({"cfg": vars(a), "plan": env.engine.plan(), "gen_s": round(t_gen, 2), "ms_per_step": round(ms, 4),
                  "agent_steps_per_s": rate, "GBps": rate * bytes_per / 1e9, "frac_6541": rate * bytes_per / 6541.5e9}))
