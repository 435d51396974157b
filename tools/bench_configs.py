"""All BASELINE.json GPU configurations with the engine's automatic plan: K steps per launch and one launch per
step (CUDA graph), as JSON lines (profiles/r01_configs.json)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig
from pogema_b200.maps import maze_map, warehouse_map

CONFIGS = [
    ("configs[1] 4096 x 32x32, 64 agents, r=5, priority/finish", 4096, dict(size=32, density=0.3, num_agents=64, obs_radius=5, collision_system="priority", on_target="finish")),
    ("configs[2] 1024 x 64x64 maze, 256 agents, r=5, soft/restart", 1024, dict(map="maze64", num_agents=256, obs_radius=5, collision_system="soft", on_target="restart")),
    ("configs[3] 512 x 256x256 warehouse, 1024 agents, r=5, block_both/finish", 512, dict(map="wh256", num_agents=1024, obs_radius=5, collision_system="block_both", on_target="finish")),
    ("configs[4] per-GPU share 2048 x 32x32, 64 agents, r=3", 2048, dict(size=32, density=0.3, num_agents=64, obs_radius=3, collision_system="priority", on_target="finish")),
    ("configs[4] per-GPU share 2048 x 32x32, 64 agents, r=5", 2048, dict(size=32, density=0.3, num_agents=64, obs_radius=5, collision_system="priority", on_target="finish")),
    ("configs[4] per-GPU share 2048 x 32x32, 64 agents, r=7", 2048, dict(size=32, density=0.3, num_agents=64, obs_radius=7, collision_system="priority", on_target="finish")),
    ("1M agents on one GPU: 16384 x 32x32, 64 agents, r=5", 16384, dict(size=32, density=0.3, num_agents=64, obs_radius=5, collision_system="priority", on_target="finish")),
]
PEAK = 6541.5e9
K = 16
for name, n, kw in CONFIGS:
    kw = dict(kw)
    if kw.get("map") == "maze64": kw["map"] = maze_map(64, 3).tolist()
    if kw.get("map") == "wh256": kw["map"] = warehouse_map(256).tolist()
    gc = GridConfig(max_episode_steps=64, **kw)
    r, A = gc.obs_radius, gc.num_agents
    D = 2 * r + 1
    h, w = gc.map_shape()
    bpa = 3 * D * D + 21 + (((h + 2 * r) * (w + 2 * r) + 7) // 8) / A
    env = BatchedPogema(gc, num_envs=n, auto_reset=True)
    env.reset()
    acts = torch.stack([env.sample_actions() for _ in range(K)])
    nring = 4 if env.engine.obs_bytes * 4 < 8e9 else 2
    ring = torch.stack([env.new_obs_buffer() for _ in range(nring)])

    def timed(fn, reps):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * K)

    reps = max(2, int(4096 // K * 262144 / (n * A)))
    ms_many = timed(lambda: env.rollout(acts, obs_out=ring), reps)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(K): env.step(acts[i], out=ring[i % nring])
    torch.cuda.current_stream().wait_stream(side)
    ms_closed = timed(g.replay, reps)
    env.check_errors()
    rate = lambda ms: n * A / (ms * 1e-3)
    print(json.dumps({"config": name, "instances": n, "agents": A, "plan": env.engine.plan(),
                      "bytes_per_agent_step": round(bpa, 2),
                      "steps_per_launch_16": {"us_per_step": round(ms_many * 1e3, 2), "agent_steps_per_s": rate(ms_many),
                                               "frac_of_measured_hbm": round(rate(ms_many) * bpa / PEAK, 3)},
                      "one_launch_per_step": {"us_per_step": round(ms_closed * 1e3, 2), "agent_steps_per_s": rate(ms_closed),
                                              "frac_of_measured_hbm": round(rate(ms_closed) * bpa / PEAK, 3)}}), flush=True)
    env.close(); del env, ring, acts, g
    torch.cuda.empty_cache()
