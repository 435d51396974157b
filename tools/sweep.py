"""Sweep team size / teams per CTA for a configuration (development aid)."""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig
from pogema_b200.maps import maze_map, warehouse_map

CONFIGS = {
    "c2": dict(n=4096, gc=dict(size=32, density=0.3, num_agents=64, obs_radius=5, collision_system="priority", on_target="finish")),
    "c3": dict(n=1024, gc=dict(map="maze64", num_agents=256, obs_radius=5, collision_system="soft", on_target="restart")),
    "c4": dict(n=512, gc=dict(map="wh256", num_agents=1024, obs_radius=5, collision_system="block_both", on_target="finish")),
    "c5r3": dict(n=2048, gc=dict(size=32, density=0.3, num_agents=64, obs_radius=3, collision_system="priority", on_target="finish")),
    "c5r5": dict(n=2048, gc=dict(size=32, density=0.3, num_agents=64, obs_radius=5, collision_system="priority", on_target="finish")),
    "c5r7": dict(n=2048, gc=dict(size=32, density=0.3, num_agents=64, obs_radius=7, collision_system="priority", on_target="finish")),
}
ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="c2")
ap.add_argument("--teams", default="0")
ap.add_argument("--tpcs", default="0")
ap.add_argument("--many", type=int, default=16)
ap.add_argument("--steps", type=int, default=512)
a = ap.parse_args()
c = CONFIGS[a.cfg]
kw = dict(c["gc"])
if kw.get("map") == "maze64": kw["map"] = maze_map(64, 3).tolist()
if kw.get("map") == "wh256": kw["map"] = warehouse_map(256).tolist()
gc = GridConfig(max_episode_steps=64, **kw)
r = gc.obs_radius; D = 2 * r + 1; A = gc.num_agents
h, w = gc.map_shape(); P2 = (h + 2 * r) * (w + 2 * r)
bpa = 3 * D * D + 21 + ((P2 + 7) // 8) / A
for team in [int(t) for t in a.teams.split(",")]:
    for tpc in [int(t) for t in a.tpcs.split(",")]:
        if tpc: os.environ["PGM_TPC"] = str(tpc)
        else: os.environ.pop("PGM_TPC", None)
        try:
            env = BatchedPogema(gc, num_envs=c["n"], auto_reset=True, team_threads=team)
        except Exception as e:
            print(a.cfg, team, tpc, "ERR", str(e)[:100]); continue
        env.reset()
        K = a.many
        acts = torch.stack([env.sample_actions() for _ in range(K)])
        ring = torch.stack([env.new_obs_buffer() for _ in range(4)])
        for _ in range(3): env.rollout(acts, obs_out=ring)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, a.steps // K)
        e0.record()
        for _ in range(reps): env.rollout(acts, obs_out=ring)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * K)
        rate = c["n"] * A / (ms * 1e-3)
        p = env.engine.plan()
        print("%-5s team %4d tpc %2d cta %4d smem %6d grid %5d : %7.2f us/step %6.2f G/s frac %.3f" % (
            a.cfg, p["team_threads"], p["teams_per_cta"], p["cta_threads"], p["smem_bytes_per_cta"], p["grid"],
            ms * 1e3, rate / 1e9, rate * bpa / 6541.5e9), flush=True)
        env.close(); del env
