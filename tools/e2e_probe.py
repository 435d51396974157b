"""Timeline of pgm_step_host with the packed transport (development aid): where the time of one host-buffer
step goes - enqueue, first/last chunk on the host, widening done, return."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pogema_b200 import BatchedPogema, GridConfig

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--mode", default="packed")
ap.add_argument("--fmt", default="u8")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--pageable", action="store_true")
ap.add_argument("--size", type=int, default=32); ap.add_argument("--agents", type=int, default=64); ap.add_argument("--r", type=int, default=5)
a = ap.parse_args()
gc = GridConfig(size=a.size, density=0.3, num_agents=a.agents, obs_radius=a.r, max_episode_steps=64, collision_system="priority", on_target="finish")
env = BatchedPogema(gc, num_envs=a.n, auto_reset=True, obs_format=a.fmt)
env.reset()
e = env.engine
e.set_host_transport(a.mode, a.threads)
N, A = a.n, a.agents
pin = (lambda t: t) if a.pageable else (lambda t: t.pin_memory())
h_act = [pin(torch.randint(0, 5, (N, A), dtype=torch.uint8)) for _ in range(4)]
h_obs = pin(torch.empty(e.obs_shape(), dtype={"f32": torch.float32, "f16": torch.float16}.get(a.fmt, torch.uint8)))
h_rew = pin(torch.empty((N, A), dtype=torch.float32)); h_te = pin(torch.empty((N, A), dtype=torch.uint8)); h_tr = pin(torch.empty((N, A), dtype=torch.uint8))
tl = []
for i in range(5):
    e.step_host(h_act[i % 4].numpy(), h_obs.numpy(), h_rew.numpy(), h_te.numpy(), h_tr.numpy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(a.steps):
    e.step_host(h_act[i % 4].numpy(), h_obs.numpy(), h_rew.numpy(), h_te.numpy(), h_tr.numpy())
    tl.append(list(e.host_transport_info()["timeline_us"].values()))
dt = (time.perf_counter() - t0) / a.steps
info = e.host_transport_info()
print(json.dumps({"mode": a.mode, "fmt": a.fmt, "threads": info["threads"], "chunks": os.environ.get("PGM_STREAM_CHUNKS", "8"),
                  "spin_us": os.environ.get("PGM_HOST_SPIN_US", "200"), "pageable": a.pageable,
                  "ms_per_step": round(dt * 1e3, 4), "M_agent_steps_per_s": round(N * A / dt / 1e6, 1),
                  "timeline_us_median[enqueued,first,last,widened,returned]": np.median(np.array(tl), axis=0).tolist()}))
