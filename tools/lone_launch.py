"""What does the lone K-step launch of the driver's invocation (bench.py --steps 20) pay for?  The same launch timed
 (a) behind a 384 MB torch fill (bench.py uses 1 GB: same L2 state, longer cover) (L2 ends up full of ORDINARY dirty lines, which outrank the kernel's
     evict-first observation lines for the whole timed region),
 (b) behind a 16-step rollout of a second engine of the same shape into its own buffers (L2 full of evict-first lines of
     foreign buffers: cold for the timed launch, but nothing squats), and
 (c) back to back with itself (steady state).
    python tools/lone_launch.py [--k 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pogema_b200 import BatchedPogema, GridConfig

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=12)
a = ap.parse_args()
gc = GridConfig(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64)
K = a.k


def make(seed0):
    env = BatchedPogema(gc, num_envs=a.n, seeds=list(range(seed0, seed0 + a.n)), auto_reset=True)
    env.reset()
    acts = torch.randint(0, 5, (K, a.n, 64), dtype=torch.uint8, device="cuda")
    ring = torch.stack([env.new_obs_buffer() for _ in range(3)])
    rew = torch.empty((K, a.n, 64), dtype=torch.float32, device="cuda")
    te = torch.empty((K, a.n, 64), dtype=torch.bool, device="cuda")
    tr = torch.empty((K, a.n, 64), dtype=torch.bool, device="cuda")
    sp = int(torch.cuda.current_stream().cuda_stream)
    return lambda k=K: env.engine.step_many(k, acts.data_ptr(), 1, ring.data_ptr(), 3, rew.data_ptr(), te.data_ptr(), tr.data_ptr(), sp)


run, other = make(0), make(100000)
flush = torch.empty(384 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    run(), other(16)
torch.cuda.synchronize()


def timed(before):
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / K)
    ts.sort()
    return ts[len(ts) // 2]


res = {"k": K,
       "behind_torch_fill_us_per_step": timed(lambda: flush.fill_(1)),
       "behind_foreign_rollout_us_per_step": timed(lambda: other(16)),
       "behind_both_fill_then_foreign_us_per_step": timed(lambda: (flush.fill_(1), other(16)))}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run(); e0.record()
for _ in range(32):
    run()
e1.record(); torch.cuda.synchronize()
res["back_to_back_us_per_step"] = e0.elapsed_time(e1) * 1e3 / (32 * K)
print(json.dumps(res))
