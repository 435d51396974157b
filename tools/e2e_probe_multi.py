"""pgm_step_host under N ranks on one host (torchrun): where does a host-buffer step go when the ranks share the
host's cores, DRAM and PCIe?  Per thread count: every rank's step time and the packed transport's timeline
(enqueued / first chunk / last chunk / widened / returned, microseconds, median), plus the concurrent fill ceiling."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pogema_b200 import BatchedPogema, GridConfig
from pogema_b200 import _native as nat

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
N, A = 4096, 64
gc = GridConfig(size=32, density=0.3, num_agents=A, obs_radius=5, max_episode_steps=64, collision_system="priority", on_target="finish")
env = BatchedPogema(gc, num_envs=N, device=dev, seeds=np.arange(rank * N, (rank + 1) * N, dtype=np.uint64), auto_reset=True)
env.reset(); e = env.engine
h_act = [torch.randint(0, 5, (N, A), dtype=torch.uint8).pin_memory() for _ in range(4)]
h_obs = torch.empty(e.obs_shape(), dtype=torch.uint8).pin_memory()
h_rew = torch.empty((N, A), dtype=torch.float32).pin_memory(); h_te = torch.empty((N, A), dtype=torch.uint8).pin_memory(); h_tr = torch.empty((N, A), dtype=torch.uint8).pin_memory()
cores = os.cpu_count() or 1
aff = len(os.sched_getaffinity(0))
res = {"rank": rank, "cpu_count": cores, "affinity": aff}
for mode, T in [("packed", max(1, cores // world)), ("packed", max(1, cores // world // 2)), ("packed", max(1, 2 * cores // world)), ("plain", 1)]:
    e.set_host_transport(mode, T)
    for i in range(8): e.step_host(h_act[i % 4].numpy(), h_obs.numpy(), h_rew.numpy(), h_te.numpy(), h_tr.numpy())
    barrier(); tl = []; t0 = time.perf_counter()
    for i in range(32):
        e.step_host(h_act[i % 4].numpy(), h_obs.numpy(), h_rew.numpy(), h_te.numpy(), h_tr.numpy())
        tl.append(list(e.host_transport_info()["timeline_us"].values()))
    dt = (time.perf_counter() - t0) / 32
    barrier()
    fill = float(nat.load().pgm_host_fill_gbps(h_obs.data_ptr(), h_obs.numel(), T, 4)) if mode == "packed" else None
    barrier()
    res[f"{mode}_{T}"] = {"ms_per_step": round(dt * 1e3, 3), "timeline_us": np.median(np.array(tl), axis=0).round(0).tolist(), "fill_GBps_this_rank": fill}
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    for r in out: print(json.dumps(r))
    for k in [k for k in out[0] if k.startswith(("packed", "plain"))]:
        ms = max(r[k]["ms_per_step"] for r in out)
        fills = [r[k]["fill_GBps_this_rank"] for r in out if r[k]["fill_GBps_this_rank"]]
        print(k, "slowest rank %.3f ms/step -> %.1f M agent-steps/s in total" % (ms, world * N * A / ms / 1e3), "| widened bytes/s %.1f GB/s | concurrent fill %.1f GB/s" % (world * e.obs_bytes / ms / 1e6, sum(fills)) if fills else "")
if world > 1: dist.destroy_process_group()
