"""Per-phase cycle timeline of the step kernel (clock64 stamps, see pgm_set_debug_buffer)."""
import argparse, sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the stamps are compiled only into the timeline build of the library (`make timeline`)
TL = os.path.join(ROOT, "pogema_b200", "_lib", "libpgm_b200_timeline.so")
if not os.environ.get("PGM_B200_LIB"):
    if not os.path.exists(TL):
        sys.exit("tools/phase_timeline.py needs the timeline build of the library: run `make timeline` first")
    os.environ["PGM_B200_LIB"] = TL
import numpy as np, torch
from pogema_b200 import BatchedPogema, GridConfig

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096); ap.add_argument("--agents", type=int, default=64)
ap.add_argument("--size", type=int, default=32); ap.add_argument("--r", type=int, default=5)
ap.add_argument("--coll", default="priority"); ap.add_argument("--ot", default="finish")
ap.add_argument("--fmt", default="u8")
ap.add_argument("--map", default="random")
ap.add_argument("--team", type=int, default=0)
a = ap.parse_args()
from pogema_b200.maps import maze_map, warehouse_map
mp = None if a.map == 'random' else (maze_map(a.size, 3) if a.map == 'maze' else warehouse_map(a.size)).tolist()
gc = GridConfig(size=a.size, density=0.3, num_agents=a.agents, obs_radius=a.r, collision_system=a.coll, on_target=a.ot, map=mp)
env = BatchedPogema(gc, num_envs=a.n, auto_reset=True, obs_format=a.fmt, team_threads=a.team)
env.reset()
acts = [env.sample_actions() for _ in range(8)]
bufs = [env.new_obs_buffer() for _ in range(4)]
for i in range(20): env.step(acts[i % 8], out=bufs[i % 4])
dbg = torch.zeros((a.n, 16), dtype=torch.int64, device="cuda")
env.engine.lib.pgm_set_debug_buffer(env.engine.handle, C.c_void_p(dbg.data_ptr()))
torch.cuda.synchronize()
for i in range(3): env.step(acts[i % 8], out=bufs[i % 4])
torch.cuda.synchronize()
env.engine.lib.pgm_set_debug_buffer(env.engine.handle, None)
d = dbg.cpu().numpy().astype(np.float64)
names = ["start", "prologue", "pdl_wait", "load+occ+mbar", "resolve", "bookkeeping", "abits+zero", "gen", "expand+store"]
if env.engine.plan().get("fast_step_kernel"):
    names = ["start", "prologue fills", "pdl_wait", "state+actions+publish", "resolve", "bookkeeping+bitmap", "batch0 bits", "batch0 expand+store", "other batches"]
# clock64 is per-SM: only differences within an instance are meaningful
print("plan", env.engine.plan())
prev = d[:, 0]
for k in range(1, 9):
    dt = d[:, k] - d[:, k - 1]
    print("%-16s mean %8.0f  p50 %8.0f  p95 %8.0f  max %8.0f cycles" % (names[k], dt.mean(), np.percentile(dt, 50), np.percentile(dt, 95), dt.max()))
tot = d[:, 8] - d[:, 0]
print("%-16s mean %8.0f  p50 %8.0f  p95 %8.0f  max %8.0f cycles" % ("total", tot.mean(), np.percentile(tot, 50), np.percentile(tot, 95), tot.max()))

# wall-clock view (globaltimer ns) of the LAST launch: when do instances start / pass the dependency wait / finish
t9, t10, t11 = d[:, 9], d[:, 10], d[:, 11]
t0 = t9.min()
print("launch wall clock (us): first CTA start 0.00 | last CTA start %.2f | dependency wait passed: first %.2f last %.2f |"
      " finish: first %.2f  p50 %.2f  p95 %.2f  last %.2f" % ((t9.max() - t0) / 1e3, (t10.min() - t0) / 1e3, (t10.max() - t0) / 1e3,
                                                         (t11.min() - t0) / 1e3, (np.percentile(t11, 50) - t0) / 1e3,
                                                         (np.percentile(t11, 95) - t0) / 1e3, (t11.max() - t0) / 1e3))
work = (t11 - t10) / 1e3
print("per-instance work after the wait (us): mean %.2f p50 %.2f p95 %.2f max %.2f" % (work.mean(), np.percentile(work, 50), np.percentile(work, 95), work.max()))

if env.engine.plan().get("fast_step_kernel"):
    first = (d[:, 13] - t10) / 1e3
    print("first observation bytes ready, after the wait (us): p5 %.2f p50 %.2f p95 %.2f max %.2f" % tuple(np.percentile(first, [5, 50, 95, 100])))
    fin = (t11 - t10.min()) / 1e3
    hist, edges = np.histogram(fin, bins=10)
    print("finish histogram (us after the wait):", " ".join("%.1f:%d" % (edges[i], hist[i]) for i in range(10)))
    sm = (d[:, 12].astype(np.int64) >> 16); wid = d[:, 12].astype(np.int64) & 0xFFFF
    per_sm = np.bincount(sm, minlength=148)
    print("teams per SM: min %d max %d; SMs used %d" % (per_sm[per_sm > 0].min(), per_sm.max(), (per_sm > 0).sum()))
    sub = np.zeros((int(sm.max()) + 1, 4), dtype=np.int64)
    np.add.at(sub, (sm, wid & 3), 1)
    print("warps per (SM, warp slot & 3): mean per slot-class", sub.mean(0).round(2), " worst SM", sub[sub.max(1).argmax()], " slot ids seen", np.unique(wid)[:40])
    # finish time by SM: is the tail specific to some SMs?
    fin_sm = np.array([fin[sm == k].max() if (sm == k).any() else 0 for k in range(int(sm.max()) + 1)])
    order = np.argsort(fin_sm)
    print("last finish per SM (us): min %.2f p50 %.2f max %.2f; slowest SMs %s" % (fin_sm[fin_sm > 0].min(), np.median(fin_sm[fin_sm > 0]), fin_sm.max(), order[-6:].tolist()))
