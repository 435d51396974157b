#!/usr/bin/env python
"""bench.py - headline benchmark of the POGEMA step path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" is one pass of the hot path (move/collision + on_target bookkeeping +
time limit + observations) over every instance of the workload.

Headline workload (BASELINE.json configs[1], the one the metric is quoted on):
    4096 instances per GPU of 32x32 random maps, density 0.3, 64 agents each, obs_radius 5,
    priority collisions, on_target='finish', max_episode_steps 64, auto reset,
    uniform random actions (pre-generated, resident in HBM).

The K timed steps are issued as launches of pgm_step_many (a launch advances every instance by 16 steps -
K <= 32: one launch of K steps - and writes every step's outputs); nothing else runs in the timed region.
The same line carries, measured after the timed region:
    closed_loop    one launch per step (pgm_step), CUDA graph replay - what an RL loop with a policy issues;
                   closed_loop.groups: the same with the instances split into 2 / 4 groups on their own streams
                   (double-buffered sampling: one group steps while the policy works on the other)
    configs        the other BASELINE.json configurations (configs[2], [3], [4] r=3/5/7 at this world size),
                   both launch forms, each against its own algorithmic bytes
    e2e            pgm_step_host with HOST buffers (copies inside the timed windows)
    roofline.lone_launch_behind_foreign_rollout   the K timed steps again, L2 flushed by a second engine's rollout
                   instead of the ordinary 1 GB fill (whose lines stay in L2 and outrank the evict-first observations)
    sharding_check rank k's results == the C oracle / a single-GPU run of the same global seeds
    cpu_baseline   the oracle port on the host cores (N=1 only)

One JSON line on stdout (rank 0).  See the task contract for the keys.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent-steps/sec incl. obs"
UNIT = "agent-steps/s"
WORKLOAD = dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
                collision_system="priority", on_target="finish")
INSTANCES_PER_GPU = 4096
WORKLOAD_NAME = ("configs[1]: 4096 instances/GPU of 32x32 random maps, density 0.3, 64 agents, obs_radius 5, "
                 "priority collisions, on_target=finish, max_episode_steps 64, auto-reset, random actions")
L2_BYTES = 126e6
STEPS_PER_LAUNCH = 16


def workload_config(world, instances):
    """The `config` object of the JSON line: the same keys and values on both arms (cuda / reference)."""
    return {"workload": WORKLOAD_NAME, "instances_per_gpu": instances,
            "agents_per_instance": WORKLOAD["num_agents"], "obs": "uint8 [N,A,3,11,11]",
            "parallelism": f"instances sharded over {world} GPU(s), no collective"}


def algorithmic_bytes_per_agent_step(r, A, PH, PW=None):
    """3(2r+1)^2 obs + 21 state/action/flags + bit-packed padded map / A."""
    D = 2 * r + 1
    PW = PH if PW is None else PW
    return 3 * D * D + 21 + ((PH * PW + 7) // 8) / A


# --------------------------------------------------------------------------- #
# CPU baseline: the oracle (restated reference), one process per core
# --------------------------------------------------------------------------- #
def _oracle_worker(args):
    seed0, n_inst, n_steps, budget_s, n_warm = args
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle import pogema_oracle as orc
    envs = []
    for k in range(n_inst):
        env = orc.pogema_v0(orc.GridConfig(seed=seed0 + k, **WORKLOAD))
        env.reset()
        envs.append(env)
    rng = np.random.default_rng(seed0)
    A = WORKLOAD["num_agents"]
    done_steps = 0
    for t in range(n_warm):
        for env in envs:
            obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=A))
            if all(term) or all(trunc):
                env.reset()
    t0 = time.perf_counter()
    for t in range(n_steps):
        for env in envs:
            obs, rew, term, trunc, info = env.step(rng.integers(0, 5, size=A))
            if all(term) or all(trunc):
                env.reset()
        done_steps += 1
        if budget_s and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done_steps * n_inst * A, dt


def oracle_throughput(cores, n_inst_per_core, n_steps, budget_s=0.0, seed0=10_000, n_warm=3):
    """Aggregate agent-steps/s of the Python/numpy oracle on `cores` processes."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(seed0 + c * 1000, n_inst_per_core, n_steps, budget_s, n_warm) for c in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return total / slowest, total, slowest, wall


def _c_oracle_worker(args):
    seed0, n_inst, budget_s = args
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle.c_driver import COracle
    co = COracle.from_python_oracle(WORKLOAD, list(range(seed0, seed0 + n_inst)))
    rng = np.random.default_rng(seed0)
    A = WORKLOAD["num_agents"]
    T = 64
    done = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        acts = rng.integers(0, 5, size=(T, n_inst, A)).astype(np.uint8)
        done += co.run(acts, auto_reset=True, want_obs=True)["agent_steps"]
    return done, time.perf_counter() - t0


def c_oracle_throughput(cores, budget_s=4.0, n_inst=32):
    """Aggregate agent-steps/s of the C restatement (oracle/step_oracle.c), one process per core."""
    import multiprocessing as mp
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_step.so")):
        return None
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_c_oracle_worker, [(20_000 + c * 1000, n_inst, budget_s) for c in range(cores)])
    total = sum(r[0] for r in res)
    return total / max(r[1] for r in res), total


# --------------------------------------------------------------------------- #
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the GPU measurements run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the upper half of the samples ~ clocks under load (idle samples before/after are lower)
        upper = sm[len(sm) // 2:]
        med = upper[len(upper) // 2] if upper else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons),
                "window": "headline timed region + closed_loop + per-config measurements (the headline region alone "
                          "is shorter than one nvidia-smi sample)"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """(dram bytes per launch, steps of that launch, source) of the step kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            t = json.load(open(path))
            return t.get("dram_bytes_per_launch"), int(t.get("steps_per_launch", 1)), t.get("source")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- #
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    A = WORKLOAD["num_agents"]
    # one bench "step" = one oracle step over cores*inst_per_core instances (a bounded sample of the workload),
    # sized for ~6 s per process at ~200 k agent-steps/s/core so that short runs are not dominated by noise
    n_steps = args.steps
    inst_per_core = int(min(256, max(8, -(-6 * 200_000 // (max(n_steps, 1) * A)))))
    t_rate, total, slowest, wall = oracle_throughput(cores, inst_per_core, n_steps, n_warm=args.warmup)
    sample = (f"{cores} processes x {inst_per_core} instances of the workload shape x {n_steps} steps "
              f"({total} agent-steps, slowest process {slowest:.1f} s), Python/numpy restatement of upstream pogema")
    line = {
        "impl": "reference", "metric": METRIC, "value": t_rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * slowest / n_steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(world, args.instances),
        "cpu_baseline": {"value": t_rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": t_rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line at the
    first communicator): keep the real stdout for the result and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def split_steps(total, per_launch):
    """K steps as ceil(K / per_launch) launches of (nearly) equal size: 20 -> [10, 10], 8192 -> 512 x [16]."""
    if total <= 0:
        return []
    n = -(-total // per_launch)
    base, rem = divmod(total, n)
    return [base + (1 if i < rem else 0) for i in range(n)]


class Harness:
    """Device-resident timing of one BatchedPogema in both launch forms."""

    def __init__(self, torch, env, dev, seed, spl=STEPS_PER_LAUNCH):
        self.torch, self.env, self.dev, self.spl = torch, env, dev, spl
        N, A = env.num_envs, env.num_agents
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.n_act = 16
        self.act_t = torch.randint(0, 5, (self.n_act, N, A), dtype=torch.uint8, device=dev, generator=gen)
        # observation ring larger than L2 so that consecutive steps cannot hit in cache
        self.nring = int(min(16, max(2, math.ceil(2.2 * L2_BYTES / env.engine.obs_bytes))))
        self.ring_t = torch.stack([env.new_obs_buffer() for _ in range(self.nring)])
        self.stream = torch.cuda.current_stream(dev)
        self.sptr = int(self.stream.cuda_stream)
        self.graph = None

    def alloc_outputs(self, max_steps):
        N, A, torch = self.env.num_envs, self.env.num_agents, self.torch
        self.rew_t = torch.empty((max_steps, N, A), dtype=torch.float32, device=self.dev)
        self.term_t = torch.empty((max_steps, N, A), dtype=torch.bool, device=self.dev)
        self.trunc_t = torch.empty((max_steps, N, A), dtype=torch.bool, device=self.dev)
        idx = torch.arange(max_steps, device=self.dev) % self.n_act
        self.act_many = self.act_t[idx].contiguous()

    def many(self, sizes):
        e = self.env.engine
        for k in sizes:
            e.step_many(k, self.act_many.data_ptr(), 1, self.ring_t.data_ptr(), self.nring, self.rew_t.data_ptr(),
                        self.term_t.data_ptr(), self.trunc_t.data_ptr(), self.sptr)

    def single_steps(self, first, count):
        for i in range(first, first + count):
            self.env.step(self.act_t[i % self.n_act], out=self.ring_t[i % self.nring])

    def capture_graph(self, steps=16):
        torch = self.torch
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(self.stream)
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self.single_steps(0, steps)
        self.stream.wait_stream(side)
        g.replay()
        torch.cuda.synchronize()
        self.graph, self.graph_steps = g, steps

    def timed(self, fn):
        """CUDA-event time of the launches fn() enqueues.  A 1 GB fill runs right before the first event: it
        flushes L2 (126 MB) and keeps the GPU busy while the CPU enqueues the timed launches, so the events
        bracket device time only (the kernels are queued by the time the fill ends), not ctypes / launch latency.  (1 GB
        = ~0.3 ms of cover: with 384 MB = ~0.12 ms one box of the pool timed the 20-step launch at 16.8 instead of
        16.3 us per step although its steady-state and behind-a-rollout figures were the usual ones - its host was late.)"""
        torch = self.torch
        if not hasattr(self, "flush"):
            self.flush = torch.empty(1024 << 20, dtype=torch.uint8, device=self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.flush.fill_(1)
        e0.record(self.stream)
        fn()
        e1.record(self.stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)


def run_cuda(args):
    out = _claim_stdout()
    import numpy as np
    import torch
    import torch.distributed as dist
    from pogema_b200 import BatchedPogema, GridConfig
    from pogema_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = measured_peak_gbs()
    N = args.instances
    A = WORKLOAD["num_agents"]
    gc = GridConfig(**WORKLOAD)
    seeds = np.arange(rank * N, (rank + 1) * N, dtype=np.uint64)  # instance k of the job <-> seed k
    env = BatchedPogema(gc, num_envs=N, device=dev, seeds=seeds, auto_reset=True)
    env.reset()
    SPL = max(1, args.steps_per_launch)
    h = Harness(torch, env, dev, 1234 + rank, SPL)
    # up to 2 x SPL steps go out as ONE launch (one ramp, one tail); longer runs as launches of SPL steps
    sizes = ([args.steps] if args.steps <= 2 * SPL else split_steps(args.steps, SPL)) if SPL > 1 else []
    h.alloc_outputs(max(sizes + [SPL, 1]))

    # ---- headline: EXACTLY args.steps steps, device resident -------------------------------------------
    warm = max(args.warmup, 3)
    if SPL > 1:
        h.many(split_steps(warm, SPL))          # W untimed warm-up steps, same launch form as the timed ones
        h.many(sizes[:1])
    else:
        h.single_steps(0, warm)
        if not args.no_graph:
            h.capture_graph(16)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = env.engine.launch_count
    if SPL > 1:
        ms_total = h.timed(lambda: h.many(sizes))
        launches = env.engine.launch_count - launches0
    elif h.graph is not None:
        reps, rem = divmod(args.steps, h.graph_steps)

        def replay():
            for _ in range(reps):
                h.graph.replay()
            h.single_steps(0, rem)
        ms_total = h.timed(replay)
        launches = reps * h.graph_steps + rem
    else:
        ms_total = h.timed(lambda: h.single_steps(0, args.steps))
        launches = env.engine.launch_count - launches0
    barrier()
    ms_total = max_over_ranks(ms_total)
    env.check_errors()
    ms_per_step = ms_total / args.steps
    value = world * N * A * args.steps / (ms_total * 1e-3)

    # ---- both launch forms of one environment, a fixed amount of work each (independent of --steps) -------
    def measure_forms(hh, n_inst, n_agents, bpa, target_ms=30.0):
        est_us = max(2.0, n_inst * n_agents * bpa / 5.0e12 * 1e6)          # at ~5 TB/s
        reps = int(min(128, max(3, target_ms * 1e3 / (est_us * 16))))
        hh.alloc_outputs(16)
        hh.many([16, 16])
        torch.cuda.synchronize()
        barrier()
        ms_many = max_over_ranks(hh.timed(lambda: hh.many([16] * reps))) / (reps * 16)
        rec = {"steps_per_launch_16": {"us_per_step": ms_many * 1e3, "agent_steps_per_s": world * n_inst * n_agents / (ms_many * 1e-3),
                                       "roofline_frac": n_inst * n_agents * bpa / (ms_many * 1e-3) / 1e9 / peak,
                                       "steps": reps * 16}}
        if not args.no_graph:
            hh.capture_graph(16)
            barrier()

            def replay():
                for _ in range(reps):
                    hh.graph.replay()
            ms_cl = max_over_ranks(hh.timed(replay)) / (reps * 16)
            rec["one_launch_per_step"] = {"us_per_step": ms_cl * 1e3, "agent_steps_per_s": world * n_inst * n_agents / (ms_cl * 1e-3),
                                          "roofline_frac": n_inst * n_agents * bpa / (ms_cl * 1e-3) / 1e9 / peak,
                                          "steps": reps * 16, "launch": "pgm_step, CUDA graph of 16 launches replayed"}
        return rec

    def measure_groups(gcc, n_inst, n_agents, bpa, groups, target_ms=30.0):
        """Closed loop over `groups` groups of instances (the double-buffered sampling of RL frameworks: the policy
        works on one group's observations while the other groups step).  Every group is its own engine on its own
        stream and advances with ONE LAUNCH PER STEP (CUDA graph of 16 single-step launches per group, replayed): a
        group's step t+1 starts only after its own step t completed, but one group's dependent front (state loads,
        move resolution, first bit assembly) overlaps the other groups' observation stores."""
        n_g = n_inst // groups
        hs = []
        for k in range(groups):
            st = torch.cuda.Stream(dev)
            with torch.cuda.stream(st):
                first = rank * n_inst + k * n_g
                eg = BatchedPogema(gcc, num_envs=n_g, device=dev, seeds=np.arange(first, first + n_g, dtype=np.uint64), auto_reset=True)
                eg.reset()
                hg = Harness(torch, eg, dev, 4321 + 16 * rank + k)
                hg.single_steps(0, 4)
                hg.capture_graph(16)
            hs.append(hg)
        est_us = max(2.0, n_inst * n_agents * bpa / 5.0e12 * 1e6)
        reps = int(min(128, max(3, target_ms * 1e3 / (est_us * 16))))
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h.flush.fill_(1)  # (h.timed ran before: the buffer exists) L2 flushed, GPU busy while the CPU enqueues
        e0.record(h.stream)
        for hg in hs:
            hg.stream.wait_event(e0)
        for _ in range(reps):
            for hg in hs:
                with torch.cuda.stream(hg.stream):
                    hg.graph.replay()
        for hg in hs:
            h.stream.wait_stream(hg.stream)
        e1.record(h.stream)
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1)) / (reps * 16)
        plan_g = hs[0].env.engine.plan()
        for hg in hs:
            hg.env.check_errors()
            hg.env.close()
        del hs
        torch.cuda.empty_cache()
        return {"groups": groups, "instances_per_group": n_g, "us_per_step": ms * 1e3,
                "agent_steps_per_s": world * n_g * groups * n_agents / (ms * 1e-3),
                "roofline_frac": n_g * groups * n_agents * bpa / (ms * 1e-3) / 1e9 / peak, "steps_per_group": reps * 16,
                "launch": "pgm_step, one launch per step and group; %d engines on %d streams, CUDA graph of 16 launches each" % (groups, groups),
                "plan": plan_g.get("fast", plan_g)}

    r = WORKLOAD["obs_radius"]
    P = WORKLOAD["size"] + 2 * r
    bpa = algorithmic_bytes_per_agent_step(r, A, P)
    # ---- what the lone K-step launch of the headline pays for (tools/lone_launch.py): the 1 GB fill in front of it
    # leaves L2 full of ORDINARY dirty lines, which outrank the kernel's evict-first observation lines for the whole
    # timed region.  The same launches behind a 16-step rollout of a second engine of the same shape (L2 flushed by
    # 1.5 GB of foreign evict-first lines instead: cold for the timed launch, nothing squats):
    lone = None
    if SPL > 1 and not args.no_configs:
        ef = BatchedPogema(gc, num_envs=N, device=dev, seeds=seeds + np.uint64(1 << 20), auto_reset=True)
        ef.reset()
        hf = Harness(torch, ef, dev, 777 + rank, SPL)
        hf.alloc_outputs(16)
        hf.many([16, 16])
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            hf.many([16])
            e0.record(h.stream)
            h.many(sizes)
            e1.record(h.stream)
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1)))
        ms_lone = sorted(ts)[len(ts) // 2] / args.steps
        lone = {"us_per_step": ms_lone * 1e3, "frac": N * A * bpa / (ms_lone * 1e-3) / 1e9 / peak,
                "how": "the same %d timed steps behind a 16-step rollout of a second engine (1.5 GB of evict-first stores to "
                       "foreign buffers flush L2) instead of the 1 GB ordinary fill of the headline, whose lines stay in L2 "
                       "and outrank the kernel's evict-first observation lines; median of 5" % args.steps}
        ef.check_errors()
        ef.close()
        del hf, ef
        torch.cuda.empty_cache()
    forms = measure_forms(h, N, A, bpa)
    closed_loop = None
    if "one_launch_per_step" in forms:
        cl = forms["one_launch_per_step"]
        closed_loop = {"value": cl["agent_steps_per_s"], "unit": UNIT, "ms_per_step": cl["us_per_step"] * 1e-3,
                       "steps": cl["steps"], "roofline_frac": cl["roofline_frac"],
                       "launch": "one kernel launch per step (pgm_step), CUDA graph of 16 launches replayed"}
        if not args.no_configs:
            closed_loop["groups"] = [measure_groups(gc, N, A, bpa, g) for g in (2, 4)]
    steady = forms["steps_per_launch_16"]
    plan_main = env.engine.plan()

    # ---- the other BASELINE.json configurations at this world size ------------------------------------------
    config_records = []
    if not args.no_configs:
        from pogema_b200.maps import maze_map, warehouse_map
        share = max(1, 16384 // world)
        table = [
            ("configs[2]: 1024 instances/GPU of 64x64 maze-like maps, 256 agents, r=5, soft collisions, on_target=restart",
             1024, dict(map=maze_map(64, 3).tolist(), num_agents=256, obs_radius=5, collision_system="soft", on_target="restart"), "weak"),
            ("configs[3]: 512 instances/GPU of 256x256 warehouse maps, 1024 agents, r=5, block_both, on_target=finish",
             512, dict(map=warehouse_map(256).tolist(), num_agents=1024, obs_radius=5, collision_system="block_both", on_target="finish"), "weak"),
        ] + [
            (f"configs[4]: 1M agents over {world} GPU(s) = {share} instances/GPU of 32x32, 64 agents, r={rr}, priority/finish",
             share, dict(size=32, density=0.3, num_agents=64, obs_radius=rr, collision_system="priority", on_target="finish"), "strong")
            for rr in (3, 5, 7)
        ]
        for name, n_inst, kw, scaling in table:
            gcc = GridConfig(max_episode_steps=64, **kw)
            hh_, ww_ = gcc.map_shape()
            rr, aa = gcc.obs_radius, gcc.num_agents
            b = algorithmic_bytes_per_agent_step(rr, aa, hh_ + 2 * rr, ww_ + 2 * rr)
            e2 = BatchedPogema(gcc, num_envs=n_inst, device=dev, seeds=np.arange(rank * n_inst, (rank + 1) * n_inst, dtype=np.uint64),
                               auto_reset=True)
            e2.reset()
            h2 = Harness(torch, e2, dev, 99 + rank)
            rec = {"config": name, "instances_per_gpu": n_inst, "agents_per_instance": aa, "scaling": scaling,
                   "algorithmic_bytes_per_agent_step": b, "plan": e2.engine.plan(), "obs_ring_slots": h2.nring}
            rec.update(measure_forms(h2, n_inst, aa, b))
            if not args.no_graph and name.startswith("configs[2]"):
                rec["one_launch_per_step_2_groups"] = measure_groups(gcc, n_inst, aa, b, 2)
            e2.check_errors()
            config_records.append(rec)
            e2.close()
            del h2, e2
            torch.cuda.empty_cache()
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the C-ABI host-buffer call (pgm_step_host), H2D actions + D2H obs/rewards/flags every step
    per_window = max(32, args.e2e_steps // 3)
    h_act = [torch.randint(0, 5, (N, A), dtype=torch.uint8).pin_memory() for _ in range(4)]
    h_obs = torch.empty(env.engine.obs_shape(), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((N, A), dtype=torch.float32).pin_memory()
    h_term = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    h_trunc = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    sptr = h.sptr

    def time_host_steps(engine, obs_np):
        """e2e rate of the host-buffer call: three back-to-back windows of `per_window` steps each (independent of
        --steps), the median window is reported (the widening threads share the host with whatever else runs on
        the box; all three window rates are kept in the JSON line)."""
        def host_step(i):
            engine.step_host(h_act[i % 4].numpy(), obs_np, h_rew.numpy(), h_term.numpy(), h_trunc.numpy(), sptr)
        for i in range(8):  # first calls allocate staging buffers and start the widening threads
            host_step(i)
        rates = []
        for w in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in range(per_window):
                host_step(i)
            torch.cuda.synchronize()
            dt = max_over_ranks(time.perf_counter() - t0)
            rates.append(world * N * A * per_window / dt)
        info = engine.host_transport_info()
        return {"value": sorted(rates)[1], "unit": UNIT, "h2d_bytes_per_step": info["h2d_bytes"],
                "d2h_bytes_per_step": info["d2h_bytes"], "steps": 3 * per_window, "window_rates": rates}

    # Two transports of the same call, same uint8 result in the caller's host buffer (pgm_set_host_transport):
    # packed = the kernel writes the observation bit stream, the copy engine moves 1/8 of the bytes, host
    # threads widen it into h_obs while later chunks are on the bus; plain = DMA of the final uint8 tensor.
    host_threads = max(1, (os.cpu_count() or 1) // world)
    env.engine.set_host_transport("packed", host_threads)
    e2e_packed = time_host_steps(env.engine, h_obs.numpy())
    e2e_packed["api"] = ("pgm_step_host (C-ABI, host buffers), packed transport: GPU-written bit stream over PCIe, "
                         "%d host threads per rank (%s) widen it to uint8 [N,A,3,11,11]" % (host_threads, env.engine.host_transport_info()["isa"]))
    env.engine.set_host_transport("plain")
    e2e_plain = time_host_steps(env.engine, h_obs.numpy())
    e2e_plain["api"] = "pgm_step_host (C-ABI, pinned host buffers), plain transport: DMA of the uint8 tensor (PCIe-bound)"
    env.engine.set_host_transport("auto")
    e2e, e2e_other = (e2e_packed, e2e_plain) if e2e_packed["value"] >= e2e_plain["value"] else (e2e_plain, e2e_packed)

    # The packed transport is bound by the host's DRAM: every rank's threads write obs_bytes of uint8 per step into
    # one host.  Ceiling = a plain non-temporal fill of the same buffer by the same number of threads on every rank
    # at the same time (pgm_host_fill_gbps); `frac_of_ceiling` = bytes/s the widening loop wrote / that.
    barrier()
    fill = float(nat.load().pgm_host_fill_gbps(h_obs.data_ptr(), h_obs.numel(), host_threads, 6))
    if world > 1:
        t = torch.tensor([fill], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        fill_total = float(t.item())
    else:
        fill_total = fill
    written_gbs = e2e_packed["value"] / (N * A) * env.engine.obs_bytes / 1e9   # all ranks
    host_dram = {"nt_fill_GBps_all_ranks": fill_total, "threads_per_rank": host_threads, "bytes": int(h_obs.numel()),
                 "packed_transport_write_GBps_all_ranks": written_gbs, "frac_of_ceiling": written_gbs / max(fill_total, 1e-9),
                 "how": "pgm_host_fill_gbps: every rank fills its own pinned observation buffer at the same time, "
                        "_mm512_stream (or AVX2 / memset), 6 passes"}

    # same call with the bit-packed observation format (48 B instead of 363 B per agent over PCIe, no widening)
    envb = BatchedPogema(gc, num_envs=N, device=dev, seeds=seeds, auto_reset=True, obs_format="bits")
    envb.reset()
    hb_obs = torch.empty(envb.engine.obs_shape(), dtype=torch.int32).pin_memory()
    e2e_bits = time_host_steps(envb.engine, hb_obs.numpy())
    e2e_bits["api"] = "pgm_step_host with obs_format=bits (uint32 [N,A,12], bit k = element k of the uint8 layout)"
    envb.close()

    # ---- sharding check (untimed): rank k's instances [k*N, (k+1)*N) give what a single-GPU run of the same
    # global seeds gives, and what the C oracle gives
    sharding = sharding_check(torch, dist, np, BatchedPogema, gc, dev, rank, world, N, A)

    if rank == 0:
        # dram bytes of ONE launch as timed here: the ncu capture is a 16-step launch, scaled to this launch's steps
        cap = ncu_traffic_bytes()
        steps_in_launch = sizes[0] if sizes else 1
        traffic = cap[0] / cap[1] * steps_in_launch if cap and cap[0] else None
        achieved = N * A * bpa / (ms_per_step * 1e-3) / 1e9
        cfg = workload_config(world, N)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "details": {"actions": "uint8 resident in HBM, 16 pre-generated tensors",
                        "l2": "obs written to a ring of %d buffers (%d x %.0f MB > 126 MB L2)" % (h.nring, h.nring, env.engine.obs_bytes / 1e6),
                        "launch": ("pgm_step_many: the %d timed steps = launches of %s steps (every step writes all its outputs)" % (args.steps, sizes if len(sizes) <= 4 else "%d x %d" % (len(sizes), sizes[0]))) if SPL > 1
                        else (("one launch per step, CUDA graph of %d launches replayed" % h.graph_steps) if h.graph is not None else "one pgm_step call per step"),
                        "steps_per_launch": SPL, "plan": plan_main},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_source": ("%s: %.1f MB per %d-step launch = %.1f MB per step, x %d steps of the timed launch"
                                            % (cap[2], cap[0] / 1e6, cap[1], cap[0] / cap[1] / 1e6, steps_in_launch)) if traffic else None,
                         "peak_source": peak_src, "frac_of_nominal_8000_GBps": achieved / 8000.0,
                         "algorithmic_bytes_per_agent_step": bpa, "kernel": ("%s (pgm_step_many, up to %d steps per launch)" % (
                             ("pgm_fast_step_kernel<%d,%d,...>" % (plan_main["fast"]["team_threads"], plan_main["fast"]["agents_per_thread"]))
                             if plan_main.get("fast_step_kernel") else "pgm_step_kernel", SPL)),
                         "algorithmic_bytes_per_launch": N * A * bpa * (sizes[0] if sizes else 1),
                         "steady_state": {"frac": steady["roofline_frac"], "us_per_step": steady["us_per_step"], "steps": steady["steps"],
                                          "note": "same kernel over a longer window (16 steps per launch), for comparison with the K-step headline"},
                         "lone_launch_behind_foreign_rollout": lone},
            "e2e": e2e,
            "e2e_other_transport": e2e_other,
            "gpu_launches": launches,
            "clocks": clocks,
            "closed_loop": closed_loop,
            "configs": config_records,
            "e2e_bits": e2e_bits,
            "host_dram": host_dram,
            "sharding_check": sharding,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, total, slowest, wall = oracle_throughput(cores, 2, 10_000, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": (f"{cores} processes x 2 instances of the workload shape for ~{args.cpu_seconds:.0f} s "
                           f"({total} agent-steps), Python/numpy restatement of upstream pogema (oracle/pogema_oracle.py)")}
            try:
                c_res = c_oracle_throughput(cores)
            except Exception:
                c_res = None
            if c_res:
                # context only: a plain-C restatement of the same path (not how the reference is implemented)
                line["cpu_baseline_c"] = {"value": c_res[0], "unit": UNIT, "cores": cores, "kind": "port (C, oracle/step_oracle.c)",
                                          "sample": f"{cores} processes x 32 instances for ~4 s ({c_res[1]} agent-steps)"}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def sharding_check(torch, dist, np, BatchedPogema, gc, dev, rank, world, N, A, T=24, samples=3):
    """Every rank steps its shard T times with actions that are a function of the GLOBAL instance index; rank 0
    then (a) re-runs `samples` instances of every rank on its own GPU and (b) runs the first sample of every
    rank on the C oracle, and compares positions, targets, active flags, reward sums and the last observation."""
    try:
        first = rank * N
        env = BatchedPogema(gc, num_envs=N, device=dev, seeds=np.arange(first, first + N, dtype=np.uint64), auto_reset=True)
        env.reset()

        def actions_for(ids):
            # [T, len(ids), A]: instance g's stream depends on g only
            return np.stack([np.random.default_rng(77_000 + int(g)).integers(0, 5, size=(T, A)) for g in ids], axis=1).astype(np.uint8)

        pick = sorted(set(int(x) for x in np.linspace(0, N - 1, samples)))
        acts = np.zeros((T, N, A), dtype=np.uint8)
        # all instances step (random actions), the sampled ones with their global streams
        acts[:] = np.random.default_rng(5 + rank).integers(0, 5, size=(T, N, A))
        acts[:, pick] = actions_for([first + k for k in pick])
        last = torch.stack([env.new_obs_buffer()])   # one observation slot: the last step's is what is compared
        obs, rew, term, trunc = env.rollout(torch.from_numpy(acts).to(dev), obs_out=last)
        rsum = rew.double().sum(0)
        mine = {"ids": [first + k for k in pick],
                "pos": env.get_agents_xy()[pick].cpu().numpy(), "tgt": env.get_targets_xy()[pick].cpu().numpy(),
                "active": env.is_active[pick].cpu().numpy(), "rsum": rsum[pick].cpu().numpy(),
                "obs": obs[-1][pick].cpu().numpy()}
        env.check_errors()
        env.close()
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, mine)
        else:
            parts = [mine]
        if rank != 0:
            return None
        ids = [g for p in parts for g in p["ids"]]
        ref = BatchedPogema(gc, num_envs=len(ids), device=dev, seeds=np.asarray(ids, dtype=np.uint64), auto_reset=True)
        ref.reset()
        robs, rrew, _, _ = ref.rollout(torch.from_numpy(actions_for(ids)).to(dev), obs_out=torch.stack([ref.new_obs_buffer()]))
        got = {k: np.concatenate([p[k] for p in parts]) for k in ("pos", "tgt", "active", "rsum", "obs")}
        same_gpu = (np.array_equal(got["pos"], ref.get_agents_xy().cpu().numpy())
                    and np.array_equal(got["tgt"], ref.get_targets_xy().cpu().numpy())
                    and np.array_equal(got["active"], ref.is_active.cpu().numpy())
                    and np.array_equal(got["rsum"], rrew.double().sum(0).cpu().numpy())
                    and np.array_equal(got["obs"], robs[-1].cpu().numpy()))
        ref.close()
        # C oracle (built from the Python oracle's reset, i.e. from numpy): the first sample of every rank
        from oracle.c_driver import COracle
        oid = [p["ids"][0] for p in parts]
        co = COracle.from_python_oracle(WORKLOAD, oid)
        o = co.run(actions_for(oid), auto_reset=True)
        sel = [ids.index(g) for g in oid]
        rr = WORKLOAD["obs_radius"]
        same_oracle = (np.array_equal(got["pos"][sel] + rr, co.pos) and np.array_equal(got["tgt"][sel] + rr, co.tgt)
                       and np.array_equal(got["active"][sel].astype(np.uint8), co.active)
                       and np.array_equal(got["rsum"][sel], o["rewards_sum"]) and np.array_equal(got["obs"][sel], o["obs"]))
        return {"status": "ok" if (same_gpu and same_oracle) else "MISMATCH", "steps": T, "instances_checked": ids,
                "vs_single_gpu_rerun": bool(same_gpu), "vs_c_oracle": bool(same_oracle), "oracle_instances": oid}
    except Exception as exc:  # the check must never take the bench line down
        return {"status": "error", "error": repr(exc)[:300]} if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8192)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--instances", type=int, default=INSTANCES_PER_GPU, help="instances per GPU")
    ap.add_argument("--e2e-steps", type=int, default=96)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config records (configs[2..4])")
    ap.add_argument("--steps-per-launch", type=int, default=STEPS_PER_LAUNCH,
                    help="most steps advanced by one kernel launch (pgm_step_many); 1 = one launch per step")
    ap.add_argument("--no-graph", action="store_true", help="launch every step from Python instead of CUDA graph replays")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
