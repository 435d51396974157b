#!/usr/bin/env python
"""bench.py - headline benchmark of the POGEMA step path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

A "step" is one pass of the hot path (move/collision + on_target bookkeeping +
time limit + observations) over every instance of the workload.

Workload (BASELINE.json configs[1], the one the metric is quoted on):
    4096 instances per GPU of 32x32 random maps, density 0.3, 64 agents each, obs_radius 5,
    priority collisions, on_target='finish', max_episode_steps 64, auto reset,
    uniform random actions (pre-generated, resident in HBM).

One JSON line on stdout (rank 0).  See the task contract for the keys.
"""
import argparse
import json
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


def algorithmic_bytes_per_agent_step(r, A, P):
    """3(2r+1)^2 obs + 21 state/action/flags + bit-packed padded map / A."""
    D = 2 * r + 1
    return 3 * D * D + 21 + ((P * P + 7) // 8) / A


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
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
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
                                          "-i", str(self.gpu_index), "-lms", "100"],
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
        # median of the upper half ~ clocks under load (idle samples before/after are lower)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """dram bytes per launch of the step kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- #
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
        "config": {"workload": WORKLOAD_NAME, "sample": sample},
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


def run_cuda(args):
    out = _claim_stdout()
    import numpy as np
    import torch
    import torch.distributed as dist
    from pogema_b200 import BatchedPogema, GridConfig

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

    N = args.instances
    A = WORKLOAD["num_agents"]
    gc = GridConfig(**WORKLOAD)
    seeds = np.arange(rank * N, (rank + 1) * N, dtype=np.uint64)  # instance k of the job <-> seed k
    env = BatchedPogema(gc, num_envs=N, device=dev, seeds=seeds, auto_reset=True)
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_act = 16
    acts = [torch.randint(0, 5, (N, A), dtype=torch.uint8, device=dev, generator=gen) for _ in range(n_act)]
    # obs ring larger than L2 (4 x 95 MB > 126 MB) so consecutive steps cannot hit in cache
    ring = [env.new_obs_buffer() for _ in range(4)]
    stream = torch.cuda.current_stream(dev)

    # The K timed steps are issued as launches of SPL consecutive steps each (pgm_step_many: one kernel
    # advances every instance by SPL steps on its own timeline, all per-step outputs are written), plus
    # K % SPL single-step launches.  --steps-per-launch 1 times the closed-loop form instead: one launch
    # per step, replayed from a CUDA graph (or launched from Python with --no-graph).
    SPL = max(1, args.steps_per_launch)
    GRAPH_STEPS = n_act
    ring_t = torch.stack(ring)                      # [4, N, A, 3, D, D] observation ring, 4 x 95 MB > L2
    ring = [ring_t[i] for i in range(4)]
    act_t = torch.stack(acts)                       # [16, N, A]
    rew_t = torch.empty((SPL, N, A), dtype=torch.float32, device=dev)
    term_t = torch.empty((SPL, N, A), dtype=torch.bool, device=dev)
    trunc_t = torch.empty((SPL, N, A), dtype=torch.bool, device=dev)
    act_many = act_t[torch.arange(SPL, device=dev) % n_act].contiguous() if SPL > 1 else None
    sptr = int(stream.cuda_stream)

    def plain_steps(first, count):
        for i in range(first, first + count):
            env.step(acts[i % n_act], out=ring[i % 4])

    def many_steps(launches):
        for _ in range(launches):
            env.engine.step_many(SPL, act_many.data_ptr(), 1, ring_t.data_ptr(), 4, rew_t.data_ptr(),
                                 term_t.data_ptr(), trunc_t.data_ptr(), sptr)

    plain_steps(0, max(args.warmup, 3))
    graph = None
    if SPL > 1:
        many_steps(2)
    elif not args.no_graph:
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                plain_steps(0, GRAPH_STEPS)
        stream.wait_stream(side)
        graph.replay()  # untimed warm-up of the instantiated graph
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = env.engine.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if SPL > 1:
        many_steps(args.steps // SPL)
        plain_steps(0, args.steps % SPL)
    elif graph is not None:
        for _ in range(args.steps // GRAPH_STEPS):
            graph.replay()
        plain_steps(0, args.steps % GRAPH_STEPS)
    else:
        plain_steps(0, args.steps)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    # graph replays launch GRAPH_STEPS kernels each without passing through pgm_step
    launches = (env.engine.launch_count - launches0) + (GRAPH_STEPS * (args.steps // GRAPH_STEPS) if graph is not None else 0)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    env.check_errors()
    ms_per_step = ms_total / args.steps
    value = world * N * A * args.steps / (ms_total * 1e-3)

    # ---- closed-loop form for comparison: one kernel launch per step (what an RL loop with a policy in
    # between would issue), replayed from a CUDA graph
    closed_loop = None
    if world == 1 and SPL > 1 and not args.no_graph:
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(stream)
        with torch.cuda.stream(side):
            with torch.cuda.graph(g2, stream=side):
                plain_steps(0, GRAPH_STEPS)
        stream.wait_stream(side)
        g2.replay()
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 2048) // GRAPH_STEPS)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(reps):
            g2.replay()
        c1.record(stream)
        torch.cuda.synchronize()
        cl_ms = c0.elapsed_time(c1) / (reps * GRAPH_STEPS)
        closed_loop = {"value": N * A / (cl_ms * 1e-3), "unit": UNIT, "ms_per_step": cl_ms, "steps": reps * GRAPH_STEPS,
                       "launch": "one kernel launch per step (pgm_step), CUDA graph of %d launches replayed" % GRAPH_STEPS}

    # ---- e2e: the C-ABI host-buffer call (pgm_step_host), H2D actions + D2H obs/rewards/flags every step
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    h_act = [torch.randint(0, 5, (N, A), dtype=torch.uint8).pin_memory() for _ in range(4)]
    h_obs = torch.empty(env.engine.obs_shape(), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty((N, A), dtype=torch.float32).pin_memory()
    h_term = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    h_trunc = torch.empty((N, A), dtype=torch.uint8).pin_memory()
    sptr = int(stream.cuda_stream)

    def host_step(i):
        env.engine.step_host(h_act[i % 4].numpy(), h_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy(), sptr)

    def time_host_steps():
        """e2e rate of the host-buffer call: three back-to-back windows of e2e_steps / 3 steps each, the median
        window is reported (the widening threads share the host with whatever else runs on the box; all three
        window rates are kept in the JSON line)."""
        for i in range(8):  # first calls allocate staging buffers and start the widening threads
            host_step(i)
        per = max(1, e2e_steps // 3)
        rates = []
        for w in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in range(per):
                host_step(i)
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            rates.append(world * N * A * per / dt)
        info = env.engine.host_transport_info()
        return {"value": sorted(rates)[1], "unit": UNIT, "h2d_bytes_per_step": info["h2d_bytes"],
                "d2h_bytes_per_step": info["d2h_bytes"], "steps": 3 * per, "window_rates": rates}

    # Two transports of the same call, same uint8 result in the caller's host buffer (pgm_set_host_transport):
    # packed = the kernel writes the observation bit stream, the copy engine moves 1/8 of the bytes, host
    # threads widen it into h_obs while later chunks are on the bus; plain = DMA of the final uint8 tensor.
    host_threads = max(1, (os.cpu_count() or 1) // world)
    env.engine.set_host_transport("packed", host_threads)
    e2e_packed = time_host_steps()
    e2e_packed["api"] = ("pgm_step_host (C-ABI, host buffers), packed transport: GPU-written bit stream over PCIe, "
                         "%d host threads (%s) widen it to uint8 [N,A,3,11,11]" % (host_threads, env.engine.host_transport_info()["isa"]))
    env.engine.set_host_transport("plain")
    e2e_plain = time_host_steps()
    e2e_plain["api"] = "pgm_step_host (C-ABI, pinned host buffers), plain transport: DMA of the uint8 tensor (PCIe-bound)"
    env.engine.set_host_transport("auto")
    e2e, e2e_other = (e2e_packed, e2e_plain) if e2e_packed["value"] >= e2e_plain["value"] else (e2e_plain, e2e_packed)
    h2d = N * A

    # same call with the bit-packed observation format (48 B instead of 363 B per agent over PCIe)
    e2e_bits = None
    if world == 1:
        envb = BatchedPogema(gc, num_envs=N, device=dev, seeds=seeds, auto_reset=True, obs_format="bits")
        envb.reset()
        hb_obs = torch.empty(envb.engine.obs_shape(), dtype=torch.int32).pin_memory()

        def host_step_bits(i):
            envb.engine.step_host(h_act[i % 4].numpy(), hb_obs.numpy(), h_rew.numpy(), h_term.numpy(), h_trunc.numpy(), sptr)

        for i in range(3):
            host_step_bits(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            host_step_bits(i)
        torch.cuda.synchronize()
        tb = time.perf_counter() - t0
        e2e_bits = {"value": N * A * e2e_steps / tb, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": envb.engine.obs_bytes + N * A * 4 + 2 * N * A, "steps": e2e_steps,
                    "api": "pgm_step_host with obs_format=bits (uint32 [N,A,12], bit k = element k of the uint8 layout)"}
        envb.close()

    if rank == 0:
        r = WORKLOAD["obs_radius"]
        P = WORKLOAD["size"] + 2 * r
        bpa = algorithmic_bytes_per_agent_step(r, A, P)
        peak, peak_src = measured_peak_gbs()
        achieved = N * A * bpa / (ms_per_step * 1e-3) / 1e9
        traffic = ncu_traffic_bytes()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "instances_per_gpu": N, "agents_per_instance": A,
                       "obs": "uint8 [N,A,3,11,11]", "actions": "uint8 resident in HBM, 16 pre-generated tensors",
                       "l2": "obs written to a ring of 4 buffers (4 x %.0f MB > 126 MB L2)" % (env.engine.obs_bytes / 1e6),
                       "launch": ("pgm_step_many: %d steps per kernel launch (every step writes all its outputs)" % SPL) if SPL > 1
                       else (("one launch per step, CUDA graph of %d launches replayed" % GRAPH_STEPS) if graph is not None else "one pgm_step call per step"),
                       "steps_per_launch": SPL,
                       "plan": env.engine.plan(), "parallelism": f"instances sharded over {world} GPU(s), no collective"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "frac_of_nominal_8000_GBps": achieved / 8000.0,
                         "algorithmic_bytes_per_agent_step": bpa, "kernel": "pgm_step_kernel (%d step(s) per launch)" % SPL,
                         "algorithmic_bytes_per_launch": N * A * bpa * SPL},
            "e2e": e2e,
            "e2e_other_transport": e2e_other,
            "gpu_launches": launches,
            "clocks": clocks,
            "closed_loop": closed_loop,
            "e2e_bits": e2e_bits,
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
    ap.add_argument("--steps-per-launch", type=int, default=16,
                    help="steps advanced by one kernel launch (pgm_step_many); 1 = one launch per step")
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
