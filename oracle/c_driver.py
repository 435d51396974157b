"""ctypes driver of oracle/liboracle_step.so (the C restatement of the step path).  TEST INFRASTRUCTURE
(see oracle/pogema_oracle.py): used by tests/ and by the CPU-baseline legs of bench.py only.
State is built from the Python oracle after reset (which calls the real numpy)."""
import ctypes as C
import os

import numpy as np

from oracle import pogema_oracle as orc

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle_step.so")
COLL = {"priority": 0, "block_both": 1, "soft": 2}
ONT = {"finish": 0, "nothing": 1, "restart": 2}


class OrcCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("PH", "PW", "A", "r", "collision", "on_target", "max_episode_steps")]


PCG_DTYPE = np.dtype([("state_hi", np.uint64), ("state_lo", np.uint64), ("inc_hi", np.uint64), ("inc_lo", np.uint64),
                      ("has_uint32", np.uint32), ("uinteger", np.uint32)])


def load():
    if not os.path.exists(LIB):
        # build the checker on the spot (gcc is part of the image); `make oracle` does the same
        import subprocess
        src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "step_oracle.c")
        try:
            subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, src, "-lm"], check=True)
        except Exception as exc:  # pragma: no cover
            raise RuntimeError("oracle/liboracle_step.so missing and gcc failed: run `make oracle`") from exc
    lib = C.CDLL(LIB)
    lib.orc_run.restype = C.c_longlong
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _build_chunk(args):
    gc_kwargs, seeds = args
    co = COracle.from_python_oracle(gc_kwargs, seeds)
    c = co.cfg
    return dict(cfg=(c.PH, c.PW, c.A, c.r, c.collision, c.on_target, c.max_episode_steps), obstacles=co.obstacles,
                pos=co.pos, tgt=co.tgt, rng=co.rng, comp_start=co.comp_start, comp_size=co.comp_size, cells=co.cells,
                cells_stride=co.cells_stride)


class COracle:
    """N instances of one config on the C oracle.  `from_python_oracle` builds the state from
    oracle.pogema_oracle resets; `from_arrays` takes explicit arrays (e.g. read back from the engine)."""

    def __init__(self, cfg, obstacles, pos, tgt, rng=None, comp_start=None, comp_size=None, cells=None,
                 cells_stride=0):
        self.lib = load()
        self.cfg = cfg
        self.N = obstacles.shape[0]
        self.A = cfg.A
        self.D = 2 * cfg.r + 1
        self.obstacles = np.ascontiguousarray(obstacles, dtype=np.uint8)
        self.pos = np.ascontiguousarray(pos, dtype=np.int32)
        self.tgt = np.ascontiguousarray(tgt, dtype=np.int32)
        self.pos0, self.tgt0 = self.pos.copy(), self.tgt.copy()
        self.active = np.ones((self.N, self.A), dtype=np.uint8)
        self.elapsed = np.zeros(self.N, dtype=np.int32)
        self.rng = rng
        self.rng0 = None if rng is None else rng.copy()
        self.comp_start, self.comp_size, self.cells, self.cells_stride = comp_start, comp_size, cells, cells_stride

    @classmethod
    def from_python_oracle(cls, gc_kwargs, seeds):
        envs = []
        for s in seeds:
            kw = dict(gc_kwargs)
            kw["seed"] = int(s)
            gc = orc.GridConfig(**kw)
            env = orc.PogemaLifeLong(gc) if gc.on_target == 'restart' else orc.Pogema(gc)
            env.reset()
            envs.append(env)
        g0 = envs[0].grid
        PH, PW = g0.obstacles.shape
        gc = envs[0].grid_config
        cfg = OrcCfg(PH, PW, gc.num_agents, gc.obs_radius, COLL[gc.collision_system], ONT[gc.on_target],
                     gc.max_episode_steps)
        obstacles = np.stack([e.grid.obstacles.astype(np.uint8) for e in envs])
        pos = np.stack([np.array(e.grid.positions_xy, dtype=np.int32) for e in envs])
        tgt = np.stack([np.array(e.grid.finishes_xy, dtype=np.int32) for e in envs])
        rng = comp_start = comp_size = cells = None
        stride = 0
        if gc.on_target == 'restart':
            A = gc.num_agents
            rng = np.zeros((len(envs), A), dtype=PCG_DTYPE)
            comp_start = np.zeros((len(envs), A), dtype=np.int32)
            comp_size = np.zeros((len(envs), A), dtype=np.int32)
            stride = PH * PW
            cells = np.zeros((len(envs), stride, 2), dtype=np.int32)
            for n, e in enumerate(envs):
                off = 0
                seen = {}
                for a in range(A):
                    st = e.random_generators[a].bit_generator.state
                    rng[n, a] = (st['state']['state'] >> 64, st['state']['state'] & (2**64 - 1),
                                 st['state']['inc'] >> 64, st['state']['inc'] & (2**64 - 1),
                                 st['has_uint32'], st['uinteger'])
                    cid = e.grid.point_to_component[e.grid.positions_xy[a]]
                    if cid not in seen:
                        pts = e.grid.component_to_points[cid]
                        seen[cid] = (off, len(pts))
                        cells[n, off:off + len(pts)] = np.array(pts, dtype=np.int32)
                        off += len(pts)
                    comp_start[n, a], comp_size[n, a] = seen[cid]
        return cls(cfg, obstacles, pos, tgt, rng, comp_start, comp_size, cells, stride)

    @classmethod
    def from_python_oracle_parallel(cls, gc_kwargs, seeds, procs=None):
        """`from_python_oracle` for many instances: the Python oracle resets (numpy RNG, BFS, placing in pure
        Python) run in a process pool, the per-chunk arrays are concatenated into one COracle."""
        import multiprocessing as mp
        seeds = [int(s) for s in seeds]
        procs = min(procs or (os.cpu_count() or 1), max(1, len(seeds) // 8))
        if procs <= 1:
            return cls.from_python_oracle(gc_kwargs, seeds)
        chunks = [seeds[i::procs] for i in range(procs)]
        with mp.get_context("spawn").Pool(procs) as pool:
            parts = pool.map(_build_chunk, [(gc_kwargs, c) for c in chunks])
        order = np.argsort(np.concatenate([np.arange(len(seeds))[i::procs] for i in range(procs)]), kind="stable")

        def cat(key):
            if parts[0][key] is None:
                return None
            return np.ascontiguousarray(np.concatenate([p[key] for p in parts])[order])
        cfg = OrcCfg(*parts[0]["cfg"])
        return cls(cfg, cat("obstacles"), cat("pos"), cat("tgt"), cat("rng"), cat("comp_start"), cat("comp_size"),
                   cat("cells"), parts[0]["cells_stride"])

    def run(self, actions, auto_reset=False, want_obs=True):
        """actions uint8 [T, N, A] -> dict with the outputs of the LAST step and the reward sums."""
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        T = actions.shape[0]
        assert actions.shape == (T, self.N, self.A)
        obs = np.zeros((self.N, self.A, 3, self.D, self.D), dtype=np.uint8) if want_obs else None
        rsum = np.zeros((self.N, self.A), dtype=np.float64)
        rew = np.zeros((self.N, self.A), dtype=np.float32)
        term = np.zeros((self.N, self.A), dtype=np.uint8)
        trunc = np.zeros((self.N, self.A), dtype=np.uint8)
        done = self.lib.orc_run(C.byref(self.cfg), self.N, T, int(auto_reset), _p(self.obstacles), _p(self.pos),
                                _p(self.tgt), _p(self.active), _p(self.elapsed), _p(self.pos0), _p(self.tgt0),
                                _p(self.rng), _p(self.rng0), _p(self.comp_start), _p(self.comp_size), _p(self.cells),
                                C.c_longlong(self.cells_stride), _p(actions), _p(obs), _p(rsum), _p(rew), _p(term),
                                _p(trunc))
        return dict(obs=obs, rewards_sum=rsum, rewards=rew, terminated=term.astype(bool), truncated=trunc.astype(bool),
                    agent_steps=done)
