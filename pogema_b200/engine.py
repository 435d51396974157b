"""Thin, torch-free Python owner of one ``pgm_engine`` handle (include/pgm_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as nat
from .grid_config import GridConfig


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """N instances of one GridConfig shape on one CUDA device."""

    def __init__(self, grid_config: GridConfig, num_envs: int, device: int = 0, auto_reset=False,
                 obs_format: str = "u8", team_threads: int = 0, reseed_stride: int = 0):
        self.lib = nat.load()
        gc = grid_config
        h, w = gc.map_shape()
        cfg = nat.PgmConfig()
        cfg.abi_version = nat.PGM_ABI_VERSION
        cfg.device = int(device)
        cfg.num_envs = int(num_envs)
        cfg.num_agents = int(gc.num_agents)
        cfg.height, cfg.width = int(h), int(w)
        cfg.obs_radius = int(gc.obs_radius)
        cfg.max_episode_steps = int(gc.max_episode_steps)
        cfg.collision_system = nat.COLLISION[gc.collision_system]
        cfg.on_target = nat.ON_TARGET[gc.on_target]
        cfg.auto_reset = 2 if auto_reset == "reseed" else (1 if auto_reset else 0)
        cfg.reserved[0] = int(reseed_stride)
        cfg.obs_format = nat.OBS_FORMAT[obs_format]
        cfg.team_threads = int(team_threads)
        self.cfg = cfg
        self.grid_config = gc
        self.num_envs, self.num_agents = int(num_envs), int(gc.num_agents)
        self.height, self.width = int(h), int(w)
        self.obs_radius = int(gc.obs_radius)
        self.D = 2 * self.obs_radius + 1
        self.obs_format = obs_format
        self.device = int(device)
        handle = C.c_void_p()
        nat.check(self.lib.pgm_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        self.obs_bytes = int(self.lib.pgm_obs_bytes(handle))
        self.obs_instance_stride = int(self.lib.pgm_obs_instance_stride(handle))

    # -- lifetime ---------------------------------------------------------- #
    def close(self):
        if getattr(self, "handle", None):
            self.lib.pgm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- task construction -------------------------------------------------- #
    def generate(self, seeds: Sequence[int], first: int = 0, num_threads: int = 0, stream: int = 0,
                 on_device: bool = False):
        """upstream Grid.__init__ for instances [first, first+len(seeds)) (see pgm_generate)."""
        gc = self.grid_config
        seeds = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        failed = C.c_int32(-1)
        if gc.agents_xy is not None and gc.targets_xy is not None:
            count = len(seeds)
            m = gc.map_array()
            if m is None:
                # explicit agents on a generated map: the obstacles still come from the seed through the
                # reference's own RNG call (upstream generator.py :: generate_obstacles), drawn on the host
                obst = np.stack([np.random.default_rng(int(s)).binomial(1, gc.density, (gc.size, gc.size))
                                 for s in seeds]).astype(np.uint8)
            else:
                obst = np.ascontiguousarray(np.broadcast_to(m, (count,) + m.shape))
            axy = np.ascontiguousarray(np.broadcast_to(np.asarray(gc.agents_xy, dtype=np.int32),
                                                       (count, self.num_agents, 2)))
            txy = np.ascontiguousarray(np.broadcast_to(np.asarray(gc.targets_xy, dtype=np.int32),
                                                       (count, self.num_agents, 2)))
            nat.check(self.lib.pgm_set_tasks(self.handle, first, count, _ptr(obst), _ptr(axy), _ptr(txy),
                                             _ptr(seeds), C.c_void_p(stream)))
            return
        m = gc.map_array()
        if on_device:
            nfb = C.c_int32(0)
            nat.check(self.lib.pgm_generate_device(self.handle, first, len(seeds), _ptr(seeds), float(gc.density),
                                                   _ptr(m), C.byref(nfb), C.c_void_p(stream)))
            self.last_host_fallbacks = int(nfb.value)
            return
        nat.check(self.lib.pgm_generate(self.handle, first, len(seeds), _ptr(seeds), float(gc.density),
                                        _ptr(m), num_threads, C.byref(failed), C.c_void_p(stream)))

    def set_tasks(self, obstacles, agents_xy, targets_xy, seeds=None, first: int = 0, stream: int = 0):
        obstacles = np.ascontiguousarray(np.asarray(obstacles, dtype=np.uint8))
        agents_xy = np.ascontiguousarray(np.asarray(agents_xy, dtype=np.int32))
        targets_xy = np.ascontiguousarray(np.asarray(targets_xy, dtype=np.int32))
        count = obstacles.shape[0]
        assert obstacles.shape == (count, self.height, self.width)
        assert agents_xy.shape == (count, self.num_agents, 2) and targets_xy.shape == agents_xy.shape
        s = None if seeds is None else np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        nat.check(self.lib.pgm_set_tasks(self.handle, first, count, _ptr(obstacles), _ptr(agents_xy),
                                         _ptr(targets_xy), _ptr(s), C.c_void_p(stream)))

    # -- device pointer API ---------------------------------------------------- #
    def reset(self, obs_ptr: int = 0, stream: int = 0):
        nat.check(self.lib.pgm_reset(self.handle, C.c_void_p(obs_ptr), C.c_void_p(stream)))

    def observe(self, obs_ptr: int, stream: int = 0):
        nat.check(self.lib.pgm_observe(self.handle, C.c_void_p(obs_ptr), C.c_void_p(stream)))

    def step(self, actions_ptr: int, itemsize: int, obs_ptr: int, rewards_ptr: int, terminated_ptr: int,
             truncated_ptr: int, stream: int = 0):
        nat.check(self.lib.pgm_step(self.handle, C.c_void_p(actions_ptr), itemsize, C.c_void_p(obs_ptr),
                                    C.c_void_p(rewards_ptr), C.c_void_p(terminated_ptr),
                                    C.c_void_p(truncated_ptr), C.c_void_p(stream)))

    def step_many(self, num_steps: int, actions_ptr: int, itemsize: int, obs_ptr: int, obs_ring: int,
                  rewards_ptr: int, terminated_ptr: int, truncated_ptr: int, stream: int = 0):
        nat.check(self.lib.pgm_step_many(self.handle, num_steps, C.c_void_p(actions_ptr), itemsize,
                                         C.c_void_p(obs_ptr), obs_ring, C.c_void_p(rewards_ptr),
                                         C.c_void_p(terminated_ptr), C.c_void_p(truncated_ptr), C.c_void_p(stream)))

    # -- host buffer API --------------------------------------------------------- #
    def obs_shape(self):
        if self.obs_format == "bits":
            return (self.num_envs, self.num_agents, (3 * self.D * self.D + 31) // 32)
        return (self.num_envs, self.num_agents, 3, self.D, self.D)

    def obs_dtype(self):
        return {"bits": np.uint32, "f32": np.float32, "f16": np.float16}.get(self.obs_format, np.uint8)

    def step_host(self, actions: np.ndarray, obs: Optional[np.ndarray], rewards: np.ndarray,
                  terminated: np.ndarray, truncated: np.ndarray, stream: int = 0,
                  active: Optional[np.ndarray] = None, was_on_goal: Optional[np.ndarray] = None):
        """pgm_step_host / pgm_step_host_ex: one step with host buffers; ``active`` / ``was_on_goal`` (uint8 [N, A],
        optional) receive upstream's ``grid.is_active`` / ``env.was_on_goal`` in the same synchronisation."""
        actions = np.ascontiguousarray(actions)
        if actions.dtype.kind not in "iu" or actions.itemsize not in (1, 2, 4, 8):
            raise TypeError(f"actions must be an integer array, got {actions.dtype}")
        if actions.size != self.num_envs * self.num_agents:
            raise ValueError(f"actions must hold {self.num_envs} x {self.num_agents} elements, got {actions.size}")
        nat.check(self.lib.pgm_step_host_ex(self.handle, _ptr(actions), actions.itemsize, _ptr(obs), _ptr(rewards),
                                            _ptr(terminated), _ptr(truncated), _ptr(active), _ptr(was_on_goal),
                                            C.c_void_p(stream)))

    def set_host_transport(self, mode="auto", num_threads: int = 0):
        """How step_host / observe_host bring observations to the host (pgm_set_host_transport):
        'plain' = DMA of the final tensor, 'packed' = DMA of the GPU-written bit stream + host threads that
        widen it into the caller's buffer, 'auto' = packed for tensors of 4 MB and more."""
        m = {"auto": -1, "plain": 0, "packed": 1}[mode]
        nat.check(self.lib.pgm_set_host_transport(self.handle, m, int(num_threads)))

    def host_transport_info(self) -> dict:
        out = (C.c_int64 * 10)()
        nat.check(self.lib.pgm_host_transport_info(self.handle, out, 10))
        return {"packed": bool(out[0]), "threads": int(out[1]), "h2d_bytes": int(out[2]), "d2h_bytes": int(out[3]),
                "isa": ["scalar", "avx2", "avx512bw"][int(out[4])],
                "timeline_us": dict(zip(("enqueued", "first_chunk", "last_chunk", "widened", "returned"),
                                        [int(v) for v in out[5:10]]))}

    def observe_host(self, stream: int = 0) -> np.ndarray:
        """Observation of the current state as a host array (pgm_observe_host)."""
        out = np.empty(self.obs_shape(), dtype=self.obs_dtype())
        nat.check(self.lib.pgm_observe_host(self.handle, _ptr(out), C.c_void_p(stream)))
        return out

    # -- state ----------------------------------------------------------------------- #
    def get_state(self, what: int, stream: int = 0) -> np.ndarray:
        n, a = self.num_envs, self.num_agents
        shapes = {
            nat.STATE_POSITIONS: ((n, a, 2), np.int32), nat.STATE_TARGETS: ((n, a, 2), np.int32),
            nat.STATE_ACTIVE: ((n, a), np.uint8), nat.STATE_ELAPSED: ((n,), np.int32),
            nat.STATE_OBSTACLES: ((n, self.height, self.width), np.uint8),
            nat.STATE_WAS_ON_GOAL: ((n, a), np.uint8), nat.STATE_EPISODE_DONE: ((n,), np.uint8),
            nat.STATE_METRICS: ((n, 4), np.int32), nat.STATE_SEEDS: ((n,), np.uint64),
            nat.STATE_SOLVE_COSTS: ((n, a), np.int32),
        }
        shape, dtype = shapes[what]
        out = np.empty(shape, dtype=dtype)
        nat.check(self.lib.pgm_get_state(self.handle, what, _ptr(out), out.nbytes, C.c_void_p(stream)))
        return out

    def state_ptr(self, what: int) -> int:
        return int(self.lib.pgm_state_ptr(self.handle, what) or 0)

    def checkpoint(self, stream: int = 0) -> np.ndarray:
        nbytes = int(self.lib.pgm_checkpoint_bytes(self.handle))
        buf = np.empty(nbytes, dtype=np.uint8)
        nat.check(self.lib.pgm_checkpoint_save(self.handle, _ptr(buf), nbytes, C.c_void_p(stream)))
        return buf

    def restore(self, buf: np.ndarray, stream: int = 0):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        nat.check(self.lib.pgm_checkpoint_load(self.handle, _ptr(buf), buf.nbytes, C.c_void_p(stream)))

    def check_errors(self, stream: int = 0):
        nat.check(self.lib.pgm_check_errors(self.handle, C.c_void_p(stream)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.pgm_launch_count(self.handle))

    def plan(self) -> dict:
        out = (C.c_int32 * 13)()
        nat.check(self.lib.pgm_plan(self.handle, out, 13))
        keys = ["team_threads", "teams_per_cta", "cta_threads", "smem_bytes_per_cta", "grid",
                "agents_per_obs_batch", "occupancy_buckets"]
        plan = dict(zip(keys, [int(v) for v in out[:7]]))
        # step launches of the common shapes run the register-resident kernel (pgm_fast.cuh) with its own geometry
        plan["fast_step_kernel"] = bool(out[7])
        if out[7]:
            plan["fast"] = dict(zip(["team_threads", "agents_per_thread", "teams_per_cta", "smem_bytes_per_cta", "grid"],
                                    [int(v) for v in out[8:13]]))
        return plan
