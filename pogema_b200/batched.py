"""``BatchedPogema`` - the batched vector env of BASELINE.json's north star:
N independent POGEMA instances advanced by ONE fused sm_100a kernel per step,
inputs and outputs as torch CUDA tensors (zero copies on the step path)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as nat
from .engine import Engine
from .grid_config import GridConfig


class _DevArray:
    """Zero-copy view of an engine-owned device array (``__cuda_array_interface__``)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}
        self._owner = owner


class BatchedPogema:
    """Vector env over ``num_envs`` instances of ``grid_config``.

    Instance k is built exactly like ``pogema_v0(GridConfig(..., seed=seeds[k]))``
    (default ``seeds[k] = (grid_config.seed or 0) + k``).

    ``reset() -> obs``; ``step(actions) -> obs, rewards, terminated, truncated``:
      obs         uint8  [N, A, 3, D, D]   (uint32 [N, A, ceil(3*D*D/32)] with obs_format='bits',
                                          float32 [N, A, 3, D, D] - the reference's dtype - with obs_format='f32',
                                          float16 with obs_format='f16': what obs.half() would give, without the pass)
      rewards     float32 [N, A]
      terminated  bool   [N, A]
      truncated   bool   [N, A]
    The returned tensors are owned by the env and overwritten by the next call
    (pass ``out=`` to step into caller-owned tensors).  With ``auto_reset=True`` an
    instance whose episode ended (all terminated or all truncated) is restored to
    its initial task inside the same step and ``obs`` holds the reset observation
    (upstream integrations/sample_factory.py :: AutoResetWrapper semantics).
    ``auto_reset="reseed"`` rebuilds such an instance on the device from a NEW seed
    instead (``seed += reseed_stride``, default ``num_envs``): a fresh random map and
    task for every episode, exactly ``pogema_v0(GridConfig(seed=new_seed)).reset()``.
    """

    def __init__(self, grid_config: Optional[GridConfig] = None, num_envs: int = 1, device="cuda",
                 seeds: Optional[Sequence[int]] = None, auto_reset=True, obs_format: str = "u8",
                 team_threads: int = 0, num_threads: int = 0, generate_on_device: bool = True,
                 reseed_stride: int = 0, **kwargs):
        if grid_config is None:
            grid_config = GridConfig(**kwargs)
        elif isinstance(grid_config, dict):
            grid_config = GridConfig(**grid_config)
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedPogema needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("device must be a CUDA device")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.grid_config = grid_config
        self.num_envs = int(num_envs)
        self.num_agents = int(grid_config.num_agents)
        self.auto_reset = auto_reset
        self.engine = Engine(grid_config, num_envs, device=dev_index, auto_reset=auto_reset,
                             obs_format=obs_format, team_threads=team_threads, reseed_stride=reseed_stride)
        if seeds is None:
            base = grid_config.seed or 0
            seeds = np.arange(base, base + self.num_envs, dtype=np.uint64)
        self.seeds = np.asarray(seeds, dtype=np.uint64)
        assert len(self.seeds) == self.num_envs
        self.generate_on_device = bool(generate_on_device) and grid_config.agents_xy is None
        self.engine.generate(self.seeds, num_threads=num_threads, stream=self._stream(),
                             on_device=self.generate_on_device)
        n, a = self.num_envs, self.num_agents
        with torch.cuda.device(self.device):
            self._obs = self._alloc_obs()
            self._rewards = torch.empty((n, a), dtype=torch.float32, device=self.device)
            self._terminated = torch.empty((n, a), dtype=torch.bool, device=self.device)
            self._truncated = torch.empty((n, a), dtype=torch.bool, device=self.device)
        self.obs_shape = tuple(self._obs.shape[1:])

    @classmethod
    def groups(cls, grid_config: Optional[GridConfig], num_envs: int, groups: int = 2, device="cuda",
               seeds: Optional[Sequence[int]] = None, **kwargs):
        """``groups`` envs over contiguous shares of ``num_envs`` instances, each with its own CUDA stream:
        ``[(env, stream), ...]``.  For closed loops (a policy between steps): step group 0 on its stream while the
        policy works on group 1's observations and vice versa (the double-buffered sampling of RL frameworks).  A
        single env's one-launch-per-step form leaves HBM idle during the dependent front of every step (state loads,
        move resolution, first bit assembly: 0.85 of the roofline on configs[1]); with two groups one group's front
        overlaps the other's observation stores (0.96, four groups 0.99 - ``closed_loop.groups`` in the bench line).
        Instance k of the job keeps seed ``seeds[k]``, so the results do not depend on the number of groups."""
        if grid_config is None:
            grid_config = GridConfig(**{k: kwargs.pop(k) for k in list(kwargs) if k in GridConfig.model_fields})
        elif isinstance(grid_config, dict):
            grid_config = GridConfig(**grid_config)
        if num_envs % groups != 0:
            raise ValueError("num_envs must be a multiple of groups")
        if seeds is None:
            base = grid_config.seed or 0
            seeds = np.arange(base, base + num_envs, dtype=np.uint64)
        seeds = np.asarray(seeds, dtype=np.uint64)
        dev = torch.device(device)
        n_g = num_envs // groups
        out = []
        for k in range(groups):
            st = torch.cuda.Stream(dev)
            with torch.cuda.stream(st):
                out.append((cls(grid_config, n_g, device=device, seeds=seeds[k * n_g:(k + 1) * n_g], **kwargs), st))
        return out

    # ------------------------------------------------------------------ #
    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def _alloc_obs(self) -> torch.Tensor:
        e = self.engine
        dtype = {"bits": torch.int32, "f32": torch.float32, "f16": torch.float16}.get(e.obs_format, torch.uint8)
        return torch.empty(e.obs_shape(), dtype=dtype, device=self.device)

    def new_obs_buffer(self) -> torch.Tensor:
        return self._alloc_obs()

    # ------------------------------------------------------------------ #
    def reset(self, out: Optional[torch.Tensor] = None, seeds: Optional[Sequence[int]] = None) -> torch.Tensor:
        """Restore every instance to its initial task (upstream: same seed -> same map/task).  With
        ``seeds`` the instances are REBUILT for the new seeds by the device-side generator
        (``pgm_generate_device``), i.e. ``pogema_v0(GridConfig(seed=seeds[k])).reset()`` for every k."""
        obs = self._obs if out is None else out
        if seeds is not None:
            self.seeds = np.asarray(seeds, dtype=np.uint64)
            assert len(self.seeds) == self.num_envs
            self.engine.generate(self.seeds, stream=self._stream(), on_device=self.generate_on_device)
        self.engine.reset(obs.data_ptr(), self._stream())
        return obs

    def observe(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        obs = self._obs if out is None else out
        self.engine.observe(obs.data_ptr(), self._stream())
        return obs

    def step(self, actions: torch.Tensor, out: Optional[torch.Tensor] = None, compute_obs: bool = True):
        if actions.device != self.device:
            raise ValueError(f"actions must live on {self.device}")
        if actions.shape != (self.num_envs, self.num_agents):
            raise ValueError(f"actions must have shape {(self.num_envs, self.num_agents)}")
        if actions.dtype not in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            raise TypeError("actions must be an integer tensor")
        actions = actions.contiguous()
        obs = self._obs if out is None else out
        self.engine.step(actions.data_ptr(), actions.element_size(), obs.data_ptr() if compute_obs else 0,
                         self._rewards.data_ptr(), self._terminated.data_ptr(), self._truncated.data_ptr(),
                         self._stream())
        return obs, self._rewards, self._terminated, self._truncated

    def step_host(self, actions: np.ndarray):
        """``step`` with HOST buffers (``pgm_step_host``): numpy actions in, numpy obs / rewards / terminated /
        truncated out (copies host<->device inside the call; page-locked result buffers are reused)."""
        if not hasattr(self, "_h_out"):
            e = self.engine
            odt = {"bits": torch.int32, "f32": torch.float32, "f16": torch.float16}.get(e.obs_format, torch.uint8)
            n, a = self.num_envs, self.num_agents
            self._h_out = (torch.empty(e.obs_shape(), dtype=odt).pin_memory(),
                           torch.empty((n, a), dtype=torch.float32).pin_memory(),
                           torch.empty((n, a), dtype=torch.uint8).pin_memory(),
                           torch.empty((n, a), dtype=torch.uint8).pin_memory())
        obs, rew, term, trunc = self._h_out
        actions = np.ascontiguousarray(actions)
        if actions.shape != (self.num_envs, self.num_agents):
            raise ValueError(f"actions must have shape {(self.num_envs, self.num_agents)}")
        self.engine.step_host(actions, obs.numpy(), rew.numpy(), term.numpy(), trunc.numpy(), self._stream())
        return obs.numpy(), rew.numpy(), term.numpy().astype(bool), trunc.numpy().astype(bool)

    def rollout(self, actions: torch.Tensor, obs_out: Optional[torch.Tensor] = None, compute_obs: bool = True):
        """K consecutive steps in ONE kernel launch (``pgm_step_many``) for actions known in advance.

        actions: int tensor [K, N, A].  Returns (obs [R, N, A, ...], rewards [K, N, A], terminated [K, N, A],
        truncated [K, N, A]); step k writes observation slot ``k % R`` where R = ``obs_out.shape[0]`` (default:
        R = K, every step's observation is kept).  Results are identical to K ``step`` calls."""
        if actions.device != self.device or actions.dim() != 3 or tuple(actions.shape[1:]) != (self.num_envs, self.num_agents):
            raise ValueError(f"actions must be a [K, {self.num_envs}, {self.num_agents}] tensor on {self.device}")
        if actions.dtype not in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            raise TypeError("actions must be an integer tensor")
        actions = actions.contiguous()
        K = actions.shape[0]
        with torch.cuda.device(self.device):
            if obs_out is None and compute_obs:
                e = self.engine
                obs_out = torch.empty((K,) + tuple(e.obs_shape()), dtype=self._obs.dtype, device=self.device)
            rew = torch.empty((K, self.num_envs, self.num_agents), dtype=torch.float32, device=self.device)
            term = torch.empty((K, self.num_envs, self.num_agents), dtype=torch.bool, device=self.device)
            trunc = torch.empty((K, self.num_envs, self.num_agents), dtype=torch.bool, device=self.device)
        ring = int(obs_out.shape[0]) if compute_obs else 1
        self.engine.step_many(K, actions.data_ptr(), actions.element_size(), obs_out.data_ptr() if compute_obs else 0,
                              ring, rew.data_ptr(), term.data_ptr(), trunc.data_ptr(), self._stream())
        return obs_out, rew, term, trunc

    def sample_actions(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        return torch.randint(0, 5, (self.num_envs, self.num_agents), dtype=torch.uint8, device=self.device,
                             generator=generator)

    # -- zero-copy state views ------------------------------------------- #
    def _view(self, what, shape, typestr):
        ptr = self.engine.state_ptr(what)
        return torch.as_tensor(_DevArray(ptr, shape, typestr, self), device=self.device)

    def _state(self) -> torch.Tensor:
        """Raw agent state words int32 [N, A, 2] (include/pgm_b200.h, PGM_STATE_POSITIONS)."""
        return self._view(nat.STATE_POSITIONS, (self.num_envs, self.num_agents, 2), "<i4")

    @property
    def is_active(self) -> torch.Tensor:
        return ((self._state()[..., 0] >> 15) & 1).bool()

    @property
    def was_on_goal(self) -> torch.Tensor:
        return self._view(nat.STATE_WAS_ON_GOAL, (self.num_envs, self.num_agents), "|u1").bool()

    @property
    def episode_done(self) -> torch.Tensor:
        return self._view(nat.STATE_EPISODE_DONE, (self.num_envs,), "|u1").bool()

    @property
    def elapsed_steps(self) -> torch.Tensor:
        return self._view(nat.STATE_ELAPSED, (self.num_envs,), "<i4")

    def get_agents_xy(self) -> torch.Tensor:
        """int32 [N, A, 2] unpadded (x, y) positions."""
        w = self._state()[..., 0]
        r = self.grid_config.obs_radius
        return torch.stack(((w & 0x7FFF) - r, ((w >> 16) & 0xFFFF) - r), dim=-1)

    def get_targets_xy(self) -> torch.Tensor:
        w = self._state()[..., 1]
        r = self.grid_config.obs_radius
        return torch.stack(((w & 0x7FFF) - r, ((w >> 16) & 0xFFFF) - r), dim=-1)

    def current_seeds(self) -> np.ndarray:
        """uint64 [N]: the seed each instance's current task was built from (advances with auto_reset='reseed')."""
        return self.engine.get_state(nat.STATE_SEEDS, self._stream())

    def get_obstacles(self) -> np.ndarray:
        """uint8 [N, H, W] unpadded obstacle maps (host array)."""
        return self.engine.get_state(nat.STATE_OBSTACLES)

    def metrics(self) -> dict:
        """Metrics of the last finished episode of every instance (upstream
        wrappers/metrics.py), computed from the raw device counters."""
        raw = self.engine.get_state(nat.STATE_METRICS, self._stream()).astype(np.float64)
        a = float(self.num_agents)
        ot = self.grid_config.on_target
        if ot == 'restart':
            return {"avg_throughput": raw[:, 0] / self.grid_config.max_episode_steps}
        if ot == 'nothing':
            # SoC / makespan: upstream wrappers/metrics.py :: SumOfCostsAndMakespanMetric (per-agent costs latched by
            # the step kernel at the end of the episode; zeros until an instance has finished one)
            cost = self.engine.get_state(nat.STATE_SOLVE_COSTS, self._stream()).astype(np.float64)
            fin = raw[:, 2] > 0
            return {"ISR": raw[:, 3] / a, "CSR": (raw[:, 3] == a).astype(np.float64), "ep_length": raw[:, 2],
                    "SoC": np.where(fin, cost.sum(axis=1) + a, 0.0), "makespan": np.where(fin, cost.max(axis=1) + 1, 0.0)}
        return {"ISR": raw[:, 0] / a, "CSR": (raw[:, 0] == a).astype(np.float64), "ep_length": raw[:, 1] / a + 1}

    # -- checkpoint / resume ------------------------------------------------ #
    def state_dict(self) -> dict:
        return {"engine": self.engine.checkpoint(self._stream()), "seeds": self.seeds.copy()}

    def load_state_dict(self, sd: dict):
        """Restore a ``state_dict()``.  The engine validates the blob's header (shape, modes) and - unless the tasks
        travel inside the blob (``auto_reset="reseed"``) - that it was taken from the same task seeds."""
        if not np.array_equal(sd["seeds"], self.seeds):
            raise ValueError("checkpoint belongs to different tasks (initial seeds differ)")
        self.engine.restore(sd["engine"], self._stream())

    def check_errors(self):
        self.engine.check_errors(self._stream())

    def close(self):
        self.engine.close()
