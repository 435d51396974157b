"""List based single-instance environments - drop-in for upstream
``pogema_v0(grid_config)`` (upstream integrations/make_pogema.py :: make_pogema,
envs.py :: Pogema / PogemaLifeLong / PogemaCoopFinish, wrappers/multi_time_limit.py,
wrappers/metrics.py).  One ``Engine`` with a single instance does all the work:
every ``step`` is one launch of the fused sm_100a kernel through the C-ABI host
buffer call (``pgm_step_host``); nothing is computed on the CPU except turning
arrays into the Python lists the upstream API returns."""
from __future__ import annotations

import os
from copy import deepcopy
from typing import Optional

import numpy as np

from . import _native as nat
from .engine import Engine
from .grid_config import GridConfig

try:  # gymnasium is optional (absent in this image): fall back to minimal space shims
    import gymnasium
    from gymnasium.spaces import Box, Discrete
    _Base = gymnasium.Env
except Exception:  # pragma: no cover - exercised when gymnasium is missing
    gymnasium = None

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.int64

        def sample(self):
            return int(np.random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

    class Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def contains(self, x):
            return tuple(np.shape(x)) == self.shape

    _Base = object


class ActionsSampler:
    """upstream envs.py :: ActionsSampler"""

    def __init__(self, num_actions, seed=42):
        self._num_actions = num_actions
        self._rnd = None
        self.update_seed(seed)

    def update_seed(self, seed=None):
        self._rnd = np.random.default_rng(seed)

    def sample_actions(self, dim=1):
        return self._rnd.integers(self._num_actions, size=dim)


class GridView:
    """Read-only stand-in for upstream ``grid.py :: Grid`` built from device state
    (``env.grid.get_agents_xy()`` etc. keep working)."""

    def __init__(self, env):
        self._env = env
        self.config = env.grid_config

    def _state(self, what):
        return self._env._engine.get_state(what)[0]

    @property
    def obstacles(self):
        """Padded obstacle array (float, like upstream after add_artificial_border)."""
        gc = self.config
        r = gc.obs_radius
        inner = self._state(nat.STATE_OBSTACLES).astype(np.float64)
        h, w = inner.shape
        full = np.zeros((h + 2 * r, w + 2 * r))
        full[r - 1, r - 1:w + r + 1] = gc.OBSTACLE
        full[r - 1:h + r + 1, r - 1] = gc.OBSTACLE
        full[h + r, r - 1:w + r + 1] = gc.OBSTACLE
        full[r - 1:h + r + 1, w + r] = gc.OBSTACLE
        full[r:h + r, r:w + r] = inner
        return full

    @property
    def positions_xy(self):
        r = self.config.obs_radius
        return [(int(x) + r, int(y) + r) for x, y in self._state(nat.STATE_POSITIONS)]

    @property
    def finishes_xy(self):
        r = self.config.obs_radius
        return [(int(x) + r, int(y) + r) for x, y in self._state(nat.STATE_TARGETS)]

    @property
    def is_active(self):
        return {i: bool(v) for i, v in enumerate(self._state(nat.STATE_ACTIVE))}

    @property
    def positions(self):
        occ = np.zeros_like(self.obstacles)
        for (x, y), act in zip(self.positions_xy, self.is_active.values()):
            if act:
                occ[x, y] = self.config.OBSTACLE
        return occ

    def get_obstacles(self, ignore_borders=False):
        if ignore_borders:
            return self._state(nat.STATE_OBSTACLES).astype(np.float64)
        return self.obstacles

    def _prepare(self, what, only_active, ignore_borders):
        r = 0 if ignore_borders else self.config.obs_radius
        pts = [[int(x) + r, int(y) + r] for x, y in self._state(what)]
        if only_active:
            act = self._state(nat.STATE_ACTIVE)
            pts = [p for p, a in zip(pts, act) if a]
        return pts

    def get_agents_xy(self, only_active=False, ignore_borders=False):
        return self._prepare(nat.STATE_POSITIONS, only_active, ignore_borders)

    def get_targets_xy(self, only_active=False, ignore_borders=False):
        return self._prepare(nat.STATE_TARGETS, only_active, ignore_borders)

    def get_agents_xy_relative(self):
        return self._relative(self.get_agents_xy())

    def get_targets_xy_relative(self):
        return self._relative(self.get_targets_xy())

    def _relative(self, pts):
        r = self.config.obs_radius
        start = [[x + r, y + r] for x, y in self._env._initial_xy]
        return [[x - sx, y - sy] for (x, y), (sx, sy) in zip(pts, start)]

    def on_goal(self, agent_id):
        return self.positions_xy[agent_id] == self.finishes_xy[agent_id]

    def is_active_agent(self, agent_id):
        return self.is_active[agent_id]

    def get_state(self, ignore_borders=False, as_dict=False):
        obstacles = self.get_obstacles(ignore_borders)
        agents_xy = self.get_agents_xy(ignore_borders=ignore_borders)
        targets_xy = self.get_targets_xy(ignore_borders=ignore_borders)
        active = [self.is_active[i] for i in range(len(agents_xy))]
        if as_dict:
            return {"obstacles": obstacles, "agents_xy": agents_xy, "targets_xy": targets_xy, "active": active}
        return obstacles, agents_xy, targets_xy, active


class PogemaBase(_Base):
    """One POGEMA instance with the upstream list based API (upstream envs.py :: PogemaBase + the step logic of
    its three subclasses, which lives in the kernel: ``grid_config.on_target`` selects it).  The time limit
    (MultiTimeLimit) and the metric wrappers are part of the same object.  Use the subclasses ``Pogema`` /
    ``PogemaLifeLong`` / ``PogemaCoopFinish`` (or ``pogema_v0``) - like upstream's, each of them fixes its own
    ``on_target`` semantics whatever the config says."""

    metadata = {"render_modes": ["ansi"]}
    _ON_TARGET = None  # subclasses: the on_target mode the class stands for

    def __init__(self, grid_config: Optional[GridConfig] = None, device: int = 0, **kwargs):
        if grid_config is None:
            grid_config = GridConfig(**kwargs)
        elif isinstance(grid_config, dict):
            grid_config = GridConfig(**grid_config)
        if self._ON_TARGET is not None and grid_config.on_target != self._ON_TARGET:
            # upstream picks the semantics by CLASS: PogemaLifeLong(GridConfig(on_target='finish')) is lifelong
            grid_config = grid_config.model_copy(update=dict(on_target=self._ON_TARGET))
        self.grid_config = grid_config
        self._device = int(device)
        self._engine = None
        self._engine_key = None
        self.grid = None
        self.was_on_goal = None
        self._initial_xy = None
        full = grid_config.obs_radius * 2 + 1
        self.action_space = Discrete(len(grid_config.MOVES))
        self.observation_space = self._make_observation_space(grid_config, full)
        self._multi_action_sampler = ActionsSampler(self.action_space.n, seed=grid_config.seed)
        self._elapsed_steps = None

    @staticmethod
    def _make_observation_space(gc, full):
        """upstream envs.py :: Pogema.__init__: Box(3, D, D) for 'default', Dict spaces for POMAPF / MAPF."""
        if gc.observation_type == 'default':
            return Box(0.0, 1.0, shape=(3, full, full), dtype=np.float32)
        spaces = dict(obstacles=Box(0.0, 1.0, shape=(full, full), dtype=np.float32),
                      agents=Box(0.0, 1.0, shape=(full, full), dtype=np.float32),
                      xy=Box(low=-1024, high=1024, shape=(2,), dtype=int),
                      target_xy=Box(low=-1024, high=1024, shape=(2,), dtype=int))
        if gc.observation_type == 'MAPF':
            h, w = gc.map_shape()
            r = gc.obs_radius
            spaces.update(global_obstacles=Box(0.0, 1.0, shape=(h + 2 * r, w + 2 * r), dtype=np.float32),
                          global_xy=Box(low=-1024, high=1024, shape=(2,), dtype=int),
                          global_target_xy=Box(low=-1024, high=1024, shape=(2,), dtype=int))
        if gymnasium is not None:
            return gymnasium.spaces.Dict(**spaces)
        return spaces

    # -- helpers --------------------------------------------------------- #
    def _ensure_engine(self):
        gc = self.grid_config
        key = (gc.num_agents, gc.map_shape(), gc.obs_radius, gc.max_episode_steps, gc.collision_system, gc.on_target)
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(gc, 1, device=self._device, auto_reset=False)
            self._engine_key = key
            n, a = 1, gc.num_agents
            self._h_obs = np.empty(self._engine.obs_shape(), dtype=np.uint8)
            self._h_rew = np.empty((n, a), dtype=np.float32)
            self._h_term = np.empty((n, a), dtype=np.uint8)
            self._h_trunc = np.empty((n, a), dtype=np.uint8)
            self._h_active = np.ones((n, a), dtype=np.uint8)
            self._h_was = np.zeros((n, a), dtype=np.uint8)
            self._h_act = np.zeros((n, a), dtype=np.int64)
            # the buffers never move: bind the C call once (building seven ctypes pointers per step costs ~7 us)
            import ctypes as C
            vp = lambda arr: arr.ctypes.data_as(C.c_void_p)
            self._step_args = (self._engine.handle, vp(self._h_act), self._h_act.itemsize, vp(self._h_obs),
                               vp(self._h_rew), vp(self._h_term), vp(self._h_trunc), vp(self._h_active),
                               vp(self._h_was), C.c_void_p(0))
        self._engine.grid_config = gc

    def _obs_list(self, obs_u8):
        """upstream envs.py :: Pogema._obs: (3, D, D) float32 arrays, or the POMAPF / MAPF dicts built from the
        same device-made crops plus the state read back from the engine."""
        n = self.grid_config.num_agents
        kind = self.grid_config.observation_type
        if kind == 'default':
            return list(obs_u8[0].astype(np.float32))  # one conversion; fresh arrays every step like upstream
        pos = self._engine.get_state(nat.STATE_POSITIONS)[0]
        tgt = self._engine.get_state(nat.STATE_TARGETS)[0]
        start = self._initial_xy
        results = []
        for i in range(n):
            results.append({'obstacles': obs_u8[0, i, 0].astype(np.float32),
                            'agents': obs_u8[0, i, 1].astype(np.float32),
                            'xy': (int(pos[i][0]) - start[i][0], int(pos[i][1]) - start[i][1]),
                            'target_xy': (int(tgt[i][0]) - start[i][0], int(tgt[i][1]) - start[i][1])})
        if kind == 'MAPF':
            r = self.grid_config.obs_radius
            global_obstacles = self.grid.get_obstacles()
            for i in range(n):
                results[i].update(global_obstacles=global_obstacles)
                results[i]['global_xy'] = (int(pos[i][0]) + r, int(pos[i][1]) + r)
                results[i]['global_target_xy'] = (int(tgt[i][0]) + r, int(tgt[i][1]) + r)
        return results

    def _get_infos(self, active=None):
        if active is None:
            active = self._engine.get_state(nat.STATE_ACTIVE)[0]
        return [dict(is_active=bool(v)) for v in active]

    # -- gymnasium style API ------------------------------------------------ #
    def reset(self, seed: Optional[int] = None, return_info: bool = True, options: Optional[dict] = None):
        if seed is not None:
            self.grid_config.seed = seed
        self._ensure_engine()
        task_seed = self.grid_config.seed
        if task_seed is None:  # upstream: default_rng(None) - fresh OS entropy every reset
            task_seed = int.from_bytes(os.urandom(7), "little")
        self._engine.generate([task_seed])
        obs = self._engine.observe_host()
        self._multi_action_sampler.update_seed(self.grid_config.seed)
        self._initial_xy = self._engine.get_state(nat.STATE_POSITIONS)[0].tolist()
        self.grid = GridView(self)

        pos = self._engine.get_state(nat.STATE_POSITIONS)[0]
        tgt = self._engine.get_state(nat.STATE_TARGETS)[0]
        self.was_on_goal = [bool((pos[i] == tgt[i]).all()) for i in range(self.grid_config.num_agents)]
        self._elapsed_steps = 0
        return self._obs_list(obs), self._get_infos()

    def step(self, action):
        assert len(action) == self.grid_config.num_agents
        act = self._h_act
        act[0, :] = action
        if ((act < 0) | (act >= self.action_space.n)).any():
            raise IndexError("action out of range [0, %d)" % self.action_space.n)
        # one C-ABI call (pgm_step_host_ex), one synchronisation: results plus upstream's is_active / was_on_goal
        nat.check(self._engine.lib.pgm_step_host_ex(*self._step_args))
        self._elapsed_steps += 1
        rewards = self._h_rew[0].tolist()
        terminated = self._h_term[0].astype(bool).tolist()
        truncated = self._h_trunc[0].astype(bool).tolist()
        self.was_on_goal = self._h_was[0].astype(bool).tolist()
        infos = self._get_infos(self._h_active[0])
        obs = self._obs_list(self._h_obs)
        if all(truncated) or all(terminated):
            infos[0]['metrics'] = self._episode_metrics()
            if self.grid_config.auto_reset:
                # upstream integrations/sample_factory.py :: AutoResetWrapper: the returned observation is the reset one
                obs, _ = self.reset()
        return obs, rewards, terminated, truncated, infos

    def _episode_metrics(self):
        """upstream wrappers/metrics.py values from the raw device counters."""
        raw = self._engine.get_state(nat.STATE_METRICS)[0]
        n = self.grid_config.num_agents
        ot = self.grid_config.on_target
        if ot == 'restart':
            return {'avg_throughput': int(raw[0]) / self.grid_config.max_episode_steps}
        if ot == 'nothing':
            cost = self._engine.get_state(nat.STATE_SOLVE_COSTS)[0]  # upstream SumOfCostsAndMakespanMetric
            return {'ISR': float(int(raw[3])) / n, 'CSR': float(int(raw[3]) == n), 'ep_length': int(raw[2]),
                    'SoC': int(cost.sum()) + n, 'makespan': int(cost.max()) + 1}
        return {'ISR': int(raw[0]) / n, 'CSR': float(int(raw[0]) == n), 'ep_length': int(raw[1]) / n + 1}

    def sample_actions(self):
        return self._multi_action_sampler.sample_actions(dim=self.grid_config.num_agents)

    def get_num_agents(self):
        return self.grid_config.num_agents

    def get_agents_xy(self, only_active=False, ignore_borders=False):
        return self.grid.get_agents_xy(only_active=only_active, ignore_borders=ignore_borders)

    def get_targets_xy(self, only_active=False, ignore_borders=False):
        return self.grid.get_targets_xy(only_active=only_active, ignore_borders=ignore_borders)

    def get_obstacles(self, ignore_borders=False):
        return self.grid.get_obstacles(ignore_borders=ignore_borders)

    def get_state(self, ignore_borders=False, as_dict=False):
        return self.grid.get_state(ignore_borders=ignore_borders, as_dict=as_dict)

    @property
    def unwrapped(self):
        return self

    def render(self, mode='ansi'):
        from .utils import render_grid
        g = self.grid
        return render_grid(g.obstacles, g.positions_xy, g.finishes_xy, g.is_active)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


class Pogema(PogemaBase):
    """upstream envs.py :: Pogema - on_target='finish': an agent that reaches its goal is rewarded once and disappears."""
    _ON_TARGET = 'finish'


class PogemaLifeLong(PogemaBase):
    """upstream envs.py :: PogemaLifeLong - on_target='restart': a reached goal is replaced by a new one."""
    _ON_TARGET = 'restart'


class PogemaCoopFinish(PogemaBase):
    """upstream envs.py :: PogemaCoopFinish - on_target='nothing': reward when all agents stand on their goals."""
    _ON_TARGET = 'nothing'


_ENV_CLASSES = {'finish': Pogema, 'restart': PogemaLifeLong, 'nothing': PogemaCoopFinish}


def _make_pogema(grid_config):
    env = _ENV_CLASSES[grid_config.on_target](grid_config)  # upstream integrations/make_pogema.py picks the class the same way
    if grid_config.persistent:  # upstream integrations/make_pogema.py :: _make_pogema
        from .wrappers import PersistentWrapper
        env = PersistentWrapper(env)
    return env


def make_pogema(grid_config=None, *args, **kwargs):
    """upstream integrations/make_pogema.py :: make_pogema (== pogema_v0)."""
    if grid_config is None:
        grid_config = GridConfig(**kwargs)
    elif isinstance(grid_config, dict):
        grid_config = GridConfig(**grid_config)
    if grid_config.integration in (None, 'gymnasium'):
        return _make_pogema(grid_config)
    if grid_config.integration == 'SampleFactory':
        # upstream integrations/sample_factory.py wrapper stack (plain Python, needs nothing from Sample Factory)
        from .wrappers import AutoResetWrapper, IsMultiAgentWrapper, MetricsForwardingWrapper
        inner_cfg = grid_config.model_copy(update=dict(auto_reset=False))
        env = IsMultiAgentWrapper(MetricsForwardingWrapper(_make_pogema(inner_cfg)))
        if grid_config.auto_reset is None or grid_config.auto_reset:
            env = AutoResetWrapper(env)
        return env
    if grid_config.integration == 'PettingZoo':
        from .integrations.pettingzoo import parallel_env
        return parallel_env(grid_config)
    raise KeyError(f"integration {grid_config.integration!r} is out of scope of this engine "
                   "(the PyMARL / rllib adapters subclass third-party base classes)")


pogema_v0 = make_pogema


def make_single_agent_gym(grid_config=None, *args, **kwargs):
    """upstream integrations/make_pogema.py :: make_single_agent_gym"""
    from .wrappers import SingleAgentWrapper
    if grid_config is None:
        grid_config = GridConfig(**kwargs)
    elif isinstance(grid_config, dict):
        grid_config = GridConfig(**grid_config)
    return SingleAgentWrapper(_make_pogema(grid_config))
