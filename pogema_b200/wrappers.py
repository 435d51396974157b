"""Host-side wrappers around the list based env: state history with
``step_back`` (upstream wrappers/persistence.py :: PersistentWrapper / AgentState) and SVG episode
animation (upstream svg_animation/ :: AnimationMonitor / AnimationConfig).

Neither evaluates a rule of the environment: every recorded state is read back from the engine
(``pgm_get_state``), and ``step_back`` restores the device state from the checkpoint taken before the
step (``pgm_checkpoint_save`` / ``pgm_checkpoint_load``)."""
from __future__ import annotations

import os
import time
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _native as nat


class _Wrapper:
    """Minimal gymnasium.Wrapper look-alike (gymnasium is optional in this image)."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


@dataclass(frozen=True)
class AgentState:
    """upstream wrappers/persistence.py :: AgentState - one agent at one step (unpadded coordinates)."""
    x: int
    y: int
    tx: int
    ty: int
    step: int
    active: bool

    def get_xy(self):
        return self.x, self.y

    def get_target_xy(self):
        return self.tx, self.ty

    def is_active(self):
        return self.active

    def get_step(self):
        return self.step


class PersistentWrapper(_Wrapper):
    """Keeps the full state history of the running episode and can undo steps.

    ``get_history()`` -> per agent, the list of ``AgentState`` from reset to now;
    ``step_back()``   -> undo the last step on the DEVICE (positions, targets, active flags, elapsed
    steps, lifelong generators, metric counters) and drop it from the history; False at the start."""

    def __init__(self, env, xy_offset=None):
        super().__init__(env)
        self._step = 0
        self._history: List[List[AgentState]] = []
        self._checkpoints: List[np.ndarray] = []

    # -- recording -------------------------------------------------------- #
    def _snapshot(self):
        eng = self.unwrapped._engine
        pos = eng.get_state(nat.STATE_POSITIONS)[0]
        tgt = eng.get_state(nat.STATE_TARGETS)[0]
        act = eng.get_state(nat.STATE_ACTIVE)[0]
        return [AgentState(int(pos[i][0]), int(pos[i][1]), int(tgt[i][0]), int(tgt[i][1]), self._step, bool(act[i]))
                for i in range(len(pos))]

    def reset(self, **kwargs):
        result = self.env.reset(**kwargs)
        self._step = 0
        self._checkpoints = []
        self._history = [[s] for s in self._snapshot()]
        return result

    def step(self, action):
        eng = self.unwrapped._engine
        before = eng.checkpoint()
        elapsed = self.unwrapped._elapsed_steps
        result = self.env.step(action)
        if self.unwrapped._elapsed_steps != elapsed + 1:
            # the env auto-reset inside the step (GridConfig.auto_reset): a new episode starts here
            self._step = 0
            self._checkpoints = []
            self._history = [[s] for s in self._snapshot()]
            return result
        self._step += 1
        self._checkpoints.append(before)
        for i, s in enumerate(self._snapshot()):
            self._history[i].append(s)
        return result

    def step_back(self) -> bool:
        if not self._checkpoints:
            return False
        inner = self.unwrapped
        inner._engine.restore(self._checkpoints.pop())
        inner._elapsed_steps -= 1
        inner.was_on_goal = [bool(v) for v in inner._engine.get_state(nat.STATE_WAS_ON_GOAL)[0]]
        self._step -= 1
        for h in self._history:
            h.pop()
        return True

    # -- access ------------------------------------------------------------ #
    def get_history(self) -> List[List[AgentState]]:
        return [list(h) for h in self._history]

    @staticmethod
    def agent_state_to_full_list(agent_states: List[AgentState]) -> List[AgentState]:
        """History of one agent with every step present (it already is: one state per step)."""
        return list(agent_states)

    @classmethod
    def decompress_history(cls, history):
        return [cls.agent_state_to_full_list(h) for h in history]


# --------------------------------------------------------------------------- #
# SVG animation
# --------------------------------------------------------------------------- #
@dataclass
class AnimationConfig:
    """upstream svg_animation :: AnimationConfig"""
    directory: str = 'renders/'
    static: bool = False
    show_agents: bool = True
    egocentric_idx: Optional[int] = None
    uid: Optional[str] = None
    save_every_idx_episode: Optional[int] = 1
    show_border: bool = True
    show_lines: bool = False


_COLORS = ['#c1433c', '#2e6f9e', '#6e81af', '#00b9c8', '#72d5c8', '#0ea08c', '#8f7b66', '#e6a23c', '#7a4ea3',
           '#4f9d3a', '#d26aa5', '#5c6b73']
_CELL = 100  # svg units per grid cell
_R = 35      # agent radius
_STEP_S = 0.25


class _Svg:
    def __init__(self, width, height):
        self.parts = [f'<?xml version="1.0" encoding="UTF-8"?>\n<svg xmlns="http://www.w3.org/2000/svg" '
                      f'viewBox="0 0 {width} {height}" width="{width // 4}" height="{height // 4}">\n'
                      '<style>.o{fill:#84a1ae}.l{stroke:#84a1ae;stroke-width:6;stroke-dasharray:18}'
                      '.t{fill:none;stroke-width:10}.a{stroke:none}</style>\n']

    def add(self, s):
        self.parts.append(s)

    def render(self):
        return ''.join(self.parts) + '</svg>\n'


def _animate(attr, values, dur, extra=''):
    vals = ';'.join(str(v) for v in values)
    return f'<animate attributeName="{attr}" dur="{dur:.2f}s" values="{vals}" repeatCount="indefinite"{extra}/>'


def render_svg(obstacles: np.ndarray, history: List[List[AgentState]], config: AnimationConfig, obs_radius: int) -> str:
    """One episode as an animated (or static: last frame of the start state) SVG.
    obstacles: uint8 [H][W] unpadded; history: per agent one AgentState per step."""
    h, w = obstacles.shape
    border = 1 if config.show_border else 0
    W, H = (w + 2 * border) * _CELL, (h + 2 * border) * _CELL
    svg = _Svg(W, H)
    ego = config.egocentric_idx
    steps = len(history[0]) if history else 1
    dur = max(steps, 1) * _STEP_S

    def cx(y):  # column -> svg x (cell centre)
        return (y + border) * _CELL + _CELL // 2

    def cy(x):  # row -> svg y
        return (x + border) * _CELL + _CELL // 2

    def visible_from_ego(t, x, y):
        if ego is None:
            return True
        e = history[ego][min(t, steps - 1)]
        return abs(e.x - x) <= obs_radius and abs(e.y - y) <= obs_radius

    # obstacles (and the wall ring of add_artificial_border when show_border)
    cells = [(x, y) for x in range(h) for y in range(w) if obstacles[x, y]]
    if border:
        cells += [(-1, y) for y in range(-1, w + 1)] + [(h, y) for y in range(-1, w + 1)]
        cells += [(x, -1) for x in range(h)] + [(x, w) for x in range(h)]
    for x, y in cells:
        rect = (f'<rect class="o" x="{(y + border) * _CELL + 5}" y="{(x + border) * _CELL + 5}" width="{_CELL - 10}" '
                f'height="{_CELL - 10}" rx="15"')
        if ego is not None and not config.static and 0 <= x < h and 0 <= y < w:
            op = [1.0 if visible_from_ego(t, x, y) else 0.3 for t in range(steps)]
            svg.add(rect + '>' + _animate('opacity', op, dur) + '</rect>\n')
        else:
            svg.add(rect + '/>\n')

    if not config.show_agents:
        return svg.render()

    for i, states in enumerate(history):
        color = _COLORS[i % len(_COLORS)]
        if ego is not None:
            color = '#c1433c' if i == ego else '#2e6f9e'
        s0 = states[0]
        # target: a ring; lifelong episodes move it
        tx = [cx(s.ty) for s in states]
        ty = [cy(s.tx) for s in states]
        ring = f'<circle class="t" stroke="{color}" cx="{tx[0]}" cy="{ty[0]}" r="{_R}"'
        if config.static or (len(set(tx)) == 1 and len(set(ty)) == 1):
            svg.add(ring + '/>\n')
        else:
            svg.add(ring + '>' + _animate('cx', tx, dur, ' calcMode="discrete"') +
                    _animate('cy', ty, dur, ' calcMode="discrete"') + '</circle>\n')
        if config.show_lines:
            line = f'<line class="l" x1="{cx(s0.y)}" y1="{cy(s0.x)}" x2="{tx[0]}" y2="{ty[0]}"'
            if config.static:
                svg.add(line + '/>\n')
            else:
                svg.add(line + '>' + _animate('x1', [cx(s.y) for s in states], dur) +
                        _animate('y1', [cy(s.x) for s in states], dur) +
                        _animate('x2', tx, dur, ' calcMode="discrete"') + _animate('y2', ty, dur, ' calcMode="discrete"') +
                        '</line>\n')
        agent = f'<circle class="a" fill="{color}" cx="{cx(s0.y)}" cy="{cy(s0.x)}" r="{_R}"'
        if config.static:
            svg.add(agent + '/>\n')
            continue
        anim = _animate('cx', [cx(s.y) for s in states], dur) + _animate('cy', [cy(s.x) for s in states], dur)
        # a finished agent disappears (upstream hide_agent); from the ego agent's view others fade outside its window
        op = []
        for t, s in enumerate(states):
            if not s.active:
                op.append(0.0)
            elif ego is not None and i != ego and not visible_from_ego(t, s.x, s.y):
                op.append(0.2)
            else:
                op.append(1.0)
        if len(set(op)) > 1:
            anim += _animate('opacity', op, dur, ' calcMode="discrete"')
        svg.add(agent + '>' + anim + '</circle>\n')
    return svg.render()


class AnimationMonitor(_Wrapper):
    """Records every episode (through a PersistentWrapper) and writes it as an SVG when the episode ends
    (upstream svg_animation :: AnimationMonitor): ``renders/pogema-ep00000.svg``, every
    ``save_every_idx_episode``-th episode; ``save_animation(name)`` writes the running episode on demand."""

    def __init__(self, env, animation_config: AnimationConfig = AnimationConfig()):
        if not isinstance(env, PersistentWrapper):
            env = PersistentWrapper(env)
        super().__init__(env)
        self.history = self.env.get_history
        self.animation_config = animation_config
        self._episode_idx = 0

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        inner = self.unwrapped
        auto = bool(inner.grid_config.auto_reset)
        snapshot = None
        if auto:
            # with GridConfig.auto_reset the env starts the next episode inside step(): keep this one's history
            snapshot = self.env.get_history()
        obs, rewards, terminated, truncated, infos = self.env.step(action)
        if all(terminated) or all(truncated):
            every = self.animation_config.save_every_idx_episode
            if every and (self._episode_idx + 1) % every == 0:
                os.makedirs(self.animation_config.directory, exist_ok=True)
                name = self.pick_name(inner.grid_config, self._episode_idx)
                # (after an auto-reset the final move is gone from the device: the animation ends one frame early)
                self.save_animation(os.path.join(self.animation_config.directory, name), history=snapshot)
            self._episode_idx += 1
        return obs, rewards, terminated, truncated, infos

    @staticmethod
    def pick_name(grid_config, episode_idx=None, zfill_ep=5):
        name = 'pogema'
        if episode_idx is not None:
            name += f'-ep{str(episode_idx).zfill(zfill_ep)}'
        if grid_config and grid_config.seed is not None:
            name += f'-seed{grid_config.seed}'
        return name + '.svg'

    def save_animation(self, name='render.svg', animation_config: Optional[AnimationConfig] = None, history=None):
        cfg = animation_config or self.animation_config
        inner = self.unwrapped
        obstacles = inner._engine.get_state(nat.STATE_OBSTACLES)[0]
        history = history if history is not None else self.env.get_history()
        svg = render_svg(np.asarray(obstacles), history, cfg, inner.grid_config.obs_radius)
        directory = os.path.dirname(name)
        if directory:
            os.makedirs(directory, exist_ok=True)
        with open(name, 'w') as f:
            f.write(svg)
        return name


class RuntimeMetricWrapper(_Wrapper):
    """upstream wrappers/metrics.py :: RuntimeMetricWrapper (opt-in upstream too): ``metrics['runtime']`` = wall-clock
    seconds of the episode spent outside ``env.step``, i.e. the policy's time."""

    def __init__(self, env):
        super().__init__(env)
        self._start_time = None
        self._env_step_time = 0.0

    def reset(self, **kwargs):
        out = self.env.reset(**kwargs)
        self._start_time = time.monotonic()
        self._env_step_time = 0.0
        return out

    def step(self, action):
        t0 = time.monotonic()
        obs, rewards, terminated, truncated, infos = self.env.step(action)
        t1 = time.monotonic()
        self._env_step_time += t1 - t0
        if all(terminated) or all(truncated):
            if self._start_time is None:  # stepped without a reset through this wrapper
                self._start_time = t0
            infos[0].setdefault('metrics', {})['runtime'] = time.monotonic() - self._start_time - self._env_step_time
            if getattr(self.env.unwrapped.grid_config, 'auto_reset', None):  # the inner env already started a new episode
                self._start_time = time.monotonic()
                self._env_step_time = 0.0
        return obs, rewards, terminated, truncated, infos


class AgentsDensityWrapper(_Wrapper):
    """upstream wrappers/metrics.py :: AgentsDensityWrapper (opt-in upstream too; for ``observation_type`` 'POMAPF' /
    'MAPF'): ``metrics['avg_agents_density']`` = the episode's mean, reset observation included, of the per-step mean
    over agents of (agents visible in the window / traversable cells of the window)."""

    def __init__(self, env):
        super().__init__(env)
        self._densities = []

    def _count(self, observations):
        if not isinstance(observations[0], dict):
            raise TypeError("AgentsDensityWrapper needs dict observations (GridConfig(observation_type='POMAPF' or 'MAPF'))")
        per_agent = [np.count_nonzero(o['agents']) / (np.size(o['obstacles']) - np.count_nonzero(o['obstacles']))
                     for o in observations]
        self._densities.append(np.mean(per_agent))

    def reset(self, **kwargs):
        self._densities = []
        observations, infos = self.env.reset(**kwargs)
        self._count(observations)
        return observations, infos

    def step(self, action):
        observations, rewards, terminated, truncated, infos = self.env.step(action)
        if all(terminated) or all(truncated):
            # (with GridConfig(auto_reset=True) `observations` is already the next episode's reset observation: it
            # opens the next episode's mean instead of closing this one's)
            auto = bool(getattr(self.env.unwrapped.grid_config, 'auto_reset', None))
            if not auto:
                self._count(observations)
            infos[0].setdefault('metrics', {})['avg_agents_density'] = float(np.mean(self._densities))
            if auto:
                self._densities = []
                self._count(observations)
        else:
            self._count(observations)
        return observations, rewards, terminated, truncated, infos


class AutoResetWrapper(_Wrapper):
    """upstream integrations/sample_factory.py :: AutoResetWrapper: when every agent is terminated or truncated the
    env is reset inside ``step`` and the returned observation is the reset one (``GridConfig(auto_reset=True)``
    does the same inside ``Pogema.step``)."""

    def step(self, action):
        obs, rewards, terminated, truncated, infos = self.env.step(action)
        if all(terminated) or all(truncated):
            obs, _ = self.env.reset()
        return obs, rewards, terminated, truncated, infos


class SingleAgentWrapper(_Wrapper):
    """upstream integrations/make_pogema.py :: SingleAgentWrapper: a single-agent gymnasium view of agent 0; the
    other agents (if any) act at random from the env's action space."""

    def step(self, action):
        inner = self.unwrapped
        others = [inner.action_space.sample() for _ in range(inner.get_num_agents() - 1)]
        observations, rewards, terminated, truncated, infos = self.env.step([action] + others)
        return observations[0], rewards[0], terminated[0], truncated[0], infos[0]

    def reset(self, seed=None, return_info=True, options=None):
        observations, infos = self.env.reset(seed=seed, options=options)
        if return_info:
            return observations[0], infos[0]
        return observations[0]


class IsMultiAgentWrapper(_Wrapper):
    """upstream integrations/sample_factory.py :: IsMultiAgentWrapper"""

    def __init__(self, env):
        super().__init__(env)
        self.is_multiagent = True

    @property
    def num_agents(self):
        return self.unwrapped.get_num_agents()


class MetricsForwardingWrapper(_Wrapper):
    """upstream integrations/sample_factory.py :: MetricsForwardingWrapper: episode metrics are copied to
    ``info['episode_extra_stats']`` where Sample Factory collects them."""

    def step(self, action):
        from copy import deepcopy
        observations, rewards, terminated, truncated, infos = self.env.step(action)
        for info in infos:
            if 'metrics' in info:
                info.update(episode_extra_stats=deepcopy(info['metrics']))
        return observations, rewards, terminated, truncated, infos
