"""PettingZoo style parallel env (upstream integrations/pettingzoo.py ::
PogemaParallel / parallel_env).  pettingzoo itself is optional: the class is
duck-typed to the ParallelEnv protocol (possible_agents / agents / reset / step
with dict keyed returns, agents named ``player_i``)."""
from __future__ import annotations

import functools

from ..envs import _make_pogema
from ..grid_config import GridConfig

try:
    from pettingzoo import ParallelEnv as _ParallelBase
except Exception:  # pragma: no cover - pettingzoo is not installed in this image
    _ParallelBase = object


class PogemaParallel(_ParallelBase):
    def __init__(self, grid_config: GridConfig, render_mode='ansi'):
        self.metadata = {'render_modes': ['ansi'], "name": "pogema"}
        self.render_mode = render_mode
        self.pogema = _make_pogema(grid_config)
        self.possible_agents = ["player_" + str(r) for r in range(self.pogema.get_num_agents())]
        self.agent_name_mapping = dict(zip(self.possible_agents, list(range(len(self.possible_agents)))))
        self.agents = None
        self.num_moves = None

    def state(self):
        return self.pogema.get_state()

    @functools.lru_cache(maxsize=None)
    def observation_space(self, agent):
        assert agent in self.possible_agents
        return self.pogema.observation_space

    @functools.lru_cache(maxsize=None)
    def action_space(self, agent):
        assert agent in self.possible_agents
        return self.pogema.action_space

    def render(self, mode="human"):
        return self.pogema.render()

    def reset(self, seed=None, options=None):
        observations, info = self.pogema.reset(seed=seed, options=options)
        self.agents = self.possible_agents[:]
        self.num_moves = 0
        anm = self.agent_name_mapping
        observations = {agent: observations[anm[agent]].astype('float32') for agent in self.agents}
        infos = {agent: info[anm[agent]] for agent in self.agents}
        return observations, infos

    def step(self, actions):
        anm = self.agent_name_mapping
        actions = [actions[agent] if agent in actions else 0 for agent in self.possible_agents]
        observations, rewards, terminated, truncated, infos = self.pogema.step(actions)
        d_observations = {agent: observations[anm[agent]].astype('float32') for agent in self.agents}
        d_rewards = {agent: rewards[anm[agent]] for agent in self.agents}
        d_terminated = {agent: terminated[anm[agent]] for agent in self.agents}
        d_truncated = {agent: truncated[anm[agent]] for agent in self.agents}
        d_infos = {agent: infos[anm[agent]] for agent in self.agents}
        for agent, idx in anm.items():
            if (terminated[idx] or truncated[idx]) and agent in self.agents:
                self.agents.remove(agent)
        self.num_moves += 1
        return d_observations, d_rewards, d_terminated, d_truncated, d_infos

    def close(self):
        self.pogema.close()


def parallel_env(grid_config: GridConfig = None, **kwargs):
    if grid_config is None:
        grid_config = GridConfig(**kwargs)
    return PogemaParallel(grid_config)
