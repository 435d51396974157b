"""``GridConfig`` - drop-in for upstream ``pogema/grid_config.py :: GridConfig``
(same field names, defaults, bounds and string-map syntax).
Upstream is a pydantic-1 model driven by field validators; this is a pydantic-2
model with one after-validator that applies the same rules."""
from __future__ import annotations

import sys
from typing import Literal, Optional, Union

from pydantic import BaseModel, ConfigDict, model_validator


class CommonSettings(BaseModel):
    model_config = ConfigDict(validate_assignment=False, extra="forbid")
    # upstream grid_config.py :: CommonSettings  (x = row, y = column)
    MOVES: list = [[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1]]
    FREE: Literal[0] = 0
    OBSTACLE: Literal[1] = 1
    empty_outside: bool = True


class GridConfig(CommonSettings):
    on_target: Literal['finish', 'nothing', 'restart'] = 'finish'
    seed: Optional[int] = None
    size: int = 8
    density: float = 0.3
    num_agents: int = 1
    obs_radius: int = 5
    agents_xy: Optional[list] = None
    targets_xy: Optional[list] = None
    collision_system: Literal['block_both', 'priority', 'soft'] = 'priority'
    persistent: bool = False
    observation_type: Literal['POMAPF', 'MAPF', 'default'] = 'default'
    map: Union[list, str, None] = None
    map_name: Optional[str] = None
    integration: Optional[Literal['SampleFactory', 'PyMARL', 'rllib', 'gymnasium', 'PettingZoo']] = None
    max_episode_steps: int = 64
    auto_reset: Optional[bool] = None

    @model_validator(mode='after')
    def _validate(self):
        assert self.seed is None or (0 <= self.seed < sys.maxsize), "seed must be in [0, " + str(sys.maxsize) + ']'
        assert 2 <= self.size <= 1024, "size must be in [2, 1024]"
        assert 0.0 <= self.density <= 1, "density must be in [0, 1]"
        assert 1 <= self.num_agents <= 10000000, "num_agents must be in [1, 10000000]"
        assert 1 <= self.obs_radius <= 128, "obs_radius must be in [1, 128]"
        assert self.max_episode_steps >= 1, "max_episode_steps must be positive"
        if not self.empty_outside:
            raise ValueError("empty_outside=False is not supported by the batched engine")
        if self.map is not None:
            v = self.map
            if isinstance(v, str):
                v, agents_xy, targets_xy = self.str_map_to_list(v, self.FREE, self.OBSTACLE)
                if agents_xy and targets_xy and self.agents_xy is not None and self.targets_xy is not None:
                    raise KeyError("Can't create task. Please provide agents_xy and targets_xy only once. "
                                   "Either with parameters or with a map.")
                if agents_xy and targets_xy:
                    self.agents_xy = agents_xy
                    self.targets_xy = targets_xy
                    self.num_agents = len(agents_xy)
            size = len(v)
            area = 0
            for line in v:
                size = max(size, len(line))
                area += len(line)
            self.size = size
            self.density = sum(sum(line) for line in v) / area
            self.map = v
        for v in (self.agents_xy, self.targets_xy):
            if v is not None:
                self.check_positions(v, self.size)
                self.num_agents = len(v)
        return self

    @staticmethod
    def check_positions(v, size):
        for position in v:
            x, y = position
            if not (0 <= x < size and 0 <= y < size):
                raise IndexError("Position is out of bounds!")

    @staticmethod
    def str_map_to_list(str_map, free, obstacle):
        obstacles = []
        agents = {}
        targets = {}
        for idx, line in enumerate(str_map.split()):
            row = []
            for char in line:
                if char == '.':
                    row.append(free)
                elif char == '#':
                    row.append(obstacle)
                elif 'A' <= char <= 'Z':
                    targets[char.lower()] = len(obstacles), len(row)
                    row.append(free)
                elif 'a' <= char <= 'z':
                    agents[char.lower()] = len(obstacles), len(row)
                    row.append(free)
                else:
                    raise KeyError(f"Unsupported symbol '{char}' at line {idx}")
            if row:
                if obstacles:
                    assert len(obstacles[-1]) == len(row), f"Wrong string size for row {idx};"
                obstacles.append(row)
        targets_xy = []
        agents_xy = []
        for _, (x, y) in sorted(agents.items()):
            agents_xy.append([x, y])
        for _, (x, y) in sorted(targets.items()):
            targets_xy.append([x, y])
        assert len(targets_xy) == len(agents_xy)
        return obstacles, agents_xy, targets_xy

    # geometry helpers used by the engine bindings
    def map_array(self):
        """The explicit map as a uint8 (H, W) array (ragged rows padded with obstacles), or None."""
        import numpy as np
        if self.map is None:
            return None
        width = max(len(line) for line in self.map)
        return np.array([list(line) + [self.OBSTACLE] * (width - len(line)) for line in self.map], dtype=np.uint8)

    def map_shape(self):
        m = self.map_array()
        return (self.size, self.size) if m is None else m.shape


class PredefinedDifficultyConfig(GridConfig):
    density: float = 0.3
    collision_system: Literal['block_both', 'priority', 'soft'] = 'priority'
    obs_radius: int = 5


class Easy8x8(PredefinedDifficultyConfig):
    size: int = 8
    max_episode_steps: int = 64
    num_agents: int = 1


class Normal8x8(PredefinedDifficultyConfig):
    size: int = 8
    max_episode_steps: int = 64
    num_agents: int = 2


class Hard8x8(PredefinedDifficultyConfig):
    size: int = 8
    max_episode_steps: int = 64
    num_agents: int = 4


class ExtraHard8x8(PredefinedDifficultyConfig):
    size: int = 8
    max_episode_steps: int = 64
    num_agents: int = 8


class Easy16x16(PredefinedDifficultyConfig):
    size: int = 16
    max_episode_steps: int = 128
    num_agents: int = 4


class Normal16x16(PredefinedDifficultyConfig):
    size: int = 16
    max_episode_steps: int = 128
    num_agents: int = 8


class Hard16x16(PredefinedDifficultyConfig):
    size: int = 16
    max_episode_steps: int = 128
    num_agents: int = 16


class ExtraHard16x16(PredefinedDifficultyConfig):
    size: int = 16
    max_episode_steps: int = 128
    num_agents: int = 32


class Easy32x32(PredefinedDifficultyConfig):
    size: int = 32
    max_episode_steps: int = 256
    num_agents: int = 16


class Normal32x32(PredefinedDifficultyConfig):
    size: int = 32
    max_episode_steps: int = 256
    num_agents: int = 32


class Hard32x32(PredefinedDifficultyConfig):
    size: int = 32
    max_episode_steps: int = 256
    num_agents: int = 64


class ExtraHard32x32(PredefinedDifficultyConfig):
    size: int = 32
    max_episode_steps: int = 256
    num_agents: int = 128


class Easy64x64(PredefinedDifficultyConfig):
    size: int = 64
    max_episode_steps: int = 512
    num_agents: int = 64


class Normal64x64(PredefinedDifficultyConfig):
    size: int = 64
    max_episode_steps: int = 512
    num_agents: int = 128


class Hard64x64(PredefinedDifficultyConfig):
    size: int = 64
    max_episode_steps: int = 512
    num_agents: int = 256


class ExtraHard64x64(PredefinedDifficultyConfig):
    size: int = 64
    max_episode_steps: int = 512
    num_agents: int = 512
