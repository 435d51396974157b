"""A second, independently recalled form of upstream's `soft` collision rule (envs.py :: Pogema.move_agents /
_revert_action), fuzzed against the oracle's form (oracle/pogema_oracle.py).  Upstream is not importable here (DESIGN.md
section 0), and SURVEY.md section 9 item 1 marks this rule LOW confidence; the two recollections differ in three places,
so their agreement on every random scenario says those places do not matter for the result:

  * where a move into an obstacle is cancelled: before the tables are built (oracle) / together with the vertex
    conflicts, `len(used_cells[target]) > 1 or has_obstacle(target)` (this file);
  * what `_revert_action` cancels: every agent heading into the reverted agent's cell, recursively (oracle) / only the
    FIRST one in that cell's list, recursively, the rest being caught by the outer loop over agents in reversed index
    order (this file);
  * whether a stay action takes part in the swap test (oracle: no; this file: yes - its edge (x,y,x,y) has one user).

    python tools/prototypes/soft_variants.py [--cases 20000]      (also run, smaller, by tests/test_oracle_rules.py)
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle.pogema_oracle as orc  # noqa: E402

MOVES = orc.MOVES


def _revert_action(grid, agent_idx, used_cells, cell, actions):
    actions[agent_idx] = 0
    used_cells[cell].remove(agent_idx)
    new_cell = tuple(grid.positions_xy[agent_idx])
    if new_cell in used_cells and len(used_cells[new_cell]) > 0:
        used_cells[new_cell].append(agent_idx)
        first = used_cells[new_cell][0]
        if first != agent_idx and actions[first] != 0:
            _revert_action(grid, first, used_cells, new_cell, actions)
    else:
        used_cells.setdefault(new_cell, []).append(agent_idx)


def soft_moves_second_form(grid, actions):
    """-> the effective action of every agent (0 = stays) under the second recollection."""
    actions = [int(a) for a in actions]
    n = len(actions)
    used_cells, used_edges = {}, {}
    agents_xy = [tuple(p) for p in grid.positions_xy]
    for i, (x, y) in enumerate(agents_xy):
        if grid.is_active[i]:
            dx, dy = MOVES[actions[i]]
            used_cells.setdefault((x + dx, y + dy), []).append(i)
            used_edges.setdefault((x, y, x + dx, y + dy), []).append(i)
            if dx != 0 or dy != 0:
                used_edges.setdefault((x + dx, y + dy, x, y), []).append(i)
    for i, (x, y) in enumerate(agents_xy):
        if grid.is_active[i]:
            dx, dy = MOVES[actions[i]]
            if len(used_edges[x, y, x + dx, y + dy]) > 1:
                used_cells[x + dx, y + dy].remove(i)
                used_cells.setdefault((x, y), []).append(i)
                actions[i] = 0
    for i in reversed(range(n)):
        if grid.is_active[i]:
            x, y = agents_xy[i]
            dx, dy = MOVES[actions[i]]
            if actions[i] != 0 and (len(used_cells[x + dx, y + dy]) > 1 or grid.has_obstacle(x + dx, y + dy)):
                _revert_action(grid, i, used_cells, (x + dx, y + dy), actions)
    return [actions[i] if grid.is_active[i] else 0 for i in range(n)]


def fuzz(cases, seed=0, verbose=False):
    rng = np.random.default_rng(seed)
    moved = conflicts = 0
    for c in range(cases):
        size = int(rng.integers(3, 9))
        agents = int(rng.integers(2, min(24, size * size // 2)))
        kw = dict(size=size, density=float(rng.choice([0.0, 0.1, 0.3])), num_agents=agents, obs_radius=1,
                  collision_system="soft", on_target=str(rng.choice(["finish", "nothing", "restart"])),
                  max_episode_steps=64, seed=int(rng.integers(0, 1 << 30)))
        try:
            env = orc.pogema_v0(orc.GridConfig(**kw))
            env.reset()
        except OverflowError:
            continue
        for t in range(int(rng.integers(1, 6))):
            acts = [int(a) for a in rng.integers(0, 5, size=agents)]
            grid = env.unwrapped.grid
            before = [tuple(p) for p in grid.positions_xy]
            active = list(grid.is_active)
            want = soft_moves_second_form(grid, acts)
            env.step(acts)
            after = [tuple(p) for p in env.unwrapped.grid.positions_xy]
            for i in range(agents):
                if not active[i]:
                    continue
                dx, dy = MOVES[want[i]]
                exp = (before[i][0] + dx, before[i][1] + dy)
                if after[i] != exp:
                    raise AssertionError(f"case {c} step {t} agent {i}: oracle {before[i]} -> {after[i]}, second form -> {exp}; "
                                         f"config {kw}, actions {acts}")
                moved += int(want[i] != 0)
                conflicts += int(want[i] == 0 and acts[i] != 0)
    if verbose:
        print(f"{cases} scenarios: the two forms agree; {moved} moves made, {conflicts} moves cancelled")
    return moved, conflicts


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    fuzz(a.cases, a.seed, verbose=True)
