"""Seeded synthetic map generators for the benchmark configurations of BASELINE.json that name a map
style upstream has no generator for: maze-like 64x64 maps (config 3) and
warehouse-style 256x256 maps (config 4).  They return uint8 (H, W) arrays (1 = obstacle) to pass as
``GridConfig(map=m.tolist())``."""
import numpy as np


def maze_map(size: int, seed: int = 0) -> np.ndarray:
    """Recursive-backtracker maze on a (size/2) x (size/2) cell lattice: corridors one cell wide."""
    rng = np.random.default_rng(seed)
    n = size // 2
    m = np.ones((size, size), np.uint8)
    seen = np.zeros((n, n), bool)
    stack = [(0, 0)]
    seen[0, 0] = True
    m[0, 0] = 0
    while stack:
        x, y = stack[-1]
        nb = [(x + dx, y + dy) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))
              if 0 <= x + dx < n and 0 <= y + dy < n and not seen[x + dx, y + dy]]
        if not nb:
            stack.pop()
            continue
        nx, ny = nb[rng.integers(len(nb))]
        seen[nx, ny] = True
        m[2 * nx, 2 * ny] = 0
        m[x + nx, y + ny] = 0
        stack.append((nx, ny))
    return m


def warehouse_map(size: int) -> np.ndarray:
    """Regular 2 x 8 shelf blocks separated by one-cell aisles."""
    m = np.zeros((size, size), np.uint8)
    for x in range(2, size - 2, 3):
        for y in range(2, size - 9, 10):
            m[x:x + 2, y:y + 8] = 1
    return m
