"""The native (C++) task generator of libpgm_b200.so against the Python oracle,
which itself calls the real numpy Generator: obstacles, starts, goals, the
lifelong per-agent PCG64 states and component sizes must match exactly.
Runs without a GPU (pgm_generate_host makes no CUDA call)."""
import ctypes as C

import numpy as np
import pytest

from oracle import pogema_oracle as orc
from pogema_b200 import _native as nat


def native_generate(size, A, r, density, seed, lifelong=False, map_=None):
    lib = nat.load()
    h, w = (size, size) if map_ is None else map_.shape
    obst = np.zeros((h, w), np.uint8)
    axy = np.zeros((A, 2), np.int32)
    txy = np.zeros((A, 2), np.int32)
    rng = np.zeros((A, 4), np.uint64)
    cs = np.zeros(A, np.int32)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.pgm_generate_host(h, w, A, r, float(density), int(lifelong), p(map_), C.c_uint64(seed), p(obst), p(axy),
                               p(txy), p(rng), p(cs))
    return rc, obst, axy, txy, rng, cs


def oracle_generate(size, A, r, density, seed, lifelong=False, map_=None):
    gc = orc.GridConfig(size=size, density=density, num_agents=A, obs_radius=r, seed=seed,
                        on_target='restart' if lifelong else 'finish', map=None if map_ is None else map_.tolist())
    env = orc.PogemaLifeLong(gc) if lifelong else orc.Pogema(gc)
    env.reset()
    return env


@pytest.mark.parametrize("size,A,density", [(8, 4, 0.3), (16, 20, 0.3), (32, 64, 0.3), (12, 10, 0.0), (10, 3, 0.6),
                                            (24, 30, 0.45), (5, 2, 0.5)])
def test_generator_matches_oracle(size, A, density):
    checked = 0
    for seed in range(40):
        try:
            env = oracle_generate(size, A, 3, density, seed)
        except OverflowError:
            rc, *_ = native_generate(size, A, 3, density, seed)
            assert rc == nat.PGM_ERR_OVERFLOW
            continue
        rc, obst, axy, txy, _, _ = native_generate(size, A, 3, density, seed)
        assert rc == 0, nat.load().pgm_last_error()
        g = env.grid
        assert np.array_equal(obst, g.get_obstacles(ignore_borders=True).astype(np.uint8)), seed
        assert np.array_equal(axy, np.array(g.get_agents_xy(ignore_borders=True))), seed
        assert np.array_equal(txy, np.array(g.get_targets_xy(ignore_borders=True))), seed
        checked += 1
    assert checked > 0 or density >= 0.5


def test_overflow_and_retry_path():
    # dense maps exercise the retry loop (Grid.rnd re-draws) and OverflowError
    seen_overflow = seen_ok = 0
    for seed in range(60):
        try:
            env = oracle_generate(5, 6, 2, 0.6, seed)
            rc, obst, axy, txy, _, _ = native_generate(5, 6, 2, 0.6, seed)
            assert rc == 0
            assert np.array_equal(obst, env.grid.get_obstacles(ignore_borders=True).astype(np.uint8)), seed
            assert np.array_equal(axy, np.array(env.grid.get_agents_xy(ignore_borders=True))), seed
            assert np.array_equal(txy, np.array(env.grid.get_targets_xy(ignore_borders=True))), seed
            seen_ok += 1
        except OverflowError:
            rc, *_ = native_generate(5, 6, 2, 0.6, seed)
            assert rc == nat.PGM_ERR_OVERFLOW
            seen_overflow += 1
    assert seen_ok and seen_overflow


def test_lifelong_generators_and_components():
    for seed in range(12):
        env = oracle_generate(16, 12, 4, 0.35, seed, lifelong=True)
        rc, obst, axy, txy, rng, cs = native_generate(16, 12, 4, 0.35, seed, lifelong=True)
        assert rc == 0
        g = env.grid
        assert np.array_equal(axy, np.array(g.get_agents_xy(ignore_borders=True)))
        assert np.array_equal(txy, np.array(g.get_targets_xy(ignore_borders=True)))
        for a in range(12):
            st = env.random_generators[a].bit_generator.state['state']
            assert (int(rng[a, 0]) << 64 | int(rng[a, 1])) == st['state']
            assert (int(rng[a, 2]) << 64 | int(rng[a, 3])) == st['inc']
            comp = g.component_to_points[g.point_to_component[g.positions_xy[a]]]
            assert cs[a] == len(comp)


def test_fixed_map_placement():
    rng = np.random.default_rng(5)
    m = (rng.random((9, 14)) < 0.2).astype(np.uint8)
    for seed in range(10):
        env = oracle_generate(14, 7, 2, 0.0, seed, map_=m)
        rc, obst, axy, txy, _, _ = native_generate(14, 7, 2, 0.0, seed, map_=m)
        assert rc == 0
        assert np.array_equal(obst, m)
        assert np.array_equal(axy, np.array(env.grid.get_agents_xy(ignore_borders=True)))
        assert np.array_equal(txy, np.array(env.grid.get_targets_xy(ignore_borders=True)))
