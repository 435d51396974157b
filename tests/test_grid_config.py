"""GridConfig keeps upstream's field names, defaults, bounds and string-map syntax."""
import pytest

from pogema_b200 import GridConfig, Hard8x8, ExtraHard64x64


def test_defaults_and_bounds():
    gc = GridConfig()
    assert (gc.size, gc.density, gc.num_agents, gc.obs_radius, gc.max_episode_steps) == (8, 0.3, 1, 5, 64)
    assert gc.collision_system == 'priority' and gc.on_target == 'finish' and gc.seed is None
    assert gc.MOVES == [[0, 0], [-1, 0], [1, 0], [0, -1], [0, 1]] and gc.FREE == 0 and gc.OBSTACLE == 1
    for bad in (dict(size=1), dict(size=1025), dict(density=1.5), dict(num_agents=0), dict(obs_radius=0),
                dict(obs_radius=129), dict(seed=-1)):
        with pytest.raises(Exception):
            GridConfig(**bad)
    with pytest.raises(Exception):
        GridConfig(collision_system='nope')
    assert Hard8x8().num_agents == 4 and ExtraHard64x64().max_episode_steps == 512


def test_string_map():
    gc = GridConfig(map="""
        .a.#
        .#.A
        b..B
    """)
    assert gc.num_agents == 2 and gc.agents_xy == [[0, 1], [2, 0]] and gc.targets_xy == [[1, 3], [2, 3]]
    assert gc.size == 4 and gc.map_shape() == (3, 4)
    assert abs(gc.density - 2 / 12) < 1e-12
    with pytest.raises(KeyError):
        GridConfig(map="a?A")
    with pytest.raises(IndexError):
        GridConfig(size=4, agents_xy=[[5, 0]], targets_xy=[[0, 0]])


def test_seed_is_mutable():
    gc = GridConfig(seed=1)
    gc.seed = 7
    assert gc.seed == 7


def test_gymnasium_registration_is_guarded():
    """`Pogema-v0` is registered when gymnasium is importable and silently skipped when it is not."""
    import importlib.util
    import pogema_b200
    have = importlib.util.find_spec("gymnasium") is not None
    assert pogema_b200.GYMNASIUM_REGISTERED == have
    if have:
        import gymnasium
        assert "Pogema-v0" in gymnasium.envs.registration.registry


def test_pin_upstream_kit_runs():
    """tools/pin_upstream.py: the probe reports, and the differential machinery agrees with itself."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "pin_upstream.py"), "--out", "/tmp/_pin_probe.txt"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "RESULT:" in out.stdout
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "pin_upstream.py"), "--self-test", "--seeds", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "episodes identical" in out.stdout, out.stdout + out.stderr
