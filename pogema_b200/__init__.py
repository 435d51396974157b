"""pogema_b200 - B200-native batched POGEMA step engine.

Drop-in surface of upstream pogema for the step path: ``GridConfig``,
``pogema_v0`` (list based gymnasium-style reset/step on one instance),
``parallel_env`` (PettingZoo-style), plus ``BatchedPogema`` (torch CUDA tensors).
All stepping and observation generation runs in the hand-written sm_100a
kernels of ``libpgm_b200.so``; there is no CPU fallback.
"""
from .grid_config import (GridConfig, Easy8x8, Normal8x8, Hard8x8, ExtraHard8x8, Easy16x16, Normal16x16,
                          Hard16x16, ExtraHard16x16, Easy32x32, Normal32x32, Hard32x32, ExtraHard32x32,
                          Easy64x64, Normal64x64, Hard64x64, ExtraHard64x64)

__version__ = "0.2.0"


def _register_gymnasium():
    """upstream pogema/__init__.py registers the id ``Pogema-v0`` (single-agent gymnasium env).  gymnasium is
    optional here (absent in the build image): register when it is importable, otherwise do nothing."""
    try:
        from gymnasium.envs.registration import register, registry
    except Exception:
        return False
    if "Pogema-v0" not in registry:
        register(id="Pogema-v0", entry_point="pogema_b200.envs:make_single_agent_gym")
    return True


GYMNASIUM_REGISTERED = _register_gymnasium()


def __getattr__(name):
    # lazy: importing the package must not require torch / a built library
    if name == "BatchedPogema":
        from .batched import BatchedPogema
        return BatchedPogema
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name in ("pogema_v0", "make_pogema", "make_single_agent_gym", "PogemaBase", "Pogema", "PogemaLifeLong", "PogemaCoopFinish"):
        from . import envs
        return getattr(envs, name)
    if name in ("AnimationMonitor", "AnimationConfig", "PersistentWrapper", "AgentState", "AutoResetWrapper",
                "SingleAgentWrapper", "IsMultiAgentWrapper", "MetricsForwardingWrapper", "RuntimeMetricWrapper", "AgentsDensityWrapper"):
        from . import wrappers
        return getattr(wrappers, name)
    if name == "parallel_env":
        from .integrations.pettingzoo import parallel_env
        return parallel_env
    raise AttributeError(name)
