"""ctypes binding of the C-ABI in ``include/pgm_b200.h`` (libpgm_b200.so).

There is no Python/CPU fallback: if the CUDA library has not been built the
import of anything that computes fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGM_B200_LIB: another build of the same library (A/B timing of two builds on one box); there is still no fallback
LIB_PATH = os.environ.get("PGM_B200_LIB") or os.path.join(_HERE, "_lib", "libpgm_b200.so")

PGM_ABI_VERSION = 1
PGM_OK = 0
PGM_ERR_INVALID, PGM_ERR_CUDA, PGM_ERR_UNSUPPORTED, PGM_ERR_OVERFLOW, PGM_ERR_STATE, PGM_ERR_ACTION = -1, -2, -3, -4, -5, -6
COLLISION = {"priority": 0, "block_both": 1, "soft": 2}
ON_TARGET = {"finish": 0, "nothing": 1, "restart": 2}
OBS_FORMAT = {"u8": 0, "bits": 1, "f32": 2, "f16": 3}
STATE_POSITIONS, STATE_TARGETS, STATE_ACTIVE, STATE_ELAPSED, STATE_OBSTACLES, STATE_WAS_ON_GOAL, STATE_EPISODE_DONE, STATE_METRICS, STATE_SEEDS, STATE_SOLVE_COSTS = range(10)

# every symbol include/pgm_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "pgm_last_error", "pgm_abi_version", "pgm_create", "pgm_destroy", "pgm_obs_bytes",
    "pgm_obs_instance_stride", "pgm_generate", "pgm_generate_device", "pgm_generate_host", "pgm_set_tasks", "pgm_reset", "pgm_observe", "pgm_step",
    "pgm_step_many", "pgm_step_host", "pgm_step_host_ex", "pgm_observe_host", "pgm_get_state", "pgm_state_ptr", "pgm_checkpoint_bytes", "pgm_checkpoint_save",
    "pgm_checkpoint_load", "pgm_check_errors", "pgm_launch_count", "pgm_plan", "pgm_set_debug_buffer",
    "pgm_set_host_transport", "pgm_host_transport_info", "pgm_expand_bits_host", "pgm_host_fill_gbps",
]


class PgmConfig(C.Structure):
    _fields_ = [(name, C.c_int32) for name in (
        "abi_version", "device", "num_envs", "num_agents", "height", "width", "obs_radius",
        "max_episode_steps", "collision_system", "on_target", "auto_reset", "obs_format", "team_threads",
    )] + [("reserved", C.c_int32 * 3)]


class PgmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"pgm error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load libpgm_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built. Run `make` at the repository root "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). pogema_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.pgm_last_error.restype = C.c_char_p
    lib.pgm_abi_version.restype = C.c_int
    lib.pgm_create.argtypes = [C.POINTER(PgmConfig), C.POINTER(vp)]
    lib.pgm_destroy.argtypes = [vp]
    lib.pgm_obs_bytes.argtypes = [vp]
    lib.pgm_obs_bytes.restype = i64
    lib.pgm_obs_instance_stride.argtypes = [vp]
    lib.pgm_obs_instance_stride.restype = i64
    lib.pgm_generate.argtypes = [vp, i32, i32, vp, C.c_double, vp, i32, C.POINTER(i32), vp]
    lib.pgm_generate_device.argtypes = [vp, i32, i32, vp, C.c_double, vp, C.POINTER(i32), vp]
    lib.pgm_generate_host.argtypes = [i32, i32, i32, i32, C.c_double, i32, vp, C.c_uint64, vp, vp, vp, vp, vp]
    lib.pgm_set_tasks.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    lib.pgm_reset.argtypes = [vp, vp, vp]
    lib.pgm_observe.argtypes = [vp, vp, vp]
    lib.pgm_step.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pgm_step_many.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp, vp, vp]
    lib.pgm_step_host.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pgm_step_host_ex.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.pgm_observe_host.argtypes = [vp, vp, vp]
    lib.pgm_get_state.argtypes = [vp, i32, vp, i64, vp]
    lib.pgm_state_ptr.argtypes = [vp, i32]
    lib.pgm_state_ptr.restype = vp
    lib.pgm_checkpoint_bytes.argtypes = [vp]
    lib.pgm_checkpoint_bytes.restype = i64
    lib.pgm_checkpoint_save.argtypes = [vp, vp, i64, vp]
    lib.pgm_checkpoint_load.argtypes = [vp, vp, i64, vp]
    lib.pgm_check_errors.argtypes = [vp, vp]
    lib.pgm_set_debug_buffer.argtypes = [vp, vp]
    lib.pgm_launch_count.argtypes = [vp]
    lib.pgm_launch_count.restype = i64
    lib.pgm_plan.argtypes = [vp, C.POINTER(i32), i32]
    lib.pgm_set_host_transport.argtypes = [vp, i32, i32]
    lib.pgm_host_transport_info.argtypes = [vp, C.POINTER(i64), i32]
    lib.pgm_expand_bits_host.argtypes = [vp, i64, vp, i32]
    lib.pgm_host_fill_gbps.argtypes = [vp, i64, i32, i32]
    lib.pgm_host_fill_gbps.restype = C.c_double
    if lib.pgm_abi_version() != PGM_ABI_VERSION:
        raise RuntimeError("libpgm_b200.so ABI version mismatch; rebuild with `make`")
    _lib = lib
    return lib


def check(code):
    if code != PGM_OK:
        msg = load().pgm_last_error().decode("utf-8", "replace")
        if code == PGM_ERR_OVERFLOW:
            raise OverflowError(msg)
        if code == PGM_ERR_ACTION:
            raise IndexError(msg)
        raise PgmError(code, msg)
