"""The C-ABI library loads (no GPU needed) and exports every symbol include/pgm_b200.h declares."""
import ctypes as C
import os
import re

from pogema_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = nat.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pgm_b200.h but not exported"
    assert sorted(nat.EXPORTS) == names, set(names) ^ set(nat.EXPORTS)
    assert lib.pgm_abi_version() == nat.PGM_ABI_VERSION


def test_config_struct_layout():
    assert C.sizeof(nat.PgmConfig) == 16 * 4


def test_errors_do_not_throw_across_the_abi():
    lib = nat.load()
    h = C.c_void_p()
    cfg = nat.PgmConfig()
    cfg.abi_version = 999
    assert lib.pgm_create(C.byref(cfg), C.byref(h)) == nat.PGM_ERR_INVALID
    assert b"abi_version" in lib.pgm_last_error()
    assert lib.pgm_step(None, None, 1, None, None, None, None, None) == nat.PGM_ERR_INVALID


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", "/nonexistent/libpgm_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nat.load()
