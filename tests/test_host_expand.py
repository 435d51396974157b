"""The widening loop of the packed host transport (pgm_expand_bits_host, no GPU): bit k of the word stream
becomes element k of the uint8 / float32 destination - for every ISA path the library has (AVX-512BW,
AVX2, scalar), every alignment of the destination, and without touching a byte outside it."""
import os
import subprocess
import sys

import numpy as np
import pytest

from pogema_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_all():
    lib = nat.load()
    rng = np.random.default_rng(0)
    for es, dt in ((1, np.uint8), (2, np.float16), (4, np.float32)):
        for nbits in (0, 1, 7, 31, 32, 63, 64, 65, 127, 363, 1000, 3 * 11 * 11 * 64, 100003):
            for off in (0, 1, 3, 17, 63):
                words = rng.integers(0, 2 ** 32, size=(nbits + 31) // 32 + 1, dtype=np.uint32)
                ref = np.unpackbits(words.view(np.uint8), bitorder="little")[:nbits].astype(dt)
                buf = np.full(nbits + 128, 7, dtype=dt)
                dst = buf[off:off + nbits]
                assert lib.pgm_expand_bits_host(words.ctypes.data, nbits, dst.ctypes.data, es) == nat.PGM_OK
                assert np.array_equal(dst, ref), (es, nbits, off)
                assert (buf[:off] == 7).all() and (buf[off + nbits:] == 7).all(), (es, nbits, off)
    assert lib.pgm_expand_bits_host(None, 8, None, 1) == nat.PGM_ERR_INVALID
    w = np.zeros(1, np.uint32)
    assert lib.pgm_expand_bits_host(w.ctypes.data, 8, w.ctypes.data, 3) == nat.PGM_ERR_INVALID


def test_expand_native_isa():
    check_all()


@pytest.mark.parametrize("isa", ["avx2", "scalar"])
def test_expand_other_isa_paths(isa):
    # the ISA is latched at first use: run the same check in a fresh process with the path forced
    env = dict(os.environ, PGM_HOST_ISA=isa, PYTHONPATH=ROOT)
    code = "import tests.test_host_expand as t; t.check_all(); print('ok')"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
