"""Golden trajectories (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py):
 - CPU: the oracle still reproduces them (pins the oracle and numpy's random streams);
 - GPU: the CUDA engine reproduces them without the oracle at run time."""
import json
import os

import numpy as np
import pytest

from tests.helpers import run_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
KEYS = ("pos", "tgt", "active", "rewards", "terminated", "truncated")


def load_golden():
    z = np.load(GOLDEN)
    cases = json.loads(str(z["cases"]))
    return z, cases, [int(s) for s in z["seeds"]], int(z["T"])


def golden_obs(z, ci, k):
    shape = tuple(z[f"c{ci}_s{k}_obs_shape"])
    return np.unpackbits(z[f"c{ci}_s{k}_obs_bits"])[:int(np.prod(shape))].reshape(shape)


def test_oracle_reproduces_golden():
    z, cases, seeds, T = load_golden()
    for ci, case in enumerate(cases):
        actions = z[f"c{ci}_actions"]
        for k, seed in enumerate(seeds):
            ref = run_oracle(case, seed, actions[:, k])
            for key in KEYS:
                assert np.array_equal(ref[key], z[f"c{ci}_s{k}_{key}"]), (ci, k, key)
            assert np.array_equal(ref["obs"], golden_obs(z, ci, k)), (ci, k)


@pytest.mark.gpu
def test_engine_reproduces_golden():
    from tests.test_gpu_parity import run_gpu
    z, cases, seeds, T = load_golden()
    for ci, case in enumerate(cases):
        actions = z[f"c{ci}_actions"]
        gpu = run_gpu(case, seeds, actions)
        for k in range(len(seeds)):
            for key in KEYS:
                assert np.array_equal(gpu[key][:, k], z[f"c{ci}_s{k}_{key}"]), (ci, k, key)
            assert np.array_equal(gpu["obs"][:, k], golden_obs(z, ci, k)), (ci, k)
