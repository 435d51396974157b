"""Generates tests/golden/golden_v1.npz from the Python oracle (oracle/pogema_oracle.py), which
calls the real numpy Generator.  Upstream pogema is not importable in this container
(/root/reference holds README.md:1-5 only), so these are golden vectors OF THE ORACLE: they make
any later change of the oracle - or of numpy's random streams - visible, and give the GPU tests a
fixture that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.helpers import make_actions, run_oracle  # noqa: E402

CASES = []
for coll, ot in itertools.product(("priority", "block_both", "soft"), ("finish", "nothing", "restart")):
    CASES.append(dict(size=8, density=0.3, num_agents=4, obs_radius=5, max_episode_steps=64, collision_system=coll,
                      on_target=ot))
    CASES.append(dict(size=10, density=0.1, num_agents=30, obs_radius=2, max_episode_steps=20, collision_system=coll,
                      on_target=ot))
CASES.append(dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
                  collision_system="priority", on_target="finish"))
SEEDS = [0, 1, 2]
T = 24


def main():
    out = {"cases": json.dumps(CASES), "seeds": np.array(SEEDS), "T": T}
    for ci, case in enumerate(CASES):
        actions = make_actions(T, len(SEEDS), case["num_agents"], seed=100 + ci)
        out[f"c{ci}_actions"] = actions
        for k, seed in enumerate(SEEDS):
            ref = run_oracle(case, seed, actions[:, k], auto_reset=False)
            for key in ("pos", "tgt", "active", "rewards", "terminated", "truncated"):
                out[f"c{ci}_s{k}_{key}"] = ref[key]
            obs = ref["obs"]
            out[f"c{ci}_s{k}_obs_shape"] = np.array(obs.shape)
            out[f"c{ci}_s{k}_obs_bits"] = np.packbits(obs.reshape(-1))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
