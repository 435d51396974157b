"""Pins the C restatement (oracle/step_oracle.c) against the Python oracle on random
scenarios for every collision_system x on_target combination, step by step."""
import itertools

import numpy as np
import pytest

from tests.helpers import make_actions, run_oracle
from tests.oracle_c import COracle


@pytest.mark.parametrize("coll,ot", list(itertools.product(("priority", "block_both", "soft"),
                                                           ("finish", "nothing", "restart"))))
@pytest.mark.parametrize("auto_reset", [False, True])
def test_c_oracle_matches_python_oracle(coll, ot, auto_reset):
    gc = dict(size=9, density=0.15, num_agents=22, obs_radius=3, max_episode_steps=12, collision_system=coll,
              on_target=ot)
    seeds = list(range(40, 46))
    T = 30
    actions = make_actions(T, len(seeds), gc["num_agents"], seed=3)
    co = COracle.from_python_oracle(gc, seeds)
    refs = [run_oracle(gc, s, actions[:, k], auto_reset=auto_reset) for k, s in enumerate(seeds)]
    r = gc["obs_radius"]
    for t in range(T):
        out = co.run(actions[t:t + 1], auto_reset=auto_reset)
        for k in range(len(seeds)):
            ref = refs[k]
            assert np.array_equal(co.pos[k] - r, ref["pos"][t + 1]), (t, k)
            assert np.array_equal(co.tgt[k] - r, ref["tgt"][t + 1]), (t, k)
            assert np.array_equal(co.active[k], ref["active"][t + 1]), (t, k)
            assert np.array_equal(out["obs"][k], ref["obs"][t + 1]), (t, k)
            assert np.array_equal(out["rewards"][k], ref["rewards"][t]), (t, k)
            assert np.array_equal(out["terminated"][k], ref["terminated"][t]), (t, k)
            assert np.array_equal(out["truncated"][k], ref["truncated"][t]), (t, k)
