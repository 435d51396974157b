"""Shared test helpers: run the CPU oracle in lockstep and collect everything
the parity contract names (positions, targets, active flags, rewards,
terminated/truncated, every observation channel)."""
import numpy as np

from oracle import pogema_oracle as orc


def oracle_config(gc_kwargs, seed):
    kw = dict(gc_kwargs)
    kw["seed"] = int(seed)
    return orc.GridConfig(**kw)


def make_actions(num_steps, num_envs, num_agents, seed=1234):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 5, size=(num_steps, num_envs, num_agents)).astype(np.uint8)


def grid_snapshot(env):
    g = env.unwrapped.grid if hasattr(env, "unwrapped") else env.grid
    r = g.config.obs_radius
    pos = np.array(g.positions_xy, dtype=np.int32) - r
    tgt = np.array(g.finishes_xy, dtype=np.int32) - r
    act = np.array([g.is_active[i] for i in range(len(g.positions_xy))], dtype=np.uint8)
    return pos, tgt, act


def run_oracle(gc_kwargs, seed, actions, auto_reset=False):
    """actions: [T, A].  Returns dict of stacked arrays; index 0 of obs/pos/tgt/active is after reset."""
    env = orc.pogema_v0(oracle_config(gc_kwargs, seed))
    obs, infos = env.reset()
    out = {k: [] for k in ("obs", "pos", "tgt", "active", "rewards", "terminated", "truncated", "metrics", "done")}

    def snap(o):
        out["obs"].append(np.stack(o).astype(np.uint8))
        p, t, a = grid_snapshot(env)
        out["pos"].append(p)
        out["tgt"].append(t)
        out["active"].append(a)

    snap(obs)
    for t in range(actions.shape[0]):
        obs, rew, term, trunc, infos = env.step(list(actions[t]))
        done = all(term) or all(trunc)
        out["metrics"].append(infos[0].get("metrics"))
        out["done"].append(done)
        if done and auto_reset:
            obs, _ = env.reset()
        snap(obs)
        out["rewards"].append(np.array(rew, dtype=np.float32))
        out["terminated"].append(np.array(term, dtype=bool))
        out["truncated"].append(np.array(trunc, dtype=bool))
    for k in ("obs", "pos", "tgt", "active", "rewards", "terminated", "truncated"):
        out[k] = np.stack(out[k])
    return out
