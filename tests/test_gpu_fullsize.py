"""BASELINE.json configurations at their FULL sizes: the CUDA path against the C oracle on every
instance (state read back from the engine's own generator, which the CPU tests pin against the
Python oracle), plus size-independent properties and cross-variant idempotence."""
import numpy as np
import pytest

from tests.oracle_c import COLL, ONT, COracle, OrcCfg

pytestmark = pytest.mark.gpu


def padded_obstacles(obst, r):
    n, h, w = obst.shape
    full = np.zeros((n, h + 2 * r, w + 2 * r), np.uint8)
    full[:, r - 1, r - 1:w + r + 1] = 1
    full[:, r - 1:h + r + 1, r - 1] = 1
    full[:, h + r, r - 1:w + r + 1] = 1
    full[:, r - 1:h + r + 1, w + r] = 1
    full[:, r:h + r, r:w + r] = obst
    return full


def c_oracle_from_engine(env):
    gc = env.grid_config
    r = gc.obs_radius
    obst = padded_obstacles(env.get_obstacles(), r)
    cfg = OrcCfg(obst.shape[1], obst.shape[2], gc.num_agents, r, COLL[gc.collision_system], ONT[gc.on_target],
                 gc.max_episode_steps)
    pos = env.get_agents_xy().cpu().numpy().astype(np.int32) + r
    tgt = env.get_targets_xy().cpu().numpy().astype(np.int32) + r
    return COracle(cfg, obst, pos, tgt)


def torch_reference_obs(env):
    """Plain torch restatement of the three observation channels from the engine's state."""
    import torch
    gc = env.grid_config
    r = gc.obs_radius
    dev = env.device
    obst = torch.from_numpy(padded_obstacles(env.get_obstacles(), r)).to(dev)
    N, PH, PW = obst.shape
    xy = env.get_agents_xy().long() + r
    txy = env.get_targets_xy().long() + r
    act = env.is_active
    occ = torch.zeros((N, PH * PW), dtype=torch.uint8, device=dev)
    lin = xy[..., 0] * PW + xy[..., 1]
    occ.scatter_(1, lin, act.to(torch.uint8))          # at most one active agent per cell
    # inactive agents must not clear a cell an active agent stands on: scatter inactive first, active last
    occ.zero_()
    nidx = torch.arange(N, device=dev)[:, None].expand_as(lin)
    occ[nidx[act], lin[act]] = 1
    occ = occ.view(N, PH, PW)
    d = torch.arange(-r, r + 1, device=dev)
    rows = (xy[..., 0, None] + d)[..., :, None]          # [N, A, D, 1]
    cols = (xy[..., 1, None] + d)[..., None, :]          # [N, A, 1, D]
    nn = torch.arange(N, device=dev)[:, None, None, None]
    ch0 = obst[nn, rows, cols]
    ch1 = occ[nn, rows, cols]
    dx = (xy[..., 0] - txy[..., 0]).clamp(-r, r)
    dy = (xy[..., 1] - txy[..., 1]).clamp(-r, r)
    D = 2 * r + 1
    ch2 = torch.zeros((N, xy.shape[1], D * D), dtype=torch.uint8, device=dev)
    ch2.scatter_(2, ((r - dx) * D + (r - dy))[..., None], 1)
    return torch.stack((ch0, ch1, ch2.view(N, -1, D, D)), dim=2)


def check_invariants(env, prev_xy=None, was_reset=None):
    import torch
    gc = env.grid_config
    r = gc.obs_radius
    xy = env.get_agents_xy().long()
    act = env.is_active
    H, W = env.get_obstacles().shape[1:]
    lin = xy[..., 0] * W + xy[..., 1]
    assert int(xy.min()) >= 0 and int(xy[..., 0].max()) < H and int(xy[..., 1].max()) < W
    # no two ACTIVE agents on one cell
    key = torch.where(act, lin, -1 - torch.arange(lin.shape[1], device=lin.device)[None, :].expand_as(lin))
    srt = key.sort(dim=1).values
    assert not bool((srt[:, 1:] == srt[:, :-1]).any())
    # no agent on an obstacle
    obst = torch.from_numpy(env.get_obstacles()).to(lin.device).view(lin.shape[0], -1)
    assert int(obst.gather(1, lin).sum()) == 0
    if prev_xy is not None:
        moved = (xy - prev_xy).abs().sum(-1)
        ok = moved <= 1
        if was_reset is not None:
            ok = ok | was_reset[:, None]
        assert bool(ok.all())
    return xy


def run_case(gc_kwargs, N, T, compare_oracle=True, team_threads=0, seed0=0):
    """compare_oracle: True = C oracle on EVERY instance (lifelong configurations: built from the Python oracle's
    own resets in a process pool - numpy generators, component tables; the others: from the engine's state after
    reset, which tests/test_native_generator.py and test_gpu_devgen.py pin to the Python oracle)."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    env = BatchedPogema(GridConfig(**gc_kwargs), num_envs=N, seeds=np.arange(seed0, seed0 + N), auto_reset=True,
                        team_threads=team_threads)
    obs = env.reset()
    assert torch.equal(obs, torch_reference_obs(env))
    co = None
    if compare_oracle and gc_kwargs.get("on_target") == "restart":
        co = COracle.from_python_oracle_parallel(gc_kwargs, list(range(seed0, seed0 + N)))
        r0 = gc_kwargs["obs_radius"]
        assert np.array_equal(env.get_agents_xy().cpu().numpy() + r0, co.pos)   # same tasks to begin with
        assert np.array_equal(env.get_targets_xy().cpu().numpy() + r0, co.tgt)
    elif compare_oracle:
        co = c_oracle_from_engine(env)
    g = torch.Generator(device="cuda").manual_seed(7)
    A = gc_kwargs["num_agents"]
    rsum = torch.zeros((N, A), dtype=torch.float64, device="cuda")
    prev = check_invariants(env)
    acts = []
    for t in range(T):
        a = env.sample_actions(g)
        acts.append(a.cpu().numpy())
        obs, rew, term, trunc = env.step(a)
        rsum += rew
        assert bool(((rew == 0) | (rew == 1)).all())
        prev = check_invariants(env, prev, env.episode_done)
        if t % 16 == 0 or t == T - 1:
            assert torch.equal(obs, torch_reference_obs(env)), f"obs differs from the torch reference at t={t}"
            assert bool((obs[:, :, 2].flatten(2).sum(-1) == 1).all())
    env.check_errors()
    if co is not None:
        out = co.run(np.stack(acts), auto_reset=True)
        r = gc_kwargs["obs_radius"]
        assert np.array_equal(env.get_agents_xy().cpu().numpy() + r, co.pos)
        assert np.array_equal(env.get_targets_xy().cpu().numpy() + r, co.tgt)
        assert np.array_equal(env.is_active.cpu().numpy().astype(np.uint8), co.active)
        assert np.array_equal(env.elapsed_steps.cpu().numpy(), co.elapsed)
        assert np.array_equal(obs.cpu().numpy(), out["obs"])
        assert np.array_equal(rsum.cpu().numpy(), out["rewards_sum"])
        assert np.array_equal(rew.cpu().numpy(), out["rewards"])
        assert np.array_equal(term.cpu().numpy(), out["terminated"])
        assert np.array_equal(trunc.cpu().numpy(), out["truncated"])
    return env, obs.clone(), rsum


def test_config2_full_size_against_c_oracle():
    """configs[1]: 4096 instances of 32x32, 64 agents, r=5, priority/finish - every instance, 70 steps."""
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
              collision_system="priority", on_target="finish")
    run_case(gc, 4096, 70)


def test_config2_team_variants_agree():
    import torch
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
              collision_system="priority", on_target="finish")
    ref = None
    for team in (32, 64, 128):
        env, obs, rsum = run_case(gc, 512, 40, compare_oracle=False, team_threads=team)
        if ref is None:
            ref = (obs, rsum)
        else:
            assert torch.equal(obs, ref[0]) and torch.equal(rsum, ref[1])


from pogema_b200.maps import maze_map, warehouse_map  # noqa: E402


def test_config3_lifelong_maze_soft():
    """configs[2]: 64x64 maze-like maps, 256 agents, soft collisions, on_target='restart': ALL 1024 instances
    against the C oracle built from the Python oracle's resets (lifelong generators and component tables come
    from numpy there), 72 steps = across the time-limit auto-reset at step 64; invariants and the torch
    observation reference on the way."""
    m = maze_map(64, 3)
    gc = dict(map=m.tolist(), num_agents=256, obs_radius=5, max_episode_steps=64, collision_system="soft",
              on_target="restart")
    env, obs, rsum = run_case(gc, 1024, 72)
    assert float(rsum.sum()) > 0                     # goals are reached and replaced


def test_config4_warehouse_block_both():
    """configs[3]: 256x256 warehouse-style maps, 1024 agents, block_both, r=5, 512 instances, 70 steps (across
    the time-limit auto-reset), every instance against the C oracle."""
    m = warehouse_map(256)
    gc = dict(map=m.tolist(), num_agents=1024, obs_radius=5, max_episode_steps=64, collision_system="block_both",
              on_target="finish")
    run_case(gc, 512, 70)


@pytest.mark.parametrize("coll", ["priority", "block_both", "soft"])
@pytest.mark.parametrize("ot", ["finish", "nothing", "restart"])
def test_all_nine_modes_at_config2_full_size(coll, ot):
    """All collision_system x on_target combinations at the configs[1] size (4096 x 64 agents, r=5), 70 steps,
    every instance against the C oracle."""
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64,
              collision_system=coll, on_target=ot)
    run_case(gc, 4096, 70, seed0=7000)


@pytest.mark.parametrize("r", [3, 5, 7])
def test_config5_per_gpu_share(r):
    """configs[4]: 1M agents over 8 GPUs = 131072 agents (2048 instances of 64) per GPU, r = 3/5/7."""
    gc = dict(size=32, density=0.3, num_agents=64, obs_radius=r, max_episode_steps=64,
              collision_system="priority", on_target="finish")
    run_case(gc, 2048, 20)
