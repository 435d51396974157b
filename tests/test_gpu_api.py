"""GPU tests of the drop-in API surface: pogema_v0 lists, metrics, PettingZoo adapter, hand-derived
scenarios, observation formats, checkpoint/resume, error behaviour."""
import numpy as np
import pytest

from oracle import pogema_oracle as orc
from tests.scenarios import SCENARIOS, clean_map

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sc", SCENARIOS, ids=[s["name"] for s in SCENARIOS])
@pytest.mark.parametrize("system", ["priority", "block_both", "soft"])
def test_collision_scenarios_on_gpu(sc, system):
    from pogema_b200 import GridConfig, pogema_v0
    env = pogema_v0(GridConfig(map=clean_map(sc["map"]), obs_radius=2, collision_system=system,
                               on_target="nothing", seed=0))
    env.reset()
    env.step(sc["actions"])
    got = [tuple(p) for p in env.get_agents_xy(ignore_borders=True)]
    assert got == sc["expect"][system]
    env.close()


@pytest.mark.parametrize("ot", ["finish", "nothing", "restart"])
@pytest.mark.parametrize("coll", ["priority", "block_both", "soft"])
def test_pogema_v0_matches_oracle_lists(coll, ot):
    """Same calls, same return types and values as the (restated) reference, incl. infos and metrics."""
    from pogema_b200 import GridConfig, pogema_v0
    kw = dict(size=8, density=0.3, num_agents=4, obs_radius=5, max_episode_steps=16, collision_system=coll,
              on_target=ot, seed=11)
    env = pogema_v0(GridConfig(**kw))
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    obs, infos = env.reset()
    robs, rinfos = ref.reset()
    assert isinstance(obs, list) and len(obs) == 4 and obs[0].dtype == np.float32 and obs[0].shape == (3, 11, 11)
    assert all(np.array_equal(a, b) for a, b in zip(obs, robs)) and infos == rinfos
    assert env.observation_space.shape == (3, 11, 11) and env.action_space.n == 5
    for episode in range(2):
        for t in range(16):
            a = env.sample_actions()
            assert np.array_equal(a, ref.sample_actions())          # same ActionsSampler stream
            o, r, te, tr, inf = env.step(a)
            ro, rr, rte, rtr, rinf = ref.step(a)
            assert all(np.array_equal(x, y) for x, y in zip(o, ro))
            assert r == rr and te == rte and tr == rtr
            assert all(isinstance(v, float) for v in r) and all(isinstance(v, bool) for v in te + tr)
            assert inf == rinf, (inf, rinf)
            assert [tuple(p) for p in env.get_agents_xy()] == [tuple(p) for p in ref.unwrapped.grid.get_agents_xy()]
            assert ([tuple(p) for p in env.get_targets_xy(ignore_borders=True)] ==
                    [tuple(p) for p in ref.unwrapped.grid.get_targets_xy(ignore_borders=True)])
            if all(te) or all(tr):
                assert "metrics" in inf[0]
                break
        obs, infos = env.reset()
        robs, rinfos = ref.reset()
        assert all(np.array_equal(a, b) for a, b in zip(obs, robs))
    assert np.array_equal(env.get_obstacles(), ref.unwrapped.grid.get_obstacles())
    env.close()


def test_parallel_env_dicts():
    from pogema_b200 import GridConfig, parallel_env
    env = parallel_env(GridConfig(size=8, num_agents=3, seed=2, max_episode_steps=5))
    obs, infos = env.reset()
    assert env.possible_agents == ["player_0", "player_1", "player_2"] and set(obs) == set(env.possible_agents)
    for t in range(5):
        obs, rew, term, trunc, infos = env.step({a: 0 for a in env.agents})
    assert all(trunc.values()) and env.agents == []
    env.close()


def test_bits_format_equals_u8():
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    for r in (2, 5, 7):
        gc = GridConfig(size=16, density=0.3, num_agents=20, obs_radius=r, max_episode_steps=32,
                        collision_system="soft", on_target="restart", seed=1)
        a = BatchedPogema(gc, num_envs=6, auto_reset=True)
        b = BatchedPogema(gc, num_envs=6, auto_reset=True, obs_format="bits")
        oa, ob = a.reset(), b.reset()
        g = torch.Generator(device="cuda").manual_seed(0)
        for t in range(40):
            D = 2 * r + 1
            bits = ob.cpu().numpy().view(np.uint32)
            unpacked = np.unpackbits(bits.view(np.uint8), bitorder="little").reshape(6, 20, -1)[:, :, :3 * D * D]
            assert np.array_equal(unpacked.reshape(6, 20, 3, D, D), oa.cpu().numpy())
            act = a.sample_actions(g)
            oa = a.step(act)[0]
            ob = b.step(act)[0]


@pytest.mark.parametrize("fmt", ["f32", "f16"])
def test_float_formats_equal_u8(fmt):
    """obs_format f32 (the reference's dtype) / f16 (half-precision policies): exactly u8.float() / u8.half(),
    for single steps, multi-step launches and odd alignments (r=2: 75 elements per agent)."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    dt = {"f32": torch.float32, "f16": torch.float16}[fmt]
    for r, agents in ((2, 7), (5, 20), (9, 5), (5, 64)):
        gc = GridConfig(size=16, density=0.3, num_agents=agents, obs_radius=r, max_episode_steps=16,
                        collision_system="priority", on_target="finish", seed=2)
        a = BatchedPogema(gc, num_envs=5, auto_reset=True)
        b = BatchedPogema(gc, num_envs=5, auto_reset=True, obs_format=fmt)
        oa, ob = a.reset(), b.reset()
        assert ob.dtype == dt and ob.shape == oa.shape
        g = torch.Generator(device="cuda").manual_seed(0)
        for t in range(30):
            assert torch.equal(ob, oa.to(dt))
            act = a.sample_actions(g)
            oa = a.step(act)[0]
            ob = b.step(act)[0]
        acts = torch.stack([a.sample_actions(g) for _ in range(6)])
        ra, rb = a.rollout(acts), b.rollout(acts)
        assert torch.equal(rb[0], ra[0].to(dt)) and torch.equal(ra[1], rb[1])


def test_checkpoint_resume_and_host_step():
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = GridConfig(size=16, density=0.3, num_agents=12, obs_radius=3, max_episode_steps=20,
                    collision_system="priority", on_target="restart", seed=4)
    env = BatchedPogema(gc, num_envs=5, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [env.sample_actions(g) for _ in range(30)]
    for t in range(10):
        env.step(acts[t])
    sd = env.state_dict()
    first = [tuple(x.clone() for x in env.step(acts[t])) for t in range(10, 30)]
    env.load_state_dict(sd)
    # resume through the C-ABI host-buffer call and compare with the device-pointer path
    n, a = 5, 12
    obs = np.empty(env.engine.obs_shape(), np.uint8)
    rew = np.empty((n, a), np.float32)
    te = np.empty((n, a), np.uint8)
    tr = np.empty((n, a), np.uint8)
    o2, r2, te2, tr2 = env.step_host(acts[10].cpu().numpy())          # convenience wrapper, same call
    assert np.array_equal(o2, first[0][0].cpu().numpy()) and np.array_equal(r2, first[0][1].cpu().numpy())
    env.load_state_dict(sd)
    for t in range(10, 30):
        env.engine.step_host(acts[t].cpu().numpy(), obs, rew, te, tr)
        o, r, term, trunc = first[t - 10]
        assert np.array_equal(obs, o.cpu().numpy()) and np.array_equal(rew, r.cpu().numpy())
        assert np.array_equal(te.astype(bool), term.cpu().numpy()) and np.array_equal(tr.astype(bool), trunc.cpu().numpy())


@pytest.mark.parametrize("ot", ["finish", "restart"])
def test_checkpoint_with_reseeding_auto_reset_carries_the_tasks(ot):
    """auto_reset='reseed' rebuilds maps / tasks / lifelong tables every episode: a checkpoint must restore them too
    (into the same env after further reseeds, and into a fresh env built from the initial seeds)."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = GridConfig(size=12, density=0.3, num_agents=9, obs_radius=3, max_episode_steps=6,
                    collision_system="soft", on_target=ot, seed=11)
    env = BatchedPogema(gc, num_envs=6, auto_reset="reseed")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    acts = [env.sample_actions(g) for _ in range(40)]
    for t in range(15):                       # two reseeds happened: tasks differ from the initial ones
        env.step(acts[t])
    assert not np.array_equal(env.current_seeds(), env.seeds)
    sd = env.state_dict()
    seeds_at_save = env.current_seeds().copy()
    obst_at_save = env.get_obstacles().copy()
    want = [tuple(x.clone() for x in env.step(acts[t])) for t in range(15, 40)]
    final = env.engine.checkpoint()
    for target in (env, BatchedPogema(gc, num_envs=6, auto_reset="reseed")):
        target.load_state_dict(sd)
        assert np.array_equal(target.current_seeds(), seeds_at_save)
        assert np.array_equal(target.get_obstacles(), obst_at_save)
        for t in range(15, 40):
            got = target.step(acts[t])
            for x, y in zip(got, want[t - 15]):
                assert torch.equal(x, y), (ot, t)
        assert np.array_equal(target.engine.checkpoint(), final)


def test_checkpoint_of_another_engine_is_rejected():
    from pogema_b200 import BatchedPogema, GridConfig
    from pogema_b200._native import PgmError
    kw = dict(size=10, density=0.2, num_agents=6, obs_radius=2, max_episode_steps=8)
    a = BatchedPogema(GridConfig(seed=1, **kw), num_envs=4)
    blob = a.engine.checkpoint()
    b = BatchedPogema(GridConfig(seed=1, collision_system="soft", **kw), num_envs=4)   # same byte size, other mode
    with pytest.raises(PgmError, match="another engine"):
        b.engine.restore(blob)
    c = BatchedPogema(GridConfig(seed=2, **kw), num_envs=4)                           # same shape, other tasks
    with pytest.raises(PgmError, match="different task seeds"):
        c.engine.restore(blob)
    with pytest.raises(PgmError, match="magic"):
        a.engine.restore(np.zeros_like(blob))
    a.engine.restore(blob)


def test_action_dtypes_are_checked():
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    env = BatchedPogema(GridConfig(size=8, density=0.1, num_agents=3, obs_radius=2, seed=0), num_envs=2, auto_reset=False)
    env.reset()
    with pytest.raises(TypeError):
        env.step_host(np.ones((2, 3), np.float64))
    with pytest.raises(TypeError):
        env.rollout(torch.ones((4, 2, 3), device="cuda"))
    # wide integer actions are read in full: 256 / -256 do not alias to 'stay'
    before = env.get_agents_xy().clone()
    env.step(torch.full((2, 3), 256, dtype=torch.int32, device="cuda"))
    with pytest.raises(IndexError):
        env.check_errors()
    env.step(torch.full((2, 3), -256, dtype=torch.int64, device="cuda"))
    with pytest.raises(IndexError):
        env.check_errors()
    env.step(torch.full((2, 3), 0x0400, dtype=torch.int16, device="cuda"))
    with pytest.raises(IndexError):
        env.check_errors()
    assert torch.equal(env.get_agents_xy(), before)        # invalid actions are treated as 'stay'


def test_batched_metrics_match_oracle_wrappers():
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    for ot in ("finish", "nothing", "restart"):
        kw = dict(size=8, density=0.2, num_agents=5, obs_radius=2, max_episode_steps=12, collision_system="priority",
                  on_target=ot)
        seeds = [0, 1, 2, 3]
        env = BatchedPogema(GridConfig(**kw), num_envs=4, seeds=seeds, auto_reset=True)
        env.reset()
        refs = [orc.pogema_v0(orc.GridConfig(seed=s, **kw)) for s in seeds]
        for r in refs:
            r.reset()
        rng = np.random.default_rng(0)
        last = [None] * 4
        for t in range(40):
            acts = rng.integers(0, 5, size=(4, 5)).astype(np.uint8)
            env.step(torch.from_numpy(acts).cuda())
            done = env.episode_done.cpu().numpy()
            for k, r in enumerate(refs):
                _, _, te, tr, info = r.step(list(acts[k]))
                fin = all(te) or all(tr)
                assert bool(done[k]) == fin
                if fin:
                    last[k] = info[0]["metrics"]
                    r.reset()
            m = env.metrics()
            for k in range(4):
                if last[k] is not None:
                    for key, val in last[k].items():
                        assert m[key][k] == val, (ot, t, k, key, m[key][k], val)


def test_error_behaviour():
    import torch
    from pogema_b200 import BatchedPogema, GridConfig, pogema_v0
    with pytest.raises(OverflowError):
        BatchedPogema(GridConfig(size=4, density=0.9, num_agents=8, seed=0), num_envs=2)
    with pytest.raises(OverflowError):
        pogema_v0(GridConfig(size=4, density=0.9, num_agents=8, seed=0)).reset()
    env = BatchedPogema(GridConfig(size=8, num_agents=2, seed=0), num_envs=2)
    env.reset()
    env.step(torch.full((2, 2), 7, dtype=torch.uint8, device="cuda"))
    with pytest.raises(IndexError):
        env.check_errors()
    with pytest.raises(ValueError):
        env.step(torch.zeros((3, 2), dtype=torch.uint8, device="cuda"))
    single = pogema_v0(GridConfig(size=8, num_agents=2, seed=0))
    single.reset()
    with pytest.raises(IndexError):
        single.step([5, 0])
    with pytest.raises(AssertionError):
        single.step([0])


@pytest.mark.parametrize("dtype", ["uint8", "int32", "int64"])
def test_action_dtypes(dtype):
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = GridConfig(size=10, num_agents=6, seed=3, obs_radius=2)
    a = BatchedPogema(gc, num_envs=3)
    b = BatchedPogema(gc, num_envs=3)
    a.reset(), b.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for t in range(10):
        act = a.sample_actions(g)
        oa = a.step(act)[0]
        ob = b.step(act.to(getattr(torch, dtype)))[0]
        assert torch.equal(oa, ob)


def test_explicit_agents_on_generated_map():
    from pogema_b200 import GridConfig, pogema_v0
    kw = dict(size=8, density=0.3, seed=9, agents_xy=[[0, 0], [7, 7]], targets_xy=[[3, 3], [4, 4]], obs_radius=3)
    env = pogema_v0(GridConfig(**kw))
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    obs, _ = env.reset()
    robs, _ = ref.reset()
    assert all(np.array_equal(a, b) for a, b in zip(obs, robs))
    for t in range(12):
        a = ref.sample_actions()
        o, r, te, tr, _ = env.step(a)
        ro, rr, rte, rtr, _ = ref.step(a)
        assert all(np.array_equal(x, y) for x, y in zip(o, ro)) and r == rr and te == rte


@pytest.mark.parametrize("kind", ["POMAPF", "MAPF"])
def test_dict_observation_types(kind):
    """upstream envs.py :: _pomapf_obs / _mapf_obs: dict observations with relative / global coordinates."""
    from pogema_b200 import GridConfig, pogema_v0
    kw = dict(size=8, density=0.3, num_agents=3, obs_radius=2, max_episode_steps=12, seed=5, observation_type=kind,
              on_target="restart")
    env = pogema_v0(GridConfig(**kw))
    ref = orc.pogema_v0(orc.GridConfig(**kw))

    def same(a, b):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert set(x) == set(y)
            for key in x:
                assert np.array_equal(np.asarray(x[key]), np.asarray(y[key])), key

    o, _ = env.reset()
    ro, _ = ref.reset()
    same(o, ro)
    for t in range(12):
        a = ref.sample_actions()
        o = env.step(a)[0]
        ro = ref.step(a)[0]
        same(o, ro)


@pytest.mark.parametrize("ot", ["finish", "restart", "nothing"])
def test_interleaved_api_calls_equal_plain_stepping(ot):
    """step / rollout / observe / checkpoint-restore / reset interleaved at random must leave the engine
    in exactly the state plain stepping produces."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    gc = GridConfig(size=12, density=0.2, num_agents=14, obs_radius=3, max_episode_steps=11,
                    collision_system="soft", on_target=ot, seed=0)
    a = BatchedPogema(gc, num_envs=9, auto_reset=True)
    b = BatchedPogema(gc, num_envs=9, auto_reset=True)
    a.reset(), b.reset()
    rng = np.random.default_rng(5)
    g = torch.Generator(device="cuda").manual_seed(3)
    saved = None
    for it in range(60):
        op = rng.integers(0, 6)
        if op <= 1:
            act = a.sample_actions(g)
            ra, rb = a.step(act), b.step(act)
            assert all(torch.equal(x, y) for x, y in zip(ra, rb))
        elif op == 2:
            K = int(rng.integers(1, 9))
            acts = torch.stack([a.sample_actions(g) for _ in range(K)])
            obs, rew, term, trunc = a.rollout(acts)
            for k in range(K):
                o, r, te, tr = b.step(acts[k])
                assert torch.equal(obs[k], o) and torch.equal(rew[k], r) and torch.equal(term[k], te) and torch.equal(trunc[k], tr)
        elif op == 3:
            assert torch.equal(a.observe(), b.observe())
        elif op == 4:
            if saved is None or rng.random() < 0.5:
                saved = (a.state_dict(), b.state_dict())
            else:
                a.load_state_dict(saved[0]), b.load_state_dict(saved[1])
        else:
            assert torch.equal(a.reset(), b.reset())
        assert torch.equal(a._state(), b._state()) and torch.equal(a.elapsed_steps, b.elapsed_steps)
        assert np.array_equal(a.engine.checkpoint(), b.engine.checkpoint())
    ma, mb = a.metrics(), b.metrics()
    assert all(np.array_equal(ma[k], mb[k]) for k in ma)


def test_list_env_auto_reset_flag():
    """GridConfig(auto_reset=True): upstream AutoResetWrapper semantics on the list API."""
    from pogema_b200 import GridConfig, pogema_v0
    kw = dict(size=8, density=0.2, num_agents=3, obs_radius=2, max_episode_steps=5, seed=4)
    env = pogema_v0(GridConfig(auto_reset=True, **kw))
    ref = orc.pogema_v0(orc.GridConfig(**kw))
    env.reset()
    ref.reset()
    for t in range(17):
        a = ref.sample_actions()
        o, r, te, tr, inf = env.step(a)
        ro, rr, rte, rtr, rinf = ref.step(a)
        if all(rte) or all(rtr):
            ro, _ = ref.reset()
        assert all(np.array_equal(x, y) for x, y in zip(o, ro)) and r == rr and te == rte and tr == rtr


def test_c_abi_misuse_returns_status_codes():
    """Call-order and argument errors come back as negative status codes with a message, never as a crash."""
    import ctypes as C
    from pogema_b200 import _native as nat
    lib = nat.load()
    cfg = nat.PgmConfig()
    cfg.abi_version = nat.PGM_ABI_VERSION
    cfg.device, cfg.num_envs, cfg.num_agents, cfg.height, cfg.width = 0, 2, 3, 8, 8
    cfg.obs_radius, cfg.max_episode_steps = 2, 8
    h = C.c_void_p()
    assert lib.pgm_create(C.byref(cfg), C.byref(h)) == nat.PGM_OK
    dummy = C.c_void_p(16)
    assert lib.pgm_step(h, dummy, 1, None, dummy, dummy, dummy, None) == nat.PGM_ERR_STATE      # no tasks yet
    assert b"before pgm_generate" in lib.pgm_last_error()
    assert lib.pgm_reset(h, None, None) == nat.PGM_ERR_STATE
    assert lib.pgm_step(h, dummy, 3, None, dummy, dummy, dummy, None) == nat.PGM_ERR_INVALID    # bad itemsize
    seeds = (C.c_uint64 * 2)(1, 2)
    assert lib.pgm_generate(h, 1, 2, seeds, C.c_double(0.3), None, 1, None, None) == nat.PGM_ERR_INVALID  # range
    assert lib.pgm_generate(h, 0, 2, seeds, C.c_double(1.5), None, 1, None, None) == nat.PGM_ERR_INVALID  # density
    assert lib.pgm_generate(h, 0, 2, seeds, C.c_double(0.3), None, 1, None, None) == nat.PGM_OK
    assert lib.pgm_get_state(h, 99, dummy, 8, None) == nat.PGM_ERR_INVALID
    small = (C.c_uint8 * 4)()
    assert lib.pgm_get_state(h, nat.STATE_POSITIONS, small, 4, None) == nat.PGM_ERR_INVALID     # buffer too small
    assert lib.pgm_destroy(h) == nat.PGM_OK
    bad = nat.PgmConfig()
    bad.abi_version = nat.PGM_ABI_VERSION
    bad.device, bad.num_envs, bad.num_agents, bad.height, bad.width, bad.obs_radius, bad.max_episode_steps = 99, 1, 1, 8, 8, 2, 8
    assert lib.pgm_create(C.byref(bad), C.byref(h)) == nat.PGM_ERR_INVALID                      # no such device


def test_env_classes_fix_their_on_target_semantics():
    """upstream: the CLASS decides the step semantics (PogemaLifeLong(GridConfig(on_target='finish')) is lifelong)."""
    from pogema_b200 import GridConfig, Pogema, PogemaBase, PogemaCoopFinish, PogemaLifeLong, pogema_v0
    gc = GridConfig(size=8, density=0.2, num_agents=3, obs_radius=2, seed=1, on_target="finish")
    assert PogemaLifeLong(gc).grid_config.on_target == "restart"
    assert PogemaCoopFinish(gc).grid_config.on_target == "nothing"
    assert Pogema(gc.model_copy(update=dict(on_target="restart"))).grid_config.on_target == "finish"
    for ot, cls in (("finish", Pogema), ("restart", PogemaLifeLong), ("nothing", PogemaCoopFinish)):
        env = pogema_v0(gc.model_copy(update=dict(on_target=ot)))
        assert type(env) is cls and isinstance(env, PogemaBase)
    # lifelong semantics really run: nobody terminates, targets are replaced
    env = PogemaLifeLong(gc)
    ref = orc.pogema_v0(orc.GridConfig(size=8, density=0.2, num_agents=3, obs_radius=2, seed=1, on_target="restart"))
    env.reset(), ref.reset()
    rng = np.random.default_rng(0)
    for t in range(40):
        a = [int(x) for x in rng.integers(0, 5, size=3)]
        o, r, te, tr, _ = env.step(a)
        ro, rr, rte, rtr, _ = ref.step(a)
        assert list(r) == list(rr) and list(te) == list(rte) and list(tr) == list(rtr)
        assert all(np.array_equal(x, y) for x, y in zip(o, ro))


def test_groups_on_their_own_streams_equal_one_env():
    """BatchedPogema.groups: G envs over contiguous shares of the instances, each stepping on its own stream
    (interleaved, as a closed loop with double-buffered sampling would) == one env over all instances."""
    import torch
    from pogema_b200 import BatchedPogema, GridConfig
    kw = dict(size=16, density=0.2, num_agents=32, obs_radius=4, max_episode_steps=9, collision_system="soft",
              on_target="restart")
    n, G = 12, 3
    one = BatchedPogema(GridConfig(**kw), num_envs=n, auto_reset=True)
    parts = BatchedPogema.groups(GridConfig(**kw), n, groups=G, auto_reset=True)
    assert len(parts) == G and len({st.cuda_stream for _, st in parts}) == G
    ref = one.reset().clone()
    got = []
    for env, st in parts:
        with torch.cuda.stream(st):
            got.append(env.reset().clone())
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(got), ref)
    rng = np.random.default_rng(4)
    for t in range(25):
        acts = torch.from_numpy(rng.integers(0, 5, size=(n, 32)).astype(np.uint8)).cuda()
        want = [x.clone() for x in one.step(acts)]
        torch.cuda.synchronize()
        outs = []
        for k, (env, st) in enumerate(parts):
            with torch.cuda.stream(st):
                outs.append([x.clone() for x in env.step(acts[k * (n // G):(k + 1) * (n // G)])])
        torch.cuda.synchronize()
        for j in range(4):
            assert torch.equal(torch.cat([o[j] for o in outs]), want[j]), (t, j)
    for env, _ in parts:
        env.check_errors()
    with pytest.raises(ValueError):
        BatchedPogema.groups(GridConfig(**kw), 10, groups=3)
