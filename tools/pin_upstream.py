#!/usr/bin/env python
"""tools/pin_upstream.py - pin the restated oracle (oracle/pogema_oracle.py) against upstream pogema.

Upstream pogema is a pure-Python package that was NOT present in the build container (SURVEY.md section 0:
/root/reference holds README.md:1-5 only, no wheel, no network).  Every "bit-exact" in this repo therefore means
"against the restated oracle" and parity with upstream is unpinned.  This script is the ready-to-run kit that
closes that gap the moment upstream is importable anywhere:

    python tools/pin_upstream.py                 # probe; if upstream imports: differential run + report
    python tools/pin_upstream.py --regen-golden  # ... and regenerate tests/golden/ from UPSTREAM

What it does when `import pogema` works (from baseline/_ref, site-packages or --path):
  1. runs all 9 collision_system x on_target modes x `--seeds` seeds (default 100) x three shapes, plus the
     hand-derived scenarios of tests/scenarios.py, on upstream and on the oracle with identical action streams;
  2. compares, per step: every observation array, rewards, terminated, truncated, infos[i]['is_active'],
     infos[0]['metrics'], agent / target positions, and the obstacle map after reset;
  3. prints the FIRST divergence per SURVEY.md section 9 item (1 soft rule, 2 placing, 3 lifelong targets,
     4 RNG instances, 5 border, 6 hidden agents, 7 square target, 8 block_both, 9 time limit, 10 GridConfig);
  4. with --regen-golden writes tests/golden/golden_v1.npz from upstream's outputs and records upstream's version.
The probe result (either way) is written to profiles/r02_pin_upstream.txt.
"""
import argparse
import importlib
import itertools
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

COLLS = ("priority", "block_both", "soft")
ONTS = ("finish", "nothing", "restart")
SHAPES = [
    dict(size=8, density=0.3, num_agents=4, obs_radius=5, max_episode_steps=64),     # BASELINE.json configs[0]
    dict(size=10, density=0.1, num_agents=30, obs_radius=2, max_episode_steps=20),   # crowded: conflicts, chains
    dict(size=32, density=0.3, num_agents=64, obs_radius=5, max_episode_steps=64),   # configs[1] instance shape
]
# which SURVEY.md section 9 item a divergence most likely belongs to
ITEMS = {1: "soft collision rule / _revert_action", 2: "placing() branch order / order list", 3: "lifelong targets (component order, choice)",
         4: "fresh vs shared default_rng(seed) instances", 5: "add_artificial_border geometry", 6: "hidden agents / terminated persistence",
         7: "get_square_target clamp / sign", 8: "block_both marking", 9: "MultiTimeLimit truncation", 10: "GridConfig defaults / bounds"}


def probe(extra_path=None):
    """Where could upstream come from?  Returns (module or None, list of report lines)."""
    lines = []
    cands = [p for p in (extra_path, os.path.join(ROOT, "baseline", "_ref")) if p]
    for p in cands:
        ok = os.path.isdir(p)
        lines.append(f"path {p}: {'present' if ok else 'absent'}")
        if ok and p not in sys.path:
            sys.path.insert(0, p)
    ref = "/root/reference"
    if os.path.isdir(ref):
        files = [os.path.join(d, f) for d, _, fs in os.walk(ref) for f in fs]
        lines.append(f"{ref}: {len(files)} file(s): {', '.join(os.path.relpath(f, ref) for f in files[:8])}")
        if os.path.isdir(os.path.join(ref, "pogema")) and ref not in sys.path:
            sys.path.insert(0, ref)
    else:
        lines.append(f"{ref}: absent")
    wh = "/opt/wheelhouse"
    if os.path.isdir(wh):
        hits = [f for f in os.listdir(wh) if any(k in f.lower() for k in ("pogema", "gymnasium", "pettingzoo"))]
        lines.append(f"{wh}: {len(os.listdir(wh))} wheels, pogema/gymnasium/pettingzoo: {hits or 'none'}")
    mod = None
    for name in ("pogema", "gymnasium", "pettingzoo"):
        try:
            m = importlib.import_module(name)
            lines.append(f"import {name}: OK, version {getattr(m, '__version__', '?')} at {os.path.dirname(getattr(m, '__file__', '?'))}")
            if name == "pogema":
                mod = m
        except Exception as exc:
            lines.append(f"import {name}: {type(exc).__name__}: {exc}")
    try:
        out = subprocess.run([sys.executable, "-m", "pip", "download", "pogema", "--no-deps", "-d", "/tmp/_pin_dl", "-q"],
                             capture_output=True, text=True, timeout=60)
        lines.append("pip download pogema: " + ("OK" if out.returncode == 0 else (out.stderr.strip().splitlines() or ["failed"])[-1][:160]))
    except Exception as exc:
        lines.append(f"pip download pogema: {type(exc).__name__}")
    return mod, lines


def snapshot(env):
    """Positions / targets / active flags / obstacles of an upstream-shaped env (works for both implementations)."""
    g = env.unwrapped.grid if hasattr(env, "unwrapped") else env.grid
    n = len(g.positions_xy)
    return dict(pos=np.array(g.positions_xy, dtype=np.int64), tgt=np.array(g.finishes_xy, dtype=np.int64),
                active=np.array([bool(g.is_active[i]) for i in range(n)]), obstacles=np.array(g.obstacles).astype(np.int64))


def classify(what, gc, t):
    if what == "obstacles":
        return 5 if t == 0 else 4
    if t == 0 and what in ("pos", "tgt"):
        return 2
    if what == "tgt":
        return 3
    if what == "truncated":
        return 9
    if what in ("pos", "active", "rewards", "terminated", "is_active"):
        return {"soft": 1, "block_both": 8}.get(gc["collision_system"], 6)
    if what.startswith("obs"):
        return 7 if what.endswith("ch2") else 6
    return 10


def compare_run(up, orc, gc, seed, actions, report):
    """One episode on both implementations; appends (item, description) of the first divergence, if any."""
    try:
        eu = up.pogema_v0(up.GridConfig(seed=seed, **gc))
        ou, iu = eu.reset()
    except OverflowError:
        eu = None
    try:
        eo = orc.pogema_v0(orc.GridConfig(seed=seed, **gc))
        oo, io = eo.reset()
    except OverflowError:
        eo = None
    if (eu is None) != (eo is None):
        report.append((2, f"OverflowError only on {'upstream' if eu is None else 'oracle'}: cfg={gc} seed={seed}"))
        return False
    if eu is None:
        return True

    def diff(t, ou, oo, extra=()):
        su, so = snapshot(eu), snapshot(eo)
        for k in ("obstacles", "pos", "tgt", "active"):
            if su[k].shape != so[k].shape or not np.array_equal(su[k], so[k]):
                return k
        for i, (a, b) in enumerate(zip(ou, oo)):
            a, b = np.asarray(a), np.asarray(b)
            if a.shape != b.shape or a.dtype != b.dtype:
                return f"obs shape/dtype {a.shape}/{a.dtype} vs {b.shape}/{b.dtype}"
            for ch in range(3):
                if not np.array_equal(a[ch], b[ch]):
                    return f"obs agent {i} ch{ch}"
        for name, a, b in extra:
            if list(a) != list(b):
                return name
        return None

    bad = diff(0, ou, oo)
    t = 0
    while bad is None and t < actions.shape[0]:
        act = [int(x) for x in actions[t]]
        ou, ru, tu, cu, iu = eu.step(act)
        oo, ro, to, co, io = eo.step(act)
        t += 1
        bad = diff(t, ou, oo, extra=(("rewards", ru, ro), ("terminated", tu, to), ("truncated", cu, co),
                                     ("is_active", [i.get("is_active") for i in iu], [i.get("is_active") for i in io]),
                                     ("metrics", *common_metrics(iu[0].get("metrics"), io[0].get("metrics"), report))))
        if all(tu) or all(cu):
            break
    if bad is not None:
        report.append((classify(bad.replace("obs agent", "obs").split(" ")[0] + ("_" + bad.split(" ")[-1] if bad.startswith("obs agent") else ""), gc, t),
                       f"{bad} differs at t={t}: cfg={gc} seed={seed}"))
        return False
    return True


_metric_key_notes = set()


def common_metrics(mu, mo, report):
    """Metrics dicts reduced to the keys both sides report (upstream versions differ in which metric wrappers
    _make_pogema stacks - SoC / makespan arrived late, 'runtime' is wall clock); a key only one side has is noted
    once as a version note (item 0), not as a divergence."""
    if mu is None or mo is None:
        return [json.dumps(mu)], [json.dumps(mo)]
    only = (set(mu) ^ set(mo)) - {"runtime"}
    for k in sorted(only - _metric_key_notes):
        _metric_key_notes.add(k)
        report.append((0, f"metric '{k}' reported only by {'upstream' if k in mu else 'the oracle'} (wrapper stacks differ between versions)"))
    keys = sorted((set(mu) & set(mo)) - {"runtime"})
    return [json.dumps({k: mu[k] for k in keys}, sort_keys=True)], [json.dumps({k: mo[k] for k in keys}, sort_keys=True)]


def differential(up, n_seeds):
    from oracle import pogema_oracle as orc
    from tests.scenarios import SCENARIOS, clean_map
    report, runs, same = [], 0, 0
    for coll, ot, shape in itertools.product(COLLS, ONTS, SHAPES):
        gc = dict(shape, collision_system=coll, on_target=ot)
        for seed in range(n_seeds):
            acts = np.random.default_rng(10_000 + seed).integers(0, 5, size=(shape["max_episode_steps"] + 6, shape["num_agents"]))
            runs += 1
            same += compare_run(up, orc, gc, seed, acts, report)
    for sc, coll in itertools.product(SCENARIOS, COLLS):
        gc = dict(map=clean_map(sc["map"]), obs_radius=2, collision_system=coll, on_target="nothing")
        runs += 1
        same += compare_run(up, orc, gc, 0, np.array([sc["actions"]]), report)
    return runs, same, report


def regen_golden(up):
    """tests/golden/golden_v1.npz from UPSTREAM (same cases / seeds / action streams as tests/golden/make_golden.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg
    from tests import helpers
    saved = helpers.orc
    helpers.orc = up                      # run_oracle() builds its env through `orc.pogema_v0(orc.GridConfig(...))`
    try:
        mg.main()
    finally:
        helpers.orc = saved
    with open(os.path.join(ROOT, "tests", "golden", "UPSTREAM_VERSION"), "w") as f:
        f.write(f"pogema {getattr(up, '__version__', '?')} numpy {np.__version__}\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path", default=None, help="directory that contains the upstream `pogema` package")
    ap.add_argument("--seeds", type=int, default=100)
    ap.add_argument("--regen-golden", action="store_true")
    ap.add_argument("--self-test", action="store_true",
                    help="run the differential machinery with the oracle standing in for upstream (checks the kit itself)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_pin_upstream.txt"))
    args = ap.parse_args()
    up, lines = probe(args.path)
    if args.self_test:
        from oracle import pogema_oracle as up_stand_in
        runs, same, report = differential(up_stand_in, args.seeds)
        print(f"self-test: {same}/{runs} episodes identical (oracle vs itself), {len(report)} divergences")
        return 0 if same == runs else 1
    out = ["tools/pin_upstream.py - probe for upstream pogema (numpy %s, python %s)" % (np.__version__, sys.version.split()[0])] + lines
    if up is None:
        out.append("RESULT: upstream pogema is NOT importable here -> parity with upstream stays UNPINNED; "
                   "the oracle is pinned to numpy known answers, hand-derived scenarios and its own golden trajectories only.")
    else:
        runs, same, report = differential(up, args.seeds)
        out.append(f"RESULT: upstream pogema {getattr(up, '__version__', '?')} imported; {same}/{runs} episodes identical to the oracle")
        first = {}
        for item, desc in report:
            if item == 0:
                out.append(f"  note: {desc}")
            else:
                first.setdefault(item, desc)
        for item in sorted(ITEMS):
            out.append(f"  SURVEY 9.{item} ({ITEMS[item]}): " + (f"FIRST DIVERGENCE: {first[item]}" if item in first else "no divergence seen"))
        if args.regen_golden:
            regen_golden(up)
            out.append("tests/golden/golden_v1.npz regenerated from upstream (tests/golden/UPSTREAM_VERSION written)")
    text = "\n".join(out) + "\n"
    sys.stdout.write(text)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
