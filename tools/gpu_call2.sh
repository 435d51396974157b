#!/bin/bash
mkdir -p gpurun_out
exec 2>&1
python tools/e2e_probe.py --mode plain
for t in 8 12 16; do python tools/e2e_probe.py --threads $t; done
for c in 2 4 16 32; do PGM_STREAM_CHUNKS=$c python tools/e2e_probe.py --threads 16; done
PGM_HOST_SPIN_US=0 python tools/e2e_probe.py --threads 16
PGM_HOST_SPIN_US=2000 python tools/e2e_probe.py --threads 16
python tools/e2e_probe.py --threads 16 --pageable
python tools/e2e_probe.py --mode plain --pageable
python tools/e2e_probe.py --threads 16 --fmt f32
python tools/e2e_probe.py --mode plain --fmt f32
python tools/e2e_probe.py --threads 16 --n 16384
echo "== obs batch, other configs (closed loop / many)"
python - <<'PY'
import os, subprocess, sys
for name, args in (("maze", "--n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze"),
                   ("wh", "--n 512 --size 256 --agents 1024 --coll block_both --map warehouse")):
    for b in ("0", "512", "256", "128", "64"):
        env = dict(os.environ)
        if b != "0": env["PGM_OBS_BATCH"] = b
        for mode in ("--graph 16", "--many 16"):
            out = subprocess.run([sys.executable, "tools/quick_bench.py", *args.split(), *mode.split(), "--steps", "512"], env=env, capture_output=True, text=True)
            import json
            try:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                print(name, "batch", b, mode, d["plan"], d["ms_per_step"], round(d["frac_6541"], 3), flush=True)
            except Exception as ex:
                print(name, b, mode, "failed", out.stderr[-300:])
PY
