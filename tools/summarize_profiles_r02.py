"""gpurun_out/r02_* (tools/make_profiles_r02.sh on a B200) -> the committed summaries under profiles/."""
import csv, json, os, re, shutil, subprocess, sys
from collections import defaultdict

G, P = "gpurun_out", "profiles"
D = lambda r: 2 * r + 1
bpa = lambda r, A, ph, pw: 3 * D(r) ** 2 + 21 + ((ph * pw + 7) // 8) / A

# launch list of `bench.py --steps 20 --warmup 5`
rows = list(csv.reader(open(f"{G}/r02_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi:
        try: d[r[ki]].append(float(r[vi].replace(',', '')))
        except ValueError: pass
tot = sum(sum(v) for v in d.values())
lines = ["kernel,launches,mean_ns,max_ns,total_ns,share_of_all_gpu_time_in_the_run"]
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append('"%s",%d,%.0f,%.0f,%.0f,%.4f' % (k[:140], len(v), sum(v) / len(v), max(v), sum(v), sum(v) / tot))
open(f"{P}/r02_launch_summary.csv", "w").write("\n".join(lines) + "\n")
shutil.copy(f"{G}/r02_launches.csv", f"{P}/r02_launches.csv")

import glob
for f in sorted(glob.glob(f"{G}/r02_ncu_*.txt")):
    out = open(f).read()
    shutil.copy(f, f"{P}/" + os.path.basename(f))
    if f.endswith("r02_ncu_c1_many.txt"):
        m = re.search(r"dram traffic \(read\+write\): ([0-9.]+) MB", out)
        if m:
            json.dump({"dram_bytes_per_launch": float(m.group(1)) * 1e6, "steps_per_launch": 16,
                       "source": "profiles/r02_ncu_c1_many.txt (ncu --set full, one 16-step launch of pgm_fast_step_kernel<64,1,0,5>)"},
                      open(f"{P}/traffic.json", "w"), indent=1)
for a, b in [("r02_bench.json", "r02_bench.json"), ("r02_bench_driver.json", "r02_bench_driver.json"), ("r02_bench_reference.json", "r02_bench_reference.json"),
             ("r02_timeline_c1.txt", "r02_phase_timeline_c1.txt"), ("r02_timeline_r3.txt", "r02_phase_timeline_r3.txt"),
             ("r02_timeline_c3.txt", "r02_phase_timeline_c3.txt"), ("r02_configs.json", "r02_configs.json"),
             ("r02_groups.jsonl", "r02_groups.jsonl"), ("r02_spl_sweep.jsonl", "r02_spl_sweep.jsonl")]:
    if os.path.exists(f"{G}/{a}"): shutil.copy(f"{G}/{a}", f"{P}/{b}")
san = []
for f in ("r02_sanitizer_fast.log", "r02_sanitizer_generic.log"):
    if os.path.exists(f"{G}/{f}"):
        san.append(f"== {f}")
        san += [ln.rstrip()[:200] for ln in open(f"{G}/{f}") if re.search(r"exit=|ERROR SUMMARY|RACECHECK SUMMARY", ln)]
open(f"{P}/r02_sanitizer.txt", "w").write("\n".join(san) + "\n")
for f in ("r02_bench_driver.json", "r02_bench.json"):
    try:
        b = json.loads(open(f"{P}/{f}").read().strip().splitlines()[-1])
        print(f, {k: b.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, b["roofline"]["frac"], b["roofline"]["steady_state"], b["closed_loop"], b["e2e"]["value"], b["sharding_check"], b["host_dram"])
        for c in b.get("configs") or []:
            print("   ", c["config"][:70], {k: (round(v["us_per_step"], 2), round(v["roofline_frac"], 3)) for k, v in c.items() if isinstance(v, dict) and "us_per_step" in v})
    except Exception as exc:
        print(f, "unparsed", exc)
