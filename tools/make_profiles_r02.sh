#!/bin/bash
# Round 2 evidence run (GPU box, under gpurun): bench lines in the driver's invocation and with defaults, the ncu
# launch list of the same command, `ncu --set full` captures of the kernels behind every benchmarked configuration,
# phase timelines, all configurations in both launch forms.  Raw outputs land in gpurun_out/r02_*; tools/
# summarize_profiles_r02.py turns them into profiles/r02_*.
set -x
O=gpurun_out
mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_driver.json 2> $O/r02_bench_driver.err
python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_reference.json 2>> $O/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > $O/r02_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -k regex:pgm_fast"
# configs[1]: the 3rd 16-step launch (after 20 single warm-up steps + 3 warm-up launches), and one single-step launch
$NCU -s 24 -c 1 -o $O/r02_prof_c1_many python tools/quick_bench.py --steps 64 --many 16 > $O/r02_ncu_c1_many.log 2>&1
$NCU -s 30 -c 1 -o $O/r02_prof_c1_single python tools/quick_bench.py --steps 32 > $O/r02_ncu_c1_single.log 2>&1
# configs[2] maze soft/restart, configs[3] warehouse block_both, configs[4] r=3 share: one 16-step launch each
$NCU -s 24 -c 1 -o $O/r02_prof_c2_many python tools/quick_bench.py --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 64 --many 16 > $O/r02_ncu_c2_many.log 2>&1
$NCU -s 24 -c 1 -o $O/r02_prof_c3_many python tools/quick_bench.py --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 64 --many 16 > $O/r02_ncu_c3_many.log 2>&1
$NCU -s 24 -c 1 -o $O/r02_prof_r3_many python tools/quick_bench.py --n 2048 --r 3 --steps 64 --many 16 > $O/r02_ncu_r3_many.log 2>&1
$NCU -s 30 -c 1 -o $O/r02_prof_r3_single python tools/quick_bench.py --n 2048 --r 3 --steps 32 > $O/r02_ncu_r3_single.log 2>&1
python tools/phase_timeline.py > $O/r02_timeline_c1.txt 2>&1
python tools/phase_timeline.py --n 2048 --r 3 > $O/r02_timeline_r3.txt 2>&1
python tools/phase_timeline.py --n 512 --size 256 --agents 1024 --coll block_both --map warehouse > $O/r02_timeline_c3.txt 2>&1
python tools/bench_configs.py > $O/r02_configs.json 2> $O/r02_configs.err
# steps per launch: 8 / 16 / 32 / 64 on configs[1] and configs[2] (do the per-team timelines drift apart in long launches?)
( for k in 8 16 32 64; do python tools/quick_bench.py --steps 2048 --many $k; done
  for k in 8 16 32 64; do python tools/quick_bench.py --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 1024 --many $k; done ) > $O/r02_spl_sweep.jsonl 2>&1
# closed loop over G groups of instances on their own streams (BatchedPogema.groups): every BASELINE configuration
( for g in 1 2 4; do python tools/two_groups.py --groups $g; done
  for g in 1 2 4; do python tools/two_groups.py --groups $g --n 1024 --size 64 --agents 256 --coll soft --ot restart --map maze --steps 512; done
  for g in 1 2; do python tools/two_groups.py --groups $g --n 512 --size 256 --agents 1024 --coll block_both --map warehouse --steps 256; done
  for g in 1 2; do python tools/two_groups.py --groups $g --n 2048 --r 3 --steps 1024; done
  for g in 1 2 4; do python tools/two_groups.py --groups $g --n 16384 --r 3 --steps 256; done ) > $O/r02_groups.jsonl 2> $O/r02_groups.err
bash tools/gpu/r02_b_san.sh > $O/r02_sanitizer_fast.log 2>&1
bash tools/sanitize.sh > $O/r02_sanitizer_generic.log 2>&1
# summarise the captures HERE (gpurun brings back at most 64 MiB): text summaries stay, of the reports only configs[1]'s
python - <<'PY'
import subprocess, sys
D = lambda r: 2 * r + 1
bpa = lambda r, A, ph, pw: 3 * D(r) ** 2 + 21 + ((ph * pw + 7) // 8) / A
caps = [("c1_many", "configs[1] (4096 x 64 agents, r=5, priority/finish): ONE 16-step launch (pgm_step_many)", 16 * 4096 * 64, bpa(5, 64, 42, 42)),
        ("c1_single", "configs[1]: ONE single-step launch (pgm_step, the closed-loop form)", 4096 * 64, bpa(5, 64, 42, 42)),
        ("c2_many", "configs[2] (1024 x 256 agents, 64x64 maze, soft/restart): ONE 16-step launch", 16 * 1024 * 256, bpa(5, 256, 74, 74)),
        ("c3_many", "configs[3] (512 x 1024 agents, 256x256 warehouse, block_both): ONE 16-step launch", 16 * 512 * 1024, bpa(5, 1024, 266, 266)),
        ("r3_many", "configs[4] r=3 share (2048 x 64 agents): ONE 16-step launch", 16 * 2048 * 64, bpa(3, 64, 38, 38)),
        ("r3_single", "configs[4] r=3 share: ONE single-step launch", 2048 * 64, bpa(3, 64, 38, 38))]
for tag, title, units, b in caps:
    out = subprocess.run([sys.executable, "tools/ncu_summary.py", f"gpurun_out/r02_prof_{tag}.ncu-rep", title, str(units), str(b)], capture_output=True, text=True)
    open(f"gpurun_out/r02_ncu_{tag}.txt", "w").write(out.stdout + (("\nSTDERR\n" + out.stderr[-2000:]) if out.returncode else ""))
PY
rm -f $O/r02_prof_c2_many.ncu-rep $O/r02_prof_c3_many.ncu-rep $O/r02_prof_r3_many.ncu-rep $O/r02_prof_r3_single.ncu-rep $O/r02_prof_c1_single.ncu-rep
ls -la $O | grep r02_; du -sh $O
