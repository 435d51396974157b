"""Turn the raw outputs of tools/make_profiles.sh (gpurun_out/) into the summaries under profiles/."""
import csv, json, shutil, subprocess, sys
from collections import defaultdict
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open('gpurun_out/launches.csv')))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = defaultdict(list)
for r in rows[hdr + 1:]:
    if len(r) > vi: d[r[ki]].append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for v in d.values())
lines = ["kernel,launches,mean_ns,max_ns,total_ns,share_of_all_gpu_time_in_the_run"]
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    lines.append('"%s",%d,%.0f,%.0f,%.0f,%.4f' % (k[:120], len(v), sum(v) / len(v), max(v), sum(v), sum(v) / tot))
open(f'profiles/{tag}_launch_summary.csv', 'w').write("\n".join(lines) + "\n")
many = [x for k, v in d.items() if 'pgm_step_kernel' in k for x in v if x > 100e3]
print('multi-step launches', len(many), 'mean us', sum(many) / len(many) / 1e3 if many else None)
out = subprocess.run("ncu -i gpurun_out/prof_step.ncu-rep --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]; u = rows[1]; v = rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_global_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum']
ki = h.index('Kernel Name') if 'Kernel Name' in h else None
txt = ["ncu --set full --clock-control none, ONE launch of %s advancing config 2 (4096 instances x 64 agents, r=5," % (v[ki] if ki is not None else 'pgm_step_kernel'),
       "priority/finish) by 16 steps (pgm_step_many) - tools/make_profiles.sh", ""]
vals = {}
for k in keep:
    if k in h:
        i = h.index(k); txt.append("%-70s %-16s %s" % (k, u[i], v[i])); vals[k] = (u[i], v[i])
for i, name in enumerate(h):
    if 'warp_issue_stalled' in name and name.endswith('_per_warp_active.pct') and v[i] and float(v[i]) > 1.0:
        txt.append("%-70s %-16s %s" % (name, u[i], v[i]))
tob = lambda u_, v_: float(v_) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u_]
rd = tob(*vals['dram__bytes_read.sum']); wr = tob(*vals['dram__bytes_write.sum'])
alg = 16 * 262144 * 387.453125
txt += ["", "algorithmic bytes of this launch: 16 steps x 262144 agents x 387.45 B = %.1f MB" % (alg / 1e6),
        "dram traffic (read+write): %.1f MB = %.3f x algorithmic" % ((rd + wr) / 1e6, (rd + wr) / alg)]
open(f'profiles/{tag}_ncu_step_kernel.txt', 'w').write("\n".join(txt) + "\n")
print("\n".join(txt))
json.dump({"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "steps_per_launch": 16,
           "source": f"profiles/{tag}_ncu_step_kernel.txt (ncu --set full, one 16-step launch of pgm_step_kernel)"},
          open('profiles/traffic.json', 'w'), indent=1)
for a, b in [('bench.json', f'{tag}_bench.json'), ('bench_reference.json', f'{tag}_bench_reference.json'),
             ('phase_timeline.txt', f'{tag}_phase_timeline.txt'), ('launches.csv', f'{tag}_launches.csv')]:
    shutil.copy('gpurun_out/' + a, 'profiles/' + b)
b = json.load(open(f'profiles/{tag}_bench.json'))
print({k: b.get(k) for k in ('value', 'ms_per_step', 'roofline', 'e2e', 'e2e_bits', 'closed_loop', 'cpu_baseline', 'cpu_baseline_c', 'gpu_launches', 'clocks')})
