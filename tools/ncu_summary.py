"""One `ncu --set full` capture (.ncu-rep) -> a text summary for profiles/: headline raw metrics, stall reasons per
warp-active, DRAM traffic against the algorithmic bytes, and warp instructions / stall samples per kernel phase.

    python tools/ncu_summary.py gpurun_out/r02_prof_c1_many.ncu-rep "title" <agents per launch x steps> <bytes per agent-step>
"""
import csv, subprocess, sys
from collections import defaultdict

rep, title, units, bpa = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'smsp__sass_inst_executed_op_global_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum']
ki = h.index('Kernel Name') if 'Kernel Name' in h else None
out = [title, "kernel: %s" % (v[ki] if ki is not None else '?'), "ncu --set full --clock-control none (cold caches, serialised: compare shares, not absolutes)", ""]
vals = {}
for k in keep:
    if k in h:
        i = h.index(k); out.append("%-70s %-16s %s" % (k, u[i], v[i])); vals[k] = (u[i], v[i])
for i, name in enumerate(h):
    if 'warp_issue_stalled' in name and name.endswith('_per_warp_active.pct') and v[i]:
        try:
            if float(v[i]) > 1.0: out.append("%-70s %-16s %s" % (name, u[i], v[i]))
        except ValueError: pass
tob = lambda u_, v_: float(v_.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u_]
rd = tob(*vals['dram__bytes_read.sum']); wr = tob(*vals['dram__bytes_write.sum'])
dur = float(vals['gpu__time_duration.sum'][1].replace(',', '')) * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(vals['gpu__time_duration.sum'][0], 1e-6)
alg = units * bpa
inst = float(vals['smsp__inst_executed.sum'][1].replace(',', ''))
out += ["", "algorithmic bytes of this launch: %.0f agent-steps x %.2f B = %.1f MB" % (units, bpa, alg / 1e6),
        "dram traffic (read+write): %.1f MB = %.3f x algorithmic (stores still dirty in L2 at kernel end are not counted)" % ((rd + wr) / 1e6, (rd + wr) / alg),
        "algorithmic bytes / duration: %.0f GB/s = %.3f of the measured 6542 GB/s (isolated, cold launch)" % (alg / dur / 1e9, alg / dur / 6542.1e9),
        "warp instructions per agent-step: %.2f" % (inst / units)]
# per-phase instruction / stall-sample shares from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
seen = set(); tables = []
for hi in hdr_idx:
    f = rows[hi - 2][1]
    if f in seen: break
    seen.add(f); tables.append(hi)
bounds = hdr_idx + [len(rows) + 2]
import os, re
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pogema_b200", "csrc", "pgm_fast.cuh")
ANCHORS = [("struct FastRows", "observation bits of an agent in registers (row windows, target bit)"),
           ("__device__ __forceinline__ void fast_store_stream", "stream store (lane shift, carry by shuffle, STS)"),
           ("__device__ __forceinline__ void fast_expand_u8", "bit -> byte expansion + 16-byte streaming stores"),
           ("__global__ void", "prologue: pointers, bulk copies, fills, state + first action loads"),
           ("// ---- actions of step k", "actions (validate, issue next step's loads), per-step setup"),
           ("// ---- phase 1", "publish on the cell grid (+ soft pass A)"),
           ("// ---- phase 2", "move resolution + pointer jumping"),
           ("// ---- block_both", "block_both: claim planes"),
           ("// ---- phase 3", "apply, counts, bookkeeping, outputs, agent bitmap"),
           ("// ---- phase 4", "observation driver loop (+ stores of the bits formats)")]
src_lines = open(SRC).read().split("\n")
marks = []
for pat, name in ANCHORS:
    for i, ln in enumerate(src_lines):
        if pat in ln:
            marks.append((i + 1, name)); break
marks.sort()
def phase_of(line):
    name = "file header"
    for ln, nm in marks:
        if line >= ln: name = nm
    return name
agg = defaultdict(lambda: [0, 0]); stalls = defaultdict(int)
for hi in tables:
    hh = rows[hi]; iI = hh.index('Instructions Executed'); iS = hh.index('# Samples')
    stall_cols = [(i, n) for i, n in enumerate(hh) if n.startswith('stall_') and 'Not Issued' not in n]
    f = rows[hi - 2][1].split('/')[-1]
    end = bounds[hdr_idx.index(hi) + 1] - 2
    cur = None
    for r in rows[hi + 1:end]:
        if len(r) <= iI: continue
        if r[0] != '':
            cur = int(r[0])
            continue
        try: n = int(r[iI]); s = int(r[iS] or 0)
        except ValueError: continue
        if 'pgm_fast' not in f: continue   # (instructions of helpers inlined from other files are listed under their pgm_fast.cuh call sites too)
        key = phase_of(cur)
        agg[key][0] += n; agg[key][1] += s
        for i, nme in stall_cols:
            try: stalls[nme] += int(r[i] or 0)
            except ValueError: pass
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
out += ["", "warp instructions / stall samples by part of pgm_fast.cuh (ncu source page; %d of the %d executed warp instructions carry a pgm_fast.cuh line):" % (tot, inst)]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    out.append("  %-78s %7.2f instr / agent-step  %5.1f%% instr  %5.1f%% samples" % (k[:78], a[0] / units, 100 * a[0] / tot, 100 * a[1] / ts))
allst = sum(stalls.values()) or 1
out.append("stall reasons (all samples): " + ", ".join("%s %.1f%%" % (n.replace('stall_', ''), 100 * c / allst) for n, c in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
print("\n".join(out))
