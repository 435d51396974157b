"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line."""
import csv, sys
from collections import defaultdict
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
# first launch only: tables until the file path repeats
seen = set(); tables = []
for hi in hdr_idx:
    f = rows[hi - 2][1]
    if f in seen: break
    seen.add(f); tables.append(hi)
bounds = hdr_idx + [len(rows) + 2]
agg = defaultdict(lambda: [0, 0, '']); tot = 0; tots = 0
for ti, hi in enumerate(tables):
    h = rows[hi]; iInst = h.index('Instructions Executed'); iS = h.index('# Samples')
    f = rows[hi - 2][1].split('/')[-1]
    end = bounds[hdr_idx.index(hi) + 1] - 2
    cur = None
    for r in rows[hi + 1:end]:
        if len(r) <= iInst: continue
        if r[0] != '':
            cur = (f, r[0]); agg[cur][2] = r[1][:100]
        else:
            try: n = int(r[iInst]); s = int(r[iS] or 0)
            except ValueError: continue
            agg[cur][0] += n; agg[cur][1] += s; tot += n; tots += s
print('total warp-instructions', tot, 'samples', tots)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-16s %5s %9d %5.1f%%  samp %5.1f%%  %s' % (k[0][:16], k[1], v[0], 100 * v[0] / tot, 100 * v[1] / max(tots, 1), v[2]))
