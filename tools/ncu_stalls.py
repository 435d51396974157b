"""Per source line: sample share and dominant stall reasons from an ncu source-page CSV."""
import csv, sys
from collections import defaultdict
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
seen = set(); tables = []
for hi in hdr_idx:
    f = rows[hi - 2][1]
    if f in seen: break
    seen.add(f); tables.append(hi)
bounds = hdr_idx + [len(rows) + 2]
agg = defaultdict(lambda: [0, defaultdict(int), '']); tots = 0
for hi in tables:
    h = rows[hi]; iS = h.index('# Samples')
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
    f = rows[hi - 2][1].split('/')[-1]
    end = bounds[hdr_idx.index(hi) + 1] - 2
    cur = None
    for r in rows[hi + 1:end]:
        if len(r) <= iS: continue
        if r[0] != '':
            cur = (f, r[0]); agg[cur][2] = r[1][:70]
        else:
            try: s = int(r[iS] or 0)
            except ValueError: continue
            agg[cur][0] += s; tots += s
            for i, n in stall_cols:
                try: agg[cur][1][n] += int(r[i] or 0)
                except ValueError: pass
print('samples', tots)
total_by = defaultdict(int)
for k, v in agg.items():
    for n, c in v[1].items(): total_by[n] += c
print('overall:', ', '.join('%s %.1f%%' % (n.replace('stall_', ''), 100 * c / max(tots, 1)) for n, c in sorted(total_by.items(), key=lambda kv: -kv[1])[:8]))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ', '.join('%s %d' % (n.replace('stall_', ''), c) for n, c in sorted(v[1].items(), key=lambda kv: -kv[1])[:3] if c)
    print('%-14s %5s %5.1f%%  %-70s | %s' % (k[0][:14], k[1], 100 * v[0] / max(tots, 1), v[2], st))
