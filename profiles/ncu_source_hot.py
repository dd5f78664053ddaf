#!/usr/bin/env python
"""Hot-loop view of an ncu source page: per SASS instruction executed count and stall samples.
usage: python profiles/ncu_source_hot.py rep.ncu-rep <kernel index> [min_exec]"""
import csv, subprocess, sys
rep, kidx = sys.argv[1], int(sys.argv[2])
if rep.endswith('.gz'):       # a source page exported on the GPU box: ncu -i rep --page source --csv | gzip
    import gzip
    out = gzip.open(rep, 'rt').read()
else:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}; blocks.append(cur)
    elif cur is not None and row:
        cur['rows'].append(row)
b = blocks[kidx]
hdr = b['rows'][0]
I = {h: i for i, h in enumerate(hdr)}
rows = b['rows'][1:]
tot_s = sum(int(r[I['# Samples']]) for r in rows)
tot_i = sum(int(r[I['Instructions Executed']]) for r in rows)
print(b['name'][:120]); print('total samples', tot_s, 'warp insts', tot_i)
mx = max(int(r[I['Instructions Executed']]) for r in rows)
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {}
cls = {}
for r in rows:
    ex = int(r[I['Instructions Executed']])
    if ex < thr * mx: continue
    s = int(r[I['# Samples']])
    op = r[I['Source']].split()[0] if not r[I['Source']].strip().startswith('@') else r[I['Source']].split()[1]
    cls[op.split('.')[0]] = cls.get(op.split('.')[0], 0) + ex
    top = sorted(((int(r[I[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print('%9d %6d  %-70s %s' % (ex, s, r[I['Source']].strip()[:70], ' '.join('%s=%d' % (n, v) for v, n in top if v)))
    for h in stalls: agg[h] = agg.get(h, 0) + int(r[I[h]])
print('--- op classes in hot region (warp insts):')
for k, v in sorted(cls.items(), key=lambda x: -x[1]): print('  %-10s %d (%.1f / hot iteration)' % (k, v, v / mx))
print('--- stall totals in hot region:', {k[6:]: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
