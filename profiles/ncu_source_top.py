#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel of an ncu source page (csv or csv.gz).
usage: python profiles/ncu_source_top.py src.csv.gz <kernel index> [n]"""
import csv, gzip, sys
rep, kidx = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = gzip.open(rep, 'rt').read() if rep.endswith('.gz') else open(rep).read()
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}; blocks.append(cur)
    elif cur is not None and row:
        cur['rows'].append(row)
b = blocks[kidx]
hdr = b['rows'][0]; I = {h: i for i, h in enumerate(hdr)}
rows = b['rows'][1:]
tot_s = sum(int(r[I['# Samples']]) for r in rows)
tot_i = sum(int(r[I['Instructions Executed']]) for r in rows)
print(b['name'][:140]); print('total samples', tot_s, 'warp insts', tot_i, 'sass lines', len(rows))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h[6:]: sum(int(r[I[h]]) for r in rows) for h in stalls}
print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
cls = {}
for r in rows:
    src = r[I['Source']].strip()
    op = src.split()[1] if src.startswith('@') else src.split()[0]
    cls[op.split('.')[0]] = cls.get(op.split('.')[0], 0) + int(r[I['Instructions Executed']])
print('op classes (warp insts, %):', ' '.join('%s=%.1f' % (k, 100.0 * v / tot_i) for k, v in sorted(cls.items(), key=lambda x: -x[1])[:24]))
ex = I.get('L1 Wavefronts Shared Excessive')
for idx, r in sorted(enumerate(rows), key=lambda x: -int(x[1][I['# Samples']]))[:n]:
    top = sorted(((int(r[I[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print('%5d %10d %6d (%.1f%%)  %-72s %s %s' % (idx, int(r[I['Instructions Executed']]), int(r[I['# Samples']]),
          100.0 * int(r[I['# Samples']]) / tot_s, r[I['Source']].strip()[:72], ' '.join('%s=%d' % (nm, v) for v, nm in top if v),
          ('xs_wave=%s' % r[ex]) if ex is not None and r[ex] not in ('0', '') else ''))
