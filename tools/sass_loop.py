#!/usr/bin/env python
"""Opcode histogram of the main loop (largest backward branch span) of one kernel in a cuobjdump -sass dump.
usage: python tools/sass_loop.py dump.sass '<substring of demangled kernel name>'"""
import re, subprocess, sys, collections
dump, key = sys.argv[1], sys.argv[2]
lines = open(dump).read().splitlines()
funcs = [(i, l) for i, l in enumerate(lines) if 'Function :' in l]
names = subprocess.run(['c++filt'], input='\n'.join(l.split('Function :')[1].strip() for _, l in funcs), capture_output=True, text=True).stdout.splitlines()
for k, ((i, l), n) in enumerate(zip(funcs, names)):
    if key in n:
        end = funcs[k + 1][0] if k + 1 < len(funcs) else len(lines)
        body = lines[i:end]
        break
else:
    sys.exit('kernel not found')
ins = []
for l in body:
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, t in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
print(n[:140]); print('total instrs', len(ins), 'loop span', best, 'instrs in span', sum(1 for a, _ in ins if best[0] <= a <= best[1]))
h = collections.Counter()
for a, t in ins:
    if best[0] <= a <= best[1]:
        op = t.split()[1] if t.startswith('@') else t.split()[0]
        h[op.split('.')[0]] += 1
print(' '.join('%s=%d' % kv for kv in h.most_common()))
