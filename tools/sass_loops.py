#!/usr/bin/env python
"""All loops (backward-branch spans) of one kernel in a cuobjdump -sass dump, with instruction counts and an opcode
histogram of each span of at least `minlen` instructions.
usage: python tools/sass_loops.py dump.sass '<substring of demangled kernel name>' [minlen]"""
import collections
import re
import subprocess
import sys

dump, key = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
lines = open(dump).read().splitlines()
funcs = [(i, l) for i, l in enumerate(lines) if 'Function :' in l]
names = subprocess.run(['c++filt'], input='\n'.join(l.split('Function :')[1].strip() for _, l in funcs),
                       capture_output=True, text=True).stdout.splitlines()
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
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(n[:160])
print('total instrs', len(ins))
spans = []
for a, t in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a:
            spans.append((tgt, a))
for tgt, a in sorted(set(spans)):
    sel = [t for x, t in ins if tgt <= x <= a]
    if len(sel) < minlen:
        continue
    h = collections.Counter()
    for t in sel:
        op = t.split()[1] if t.startswith('@') else t.split()[0]
        h[op.split('.')[0]] += 1
    fp64 = sum(v for k2, v in h.items() if k2 in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMNMX'))
    print('loop 0x%x..0x%x: %d instrs, %d fp64 | %s' % (tgt, a, len(sel), fp64, ' '.join('%s=%d' % kv for kv in h.most_common(14))))
