#!/usr/bin/env python
"""Registers / spills per kernel from the -Xptxas -v build logs (levelsetpy_b200/build/*.log)."""
import glob, re, subprocess, sys
for log in sorted(glob.glob('levelsetpy_b200/build/*.cu.log')):
    txt = open(log).read()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            try:
                cur = subprocess.run(['c++filt', cur], capture_output=True, text=True).stdout.strip()
            except Exception:
                pass
            cur = re.sub(r'\(anonymous namespace\)::', '', cur)
            cur = re.sub(r'\(CUtensorMap_st.*|\(KGrid.*', '', cur)
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m:
            spill = (m.group(2), m.group(3))
        m = re.search(r'Used (\d+) registers', line)
        if m and cur:
            print('%-4s regs  spill st/ld %5s/%-5s  %s' % (m.group(1), spill[0], spill[1], cur[:150]))
