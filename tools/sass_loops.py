#!/usr/bin/env python
"""Loops of one kernel in a cuobjdump -sass dump: size, spill (LDL/STL), barrier and memory instruction counts.
  cuobjdump -sass -fun <mangled> lib.so > f.sass; python tools/sass_loops.py f.sass"""
import re, sys
ins = []
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
print(len(ins), "instructions")
for s, e in sorted(loops):
    body = [t for a, t in ins if s <= a <= e]
    if len(body) < 100:
        continue
    c = lambda k: sum(1 for t in body if re.match(r'(@!?U?P\d+\s+)?(' + k + r')\b', t))
    print(f"loop {s:#x}..{e:#x}: {len(body):5d} instr  LDL {c('LDL')} STL {c('STL')} BAR {c('BAR')} LDS {c('LDS[.A-Z0-9]*')} STS {c('STS[.A-Z0-9]*')} "
          f"ST {c('ST[.A-Z0-9]*')} LDG {c('LDG[.A-Z0-9]*')} SYNCS {c('SYNCS[.A-Z0-9]*')} MEMBAR {c('MEMBAR[.A-Z0-9]*')} UCGABAR {c('UCGABAR_[A-Z]*')}")
