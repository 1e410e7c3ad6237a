#!/usr/bin/env python
"""SASS evidence for profiles/: for one kernel of libsangnom_cuda.so, the mnemonic counts that prove the sm_100a data
path (bulk async copies, mbarriers, async remote stores, TMA tensor loads/stores, packed video instructions) and the
code around every such instruction.
  python tools/sass_excerpt.py <substring of the mangled kernel name> [context lines] > profiles/sass_<name>.txt"""
import re, subprocess, sys, collections
LIB = "avisynth-sangnom2_b200/libsangnom_cuda.so"
pat, ctx = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 6
names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.findall(r"Function : (\S+)", names)
match = [f for f in funcs if pat in f]
if not match:
    sys.exit(f"no kernel matching {pat}; have: {funcs}")
fn = match[0]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout
demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
ins = []
for l in out.splitlines():
    m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(1), m.group(2).strip()))
EVID = r"UBLKCP|UTMALDG|UTMASTG|UTMACMDFLUSH|STAS|SYNCS|UCGABAR|MEMBAR|VABSDIFF4|VIMNMX3?(\.U16x2)?|LDGDEPBAR|BAR\.SYNC|LDS\.128|STS\.128"
cnt = collections.Counter()
for _, t in ins:
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    cnt[op.split(".")[0] if not re.match(EVID, op) else op] += 1
print(f"# {demangled}\n# {len(ins)} instructions in {LIB} (cuobjdump -sass); evidence mnemonics:")
for op, n in sorted(cnt.items(), key=lambda kv: -kv[1]):
    if re.match(EVID, op):
        print(f"#   {op:28s} {n}")
print("#\n# code around the first occurrences of the async-copy / barrier / tensor-map instructions:")
shown, last = set(), -10
for i, (a, t) in enumerate(ins):
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    key = op.split(".")[0]
    if re.match(r"UBLKCP|UTMALDG|UTMASTG|STAS|SYNCS|UTMACMDFLUSH|UCGABAR", op) and (key, t.split()[0]) not in shown and len([k for k in shown if k[0] == key]) < 2:
        shown.add((key, t.split()[0]))
        lo, hi = max(0, i - ctx), min(len(ins), i + ctx + 1)
        if lo <= last:
            lo = last + 1
        if lo < hi:
            print("    ...")
            for a2, t2 in ins[lo:hi]:
                print(f"    /*{a2}*/ {t2}")
            last = hi - 1
