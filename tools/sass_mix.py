#!/usr/bin/env python
"""Opcode mix of one kernel launch weighted by executed warp-instructions, from `ncu --page source --csv --print-source sass`.
  ncu -i X.ncu-rep --page source --csv --print-source sass --launch-count 1 > /tmp/src.csv; python tools/sass_mix.py /tmp/src.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
mix = collections.Counter(); samp = collections.Counter(); stalls = collections.Counter()
tot = 0
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break          # first launch only
    if len(r) <= iN or r[iN] == 'Instructions Executed': continue
    op = r[iS].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", op).split()[0].rstrip(";")
    base = op.split(".")[0]
    key = base if base not in ("IMAD", "LDS", "STS", "VIMNMX", "LOP3", "SHF", "LEA") else ".".join(op.split(".")[:2])
    n = int(r[iN] or 0); mix[key] += n; tot += n; samp[key] += int(r[iSamp] or 0)
    for h, i in stall_cols: stalls[h] += int(r[i] or 0)
print(f"total warp-instructions {tot}")
for k, n in mix.most_common(40): print(f"{k:18s} {n:12d} {100*n/tot:5.1f}%  samples {samp[k]}")
ts = sum(stalls.values())
print("stall samples:", ", ".join(f"{h[6:]} {100*v/ts:.1f}%" for h, v in stalls.most_common(10)))
