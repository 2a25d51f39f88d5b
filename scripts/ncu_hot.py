#!/usr/bin/env python
"""Hot SASS instructions of the first kernel in an .ncu-rep: samples, executed count, dominant stall reason.
Usage: ncu_hot.py file.ncu-rep [top N]"""
import csv, subprocess, sys
from collections import Counter
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []; k = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        if k > 1: break
        continue
    if hdr is None and r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
st = [i for i, n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[iN]) for r in data)
cnt = Counter(int(r[iE]) for r in data if 'MUFU.EX2' in r[iS])
hot = cnt.most_common(1)[0][0]
loop = [(i, r) for i, r in enumerate(data) if abs(int(r[iE]) - hot) <= hot * 0.02]
ls = sum(int(r[iN]) for _, r in loop)
print(f'loop instrs {len(loop)} (exec {hot} each), samples in loop {ls} of {tot} = {100*ls/tot:.1f}%')
agg = Counter()
for _, r in loop:
    for i in st: agg[hdr[i]] += int(r[i])
print('loop stalls:', ', '.join(f'{k}={100*v/max(ls,1):.1f}%' for k, v in agg.most_common(8)))
ops = Counter(); opn = Counter()
for _, r in loop:
    t = r[iS].split(); o = t[1] if t[0].startswith('@') else t[0]
    ops[o] += int(r[iN]); opn[o] += 1
for o, v in ops.most_common(14): print(f'   {o:22s} n={opn[o]:4d} samples {100*v/max(ls,1):5.1f}%')
print('hottest loop instructions:')
for i, r in sorted(loop, key=lambda x: -int(x[1][iN]))[:top]:
    reasons = sorted(((int(r[j]), hdr[j]) for j in st), reverse=True)[:2]
    print(f'  #{i:5d} {r[iS].strip():60s} samples {r[iN]:>4s}  ' + ' '.join(f'{n}:{v}' for v, n in reasons if v))
