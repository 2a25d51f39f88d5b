#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions executed, stall samples) of the first kernel of an .ncu-rep captured
with --import-source on.  Usage: ncu_lines.py file.ncu-rep [top N]"""
import csv, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(out.splitlines()))
per = collections.OrderedDict(); cur = None; hdr = None; kernels = 0
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name':
        kernels += 1
        if kernels > 1: break
        continue
    if r[0] == 'Line No' or (len(r) > 2 and r[2] == 'Address'):
        hdr = r; continue
    if r[0] not in ('', 'File Name') and r[0].isdigit():
        cur = (int(r[0]), r[1]); per.setdefault(cur, [0, 0]); continue
    if cur and len(r) > 8 and r[2].startswith('0x'):
        try:
            per[cur][0] += int(r[7]); per[cur][1] += int(r[6])
        except ValueError:
            pass
totI = sum(v[0] for v in per.values()); totS = sum(v[1] for v in per.values())
print('total instr', totI, 'samples', totS)
for (ln, src), (i, s) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{ln:5d} instr {100*i/max(totI,1):5.1f}% samples {100*s/max(totS,1):5.1f}%  {src.strip()[:110]}')
