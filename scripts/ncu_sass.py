#!/usr/bin/env python
"""Per-opcode and per-region summary of the SASS page of an .ncu-rep (first kernel): instructions executed,
stall samples.  Usage: tools_ncu_sass.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for row in csv.reader(out.splitlines()):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': [], 'hdr': None}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = row; continue
    cur['rows'].append(row)
b = blocks[kidx]; h = b['hdr']
iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[iE]) for r in b['rows']); totS = sum(int(r[iN]) for r in b['rows'])
print(b['name'], 'instr', tot, 'samples', totS, 'sass lines', len(b['rows']))
op = collections.Counter(); ops = collections.Counter()
for r in b['rows']:
    toks = r[iS].split()
    o = toks[1] if toks[0].startswith('@') else toks[0]
    o = o.split('.')[0] if not o.startswith('MUFU') else o
    op[o] += int(r[iE]); ops[o] += int(r[iN])
for o, c in op.most_common(25): print(f'  {o:12s} {c:10d} {100*c/tot:5.1f}%  samples {100*ops[o]/max(totS,1):5.1f}%')
st = collections.Counter()
for r in b['rows']:
    for i in stall_cols: st[h[i]] += int(r[i])
print('stalls:', ', '.join(f'{k}={100*v/max(totS,1):.1f}%' for k, v in st.most_common(8)))
# regions: chunks of 256 sass lines
n = len(b['rows']); step = max(1, n // 24)
for s in range(0, n, step):
    rr = b['rows'][s:s + step]
    e = sum(int(r[iE]) for r in rr); sm = sum(int(r[iN]) for r in rr)
    print(f'  lines {s:5d}-{s+len(rr):5d}: instr {100*e/tot:5.1f}%  samples {100*sm/max(totS,1):5.1f}%')
