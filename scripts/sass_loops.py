"""Instruction mix of the loops of one kernel in libwsdl_b200.so (no GPU needed): python scripts/sass_loops.py <regex> [min_body]"""
import re, subprocess, sys
from collections import Counter
pat = sys.argv[1]
minb = int(sys.argv[2]) if len(sys.argv) > 2 else 150
out = subprocess.run(["cuobjdump", "-sass", "weaklysuperviseddl_b200/libwsdl_b200.so"], capture_output=True, text=True).stdout
cur, funcs = None, {}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if not re.search(pat, name):
        continue
    print(name, len(ins), "instructions")
    for a, t in ins:
        m = re.search(r"BRA(?:\.[A-Z.]+)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                body = [x for x in ins if tgt <= x[0] <= a]
                if len(body) < minb or len(body) > 1400:
                    continue
                c = Counter([w for w in x.split() if not w.startswith("@")][0].split(".")[0] for _, x in body)
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr", c.most_common(12))
