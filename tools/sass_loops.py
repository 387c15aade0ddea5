#!/usr/bin/env python3
"""List the loops (backward branches) of a kernel's SASS with their opcode mix.
usage: sass_loops.py <object-or-so> <mangled-kernel-substring> [min_instr]"""
import collections
import re
import subprocess
import sys

obj, kern = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = next(b for b in blocks if b.split("\n")[0].find(kern) >= 0)
ins = []
for ln in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)([.A-Z0-9_]*)\s*(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(5)))
fp = ("DFMA", "DMUL", "DADD")
print("total instructions", len(ins))
for a, op, suf, args in ins:
    if op == "BRA":
        m = re.search(r"0x([0-9a-f]+)", args)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            reg = [i for i in ins if t <= i[0] <= a]
            c = collections.Counter(i[1] for i in reg)
            if len(reg) >= minlen:
                print(hex(t), hex(a), "instr", len(reg), "fp64", sum(c[k] for k in fp), c.most_common(14))
