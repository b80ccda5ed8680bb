#!/usr/bin/env python
"""Static instruction counts of the loops of one kernel from `cuobjdump -sass` (no GPU needed).
    cuobjdump -sass lib.so > all.sass;  python tools/sass_loops.py all.sass <substring of the mangled kernel name>
A loop = a backward branch; prints, per loop, its length in instructions and the mix (fp64, shuffles, shared / local /
global memory).  Used to track the warp-instructions per stage of the K3 sweeps between GPU runs."""
import re, sys
from collections import Counter

def classify(op):
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")): return "fp64"
    if op.startswith("MUFU"): return "mufu"
    if op.startswith("SHFL"): return "shfl"
    if op.startswith(("LDS", "STS")): return "smem"
    if op.startswith(("LDL", "STL")): return "local"
    if op.startswith(("LDG", "STG", "LD.", "ST.", "LDGSTS", "LDGDEPBAR", "DEPBAR")): return "global"
    if op.startswith(("BRA", "BSSY", "BSYNC", "WARPSYNC", "CALL", "RET", "EXIT")): return "ctrl"
    return "other"

def main(path, pat):
    txt = open(path).read().split("Function : ")
    for blk in txt[1:]:
        name = blk.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for ln in blk.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2)))
        print(f"== {name[:110]}  {len(ins)} instructions")
        addr = [a for a, _ in ins]
        loops = []
        for a, t in ins:
            m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d,?\s+)?`?\(?\.?L?_?x?_?\w*\)?\s*(0x[0-9a-f]+)?", t)
            mm = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
            if mm:
                tgt = int(mm.group(1), 16)
                if tgt < a:
                    loops.append((tgt, a))
        for (t0, t1) in sorted(loops):
            body = [t for a, t in ins if t0 <= a <= t1]
            c = Counter()
            for t in body:
                op = t.split()[0] if not t.startswith("@") else t.split()[1]
                c[classify(op)] += 1
            print(f"  loop {t0:#07x}-{t1:#07x}: {len(body):5d} instr  " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
