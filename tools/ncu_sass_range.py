#!/usr/bin/env python
"""SASS of one kernel in program order with executed counts and warp-stall samples by reason, restricted to the
instructions whose inline chain passes through a source-line range (inlined callees included).

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all lib.so; nvdisasm -gi -c x.cubin > all_gi.sass
    python tools/ncu_sass_range.py sass.csv all_gi.sass <kernel-substring> <file-substring> <line0> <line1> [-q]
"""
import csv, re, sys, collections

def main():
    sass_csv, disasm, kern, fsub = sys.argv[1:5]
    l0, l1 = int(sys.argv[5]), int(sys.argv[6])
    quiet = "-q" in sys.argv
    amap, chain, infn = {}, [], False
    fresh = True
    for ln in open(disasm, errors="ignore"):
        if ln.startswith(".text.") or ln.lstrip().startswith(".section"):
            infn = kern in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain = []; fresh = False
            chain.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/', ln)
        if m:
            if infn:
                amap[int(m.group(1), 16)] = list(chain)
            fresh = True
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]; ci = {n: i for i, n in enumerate(hdr)}
    reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    base = None; tot = 0; sel = 0; nex = 0
    agg = collections.Counter(); byline = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        a = int(r[ci["Address"]], 16)
        if base is None:
            base = a
        ch = amap.get(a - base) or []
        smp = int(r[ci["# Samples"]] or 0); tot += smp
        hit = [c for c in ch if fsub in c[0] and l0 <= c[1] <= l1]
        if not hit:
            continue
        rs = {n[6:]: int(r[ci[n]] or 0) for n in reasons if int(r[ci[n]] or 0)}
        ex = int(r[ci["Instructions Executed"]] or 0)
        sel += smp; nex += ex
        for k, v in rs.items():
            agg[k] += v
        b = byline[hit[0][1]]; b[0] += ex; b[1] += smp
        for k, v in rs.items():
            b[2][k] += v
        if not quiet:
            print(f"{a - base:6x} L{hit[0][1]:4d}<{ch[0][1]:4d} ex={ex:6d} s={smp:5d} {r[ci['Source']].strip()[:64]:64s} {rs if rs else ''}")
    print("by line of the range (executed warp-instructions, samples, reasons):")
    for ln in sorted(byline):
        b = byline[ln]
        print(f"  L{ln:4d} ex={b[0]:7d} s={b[1]:6d} ({100.0 * b[1] / max(sel, 1):5.1f} %) {dict(b[2].most_common(4))}")
    print(f"range: executed {nex}, samples {sel} of {tot} ({100.0 * sel / max(tot, 1):.1f} %), by reason: {dict(agg.most_common())}")

main()
