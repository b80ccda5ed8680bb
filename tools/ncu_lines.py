#!/usr/bin/env python
"""Aggregate an ncu per-SASS-instruction export by source line.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c x.cubin > all_g.sass
    python tools/ncu_lines.py sass.csv all_g.sass <kernel-substring> [file-substring]

Prints, per source line of the chosen file, warp-instructions executed and stall samples."""
import csv, re, sys, collections

def main():
    sass_csv, disasm, kern = sys.argv[1:4]
    fsub = sys.argv[4] if len(sys.argv) > 4 else "rti_coop.cuh"
    # address -> (file, line) from nvdisasm -g
    amap = {}
    cur = None
    infn = False
    for ln in open(disasm, errors="ignore"):
        if ln.startswith(".text.") or ln.lstrip().startswith(".section"):
            infn = kern in ln
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # keep the innermost (first listed) location; "inlined at" lines follow
            if "inlined at" not in ln:
                cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]+)\*/', ln)
        if m and infn:
            amap[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {n: i for i, n in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_i = tot_s = 0
    base = None
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None:
            base = a
        loc = amap.get(a - base)
        ie = int(r[ci["Instructions Executed"]] or 0)
        sm = int(r[ci["# Samples"]] or 0)
        tot_i += ie; tot_s += sm
        key = loc if loc else ("?", 0)
        agg[key][0] += ie; agg[key][1] += sm
        op = r[ci["Source"]].split()[0] if r[ci["Source"]] else "?"
        if op.startswith("@"):
            op = r[ci["Source"]].split()[1]
        agg[key][2][op.split(".")[0]] += ie
    print(f"total warp-instructions {tot_i:.4g}, samples {tot_s}")
    items = sorted(agg.items(), key=lambda kv: -kv[1][0])
    for (f, l), (ie, sm, ops) in items[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
        if fsub not in f and f != "?":
            tag = f.split("/")[-1]
        else:
            tag = f.split("/")[-1]
        top = ",".join(f"{k}:{v/ie:.0%}" for k, v in ops.most_common(4)) if ie else ""
        print(f"{tag}:{l:5d}  inst {ie/tot_i:6.2%}  samples {sm/max(tot_s,1):6.2%}  {top}")

if __name__ == "__main__":
    main()
