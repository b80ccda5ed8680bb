#!/usr/bin/env python
"""Attribute an ncu per-SASS export of k_ipm_group to the phases of rti_group.cuh (by address order:
an instruction belongs to the phase of the last phase-body source line seen before it)."""
import csv, re, sys, collections, bisect
sass_csv, disasm, kern, src = sys.argv[1:5]
marks = []   # (line, name)
for i, ln in enumerate(open(src), 1):
    m = re.search(r'// ---- (\w+):', ln)
    if m: marks.append((i, m.group(1)))
    m = re.search(r'NMPC_HD static \w+\*? (\w+)\(', ln)
    if m: marks.append((i, "fn:" + m.group(1)))
    m = re.search(r'template <bool DELTA>', ln)
marks.sort()
first_body = next(l for l, n in marks if n == "fn:sweep_B")
def phase_of(line):
    j = bisect.bisect_right([l for l, _ in marks], line) - 1
    return marks[j][1]
amap = {}; cur = None; infn = False
for ln in open(disasm, errors="ignore"):
    if ln.startswith(".text."): infn = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if m.group(1).endswith(src.split("/")[-1]) and int(m.group(2)) >= first_body: cur = int(m.group(2))
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/', ln)
    if m and infn: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ci = {n: i for i, n in enumerate(rows[hi])}
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
base = None; ti = ts = 0
stall_cols = [n for n in rows[hi] if n.startswith("stall_") and "Not Issued" not in n]
for r in rows[hi + 1:]:
    if len(r) < len(rows[hi]): continue
    a = int(r[ci["Address"]], 16)
    if base is None: base = a
    ln = amap.get(a - base)
    ph = phase_of(ln) if ln else "?"
    ie = int(r[ci["Instructions Executed"]] or 0); sm = int(r[ci["# Samples"]] or 0)
    ti += ie; ts += sm
    e = agg[ph]; e[0] += ie; e[1] += sm
    toks = r[ci["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    e[2][op.split(".")[0]] += ie
    for sc in stall_cols:
        v = r[ci[sc]]
        if v: e[3][sc[6:]] += int(v)
print(f"total warp-instructions {ti:.4g}, samples {ts}")
for ph, (ie, sm, ops, st) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ph:18s} inst {ie/ti:6.2%} samples {sm/ts:6.2%} | " + " ".join(f"{k}:{v/max(ie,1):.0%}" for k, v in ops.most_common(6)) + " | " +
          " ".join(f"{k}:{v/max(sm,1):.0%}" for k, v in st.most_common(5)))
