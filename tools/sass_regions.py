#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: runs of SASS instructions with the same execution
count (= the same basic-block nest), with their share of the kernel's warp instructions and lanes per instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iA = hdr.index("Avg. Threads Executed"); iT = hdr.index("Thread Instructions Executed")
tot = sum(int(r[iI]) for r in data); totT = sum(int(r[iT]) for r in data); totS = sum(int(r[iS]) for r in data)
print(f"# {rows[0][1][:90]}\n# warp instructions {tot}, thread instructions {totT}, lanes/instruction {totT / tot:.2f}, samples {totS}")
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
prev = None
def flush():
    if acc / tot * 100 >= minshare:
        print(f"[{start:4d}-{k - 1:4d}] n={k - start:3d} exec/inst={prev / 1e6:6.2f}M lanes={accT / max(acc, 1):5.1f} share={acc / tot * 100:5.1f}% samples={accS / totS * 100:5.1f}%  {first[:60]}")
for k, r in enumerate(data):
    c = int(r[iI])
    if prev is None or abs(c - prev) > 0.02 * max(prev, 1):
        if prev is not None: flush()
        start = k; acc = 0; accS = 0; accT = 0; prev = c; first = r[1].strip()
    acc += c; accS += int(r[iS]); accT += int(r[iT])
k = len(data); flush()
