#!/usr/bin/env python
"""Join an `ncu --page source --csv --print-source sass` dump with `nvdisasm -g` line info of the same build and sum the
executed warp instructions / stall samples per source line.
usage: sass_by_line.py <sass.csv> <kernel mangled-name substring> [top N] [smp]   (run after `make`, same libb200pt.so;
"smp" sorts by stall samples instead of executed instructions)"""
import csv, os, re, subprocess, sys, tempfile, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
csvf, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
key = 2 if len(sys.argv) > 4 and sys.argv[4] == "smp" else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "rtx-pathtracer_b200", "libb200pt.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = {}
for cub in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
    inside, cur = False, ("?", 0)
    for l in dis:
        if l.startswith(".text."):
            inside = kern in l
            cur = ("?", 0)
            continue
        if not inside: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', l)
        if m: lines[int(m.group(1), 16)] = cur
    if lines: break
rows = list(csv.reader(open(csvf)))
hdr, data = rows[1], rows[2:]
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
a0 = int(data[0][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    k = lines.get(int(r[0], 16) - a0, ("?", 0))
    a = agg[k]; a[0] += int(r[iI]); a[1] += int(r[iT]); a[2] += int(r[iS])
tot = sum(a[0] for a in agg.values()); totS = sum(a[2] for a in agg.values())
print("# %s: %d SASS instructions mapped, warp instructions %d" % (kern, len(lines), tot))
src = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    if f not in src:
        p = os.path.join(ROOT, "rtx-pathtracer_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print("%5.1f%% inst %5.1f%% smp lanes %4.1f  %s:%d  %s" % (100.0 * a[0] / tot, 100.0 * a[2] / max(totS, 1), a[1] / max(a[0], 1), f, ln, text))
