#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing with `nvdisasm -g` line info of the same cubin and
aggregate executed instructions / stall samples per CUDA source line.
usage: ncu_by_line.py <source_page.csv> <nvdisasm.txt> <mangled function name> [top N]"""
import csv, re, sys, collections
src_csv, dis, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = rows[2:]
base = int(sass[0][ci["Address"]], 16)
lines = {}
cur = None
infn = False
for l in open(dis):
    if l.startswith(".text."):
        infn = l.strip() == ".text.%s:" % fn
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m:
        lines[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r in sass:
    off = int(r[ci["Address"]], 16) - base
    key = lines.get(off)
    key = (key[0], key[1]) if key else ("?", 0)
    ie = int(r[ci["Instructions Executed"]] or 0)
    sm = int(r[ci["# Samples"]] or 0)
    agg[key][0] += ie
    agg[key][1] += sm
    agg[key][2] += 1
    tot[0] += ie
    tot[1] += sm
print("total warp-instructions %d, samples %d" % tuple(tot))
srcs = {}
for (f, ln), (ie, sm, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/vgsim_b200/csrc/" + f).read().split("\n")
        except OSError:
            srcs[f] = []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%% inst %5.1f%% smp %4d sass  %s:%d  %s" % (100.0 * ie / tot[0], 100.0 * sm / max(tot[1], 1), n, f, ln, text))
