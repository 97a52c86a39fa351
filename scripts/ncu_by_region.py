#!/usr/bin/env python
"""Like ncu_by_line.py but aggregates into named line ranges of tau_kernel.cu + other files; also reports the
dominant stall reason per region.  usage: ncu_by_region.py <source_page.csv> <nvdisasm.txt> <function>"""
import csv, re, sys, collections
src_csv, dis, fn = sys.argv[1:4]
REG = [("channel helpers", 119, 212), ("block_min", 213, 225), ("list/zero lists", 226, 252), ("drift pass1 Q,F", 253, 272),
       ("drift pass2 dI", 273, 298), ("drift dS", 299, 328), ("load_replicate", 332, 379), ("lockdown", 380, 406),
       ("book", 407, 431), ("wipe_leap", 432, 449), ("geom/mig_total", 450, 487), ("primary_draw", 488, 506),
       ("split_total", 507, 569), ("process_entry", 570, 609), ("kernel prologue/loop", 610, 666), ("3a primary", 667, 729),
       ("3b drain", 730, 787), ("feasibility", 788, 813), ("apply", 814, 839), ("restart/commit", 840, 900)]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
sass = rows[2:]; base = int(sass[0][ci["Address"]], 16)
lines = {}; cur = None; infn = False
for l in open(dis):
    if l.startswith(".text."):
        infn = l.strip() == ".text.%s:" % fn; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m: lines[int(m.group(1), 16)] = cur
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r in sass:
    key = lines.get(int(r[ci["Address"]], 16) - base) or ("?", 0)
    name = key[0]
    if key[0] == "tau_kernel.cu":
        name = next((n for n, a, b in REG if a <= key[1] <= b), "tau_kernel.cu:other")
    a = agg[name]
    a["inst"] += int(r[ci["Instructions Executed"]] or 0)
    a["smp"] += int(r[ci["# Samples"]] or 0)
    for sname in stalls:
        a[sname] += int(r[ci[sname]] or 0)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["smp"] for a in agg.values())
print("total warp-inst %d samples %d" % (ti, ts))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
    top = sorted(((a[s], s) for s in stalls), reverse=True)[:3]
    print("%5.1f%% inst %5.1f%% smp  %-24s %s" % (100 * a["inst"] / ti, 100 * a["smp"] / ts, name,
          " ".join("%s=%.0f%%" % (s.replace("stall_", ""), 100 * v / max(a["smp"], 1)) for v, s in top)))
