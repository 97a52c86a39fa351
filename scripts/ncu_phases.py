#!/usr/bin/env python
"""Per-phase digest of an `ncu --page source --csv` SASS listing of tau_warp_kernel: every SASS instruction is
attributed through the inlining chain of `nvdisasm -gi` to (outermost line inside the kernel body -> phase) and to the
helper function it was inlined from.
usage: ncu_phases.py <source_page.csv> <nvdisasm_gi.txt> <mangled function> [regions.json]"""
import csv, re, sys, collections, json

src_csv, dis, fn = sys.argv[1:4]
# phases of tau_warp_kernel by line range of tau_warp.cuh (outermost frame); override with a json [[name, lo, hi], ...]
PH = [("prologue/load", 642, 763), ("leap top + barrier + wipe", 764, 787), ("drifts+tau", 788, 793),
      ("draw setup", 794, 819), ("3a primary", 820, 890), ("3b drain", 891, 896), ("feasibility", 897, 937),
      ("apply+lists", 938, 961), ("lockdown", 962, 966), ("restart/commit/tail", 967, 1045)]
if len(sys.argv) > 4:
    PH = [tuple(x) for x in json.load(open(sys.argv[4]))]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
sass = rows[2:]
base = int(sass[0][ci["Address"]], 16)
chain_of = {}
chain = []
infn = False
prev_was_file = False
for l in open(dis):
    if l.startswith(".text."):
        infn = l.strip() == ".text.%s:" % fn
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not prev_was_file:
            chain = []
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        prev_was_file = True
        continue
    prev_was_file = False
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m:
        chain_of[int(m.group(1), 16)] = list(chain)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def phase(ch):
    if not ch:
        return "?"
    f, ln = ch[-1]
    if f != "tau_warp.cuh":
        return "?%s" % f
    for n, a, b in PH:
        if a <= ln <= b:
            return n
    return "other:%d" % ln


def helper(ch):
    # first frame below the outermost one = the helper called from the kernel body
    if len(ch) < 2:
        return "(kernel body)"
    f, ln = ch[-2]
    return "%s:%d" % (f, ln)


agg = collections.defaultdict(collections.Counter)
agg2 = collections.defaultdict(collections.Counter)
for r in sass:
    ch = chain_of.get(int(r[ci["Address"]], 16) - base, [])
    for key, A in ((phase(ch), agg), ((phase(ch), helper(ch)), agg2)):
        a = A[key]
        a["inst"] += int(r[ci["Instructions Executed"]] or 0)
        a["thr"] += int(r[ci["Thread Instructions Executed"]] or 0)
        a["smp"] += int(r[ci["# Samples"]] or 0)
        a["sass"] += 1
        for s in stalls:
            a[s] += int(r[ci[s]] or 0)
ti = sum(a["inst"] for a in agg.values())
ts = sum(a["smp"] for a in agg.values())
print("total warp-inst %d samples %d sass %d" % (ti, ts, len(sass)))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"]):
    top = sorted(((a[s], s) for s in stalls), reverse=True)[:4]
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  sass %5d  %-28s %s" % (
        100 * a["inst"] / ti, 100 * a["smp"] / ts, a["thr"] / max(a["inst"], 1), a["sass"], name,
        " ".join("%s=%.0f%%" % (s.replace("stall_", ""), 100 * v / max(a["smp"], 1)) for v, s in top)))
print("---- by (phase, helper frame) top 40")
for name, a in sorted(agg2.items(), key=lambda kv: -kv[1]["smp"])[:40]:
    top = sorted(((a[s], s) for s in stalls), reverse=True)[:3]
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  sass %5d  %-24s %-22s %s" % (
        100 * a["inst"] / ti, 100 * a["smp"] / ts, a["thr"] / max(a["inst"], 1), a["sass"], name[0], name[1],
        " ".join("%s=%.0f%%" % (s.replace("stall_", ""), 100 * v / max(a["smp"], 1)) for v, s in top)))
