#!/usr/bin/env python
"""Per source line: stall samples split into barrier / non-barrier (from an ncu --page source --csv SASS listing
joined with nvdisasm -g line info).  usage: ncu_stalls.py <source_page.csv> <nvdisasm.txt> <function> [top]"""
import csv, re, sys, collections
src_csv, dis, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
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
bar = collections.Counter(); non = collections.Counter(); inst = collections.Counter()
for r in sass:
    key = lines.get(int(r[ci["Address"]], 16) - base) or ("?", 0)
    b = int(r[ci["stall_barrier"]] or 0); sm = int(r[ci["# Samples"]] or 0)
    bar[key] += b; non[key] += sm - b; inst[key] += int(r[ci["Instructions Executed"]] or 0)
tb, tn = sum(bar.values()), sum(non.values())
src = open("/root/repo/vgsim_b200/csrc/tau_kernel.cu").read().split("\n")
def text(k): return src[k[1] - 1].strip()[:80] if k[0] == "tau_kernel.cu" and 0 < k[1] <= len(src) else ""
print("samples: barrier %d (%.0f%%), other %d" % (tb, 100.0 * tb / (tb + tn), tn))
print("-- where warps wait at barriers")
for k, v in bar.most_common(10): print("  %5.1f%%  %s:%d  %s" % (100.0 * v / tb, k[0], k[1], text(k)))
print("-- non-barrier stall samples by line")
for k, v in non.most_common(top): print("  %5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * v / tn, 100.0 * inst[k] / sum(inst.values()), k[0], k[1], text(k)))
