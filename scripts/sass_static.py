#!/usr/bin/env python
"""Static SASS instruction count of one kernel by (outermost source line inside the kernel body, helper frame below it).
usage: sass_static.py <cubin> <mangled function> [top N]"""
import re, sys, subprocess, collections
cubin, fn = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
infn = False; chain = []; prevf = False
by_outer = collections.Counter(); by_helper = collections.Counter(); total = 0
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = l.strip() == ".text.%s:" % fn; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not prevf: chain = []
        chain.append((m.group(1).split("/")[-1], int(m.group(2)))); prevf = True; continue
    prevf = False
    if re.match(r"\s+/\*[0-9a-f]+\*/", l):
        total += 1
        outer = chain[-1] if chain else ("?", 0)
        helper = chain[-2] if len(chain) > 1 else ("(body)", 0)
        by_outer["%s:%d" % outer] += 1
        by_helper["%s:%d <- %s:%d" % (helper + outer)] += 1
print("total", total)
for k, v in by_outer.most_common(top): print("%6d  %s" % (v, k))
print("---- helper frames")
for k, v in by_helper.most_common(top): print("%6d  %s" % (v, k))
