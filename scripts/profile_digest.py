#!/usr/bin/env python
"""Digest of one `ncu --set full` capture for profiles/: headline metrics from the raw page, stall mix and the top
source lines from the source page (joined with `nvdisasm -g` of the same cubin).
usage: profile_digest.py <raw.csv> <source.csv> <nvdisasm.txt> <mangled kernel name> <replicates> <leaps> > profiles/<name>.txt
       (also prints a JSON line `TRAFFIC {...}` on stderr for profiles/tau_kernel_traffic.json)"""
import csv, json, re, sys, collections
raw, src_csv, dis, fn, R, L = sys.argv[1:7]
R, L = int(R), int(L)
rows = list(csv.reader(open(raw))); hdr, units, r = rows[0], rows[1], rows[2]
d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
def num(k):
    v = d.get(k, "")
    try: return float(v.replace(",", ""))
    except ValueError: return float("nan")
def scale(k):  # to base units
    un = u.get(k, ""); v = num(k)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(un, 1)
rd, wr, t = scale("dram__bytes_read.sum"), scale("dram__bytes_write.sum"), scale("gpu__time_duration.sum")
print("kernel: %s" % d.get("Kernel Name"))
print("grid %s block %s  registers/thread %s  dynamic smem/block %s %s" % (d.get("Grid Size"), d.get("Block Size"),
      d.get("launch__registers_per_thread"), d.get("launch__shared_mem_per_block_dynamic"), u.get("launch__shared_mem_per_block_dynamic")))
print("duration under ncu (cold, serialised): %.3f ms   [the bench number is CUDA-event timed, not this]" % (t * 1e3))
print("dram read %.3f GB  write %.3f GB  -> traffic %.3f GB per launch (%d replicates x %d leaps)" % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, R, L))
inst = num("smsp__inst_executed.sum")
print("warp instructions %.4g (%.0f per leap)   thread-instructions per instruction %.2f / 32" % (inst, inst / (R * L), num("smsp__thread_inst_executed_per_inst_executed.ratio")))
for k in ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
          "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
          "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct"]:
    if k in d: print("  %-88s %s %s" % (k, d[k], u.get(k, "")))
print("warps stalled per issue (smsp__average_warps_issue_stalled_*_per_issue_active):")
ks = [k for k in hdr if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
for v, k in sorted(((num(k), k) for k in ks), reverse=True)[:9]:
    print("  %6.3f  %s" % (v, k.split("issue_stalled_")[1].split("_per_issue")[0]))
sys.stderr.write("TRAFFIC " + json.dumps({"kernel": "tau_warp_kernel", "replicates": R, "leaps": L, "dram_bytes_per_launch": rd + wr,
                 "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes_per_launch": None}) + "\n")
# ---- source page
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
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
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(collections.Counter)
for r in sass:
    a = agg[lines.get(int(r[ci["Address"]], 16) - base) or ("?", 0)]
    a["inst"] += int(r[ci["Instructions Executed"]] or 0); a["smp"] += int(r[ci["# Samples"]] or 0)
    a["thr"] += int(r[ci["Thread Instructions Executed"]] or 0)
    for s in stall_cols: a[s] += int(r[ci[s]] or 0)
ti = sum(a["inst"] for a in agg.values()); ts = sum(a["smp"] for a in agg.values())
tot = collections.Counter()
for a in agg.values():
    for s in stall_cols: tot[s] += a[s]
print("SASS instructions in the kernel: %d; warp-state samples %d" % (len(sass), ts))
print("sample mix: " + ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / ts) for s, v in tot.most_common(9)))
print("top source lines by samples (share of samples, share of instructions, mean active lanes, top stall reasons):")
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:30]:
    if f not in srcs:
        try: srcs[f] = open("/root/repo/vgsim_b200/csrc/" + f).read().split("\n")
        except OSError: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:64] if 0 < ln <= len(srcs[f]) else ""
    top = sorted(((a[s], s) for s in stall_cols), reverse=True)[:2]
    print("  %5.1f%% %5.1f%% %5.1f  %-34s %s:%d  %s" % (100.0 * a["smp"] / ts, 100.0 * a["inst"] / ti, a["thr"] / max(a["inst"], 1),
          " ".join("%s:%d%%" % (s[6:], 100 * v / max(a["smp"], 1)) for v, s in top), f, ln, text))
