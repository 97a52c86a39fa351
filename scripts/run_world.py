#!/usr/bin/env python
"""BASELINE.json configs[3] ("world-scale tau-leap: 100 demes with full migration matrix, 1e8 individuals, 1e5-sample
genealogy") end to end on one GPU, with size-independent checks.  Prints one JSON line.

    python scripts/run_world.py [replicates] [max_leaps] [samples] [t_direct] [leap_block]

Model: SURVEY §8(d) config 4 = data/Table 3/Table 3.py with K=100 (tests/scenarios.py "w"): direct method until the
epidemic has taken off (t = 40; the reference restarts any run with <= 100 log rows), then tau-leaping until
`samples` cases are sampled, genealogy over the mixed log, epidemic curves.  The dense log of a replicate is
leaps x 1.97 MB (32 replicates x 1,200 leaps = 76 GB), so the stated size (256 replicates, 1e5 samples = ~2,000 leaps
each) runs in leap blocks: finished blocks are kept as the sparse archive (vgsim_simulate_tau_blocks), leap_block 0 =
one dense call."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402
from vgsim_b200._engine import BirthDeathModel as Eng  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MAXL = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
NS = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
T_DIRECT = float(sys.argv[4]) if len(sys.argv) > 4 else 40.0
BLOCK = int(sys.argv[5]) if len(sys.argv) > 5 else 101
(U, K, S), setup = SCENARIOS["w"]
e = Eng(U, K, S, 4242, False, False, int(1e6), 0.0, replicates=R)
setup(e)
def note(msg):
    sys.stderr.write("[%7.1f s] %s\n" % (time.time() - T00, msg))
    sys.stderr.flush()


T00 = time.time()
out = {"config": "world: K=%d demes x H=%d x S=%d, 1e6 per deme, %d replicates" % (K, 4 ** U, S, R)}
t0 = time.time()
e.SimulatePopulation(200000, 10 ** 9, T_DIRECT, 200)
h = e._handle
out["direct_s"] = time.time() - t0
note("direct phase done")
out["direct_kernel_ms"] = h.last_kernel_ms()
c = e.counters()
out["direct_events_mean"] = float(c["events"].mean())
out["P"] = int(h.P)
t0 = time.time()
e.SimulatePopulation_tau(MAXL, NS, -1, 200, leap_block=BLOCK if BLOCK > 0 else None)
out["tau_s"] = time.time() - t0
out["tau_kernel_ms"] = h.last_kernel_ms()  # blocks: first block's start to last block's end, archive passes included
out["leap_block"] = BLOCK
note("tau phase done: %.0f ms" % out["tau_kernel_ms"])
st = h.archive_stats()
out["archive"] = dict(st, bytes_total=8 * st["entries_total"], dense_bytes_replaced=st["leaps_archived"] * 4 * int(h.P))
c = e.counters()
ev = sum(int(c[k].sum()) for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"))
leaps = int(c["leaps"].sum())
out.update(leaps_mean=float(c["leaps"].mean()), leaps_max=int(c["leaps"].max()), samples_min=int(c["sCounter"].min()),
           samples_mean=float(c["sCounter"].mean()), time_mean=float(c["time"].mean()),
           events_total=ev, tau_events_per_s=ev / (out["tau_kernel_ms"] * 1e-3), tau_leaps_per_s=leaps / (out["tau_kernel_ms"] * 1e-3),
           tau_log_GBps=leaps * (4.0 * h.P + 16) / (out["tau_kernel_ms"] * 1e-3) / 1e9,
           infectious_mean=float(c["globalInfectious"].mean()), device_errors=int(h.synchronize(strict=False)))
# size-independent checks on the full-size log: curves replay the log to the final state, deme sizes conserved
Sx_f, I_f = h.get_state()
t0 = time.time()
cv = e.epidemic_curves(16, want=("infectious", "susceptible", "sampled"))
out["curves_s"] = time.time() - t0
note("curves done")
out["curves_kernel_ms"] = h.last_kernel_ms()
assert np.array_equal(cv["infectious"][:, -1], I_f) and np.array_equal(cv["susceptible"][:, -1], Sx_f)
tot = cv["infectious"].sum(axis=3) + cv["susceptible"].sum(axis=3)
assert np.all(tot == 1000000)
assert np.array_equal(cv["sampled"][:, -1].sum(axis=(1, 2)), c["sCounter"])
out["checks"] = ["curves replay every log to its final state", "every deme keeps 1e6 individuals at all 17 grid points",
                 "sampled tallies equal sCounter"]
t0 = time.time()
h.genealogy(99, sync=False)
flags = int(h.synchronize(strict=False))
out["genealogy_s"] = time.time() - t0
note("genealogy done")
# bit 16: a MULTITYPE BIRTH record asked for more coalescences than the cell has lineage pairs; the reference reads
# past the end of a vector there (src/_BirthDeath.pyx:885-915), the kernel clamps and says so
out["genealogy_flags"] = flags
sm = h.summaries()
nodes, roots = sm[:, 13], sm[:, 16]
# every replicate's tree: n leaves (the sampled cases), every internal node has exactly two children, a parent is created
# after its children (larger index) and is not later in time, so `roots` trees over n leaves have 2n - roots nodes
bad_trees = []
for r in range(R):
    parent, pop, tm = h.get_tree(r)
    n_used = int(nodes[r])
    parent, tm = parent[:n_used], tm[:n_used]
    idx = np.nonzero(parent >= 0)[0]
    kids = np.bincount(parent[idx], minlength=n_used)
    ok = (np.all(parent[idx] > idx) and np.all(tm[parent[idx]] <= tm[idx] + 1e-12) and set(np.unique(kids)) <= {0, 2}
          and int((kids == 0).sum()) == int(c["sCounter"][r]) and int((parent < 0).sum()) == int(roots[r])
          and n_used == 2 * int(c["sCounter"][r]) - int(roots[r]))
    if not ok:
        bad_trees.append({"replicate": r, "nodes": n_used, "samples": int(c["sCounter"][r]), "roots": int(roots[r]),
                          "root_count_host": int((parent < 0).sum()), "children_hist": np.bincount(kids).tolist(),
                          "parent_not_after_child": int((parent[idx] <= idx).sum())})
if bad_trees:
    out["FAILED_trees"] = bad_trees[:8]
    out["n_bad_trees"] = len(bad_trees)
    print(json.dumps(out))
    sys.exit(1)
note("tree checks done")
out.update(tree_nodes_mean=float(nodes.mean()), fully_coalesced=int((roots == 1).sum()), tree_height_mean=float(sm[:, 14].mean()),
           mutation_rows_mean=float(sm[:, 17].mean()), migration_rows_mean=float(sm[:, 18].mean()))
out["checks"] += ["every replicate: leaves = samples, internal nodes have two children, parent index > child index, parent time <= child time, 2n - roots nodes"]
print(json.dumps(out))
