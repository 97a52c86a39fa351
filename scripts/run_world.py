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
out = {"config": "world: K=%d demes x H=%d x S=%d, 1e6 per deme, %d replicates" % (K, 4 ** U, S, R)}
t0 = time.time()
e.SimulatePopulation(200000, 10 ** 9, T_DIRECT, 200)
h = e._handle
out["direct_s"] = time.time() - t0
out["direct_kernel_ms"] = h.last_kernel_ms()
c = e.counters()
out["direct_events_mean"] = float(c["events"].mean())
out["P"] = int(h.P)
t0 = time.time()
e.SimulatePopulation_tau(MAXL, NS, -1, 200, leap_block=BLOCK if BLOCK > 0 else None)
out["tau_s"] = time.time() - t0
out["tau_kernel_ms"] = h.last_kernel_ms()  # blocks: first block's start to last block's end, archive passes included
out["leap_block"] = BLOCK
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
# bit 16: a MULTITYPE BIRTH record asked for more coalescences than the cell has lineage pairs; the reference reads
# past the end of a vector there (src/_BirthDeath.pyx:885-915), the kernel clamps and says so
out["genealogy_flags"] = flags
sm = h.summaries()
nodes, roots = sm[:, 13], sm[:, 16]
# n sampled leaves coalesced into `roots` trees have 2n - roots nodes (2n-1 when fully coalesced; a replicate whose replay
# clamped a coalescence count, genealogy flag 16, can be left with a few roots)
if not np.array_equal(nodes, 2 * c["sCounter"] - roots):
    bad = np.nonzero(nodes != 2 * c["sCounter"] - roots)[0]
    out["FAILED_tree_size"] = {"replicates": bad[:8].tolist(), "nodes": nodes[bad[:8]].tolist(), "roots": roots[bad[:8]].tolist(),
                               "samples": c["sCounter"][bad[:8]].tolist(), "n_bad": int(len(bad))}
    print(json.dumps(out))
    sys.exit(1)
out.update(tree_nodes_mean=float(nodes.mean()), fully_coalesced=int((roots == 1).sum()), tree_height_mean=float(sm[:, 14].mean()),
           mutation_rows_mean=float(sm[:, 17].mean()), migration_rows_mean=float(sm[:, 18].mean()))
parent, pop, tm = h.get_tree(0)
assert (parent == -1).sum() == roots[0] and np.all(parent[parent >= 0] > np.nonzero(parent >= 0)[0]), "parents are created after children"
assert np.all(tm[parent[parent >= 0]] <= tm[np.nonzero(parent >= 0)[0]] + 1e-12), "a parent is not later than its child"
out["checks"] += ["every tree has 2n - roots nodes", "parent index > child index and parent time <= child time (replicate 0)"]
print(json.dumps(out))
