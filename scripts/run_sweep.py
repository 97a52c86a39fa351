#!/usr/bin/env python
"""One GPU's share of BASELINE.json configs[4]: the R0 x migration sweep of SURVEY §8(d) config 5 -- 256 x 256 grid,
model of config 4 with K = 10, one replicate per point, 8,192 points per GPU (rank `r` of 8 takes rows 32r..32r+31
of the R0 grid).  Direct method for `iters` events, then `leaps` tau leaps, genealogy, summaries.  One JSON line.

    python scripts/run_sweep.py [rank] [world] [iters] [leaps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402
from vgsim_b200.sweep import Sweep  # noqa: E402

rank = int(sys.argv[1]) if len(sys.argv) > 1 else 0
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
leaps = int(sys.argv[4]) if len(sys.argv) > 4 else 64
dims, setup = SCENARIOS["table3_k10"]
B = np.linspace(1.2, 4.0, 256) * (0.099 + 0.001)
MIG = np.logspace(-4, -1, 256)


def mk(b, m):
    def f(e):
        e.set_transmission_rate(float(b), None)
        e.set_total_migration_probability(float(m))
    return f


pts = [mk(b, m) for b in B for m in MIG]
out = {"config": "sweep 256x256 (R0 x total migration), table-3 model K=10, rank %d of %d" % (rank, world)}
t0 = time.time()
sw = Sweep(dims, setup, pts, 1, seed=31337, rank=rank, world=world)
out["setup_s"] = time.time() - t0
out["replicates"] = sw.R
t0 = time.time(); sw.simulate(iters, 10 ** 9, -1, "direct"); out["direct_s"] = time.time() - t0
out["direct_kernel_ms"] = sw.h.last_kernel_ms()
c = sw.counters()
out["direct_events"] = int(c["events"].sum())
t0 = time.time(); sw.simulate(leaps, 10 ** 9, -1, "tau"); out["tau_s"] = time.time() - t0
out["tau_kernel_ms"] = sw.h.last_kernel_ms()
c2 = sw.counters()
ev = sum(int(c2[k].sum() - c[k].sum()) for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"))
out.update(tau_leaps=int(c2["leaps"].sum()), tau_events=ev, tau_events_per_s=ev / (out["tau_kernel_ms"] * 1e-3),
           samples_total=int(c2["sCounter"].sum()), device_errors=int(sw.h.synchronize(strict=False)))
t0 = time.time(); sw.genealogy(); out["genealogy_s"] = time.time() - t0
s = sw.summaries()[:, 0]
rows = s.reshape(-1, 256, s.shape[1])          # [R0 rows of this rank][migration][summary]
out["tree_nodes_total"] = float(s[:, 13].sum())
out["final_time_by_R0_row_first_last"] = [float(np.median(rows[0, :, 12])), float(np.median(rows[-1, :, 12]))]
out["migration_rows_by_mig_first_last"] = [float(rows[:, :16, 18].mean()), float(rows[:, -16:, 18].mean())]
assert out["final_time_by_R0_row_first_last"][0] > out["final_time_by_R0_row_first_last"][1], "higher R0 must get there sooner"
assert out["migration_rows_by_mig_first_last"][0] < out["migration_rows_by_mig_first_last"][1], "more migration, more migrant lineages"
print(json.dumps(out))
