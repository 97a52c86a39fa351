#!/usr/bin/env python
"""Calibration data for the group cost model of tau_sched_kernel: per CTA of a boustrophedon launch, its
duration and the cell counts of the replicates of its two groups.  Prints one JSON line."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
R, T = 4096, float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
(U, K, S), setup = SCENARIOS["t3"]
e = Eng(U, K, S, 1000, False, False, int(1e6), 0.0, replicates=R, device=0)
setup(e)
h = e._sync_params()
h.simulate_direct(250000, -1, T, 200)
if len(sys.argv) > 2:
    h.set_tau_variant(0); h.simulate_tau(int(sys.argv[2]), -1, -1.0, 1); h.recycle_log()
Sx, I = h.get_state()
w = (I.reshape(R, -1) != 0).sum(axis=1)
ev0 = h.get_counters()
h.set_tau_variant(2)
h.tau_phase_cycles(reset=True); h.tau_cta_end(reset=True)
h.simulate_tau(32, -1, -1.0, 1)
ms = h.last_kernel_ms()
ce = h.tau_cta_end(reset=True)[:148].astype(np.int64)
c1 = h.get_counters()
evs = sum(c1[k] - ev0[k] for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"))
idle = (ce[:, 1].max() - ce[:, 1]) * 1e-6
order = np.argsort(-w, kind="stable")
nw, G = 14, 148
out = {"kernel_ms": ms, "cta": []}
for b in range(G):
    gs = [b, 2 * G - 1 - b]
    rec = {"b": b, "dur_ms": ms - float(idle[b]), "groups": []}
    for g in gs:
        rr = order[g * nw:(g + 1) * nw]
        if len(rr):
            rec["groups"].append({"wmax": int(w[rr].max()), "wmean": float(w[rr].mean()), "n": int(len(rr)), "ev_max": int(evs[rr].max()), "ev_mean": float(evs[rr].mean())})
    out["cta"].append(rec)
print(json.dumps(out))
