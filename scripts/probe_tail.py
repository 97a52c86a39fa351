#!/usr/bin/env python
"""How long before the end of a tau launch do its warps run out of work?  T3 bench shape, one launch of 32 leaps from the
t=60 state with the timing-tap variant of the kernel (bit 1): spread of the warps' finish times on the global timer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
(U, K, S), setup = SCENARIOS["t3"]
e = Eng(U, K, S, 1000, False, False, int(1e6), 0.0, replicates=R, device=0)
setup(e)
h = e._sync_params()
h.simulate_direct(250000, -1, T, 200)
h.set_tau_variant(2)
for rep in range(3):
    h.tau_phase_cycles(reset=True)
    h.tau_cta_end(reset=True)
    h.simulate_tau(32, -1, -1.0, 1)
    ms = h.last_kernel_ms()
    pc = h.tau_phase_cycles(reset=True)
    latest, s40, n, earliest = int(pc[8]), int(pc[9]), int(pc[10]), int(pc[11])
    mean = (s40 / n) if n else 0
    base = earliest & 0xffffffffff
    print("launch %d: kernel %.3f ms; warps %d; finish spread: earliest -%.3f ms, mean -%.3f ms before the last warp" % (
        rep, ms, n, (latest - earliest) * 1e-6, ((latest & 0xffffffffff) - mean) * 1e-6))
    ce = h.tau_cta_end(reset=True)[:148].astype(np.int64)
    last = ce[:, 1].max()
    d = (last - ce[:, 1]) * 1e-6           # how long before the end each CTA's last warp finished (ms)
    print("   CTA idle tail (ms): mean %.3f, quantiles 10/50/90/100%%: %s; CTAs 0-3: %s; CTAs 144-147: %s" % (
        d.mean(), np.round(np.quantile(d, [0.1, 0.5, 0.9, 1.0]), 3), np.round(d[:4], 3), np.round(d[-4:], 3)))
    print("   within-CTA spread first->last warp (ms): mean %.3f max %.3f" % (((ce[:, 1] - ce[:, 0]) * 1e-6).mean(), ((ce[:, 1] - ce[:, 0]) * 1e-6).max()))
    srt = np.argsort(d)
    print("   busiest CTAs (finish last):", srt[:8].tolist(), " idlest:", srt[-8:].tolist())
    h.recycle_log()
