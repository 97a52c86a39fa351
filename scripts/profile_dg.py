#!/usr/bin/env python
"""Driver for ncu captures of direct_kernel and genealogy_kernel: T3 shape, R replicates, device direct method from one
infected host to t = T, then the genealogy of every replicate.  usage: profile_dg.py [R] [T] [scenario]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = float(sys.argv[2]) if len(sys.argv) > 2 else 55.0
name = sys.argv[3] if len(sys.argv) > 3 else "t3"
(U, K, S), setup = SCENARIOS[name]
e = Eng(U, K, S, 1000, False, False, int(1e6), 0.0, replicates=R, device=0)
setup(e)
h = e._sync_params()
h.simulate_direct(250000, -1, T, 200)
ms_d = h.last_kernel_ms()
c = h.get_counters()
t0 = time.perf_counter()
h.genealogy(seed=np.arange(R, dtype=np.uint64) + 7)
h.synchronize(strict=False)
ms_g = 1e3 * (time.perf_counter() - t0)
print("direct: %d replicates, %.0f mean / %d max events, %.2f ms, %.3g events/s; genealogy %.1f ms wall, %.0f mean samples" % (
    R, c["events"].mean(), c["events"].max(), ms_d, c["events"].sum() / (ms_d * 1e-3), ms_g, c["sCounter"].mean()))
