#!/usr/bin/env python
"""Where the time of a blocked world-shape tau run goes: per block, the tau kernel (device ms and wall) and the archive pass (wall)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 6
BLOCK = int(sys.argv[3]) if len(sys.argv) > 3 else 101
(U, K, S), setup = SCENARIOS["w"]
e = Eng(U, K, S, 4242, False, False, int(1e6), 0.0, replicates=R)
setup(e)
e.SimulatePopulation(10 ** 7, 10 ** 9, 40.0, 200)
h = e._handle
for b in range(NB):
    t0 = time.time(); h.simulate_tau(BLOCK, 100000, -1.0, 1); t1 = time.time()
    kms = h.last_kernel_ms()
    h.archive_tau_log(); h.wait(); t2 = time.time()
    print("block %d: tau wall %.1f ms (kernel %.1f ms), archive wall %.1f ms, stats %s" % (b, (t1 - t0) * 1e3, kms, (t2 - t1) * 1e3, h.archive_stats()), flush=True)
