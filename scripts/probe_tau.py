"""Quick throughput probe of the tau kernel (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from scenarios import SCENARIOS
from vgsim_b200._engine import BirthDeathModel as Eng
from oracle import oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "t3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = int(sys.argv[3]) if len(sys.argv) > 3 else 32
t_warm = float(sys.argv[4]) if len(sys.argv) > 4 else 70.0
(U, K, S), setup = SCENARIOS[name]
e0 = Eng(U, K, S, 11, False, False, int(1e6), 0.0); setup(e0)
om = O.OracleModel.from_engine(e0)
t = time.time(); om.simulate(10**7, sample_size=10**9, epidemic_time=t_warm); print("oracle warm-up %.2fs" % (time.time() - t), om.counters())
Sx0, I0 = om.get_state()
print("infectious total", I0.sum(), "nonzero cells", (I0 > 0).sum(), "of", I0.size)
e = Eng(U, K, S, 1000, False, False, int(1e6), 0.0, replicates=R); setup(e)
e._susceptible[...] = Sx0; e._infectious[...] = I0
h = e._sync_params()
h.set_stream(torch.cuda.current_stream().cuda_stream)
P = h.P
for it in range(3):
    c0 = h.get_counters()
    ev0 = sum(c0[k].sum() for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"))
    l0 = c0["leaps"].sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    h.simulate_tau(L, -1, -1.0, 1, sync=False)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    c1 = h.get_counters()
    ev1 = sum(c1[k].sum() for k in ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus"))
    leaps = c1["leaps"].sum() - l0
    print("iter %d: %.2f ms, leaps %d, events %d -> %.3e events/s, %.3e leaps/s, %.3e channel-draws/s, log %.1f GB/s"
          % (it, ms, leaps, ev1 - ev0, (ev1 - ev0) / ms * 1e3, leaps / ms * 1e3, leaps * P / ms * 1e3,
             leaps * (4 * P + 16) / ms / 1e6))
print("err", h.synchronize(strict=False))
