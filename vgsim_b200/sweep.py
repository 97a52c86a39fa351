"""Parameter sweeps: one replicate batch in which every replicate (or group of replicates) has its own parameter
point (BASELINE.json configs[4]: "ABC-style sweep: 65,536 replicates over R0/migration grid sharded across 8xB200").

The reference runs one `Simulator` per parameter point and process.  Here a sweep is ONE handle: `n_points`
parameter blobs are uploaded once (`vgsim_upload_params`), `vgsim_set_replicate_params` maps replicates to points,
and a single launch of each kernel advances all of them; a rank owns a contiguous range of the global point list
(`_shard.replicate_range`), replicate seeds are functions of the GLOBAL replicate id, and the only exchange is the
all-gather of the fixed-size summaries.
"""
import numpy as np

from . import _capi, _shard
from ._engine import BirthDeathModel


class Sweep:
    """`points`: list of callables, each taking a configured engine and changing it through the reference's setters
    (e.g. ``lambda e: (e.set_transmission_rate(0.3, None), e.set_total_migration_probability(1e-3))``)."""

    def __init__(self, dims, base_setup, points, replicates_per_point=1, seed=0, rank=0, world=1, device=None):
        U, K, S = dims
        self.n_global = len(points)
        self.lo, self.hi = _shard.replicate_range(rank, world, self.n_global)
        self.rpp = int(replicates_per_point)
        n_pts = self.hi - self.lo
        self.R = n_pts * self.rpp
        self.engine = BirthDeathModel(U, K, S, seed, False, False, int(1e6), 0.0)   # host-side validation + arrays only
        base_setup(self.engine)
        self.h = _capi.Handle(U, K, S, self.R, n_pts, device)
        # the map first: an upload seeds the live contact density of the replicates that are mapped to its point
        self.h.set_replicate_params(np.repeat(np.arange(n_pts, dtype=np.int32), self.rpp))
        for i in range(n_pts):
            e = self.engine
            saved = {k: v.copy() for k, v in e.param_arrays().items()}
            points[self.lo + i](e)
            self.h.upload_params(i, e.param_arrays())
            self._restore(e, saved)
        gid = np.arange(self.lo * self.rpp, self.hi * self.rpp, dtype=np.uint64)
        self.h.set_seeds((np.uint64(seed) + gid).astype(np.uint64))
        Sx, I = self.engine._susceptible, self.engine._infectious
        self.h.set_state(np.ascontiguousarray(np.broadcast_to(Sx, (self.R,) + Sx.shape)),
                         np.ascontiguousarray(np.broadcast_to(I, (self.R,) + I.shape)))

    @staticmethod
    def _restore(e, saved):
        """Put the base parameter arrays back (the engine's arrays are the live views the setters write into)."""
        live = e.param_arrays()
        names = dict(b="bRate", d="dRate", s="sRate", mRate="mRate", hapMutType="hapMutType", sigma="_susceptibility",
                     suscType="suscType", T="suscepTransition", m="migrationRates", cd="contactDensity",
                     cdBefore="contactDensityBeforeLockdown", cdAfter="contactDensityAfterLockdown", startLD="startLD",
                     endLD="endLD", sm="samplingMultiplier", sizes="sizes")
        for k, attr in names.items():
            getattr(e, attr)[...] = saved[k]
        del live

    def simulate(self, iterations, sample_size=None, epidemic_time=-1, method="direct", attempts=200):
        ss = iterations if sample_size is None else sample_size
        if method == "direct":
            self.h.simulate_direct(int(iterations), int(ss), float(epidemic_time), int(attempts))
        elif method == "tau":
            self.h.simulate_tau(int(iterations), int(ss), float(epidemic_time), int(attempts))
        else:
            raise ValueError("Unknown method. Choose between 'direct' and 'tau'.")

    def genealogy(self, seed=None):
        self.h.genealogy(seed, None)

    def summaries(self):
        """[points on this rank, replicates_per_point, NSUMMARY] (include/vgsim_b200.h: vgsim_summaries)."""
        return self.h.summaries().reshape(self.hi - self.lo, self.rpp, _capi.NSUMMARY)

    def counters(self):
        return self.h.get_counters()
