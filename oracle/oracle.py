"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of the CPU oracle (oracle/libvgsim_oracle.so) and a loader
for the out-of-tree build of the unmodified reference engine (oracle/_ref, see build_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (vgsim_b200/) never does.
"""
import contextlib
import ctypes
import io
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libvgsim_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

c_void_p, c_int, c_int64, c_uint64, c_uint32, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                        ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _load():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "vgsim_oracle.cpp")):
        build()
    lib = ctypes.CDLL(LIB)
    P = c_void_p
    sig = {
        "vgo_create": (P, [c_int, c_int, c_int, c_uint64]),
        "vgo_destroy": (None, [P]),
        "vgo_set_params": (None, [P] * 17),
        "vgo_set_state": (None, [P, P, P]),
        "vgo_get_state": (None, [P, P, P]),
        "vgo_simulate_direct": (c_int, [P, c_int64, c_int64, c_float, c_int64]),
        "vgo_simulate_tau": (c_int, [P, c_int64, c_int64, c_float, c_int64]),
        "vgo_get_counters": (None, [P, P, P]),
        "vgo_num_events": (c_int64, [P]),
        "vgo_num_multievents": (c_int64, [P]),
        "vgo_prop_num": (c_int64, [P]),
        "vgo_get_events": (None, [P, P]),
        "vgo_set_events": (None, [P, P, c_int64]),
        "vgo_get_multievents": (None, [P] * 8),
        "vgo_append_tau_log": (None, [P, P, P, c_int64]),
        "vgo_update_all_rates": (None, [P]),
        "vgo_propensities": (None, [P, P, P, P, P]),
        "vgo_rates": (None, [P] * 9),
        "vgo_inject_uniforms": (None, [P, P, c_int64]),
        "vgo_injected_used": (c_int64, [P]),
        "vgo_genealogy": (c_int, [P, c_int, c_uint64]),
        "vgo_tree_size": (c_int64, [P]),
        "vgo_get_tree": (None, [P, P, P, P]),
        "vgo_num_mutations": (c_int64, [P]),
        "vgo_get_mutations": (None, [P] * 6),
        "vgo_num_migrations": (c_int64, [P]),
        "vgo_get_migrations": (None, [P] * 5),
        "vgo_num_lockdowns": (c_int64, [P]),
        "vgo_get_lockdowns": (None, [P] * 4),
        "vgo_clamped": (c_int64, [P]),
        "vgo_multievents_overflowed": (c_int, [P]),
        "vgo_rng_create": (P, [c_uint64, c_uint32]),
        "vgo_rng_destroy": (None, [P]),
        "vgo_rng_doubles": (None, [P, P, c_int64]),
        "vgo_rng_raw": (None, [P, P, c_int64]),
        "vgo_rng_poisson": (None, [P, P, P, c_int64]),
        "vgo_rng_hypergeometric": (None, [P, P, P, P, P, c_int64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
COUNTER_NAMES = ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus", "migNonPlus",
                 "swapLockdown", "good_attempt", "events", "multievents", "globalInfectious")


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


class OracleModel:
    """One replicate of the reference algorithm on the CPU."""

    def __init__(self, sites, K, S, seed):
        self.sites, self.K, self.S, self.H = sites, K, S, 4 ** sites
        self._h = c_void_p(lib.vgo_create(sites, K, S, seed))
        self._keep = []

    @classmethod
    def from_engine(cls, eng, seed=None):
        """Build from anything exposing the engine's parameter arrays (vgsim_b200._engine.BirthDeathModel)."""
        m = cls(eng.sites, eng.popNum, eng.susNum, eng.user_seed if seed is None else seed)
        m.set_params(eng.param_arrays())
        m.set_state(eng.susceptible, eng.infectious)
        return m

    def __del__(self):
        if getattr(self, "_h", None):
            lib.vgo_destroy(self._h)
            self._h = None

    def set_params(self, a):
        f, i = np.float64, np.int64
        order = [("b", f), ("d", f), ("s", f), ("mRate", f), ("hapMutType", f), ("sigma", f), ("suscType", i), ("T", f),
                 ("m", f), ("cd", f), ("cdBefore", f), ("cdAfter", f), ("startLD", f), ("endLD", f), ("sm", f), ("sizes", i)]
        arrs = [None if a.get(k) is None else np.ascontiguousarray(a[k], dtype=t) for k, t in order]
        lib.vgo_set_params(self._h, *[_p(x) for x in arrs])

    def set_state(self, Sx, I):
        Sx = None if Sx is None else np.ascontiguousarray(Sx, np.int64)
        I = None if I is None else np.ascontiguousarray(I, np.int64)
        lib.vgo_set_state(self._h, _p(Sx), _p(I))

    def get_state(self):
        Sx = np.zeros((self.K, self.S), np.int64)
        I = np.zeros((self.K, self.H), np.int64)
        lib.vgo_get_state(self._h, _p(Sx), _p(I))
        return Sx, I

    def simulate(self, iterations, sample_size=None, epidemic_time=-1, method="direct", attempts=200):
        if sample_size is None:
            sample_size = iterations
        fn = lib.vgo_simulate_direct if method == "direct" else lib.vgo_simulate_tau
        rc = fn(self._h, iterations, sample_size, epidemic_time, attempts)
        if rc:
            raise RuntimeError("oracle error %d" % rc)

    def counters(self):
        c = np.zeros(12, np.int64)
        t = np.zeros(1, np.float64)
        lib.vgo_get_counters(self._h, _p(c), _p(t))
        d = {k: int(v) for k, v in zip(COUNTER_NAMES, c)}
        d["time"] = float(t[0])
        return d

    @property
    def P(self):
        return int(lib.vgo_prop_num(self._h))

    def events(self):
        n = int(lib.vgo_num_events(self._h))
        out = np.zeros((6, n), np.float64)
        lib.vgo_get_events(self._h, _p(out))
        return out

    def set_events(self, chain6xN):
        chain = np.ascontiguousarray(chain6xN, np.float64)
        lib.vgo_set_events(self._h, _p(chain), chain.shape[1])

    def multievents(self):
        n = int(lib.vgo_num_multievents(self._h))
        i = np.int64
        num, typ, hap, pop, nhap, npop = (np.zeros(n, i) for _ in range(6))
        t = np.zeros(n, np.float64)
        lib.vgo_get_multievents(self._h, _p(num), _p(t), _p(typ), _p(hap), _p(pop), _p(nhap), _p(npop))
        return dict(num=num, time=t, type=typ, hap=hap, pop=pop, nhap=nhap, npop=npop)

    def append_tau_log(self, counts, times):
        counts = np.ascontiguousarray(counts, np.int64)
        times = np.ascontiguousarray(times, np.float64)
        lib.vgo_append_tau_log(self._h, _p(counts), _p(times), counts.shape[0])

    def propensities(self):
        out = np.zeros(self.P, np.float64)
        dI = np.zeros((self.K, self.H), np.float64)
        dS = np.zeros((self.K, self.S), np.float64)
        tau = np.zeros(1, np.float64)
        lib.vgo_propensities(self._h, _p(out), _p(dI), _p(dS), _p(tau))
        return out, dI, dS, float(tau[0])

    def rates(self):
        K, H = self.K, self.H
        r = dict(A=np.zeros(K), eff=np.zeros((K, K)), maxEBM=np.zeros(K), ev=np.zeros((K, H, 4)), hp=np.zeros((K, H)),
                 popRate=np.zeros(K), migPop=np.zeros(K), totals=np.zeros(2))
        lib.vgo_rates(self._h, *[_p(r[k]) for k in ("A", "eff", "maxEBM", "ev", "hp", "popRate", "migPop", "totals")])
        return r

    def inject_uniforms(self, u):
        u = np.ascontiguousarray(u, np.float64)
        self._keep.append(u)
        lib.vgo_inject_uniforms(self._h, _p(u), len(u))

    def injected_used(self):
        return int(lib.vgo_injected_used(self._h))

    def genealogy(self, seed=None):
        rc = lib.vgo_genealogy(self._h, 0 if seed is None else 1, 0 if seed is None else seed)
        if rc:
            raise RuntimeError("oracle genealogy error %d" % rc)

    def tree(self):
        n = int(lib.vgo_tree_size(self._h))
        tree, pop, t = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float64)
        lib.vgo_get_tree(self._h, _p(tree), _p(pop), _p(t))
        return tree, pop, t

    def mutations(self):
        n = int(lib.vgo_num_mutations(self._h))
        node, AS, DS, site = (np.zeros(n, np.int64) for _ in range(4))
        t = np.zeros(n, np.float64)
        lib.vgo_get_mutations(self._h, _p(node), _p(AS), _p(DS), _p(site), _p(t))
        return node, AS, DS, site, t

    def migrations(self):
        n = int(lib.vgo_num_migrations(self._h))
        node, oldp, newp = (np.zeros(n, np.int64) for _ in range(3))
        t = np.zeros(n, np.float64)
        lib.vgo_get_migrations(self._h, _p(node), _p(t), _p(oldp), _p(newp))
        return node, t, oldp, newp

    def lockdowns(self):
        n = int(lib.vgo_num_lockdowns(self._h))
        st, pop = np.zeros(n, np.int64), np.zeros(n, np.int64)
        t = np.zeros(n, np.float64)
        lib.vgo_get_lockdowns(self._h, _p(st), _p(pop), _p(t))
        return st, pop, t

    def clamped(self):
        return int(lib.vgo_clamped(self._h))


class OracleRng:
    def __init__(self, entropy, num):
        self._g = c_void_p(lib.vgo_rng_create(entropy, num))

    def __del__(self):
        if getattr(self, "_g", None):
            lib.vgo_rng_destroy(self._g)
            self._g = None

    def doubles(self, n):
        out = np.zeros(n, np.float64)
        lib.vgo_rng_doubles(self._g, _p(out), n)
        return out

    def raw(self, n):
        out = np.zeros(n, np.uint64)
        lib.vgo_rng_raw(self._g, _p(out), n)
        return out

    def poisson(self, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        out = np.zeros(len(lam), np.int64)
        lib.vgo_rng_poisson(self._g, _p(lam), _p(out), len(lam))
        return out

    def hypergeometric(self, good, bad, sample):
        good, bad, sample = (np.ascontiguousarray(x, np.int64) for x in (good, bad, sample))
        out = np.zeros(len(good), np.int64)
        lib.vgo_rng_hypergeometric(self._g, _p(good), _p(bad), _p(sample), _p(out), len(good))
        return out


# --------------------------------------------------------------------------------------------------
# The real reference (built by build_ref.py from /root/reference sources; only the built artefacts are here)
def choose_tau(dI, dS, I, Sx):
    """ChooseTau (src/_BirthDeath.pyx:2432-2450) on GIVEN drifts, with the reference's float epsilon product (quirk Q1).
    Returns (tau, |drift| of the compartment that binds it, or 0 when tau stayed 1).  TEST INFRASTRUCTURE: splits a
    tau comparison into 'same formula on the same drifts' (exact) and 'same drifts' (summation-order tolerance)."""
    eps = np.float32(0.03)
    tau, bind = 1.0, 0.0
    for cnt, d in ((np.asarray(I), np.asarray(dI)), (np.asarray(Sx), np.asarray(dS))):
        x = (eps * cnt.astype(np.float32)).astype(np.float64) / 2.0
        ad = np.abs(d.astype(np.float64))
        ok = ad >= 1e-8
        if not ok.any():
            continue
        cand = np.where(ok, np.maximum(1.0, x) / np.where(ok, ad, 1.0), np.inf)
        k = np.unravel_index(np.argmin(cand), cand.shape)
        if cand[k] < tau:
            tau, bind = float(cand[k]), float(ad[k])
    return tau, bind


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "VGsim"))


def load_reference():
    """Returns the reference's BirthDeathModel class (unmodified engine, out-of-tree build)."""
    if not reference_available():
        raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from VGsim import BirthDeathModel  # noqa
    return BirthDeathModel


@contextlib.contextmanager
def quiet():
    """The reference prints progress from C level print(); swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def make_reference(sites=0, K=1, S=1, seed=0):
    BD = load_reference()
    return BD(number_of_sites=sites, populations_number=K, number_of_susceptible_groups=S, seed=seed,
              sampling_probability=False, memory_optimization=False, genome_length=int(1e6),
              recombination_probability=0.0)


# ---------------------------------------------------------------------------------------------------------
# Epidemic curves: literal restatement of get_data_infectious / get_data_susceptible
# (reference src/_BirthDeath.pyx:1967-2005 and :2008-2045), including the operator-precedence quirk of the DEATH /
# SAMPLING branch (`a or b or c and d and e` -> every DEATH / SAMPLING of ANY cell decrements Data and, for
# SAMPLING, increments Sample) and the multi-event MIGRATION branch of the susceptible curve comparing
# `haplotypes` (not `newHaplotypes`) with the group index.  chain = 6 x N export_chain_events layout (used rows
# only); multi = dict of the multiEvents SoA (num, type, hap, pop, nhap, npop) or None.  TEST INFRASTRUCTURE.
BIRTH, DEATH, SAMPLING, MUTATION, SUSCCHANGE, MIGRATION, MULTITYPE = range(7)


def ref_data_infectious(chain, multi, initial, current_time, pop, hap, step_num):
    tp = [i * current_time / step_num for i in range(step_num + 1)]
    Data = np.zeros(step_num + 1)
    Sample = np.zeros(step_num + 1)
    Data[0] = initial
    t_, ty_, h_, p_, nh_, np_ = (chain[k] for k in range(6))
    point = 0
    for i in range(chain.shape[1]):
        while point != step_num and tp[point] < t_[i]:
            Data[point + 1] = Data[point]
            Sample[point + 1] = Sample[point]
            point += 1
        ty = int(ty_[i])
        if ty == BIRTH and p_[i] == pop and h_[i] == hap:
            Data[point] += 1
        elif ty == DEATH or ty == SAMPLING or ty == MUTATION and p_[i] == pop and h_[i] == hap:
            Data[point] -= 1
            if ty == SAMPLING:
                Sample[point] += 1
        elif ty == MUTATION and nh_[i] == hap and p_[i] == pop:
            Data[point] += 1
        elif ty == MIGRATION and np_[i] == pop and h_[i] == hap:
            Data[point] += 1
        elif ty == MULTITYPE:
            lo, hi = int(h_[i]), int(p_[i])
            nz = lo + np.nonzero(multi["num"][lo:hi])[0]  # records with num == 0 add nothing
            for j in nz:
                mt, n = int(multi["type"][j]), int(multi["num"][j])
                if mt == BIRTH and multi["hap"][j] == hap and multi["pop"][j] == pop:
                    Data[point] += n
                elif mt == DEATH or mt == SAMPLING or mt == MUTATION and multi["hap"][j] == hap and multi["pop"][j] == pop:
                    Data[point] -= n
                    if mt == SAMPLING:
                        Sample[point] += n
                elif mt == MUTATION and multi["nhap"][j] == hap and multi["pop"][j] == pop:
                    Data[point] += n
                elif mt == MIGRATION and multi["npop"][j] == pop and multi["hap"][j] == hap:
                    Data[point] += n
    return Data, Sample, tp


def ref_data_susceptible(chain, multi, initial, current_time, pop, sus, step_num):
    tp = [i * current_time / step_num for i in range(step_num + 1)]
    Data = np.zeros(step_num + 1)
    Data[0] = initial
    t_, ty_, h_, p_, nh_, np_ = (chain[k] for k in range(6))
    point = 0
    for i in range(chain.shape[1]):
        while point != step_num and tp[point] < t_[i]:
            Data[point + 1] = Data[point]
            point += 1
        ty = int(ty_[i])
        if ty == BIRTH and p_[i] == pop and nh_[i] == sus:
            Data[point] -= 1
        elif (ty == DEATH or ty == SAMPLING or ty == SUSCCHANGE) and p_[i] == pop and nh_[i] == sus:
            Data[point] += 1
        elif ty == SUSCCHANGE and h_[i] == sus and p_[i] == pop:
            Data[point] -= 1
        elif ty == MIGRATION and np_[i] == pop and nh_[i] == sus:
            Data[point] -= 1
        elif ty == MULTITYPE:
            lo, hi = int(h_[i]), int(p_[i])
            nz = lo + np.nonzero(multi["num"][lo:hi])[0]
            for j in nz:
                mt, n = int(multi["type"][j]), int(multi["num"][j])
                if mt == BIRTH and multi["nhap"][j] == sus and multi["pop"][j] == pop:
                    Data[point] -= n
                elif (mt == DEATH or mt == SAMPLING or mt == SUSCCHANGE) and multi["nhap"][j] == sus and multi["pop"][j] == pop:
                    Data[point] += n
                elif mt == SUSCCHANGE and multi["hap"][j] == sus and multi["pop"][j] == pop:
                    Data[point] -= n
                elif mt == MIGRATION and multi["npop"][j] == pop and multi["hap"][j] == sus:
                    Data[point] -= n
    return Data, tp


def ref_epidemiology_timelines(chain, sizes, K, S, H, current_time, step_num):
    """Literal restatement of output_epidemiology_timelines (reference src/_BirthDeath.pyx:1765-1847, dict branch).
    Upstream the method cannot run (it reads `self.susceptible_num`, an attribute that does not exist, :1767), so there
    is no reference output to pin against: this loop IS the specification the vectorised product code is checked
    with.  Returns (times, sus[pts][K][S], inf[pts][K][H])."""
    tp = [i * current_time / step_num for i in range(step_num + 1)]
    sus = np.zeros((K, S), np.int64)
    inf = np.zeros((K, H), np.int64)
    for i in range(K):
        sus[i, 0] = sizes[i]
    inf[0, 0] += 1
    sus[0, 0] -= 1
    t_, ty_, h_, p_, nh_, np_ = (chain[k] for k in range(6))
    times, out_s, out_i = [], [], []
    point = 0
    for j in range(chain.shape[1]):
        ty, h, p, nh, npp = int(ty_[j]), int(h_[j]), int(p_[j]), int(nh_[j]), int(np_[j])
        if ty == BIRTH:
            inf[p, h] += 1; sus[p, nh] -= 1
        elif ty == DEATH or ty == SAMPLING:
            inf[p, h] -= 1; sus[p, nh] += 1
        elif ty == MUTATION:
            inf[p, h] -= 1; inf[p, nh] += 1
        elif ty == SUSCCHANGE:
            sus[p, h] -= 1; sus[p, nh] += 1
        elif ty == MIGRATION:
            sus[npp, nh] -= 1; inf[npp, h] += 1
        if point <= step_num and tp[point] <= t_[j]:   # (the reference would raise IndexError past the last point)
            times.append(tp[point]); out_s.append(sus.copy()); out_i.append(inf.copy())
            point += 1
    return times, np.array(out_s).reshape(-1, K, S), np.array(out_i).reshape(-1, K, H)
