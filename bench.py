#!/usr/bin/env python
"""bench.py — simulated events/sec of the replicate-batched tau-leap path (BASELINE.json metric).

Workload ("T3"): 3 sites (64 haplotypes) x 10 demes x 3 susceptibility
groups, 1e6 hosts per deme, R = 4096 replicates PER GPU (weak scaling: replicates shard over ranks with
no data-path collective; one NCCL all-gather of the per-replicate summaries at the end).
Phase A (setup, untimed): every replicate runs the direct method on the device to t = 60 (the
reference's intended "direct warm-up, then tau" usage) and the resulting compartment
states are snapshotted.  A STEP = one batch: vgsim_reset + state upload + vgsim_simulate_tau of
L leaps for all R replicates (fresh seeds per step), i.e. R*L leaps, R*L*P Poisson channel draws and
R*L*(4P+16) bytes of dense event log written to HBM (13.8 GB per step at the defaults >> 126 MB L2,
so no L2 flush is needed between steps).

  value  : events/s with the batch's input state already resident in HBM (device->device restore).
  e2e    : the same step through the C ABI with HOST buffers: seeds + states are copied host->device
           from pinned memory and counters + final states are read back device->host every step.
           Two handles alternate (double buffering): the copies of one overlap the kernel of the other.
           The event log stays in HBM, as it stays inside the engine object in the reference; it is
           what vgsim_genealogy consumes.
  roofline: dominant kernel = tau_warp_kernel; achieved = leaps * (4P+16) B / its CUDA-event duration
           (events recorded inside vgsim_simulate_tau on the launching stream).
  windows: the same step measured from snapshots of the SAME trajectories further into the epidemic
           (SURVEY 8(d) config 3 times tau from t=60 to t=150): the replicates are advanced by tau-leaping
           in leap blocks (vgsim_recycle_log between blocks) to t=90 and t=120 and every window reports
           leaps/s, events per leap, active cells and its own roofline fraction; `value` stays the t=60
           window, `roofline_min_frac` is the smallest fraction over the windows.
  direct : phase A's device direct method (16 B of log per event) with its own roofline and CPU baseline.
  cpu_baseline / --impl reference: the UNMODIFIED reference engine (oracle/_ref, Cython build of
           /root/reference made by oracle/build_ref.py) on the host cores, one process per core,
           each process running whole replicates of the same workload (direct warm-up to t = 60
           untimed, then L tau leaps timed).  If oracle/_ref is absent the CPU oracle port is used.

Launch: python bench.py [--gpus N --steps K --warmup W]   (N > 1: under torch.distributed.run)
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WARM_CAP = 600000   # log capacity of the reference's direct warm-up (the longest T3 epidemics have ~3e5 events by t = 60)
T_WARM = 60.0
SEED0 = 1000
WORKLOAD = "T3 tau-leap: 3 sites (64 haplotypes) x 10 demes x 3 susceptibility groups, 1e6/deme"
CPU_SAMPLE_TIMEOUT = 180.0   # wall seconds allowed for one bounded CPU sample (all workers)
KERNEL = "tau_warp_kernel"   # 4,096 replicates per GPU run on the warp-per-replicate kernel
EVENT_KEYS = ("bCounter", "dCounter", "sCounter", "mCounter", "iCounter", "migPlus")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicates", type=int, default=4096, help="replicates per GPU")
    ap.add_argument("--leaps", type=int, default=32, help="tau leaps per replicate per step")
    ap.add_argument("--scenario", default="t3")
    ap.add_argument("--profile-window", type=float, default=None,
                    help="cudaProfilerStart/Stop around the timed steps of this window (ncu --profile-from-start off)")
    ap.add_argument("--windows", default="60,90,120", help="epidemic times of the measured windows (first = the metric's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-curves", action="store_true", help="skip the (untimed) epidemic-curves pass reported beside the metric")
    ap.add_argument("--phases", action="store_true", help="add the tau kernel's per-phase critical-path cycles (timing tap)")
    ap.add_argument("--cpu-seconds", type=float, default=4.0, help="target timed CPU seconds per worker")
    ap.add_argument("--cpu-worker", nargs=4, metavar=("SEED", "REPS", "LEAPS", "SCENARIO"), default=None)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) or, failing that, the oracle port.  TEST/BASELINE
# infrastructure: the only place bench.py touches oracle/.
def _ref_counters(ref):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.PrintCounters()
    tot = 0
    for line in buf.getvalue().splitlines():
        if ":" in line:
            tot += int(line.rsplit(":", 1)[1])
    return tot


def cpu_worker(seed, reps, leaps, scenario):
    """One host core: `reps` replicates, each direct to T_WARM (untimed) then `leaps` tau leaps (timed)."""
    import numpy as np  # noqa
    from oracle import oracle as O
    from scenarios import SCENARIOS
    (U, K, S), setup = SCENARIOS[scenario]
    H = 4 ** U
    P = K * ((K - 1) * H * S + S * (S - 1) + H * (2 + 3 * U + S))
    use_ref = O.reference_available()
    t_tau = 0.0
    events = 0
    n_leaps = 0
    t_direct = 0.0
    n_direct_ev = 0
    t_loop0 = time.perf_counter()
    for r in range(reps):
        if use_ref:
            def fresh():
                m = O.make_reference(U, K, S, seed + r)
                setup(m)
                return m
            with O.quiet():
                # pass 1 learns the warm-up's row count; pass 2 repeats it with iterations == that count so the
                # tau call's log allocation is exactly `leaps` rows (avoids reference quirk Q3)
                m = fresh()
                m.SimulatePopulation(WARM_CAP, 10 ** 9, T_WARM, 200)
                import tempfile
                with tempfile.TemporaryDirectory() as d:
                    m.export_chain_events(os.path.join(d, "c"))
                    n_direct = int(np.count_nonzero(np.load(os.path.join(d, "c.npy"))[0]))
                del m
                m = fresh()
                t0 = time.perf_counter()
                m.SimulatePopulation(n_direct, 10 ** 9, T_WARM, 200)
                t_direct += time.perf_counter() - t0
                n_direct_ev += n_direct
                c0 = _ref_counters(m)
                t0 = time.perf_counter()
                m.SimulatePopulation_tau(leaps, 10 ** 9, -1, 1)
                t_tau += time.perf_counter() - t0
                events += _ref_counters(m) - c0
                n_leaps += leaps
        else:
            from vgsim_b200._engine import BirthDeathModel as Eng
            e = Eng(U, K, S, seed + r, False, False, int(1e6), 0.0)
            setup(e)
            om = O.OracleModel.from_engine(e)
            t0 = time.perf_counter()
            om.simulate(2000000, sample_size=10 ** 9, epidemic_time=T_WARM)
            t_direct += time.perf_counter() - t0
            c0 = om.counters()
            n_direct_ev += c0["events"]
            t0 = time.perf_counter()
            om.simulate(leaps, sample_size=10 ** 9, epidemic_time=-1, method="tau", attempts=1)
            t_tau += time.perf_counter() - t0
            c1 = om.counters()
            events += sum(c1[k] - c0[k] for k in EVENT_KEYS)
            n_leaps += c1["events"] - c0["events"]
    # what the reference's per-call log allocation costs (multievents.CreateEvents(iterations * propNum) inside every
    # SimulatePopulation_tau call, src/_BirthDeath.pyx:2305: 7 arrays of leaps*P 8-byte zeros, first-touch page faults):
    # measured on equal arrays so that the rate can also be quoted without it.  The GPU arm reuses its log capacity.
    t_alloc = 0.0
    if use_ref:
        t0 = time.perf_counter()
        for _ in range(3):
            z = [np.zeros(leaps * P, dtype=np.int64) for _ in range(7)]
            for a_ in z:
                a_[::512] = 1
            del z
        t_alloc = (time.perf_counter() - t0) / 3 * reps
    print(json.dumps({"events": int(events), "leaps": int(n_leaps), "seconds": t_tau, "P": P,
                      "kind": "reference" if use_ref else "port", "direct_events": int(n_direct_ev),
                      "direct_seconds": t_direct, "alloc_seconds": t_alloc,
                      "wall_seconds": time.perf_counter() - t_loop0}))
    sys.stdout.flush()
    # the reference's destructor can abort at interpreter teardown (free(): invalid pointer): run the registered exit
    # hooks (the driver's record of loaded .so files among them) explicitly, then leave without teardown
    try:
        import atexit
        atexit._run_exitfuncs()
    except Exception:
        pass
    os._exit(0)


def run_cpu_sample(scenario, leaps, reps_per_worker, seed_base):
    """All host cores at once, one process per core; returns aggregate events/s etc."""
    cores = os.cpu_count() or 1
    procs = []
    t0 = time.perf_counter()
    for c in range(cores):
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(seed_base + c * reps_per_worker),
               str(reps_per_worker), str(leaps), scenario]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    outs = []
    deadline = time.perf_counter() + CPU_SAMPLE_TIMEOUT
    for p in procs:
        try:
            out, _ = p.communicate(timeout=max(1.0, deadline - time.perf_counter()))
        except subprocess.TimeoutExpired:  # a stuck worker must not stall the bench: drop it, keep the others
            p.kill()
            out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("{"):
                outs.append(json.loads(line))
    wall = time.perf_counter() - t0
    if not outs:
        raise RuntimeError("CPU baseline workers produced no output")
    ev = sum(o["events"] for o in outs)
    lp = sum(o["leaps"] for o in outs)
    # workers run concurrently; aggregate rate = sum of per-worker rates over their timed (tau) sections
    rate = sum(o["events"] / o["seconds"] for o in outs if o["seconds"] > 0)
    lrate = sum(o["leaps"] / o["seconds"] for o in outs if o["seconds"] > 0)
    drate = sum(o["direct_events"] / o["direct_seconds"] for o in outs if o.get("direct_seconds", 0) > 0)
    rate_xa = sum(o["events"] / max(o["seconds"] - o.get("alloc_seconds", 0.0), 1e-9) for o in outs if o["seconds"] > 0)
    return dict(events_per_s=rate, leaps_per_s=lrate, events=ev, leaps=lp, cores=len(outs), wall=wall,
                kind=outs[0]["kind"], P=outs[0]["P"], timed_seconds=max(o["seconds"] for o in outs),
                direct_events_per_s=drate, events_per_s_excl_alloc=rate_xa)


def calibrate_cpu_reps(scenario, leaps, target_seconds, wall_budget=None):
    """Two replicates on one core to size the sample (reference T3: ~2 ms per leap on one core).  `target_seconds` bounds
    the TIMED tau seconds per worker, `wall_budget` (optional) everything a worker does for one sample -- the untimed
    direct warm-up of every replicate included -- plus the start-up of the worker process itself."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(SEED0), "2", str(leaps), scenario]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    wall = time.perf_counter() - t0
    o = [json.loads(l) for l in out.splitlines() if l.startswith("{")][0]
    reps = max(1, int(round(target_seconds / max(o["seconds"] / 2, 1e-3))))
    if wall_budget is not None:
        per_rep = max(o.get("wall_seconds", o["seconds"]) / 2, 1e-3)
        startup = max(wall - o.get("wall_seconds", o["seconds"]), 0.0)
        reps = max(1, min(reps, int((wall_budget - startup) / per_rep)))
    return reps


def reference_arm(args, rank):
    if rank != 0:
        return
    # Every step is a bounded sample of the workload, sized so that the whole --steps K --warmup W run ends within a few
    # minutes on the box's host cores (VGSIM_REF_BUDGET_S, default 150 s of wall clock for the K + W samples together).
    budget = float(os.environ.get("VGSIM_REF_BUDGET_S", "150"))
    reps = calibrate_cpu_reps(args.scenario, args.leaps, args.cpu_seconds, budget / max(args.steps + args.warmup, 1))
    for _ in range(args.warmup):
        run_cpu_sample(args.scenario, args.leaps, 1, SEED0 + 500000)
    t_events = 0
    t_rate = []
    walls = []
    res = None
    for s in range(args.steps):
        res = run_cpu_sample(args.scenario, args.leaps, reps, SEED0 + s * 100000)
        t_events += res["events"]
        t_rate.append(res["events_per_s"])
        walls.append(res["timed_seconds"])
    value = sum(t_rate) / len(t_rate)
    sample = "%d cores x %d replicates x %d tau leaps per step (direct warm-up to t=%g untimed)" % (
        res["cores"], reps, args.leaps, T_WARM)
    line = {
        "impl": "reference", "metric": "simulated events/sec (tau-leap, replicate-batched)", "value": value,
        "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "scenario": args.scenario, "leaps_per_step": args.leaps, "channels_P": res["P"]},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": res["cores"], "kind": res["kind"], "sample": sample,
                         "leaps_per_s": res["leaps_per_s"],
                         "note": "the timed call includes the reference's own per-call log allocation (CreateEvents, "
                                 "src/_BirthDeath.pyx:2305); excluding an equal allocation measured beside it the rate "
                                 "would be %.4g events/s" % res["events_per_s_excl_alloc"],
                         "direct_events_per_s": res["direct_events_per_s"]},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (so torch can address library buffers)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def gpu_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from scenarios import SCENARIOS
    from vgsim_b200 import _capi, _shard
    from vgsim_b200._engine import BirthDeathModel as Eng

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the vgsim_b200 hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    R, L = args.replicates, args.leaps
    (U, K, S), setup = SCENARIOS[args.scenario]
    lo, hi = _shard.replicate_range(rank, world, world * R)   # weak scaling: R replicates per GPU
    seed_base = SEED0 + lo
    windows_t = [float(x) for x in args.windows.split(",") if x]

    def make_handle():
        eng = Eng(U, K, S, seed_base, False, False, int(1e6), 0.0, replicates=R, device=local_rank)
        setup(eng)
        h = eng._sync_params()
        st = torch.cuda.Stream(device=dev)
        h.set_stream(st.cuda_stream)
        return eng, h, st

    eng, h, stream = make_handle()
    P, H = h.P, h.H
    if args.phases:
        h.set_tau_variant(2)
    b_leap = 4 * P + 16
    pSx, pI = h.state_dev_ptrs()
    live_Sx = torch.as_tensor(_DevArray(pSx, (R, K, S), "<i8"), device=dev)
    live_I = torch.as_tensor(_DevArray(pI, (R, K, H), "<i8"), device=dev)

    # ---- Phase A (untimed for the metric; reported as the `direct` block): device direct method to T_WARM
    h.simulate_direct(250000, -1, windows_t[0], 200)
    direct_ms = h.last_kernel_ms()
    cA = h.get_counters()

    def snapshot():
        with torch.cuda.stream(stream):
            sx, ii = live_Sx.clone(), live_I.clone()
        stream.synchronize()
        return sx, ii

    # ---- the same trajectories further into the epidemic: tau-leaping in leap blocks up to each window's time
    snaps = []
    t_now = windows_t[0]
    for wt in windows_t:
        blocks = 0
        if wt > t_now:
            for blocks in range(1, 400):
                h.recycle_log()
                h.simulate_tau(L, -1, wt, 1, sync=False)
                c = h.get_counters()
                if bool(np.all((c["time"] >= np.float32(wt)) | (c["globalInfectious"] == 0))):
                    break
            t_now = wt
        c = h.get_counters()
        sx, ii = snapshot()
        iin = ii.cpu().numpy()
        snaps.append({"t": wt, "Sx": sx, "I": ii, "mean_time": float(np.mean(c["time"])),
                      "mean_infectious": float(iin.sum() / R), "mean_active_cells": float((iin != 0).sum() / R),
                      "extinct": int((iin.reshape(R, -1).sum(axis=1) == 0).sum()), "advance_blocks": blocks})
    Sx0, I0 = snaps[0]["Sx"].cpu().numpy(), snaps[0]["I"].cpu().numpy()
    hSx = torch.from_numpy(Sx0).pin_memory()
    hI = torch.from_numpy(I0).pin_memory()

    cptr, _tptr = h.counters_dev_ptrs()
    dev_counters = torch.as_tensor(_DevArray(cptr, (R, _capi.NCOUNTERS), "<i8"), device=dev)
    n_total = args.warmup + args.steps
    acc = torch.zeros((n_total, R, _capi.NCOUNTERS), dtype=torch.int64, device=dev)
    # one pinned seed buffer per step of a window: in async mode the H2D copy of a step's seeds is only enqueued, so its
    # source must stay untouched until the window has been synchronised
    seeds_pinned = torch.empty((n_total, R), dtype=torch.int64).pin_memory()

    def seeds_for(step, slot):
        # fresh Philox keys per step and per replicate, disjoint across ranks
        s = _shard.replicate_seeds(SEED0, lo, hi, batch=step + 1)
        seeds_pinned[slot].numpy()[:] = s.view(np.int64)
        return seeds_pinned[slot].numpy().view(np.uint64)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident(i, snap, slot):
        h.reset()
        h.set_seeds(seeds_for(i, slot))
        h.set_state_dev(snap["Sx"].data_ptr(), snap["I"].data_ptr())
        h.simulate_tau(L, -1, -1.0, 1, sync=False)
        with torch.cuda.stream(stream):
            acc[slot].copy_(dev_counters, non_blocking=True)

    def measure_window(snap, seed_off, clocks=None):
        """W warm-up + K timed device-resident steps from one snapshot; CUDA events on the launching stream."""
        kernel_ms = []
        h.set_async(True)    # nothing in a resident step needs the host to wait: the steps queue back to back
        for i in range(args.warmup):
            step_resident(seed_off + i, snap, i)
        barrier()
        if clocks is not None:
            clocks.start()
        launches0 = h.launch_count()
        if args.phases:
            h.tau_phase_cycles(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.warmup, n_total):
            step_resident(seed_off + i, snap, i)
            kernel_ms.append(h.kernel_ms_async())      # per-step event pair, read back after the window
        e1.record(stream)
        barrier()
        h.set_async(False)
        ms = e0.elapsed_time(e1)
        cnt = acc[args.warmup:n_total].cpu().numpy()
        kernel_ms = [k() for k in kernel_ms]
        return {"ms": ms, "events": int(cnt[:, :, :6].sum()), "leaps": int(cnt[:, :, 10].sum()),
                "launches": h.launch_count() - launches0, "kernel_ms_sum": sum(kernel_ms),
                "phase_cycles": h.tau_phase_cycles(reset=True) if args.phases else None}

    # ---- device-resident timing: the metric's window first (clocks sampled there), then the later windows
    clocks = ClockSampler(local_rank) if rank == 0 else None
    res_w = []
    for wi, snap in enumerate(snaps):
        prof_on = args.profile_window is not None and abs(args.profile_window - snap["t"]) < 1e-9
        if prof_on:
            torch.cuda.synchronize(dev)
            torch.cuda.cudart().cudaProfilerStart()
        res_w.append(measure_window(snap, wi * 10 * n_total, clocks if wi == 0 else None))
        if prof_on:
            torch.cuda.cudart().cudaProfilerStop()
    m0 = res_w[0]
    ms, events, leaps, launches, phase_cycles = m0["ms"], m0["events"], m0["leaps"], m0["launches"], m0["phase_cycles"]
    err = h.synchronize(strict=False)

    # ---- (reported beside the metric, outside the timed regions) epidemic curves of every replicate over the log the
    #      last step left in HBM: one read of R*L dense rows (B_leap each) by curves_kernel
    curves = None
    if rank == 0 and not args.no_curves:
        try:
            step_resident(5 * 10 * n_total, snaps[0], 0)
            h.wait()
            t_c = []
            for _ in range(3):
                cv = h.epidemic_curves(8, want=("infectious",))
                t_c.append(h.last_kernel_ms())
            Sx_l, I_l = h.get_state()
            assert np.array_equal(cv["infectious"][:, -1], I_l), "curves: last grid point != final state"
            log_bytes = float(R) * L * b_leap
            curves = {"kernel": "curves_kernel", "kernel_ms": min(t_c), "log_bytes_read": log_bytes,
                      "achieved_GBps": log_bytes / (min(t_c) * 1e-3) / 1e9, "grid_points": 9,
                      "check": "last grid point equals the final state of all %d replicates" % R}
        except Exception as ex:
            curves = {"error": str(ex)}

    # ---- end-to-end timing (host buffers through the C ABI).  Two handles alternate in async mode (vgsim_set_async):
    #      every call below only enqueues on its handle's stream, so the upload of one batch and the read-back of the
    #      other overlap the running kernel; a batch's results are consumed after vgsim_wait, two steps later.
    eng2, h2, stream2 = make_handle()
    hs = [h, h2]
    outs = []
    for hh in hs:
        hh.set_async(True)
        outs.append({"Sx": torch.empty_like(hSx).pin_memory(), "I": torch.empty_like(hI).pin_memory(),
                     "cnt": torch.empty((R, _capi.NCOUNTERS), dtype=torch.int64).pin_memory(),
                     "time": torch.empty(R, dtype=torch.float64).pin_memory(),
                     "seeds": torch.empty(R, dtype=torch.int64).pin_memory(), "pending": False})

    def harvest(k):
        o = outs[k]
        if not o["pending"]:
            return 0
        hs[k].wait()                                                       # the batch and its D2H copies are complete
        o["pending"] = False
        return int(o["cnt"][:, :6].sum())

    def step_e2e(i):
        k = i & 1
        o = outs[k]
        got = harvest(k)                                                   # results of step i-2
        o["seeds"].numpy()[:] = _shard.replicate_seeds(SEED0, lo, hi, batch=i + 1).view(np.int64)
        # uploads first: they overlap the other handle's kernel, which leaves no room on the SMs for anything else
        # (its CTAs hold all of an SM's shared memory), so the reset kernel behind them waits for it anyway
        hs[k].set_seeds(o["seeds"].numpy().view(np.uint64))                # H2D  R*8
        hs[k].set_state(hSx.numpy(), hI.numpy())                           # H2D  R*K*(S+H)*8 from pinned memory
        hs[k].reset()
        hs[k].simulate_tau(L, -1, -1.0, 1, sync=False)
        hs[k].get_counters(out=(o["cnt"].numpy(), o["time"].numpy()))      # D2H  R*(12+1)*8
        hs[k].get_state(out=(o["Sx"].numpy(), o["I"].numpy()))             # D2H  R*K*(S+H)*8
        o["pending"] = True
        return got

    for i in range(max(args.warmup, 2)):
        step_e2e(7 * 10 * n_total + i)
    harvest(0), harvest(1)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    f0.record(stream)
    ev_e2e = 0
    for i in range(args.steps):
        ev_e2e += step_e2e(8 * 10 * n_total + i)
    ev_e2e += harvest(0) + harvest(1)
    f1.record(stream)
    barrier()
    # the host-blocking copies make the host clock the honest one here: take the larger of the two
    ms_e2e = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - w0))
    clk = clocks.stop() if rank == 0 else None
    h.set_async(False)
    h2.close()
    del eng2

    # ---- final all-gather of per-replicate summaries (the only collective of the path)
    sptr = h.summaries_dev_ptr()
    summ = torch.as_tensor(_DevArray(sptr, (R, _capi.NSUMMARY), "<f8"), device=dev)
    stream.synchronize()
    wstats = [[w["ms"], float(w["events"]), float(w["leaps"]), w["kernel_ms_sum"]] for w in res_w]
    if world > 1:
        gathered = _shard.gather_summaries(summ.clone(), world)
        assert gathered.shape == (world * R, _capi.NSUMMARY)
        t = torch.tensor([ms, ms_e2e, float(events), float(leaps), float(ev_e2e), float(launches), float(err)] +
                         [x for w in wstats for x in w], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(tmax[0]), float(tmax[1])
        events, leaps, ev_e2e, launches = int(tsum[2]), int(tsum[3]), int(tsum[4]), int(tsum[5])
        err = int(tmax[6])
        for wi in range(len(wstats)):   # times: max over ranks; events / leaps: sums
            o = 7 + 4 * wi
            wstats[wi] = [float(tmax[o]), float(tsum[o + 1]), float(tsum[o + 2]), float(tmax[o + 3])]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        wlines = []
        for snap, (w_ms, w_ev, w_lp, w_kms) in zip(snaps, wstats):
            k_ms = w_kms / args.steps
            ach = (w_lp / max(world, 1) / args.steps) * b_leap / (k_ms * 1e-3) / 1e9
            wlines.append({"t": snap["t"], "mean_time": snap["mean_time"], "mean_infectious": snap["mean_infectious"],
                           "mean_active_cells": snap["mean_active_cells"], "extinct_replicates": snap["extinct"],
                           "events_per_s": w_ev / (w_ms * 1e-3), "leaps_per_s": w_lp / (w_ms * 1e-3),
                           "events_per_leap": w_ev / max(w_lp, 1.0), "ms_per_step": w_ms / args.steps, "kernel_ms": k_ms,
                           "achieved_GBps": ach, "frac": ach / peak})
        if args.phases:  # timing tap: per-window critical-path cycles of the kernel's phases (lane 0 of every warp)
            pn = (["wipe+lists+Q", "drifts+tau", "primary draws", "slow-path drain", "feasibility", "apply", "lockdown vote"]
                  if KERNEL == "tau_kernel" else
                  ["-", "row wipe issue + drifts + tau", "primary draws + drain", "feasibility", "apply + lists", "-", "-"])
            for wl, rw in zip(wlines, res_w):
                pcw = rw["phase_cycles"]
                nlw = max(int(pcw[7]), 1)
                wl["phase_cycles_per_leap"] = {n: round(float(pcw[i]) / nlw, 1) for i, n in enumerate(pn) if n != "-"}
        k_ms, achieved = wlines[0]["kernel_ms"], wlines[0]["achieved_GBps"]
        h2d = R * 8 + R * K * (S + H) * 8
        d2h = R * (_capi.NCOUNTERS + 1) * 8 + R * K * (S + H) * 8
        n_direct = float(np.sum(cA["events"]))
        line = {
            "metric": "simulated events/sec (tau-leap, replicate-batched)",
            "value": events / (ms * 1e-3), "unit": "events/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "scenario": args.scenario, "replicates_per_gpu": R, "leaps_per_step": L,
                       "channels_P": P, "phase_a": "device direct method to t=%g per replicate (untimed)" % windows_t[0],
                       "window": "t=%g (value); later windows of the same trajectories under `windows`" % windows_t[0],
                       "l2": "each step writes %.1f GB of event log per GPU (> 126 MB L2): no flush needed" % (R * L * b_leap / 1e9),
                       "parallelism": "replicates sharded %dx%d, no data-path collective" % (world, R),
                       "e2e_pipeline": "two handles in async mode alternate: H2D/D2H of one batch overlap the kernel of the other"},
            "leaps_per_s": leaps / (ms * 1e-3), "channel_draws_per_s": leaps * P / (ms * 1e-3),
            "events_per_leap": events / max(leaps, 1),
            "roofline": {"bound": "hbm", "kernel": KERNEL, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_leap": b_leap, "kernel_ms": k_ms},
            "windows": wlines, "roofline_min_frac": min(w["frac"] for w in wlines),
            "e2e": {"value": ev_e2e / (ms_e2e * 1e-3), "unit": "events/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "device_error_flags": err,
            "direct": {"kernel": "direct_kernel", "what": "phase A: %d replicates from one infected host to t=%g" % (R, windows_t[0]),
                       "mean_events": float(np.mean(cA["events"])), "max_events": float(np.max(cA["events"])),
                       "mean_time": float(np.mean(cA["time"])), "mean_infectious": snaps[0]["mean_infectious"],
                       "kernel_ms": direct_ms, "events_per_s": n_direct / (direct_ms * 1e-3),
                       "roofline": {"bound": "hbm", "achieved": n_direct * 16 / (direct_ms * 1e-3) / 1e9, "peak": peak,
                                    "unit": "GB/s", "frac": n_direct * 16 / (direct_ms * 1e-3) / 1e9 / peak,
                                    "bytes_per_event": 16,
                                    "note": "latency-bound: one dependent chain per replicate, the batch ends with its longest replicate"}},
        }
        line["phase_a"] = {"mean_events": line["direct"]["mean_events"], "mean_time": line["direct"]["mean_time"],
                           "mean_infectious": line["direct"]["mean_infectious"], "direct_kernel_ms": direct_ms,
                           "direct_events_per_s": line["direct"]["events_per_s"]}
        if curves is not None:
            line["epidemic_curves"] = curves
            if "achieved_GBps" in curves:
                curves["frac_of_hbm_peak"] = curves["achieved_GBps"] / peak
        if phase_cycles is not None:
            names = (["wipe+lists+Q", "drifts+tau", "primary draws", "slow-path drain", "feasibility", "apply", "lockdown vote"]
                     if KERNEL == "tau_kernel" else
                     ["-", "row wipe issue + drifts + tau", "primary draws + drain", "feasibility", "apply + lists", "-", "-"])
            nl = max(int(phase_cycles[7]), 1)
            line["tau_phase_cycles_per_leap"] = {n: float(phase_cycles[i]) / nl for i, n in enumerate(names)}
        prof = os.path.join(ROOT, "profiles", "tau_kernel_traffic.json")
        if os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                if tj.get("kernel") == KERNEL and tj.get("replicates") == R and tj.get("leaps") == L:
                    line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            try:
                reps = calibrate_cpu_reps(args.scenario, L, args.cpu_seconds)
                res = run_cpu_sample(args.scenario, L, reps, SEED0)
                line["cpu_baseline"] = {
                    "value": res["events_per_s"], "unit": "events/s", "cores": res["cores"], "kind": res["kind"],
                    "sample": "%d cores x %d replicates x %d tau leaps (direct warm-up to t=%g untimed)" % (
                        res["cores"], reps, L, T_WARM), "leaps_per_s": res["leaps_per_s"],
                    "note": "includes the reference's per-call log allocation (CreateEvents); %.4g events/s without an equal "
                            "allocation measured beside it" % res["events_per_s_excl_alloc"]}
                line["direct"]["cpu_baseline"] = {
                    "value": res["direct_events_per_s"], "unit": "events/s", "cores": res["cores"], "kind": res["kind"],
                    "sample": "the same workers' direct warm-up to t=%g (%d replicates per core)" % (T_WARM, reps)}
            except Exception as ex:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "events/s", "cores": 0, "kind": "unavailable",
                                        "sample": "failed: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    wd = float(os.environ.get("VGSIM_BENCH_WATCHDOG", "0") or 0)
    if wd > 0:  # diagnostics: dump every thread's Python stack to stderr if the run is still going after wd seconds
        import faulthandler
        faulthandler.dump_traceback_later(wd, repeat=True, file=sys.stderr)
    if args.cpu_worker:
        seed, reps, leaps, scenario = args.cpu_worker
        cpu_worker(int(seed), int(reps), int(leaps), scenario)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
