// Batched direct-method Gillespie (reference SimulatePopulation, src/_BirthDeath.pyx:396-738).
//
// One WARP per replicate.  The rate hierarchy of the reference
//     totalRate -> popRate[p] -> {immunePopRate[p], infectPopRate[p]} -> hapPopRate[p,h] -> eventHapPopRate[p,h,0:4]
// lives in shared memory (per warp); the two-level cumulative search of fastChoose
// (src/fast_choose.pxi:18-52) is a lane-parallel inclusive scan (__shfl_up_sync) + __ballot_sync that
// returns the same (index, residual) pair, so ONE uniform is recycled through all levels exactly like
// the reference does.  The per-deme contact factor c[p] = sum_r m[p,r]^2 cd[r]/A[r] is hoisted, so an
// event refreshes one deme in O(H*S/32) instead of the reference's O(H*S*K) (UpdateRates, :516-546).
// Events are logged as 16 bytes (fp64 time + packed 64-bit descriptor).
#include "common.cuh"
#include "handle.h"
#include "rates.cuh"

namespace vg {

struct DirShared {
    double *hp, *pr, *inf, *imm, *mp, *cd, *c, *maxEBM;
    int *I, *Sx, *totInf, *totSus, *lock;
};

__host__ __device__ inline size_t dir_warp_bytes(const Dims &D) {
    size_t nd = (size_t)D.K * D.H + (size_t)D.K * 7;
    size_t ni = (size_t)D.K * D.H + (size_t)D.K * D.S + (size_t)D.K * 3;
    return nd * 8 + ((ni + 3) & ~(size_t)3) * 4;
}

__device__ inline void dir_carve(DirShared &s, const Dims &D, unsigned char *base) {
    double *p = reinterpret_cast<double *>(base);
    s.hp = p; p += D.K * D.H;
    s.pr = p; p += D.K;
    s.inf = p; p += D.K;
    s.imm = p; p += D.K;
    s.mp = p; p += D.K;
    s.cd = p; p += D.K;
    s.c = p; p += D.K;
    s.maxEBM = p; p += D.K;
    int *q = reinterpret_cast<int *>(p);
    s.I = q; q += D.K * D.H;
    s.Sx = q; q += D.K * D.S;
    s.totInf = q; q += D.K;
    s.totSus = q; q += D.K;
    s.lock = q;
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fastChoose over n weights produced by `wf(i)`: x in [0, total) -> (index, residual in [0,1)).
// First index whose inclusive prefix sum reaches x (zero weights can never be selected); if rounding
// leaves x above the total, the last positive weight is the catch-all, like the reference's loop bound.
template <class WF>
__device__ __forceinline__ int warp_choose(WF wf, int n, double x, double &resid, int skip = -1) {
    const int lane = threadIdx.x & 31;
    double base = 0.0;
    int last_i = -1;
    double last_w = 0.0, last_prev = 0.0;
    for (int off = 0; off < n; off += 32) {
        int i = off + lane;
        double w = (i < n && i != skip) ? wf(i) : 0.0;
        double c = w;
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= o) c += t;
        }
        c += base;
        unsigned hit = __ballot_sync(0xffffffffu, w > 0.0 && c >= x);
        if (hit) {
            int l = __ffs(hit) - 1;
            double cl = __shfl_sync(0xffffffffu, c, l), wl = __shfl_sync(0xffffffffu, w, l);
            double r = (x - (cl - wl)) / wl;
            resid = r < 0.0 ? 0.0 : (r >= 1.0 ? 0.9999999999999999 : r);
            return off + l;
        }
        unsigned pos = __ballot_sync(0xffffffffu, w > 0.0);
        if (pos) {
            int l = 31 - __clz(pos);
            last_i = off + l;
            last_w = __shfl_sync(0xffffffffu, w, l);
            last_prev = __shfl_sync(0xffffffffu, c, l) - last_w;
        }
        base = __shfl_sync(0xffffffffu, c, 31);
    }
    if (last_i >= 0) {
        double r = (x - last_prev) / last_w;
        resid = r < 0.0 ? 0.0 : (r >= 1.0 ? 0.9999999999999999 : r);
    }
    return last_i;  // -1: every weight was zero (reference: "0-weight sampled", sys.exit)
}

// refresh of one deme after its compartments changed (UpdateRates(pi, infect, immune, migration))
__device__ __forceinline__ void refresh_deme(const Dims &D, const double *__restrict__ pp, const DirShared &s, int p,
                                             bool infect) {
    const int lane = threadIdx.x & 31;
    const int H = D.H, S = D.S;
    if (infect) {
        double acc = 0.0;
        const double cp = s.c[p], smp = pp[D.o_sm + p];
        for (int h = lane; h < H; h += 32) {
            double Q = 0.0;
            for (int sn = 0; sn < S; sn++) Q += (double)s.Sx[p * S + sn] * pp[D.o_sigT + sn * H + h];
            double tm = 0.0;
            for (int u = 0; u < D.U; u++) tm += pp[D.o_mu + h * D.U + u];
            double te = pp[D.o_b + h] * (Q * cp) + pp[D.o_d + h] + pp[D.o_sr + h] * smp + tm;
            double v = te * (double)s.I[p * H + h];
            s.hp[p * H + h] = v;
            acc += v;
        }
        acc = warp_sum(acc);
        if (lane == 0) s.inf[p] = acc;
    }
    if (lane == 0) {
        double im = 0.0;
        for (int sn = 0; sn < S; sn++) im += pp[D.o_Tc + sn] * (double)s.Sx[p * S + sn];
        s.imm[p] = im;
    }
    __syncwarp();
    if (lane == 0) s.pr[p] = s.inf[p] + s.imm[p];
    __syncwarp();
}

// the four event rates of (p,h): {birth, death, sampling, mutation} (eventHapPopRate, :310-314)
__device__ __forceinline__ void event_rates(const Dims &D, const double *__restrict__ pp, const DirShared &s, int p,
                                            int h, double ev[4]) {
    double Q = 0.0;
    for (int sn = 0; sn < D.S; sn++) Q += (double)s.Sx[p * D.S + sn] * pp[D.o_sigT + sn * D.H + h];
    double tm = 0.0;
    for (int u = 0; u < D.U; u++) tm += pp[D.o_mu + h * D.U + u];
    ev[0] = pp[D.o_b + h] * (Q * s.c[p]);
    ev[1] = pp[D.o_d + h];
    ev[2] = pp[D.o_sr + h] * pp[D.o_sm + p];
    ev[3] = tm;
}

__device__ __forceinline__ void refresh_migration(const Dims &D, const DirShared &s, long long ginf) {
    const int lane = threadIdx.x & 31;
    for (int q = lane; q < D.K; q += 32)
        s.mp[q] = s.maxEBM[q] * (double)s.totSus[q] * (double)(ginf - s.totInf[q]);
    __syncwarp();
}

__device__ void refresh_all(const Dims &D, const double *__restrict__ pp, const DirShared &s, long long &ginf) {
    const int lane = threadIdx.x & 31;
    long long g = 0;
    for (int p = 0; p < D.K; p++) {
        long long ti = 0, ts = 0;
        for (int h = lane; h < D.H; h += 32) ti += s.I[p * D.H + h];
        for (int sn = lane; sn < D.S; sn += 32) ts += s.Sx[p * D.S + sn];
        ti = warp_sum_ll(ti);
        ts = warp_sum_ll(ts);
        if (lane == 0) {
            s.totInf[p] = (int)ti;
            s.totSus[p] = (int)ts;
        }
        g += ti;
    }
    ginf = g;
    __syncwarp();
    for (int p = 0; p < D.K; p++) refresh_deme(D, pp, s, p, true);
    refresh_migration(D, s, ginf);
}

// sequential categorical draw over a handful of weights, executed identically by every lane
template <class WF>
__device__ __forceinline__ int small_choose(WF wf, int n, double total, double &rn) {
    double x = rn * total, acc = 0.0;
    int pick = -1;
    double wprev = 0.0, wpick = 0.0;
    int last = -1;
    double lastw = 0.0, lastprev = 0.0;
    for (int i = 0; i < n; i++) {
        double w = wf(i);
        if (w > 0.0) {
            last = i;
            lastw = w;
            lastprev = acc;
        }
        acc += w;
        if (pick < 0 && w > 0.0 && acc >= x) {
            pick = i;
            wpick = w;
            wprev = acc - w;
        }
    }
    if (pick < 0) {
        pick = last;
        wpick = lastw;
        wprev = lastprev;
    }
    if (pick >= 0) {
        double r = (x - wprev) / wpick;
        rn = r < 0.0 ? 0.0 : (r >= 1.0 ? 0.9999999999999999 : r);
    }
    return pick;
}

// 16 warps per CTA at 128 registers: the event loop is one dependent chain per replicate, so what the SM needs is
// more replicates in flight (the unbounded build used 168 registers = 8 warps per SM)
__global__ void __launch_bounds__(512, 1) direct_kernel(DevState st, SimArgs a, int warps_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Dims D = st.D;
    const int K = D.K, H = D.H, S = D.S, U = D.U;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t wbytes = (dir_warp_bytes(D) + 15) & ~(size_t)15;
    DirShared s;
    dir_carve(s, D, smem_raw + wib * wbytes);
    const int nwarps = gridDim.x * warps_per_cta;

    for (int r = blockIdx.x * warps_per_cta + wib; r < st.R; r += nwarps) {
        const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
        double *eff = st.eff + (size_t)r * K * K;
        long long *ctr = st.counters + (size_t)r * NCOUNT;
        const uint64_t seed = st.seeds[r];
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        // ---- load
        int ovf = 0;
        for (int i = lane; i < K * H; i += 32) {
            long long v = st.I[(size_t)r * K * H + i];
            if (v < 0 || v > 2147483647LL) ovf = 1;
            s.I[i] = (int)v;
        }
        for (int i = lane; i < K * S; i += 32) {
            long long v = st.Sx[(size_t)r * K * S + i];
            if (v < 0 || v > 2147483647LL) ovf = 1;
            s.Sx[i] = (int)v;
        }
        for (int i = lane; i < K; i += 32) {
            s.cd[i] = st.cd[(size_t)r * K + i];
            s.c[i] = st.ceff[(size_t)r * K + i];
            s.maxEBM[i] = st.maxEBM[(size_t)r * K + i];
            s.lock[i] = st.lock[(size_t)r * K + i];
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, ovf)) {
            if (lane == 0) st.err[r] |= ERR_COUNT_OVERFLOW;
            continue;
        }
        long long ginf = 0;
        refresh_all(D, pp, s, ginf);

        long long cB = ctr[C_B], cD = ctr[C_D], cS = ctr[C_S], cM = ctr[C_M], cI = ctr[C_I], cGp = ctr[C_MIGP],
                  cGn = ctr[C_MIGN];
        long long evptr = ctr[C_EVPTR], leaps = ctr[C_LEAPS], good_attempt = ctr[C_GOOD];
        int swaps = 0, errbits = 0;
        double t = st.time[r];
        unsigned epoch = st.epoch[r];
        const long long ev_limit = evptr + a.iterations;
        double *ev_time = st.ev_time + (size_t)r * st.ev_cap;
        unsigned long long *ev_desc = st.ev_desc + (size_t)r * st.ev_cap;
        int *loc_sp = st.loc_sp + (size_t)r * st.loc_cap;
        double *loc_t = st.loc_t + (size_t)r * st.loc_cap;

        for (long long attempt = 0; attempt < a.attempts; attempt++) {
            epoch++;
            unsigned long long iter = 0;
            double Rt = 0.0, Rm = 0.0;
            {
                double x = 0.0, y = 0.0;
                for (int p = lane; p < K; p += 32) {
                    x += s.pr[p];
                    y += s.mp[p];
                }
                Rt = warp_sum(x);
                Rm = warp_sum(y);
            }
            if (Rt + Rm != 0.0 && ginf != 0) {
                while (evptr < ev_limit && evptr < st.ev_cap && (a.sample_size == -1 || cS <= a.sample_size) &&
                       (!a.has_time || t < (double)a.time)) {
                    // ---- SampleTime + GenerateEvent (:476-512): two uniforms per iteration
                    uint4 w = philox4x32_10(make_uint4((uint32_t)iter, (uint32_t)(iter >> 32), epoch, 0x44495245u), key);
                    iter++;
                    double u1 = u53(w.x, w.y), rn = u53(w.z, w.w);
                    if (u1 <= 0.0) u1 = 1.0 / 9007199254740992.0;
                    t += -log(u1) / (Rt + Rm);
                    double choose = rn * (Rt + Rm);
                    int touched = -1;      // deme whose rates changed
                    bool infect = false;   // hapPopRate of `touched` must be rebuilt
                    unsigned long long desc = 0;
                    bool logged = false;
                    if (Rt > choose) {
                        rn = choose / Rt;
                        int p = warp_choose([&](int i) { return s.pr[i]; }, K, rn * Rt, rn);
                        if (p < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                        double ch2 = rn * s.pr[p];
                        if (s.imm[p] > ch2) {
                            // ---- ImmunityTransition (:550-564)
                            rn = ch2 / s.imm[p];
                            int ssi = small_choose([&](int i) { return pp[D.o_Tc + i] * (double)s.Sx[p * S + i]; }, S,
                                                   s.imm[p], rn);
                            if (ssi < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                            int tsi = small_choose([&](int i) { return pp[D.o_T + ssi * S + i]; }, S, pp[D.o_Tc + ssi], rn);
                            if (tsi < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                            if (lane == 0) {
                                s.Sx[p * S + ssi] -= 1;
                                s.Sx[p * S + tsi] += 1;
                            }
                            __syncwarp();
                            cI++;
                            desc = pack_event(EV_SUSCCHANGE, ssi, p, tsi, 0);
                            logged = true;
                            touched = p;
                            infect = true;  // susceptible composition changed -> birth rates of the deme change
                        } else {
                            rn = (ch2 - s.imm[p]) / s.inf[p];
                            int h = warp_choose([&](int i) { return s.hp[p * H + i]; }, H, rn * s.inf[p], rn);
                            if (h < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                            double ev[4];
                            event_rates(D, pp, s, p, h, ev);
                            int e = small_choose([&](int i) { return ev[i]; }, 4, ev[0] + ev[1] + ev[2] + ev[3], rn);
                            if (e == 0) {
                                // ---- Birth (:568-605)
                                double ws = 0.0;
                                for (int i = 0; i < S; i++) ws += (double)s.Sx[p * S + i] * pp[D.o_sigT + i * H + h];
                                int si = small_choose([&](int i) { return (double)s.Sx[p * S + i] * pp[D.o_sigT + i * H + h]; },
                                                      S, ws, rn);
                                if (si < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                                if (lane == 0) {
                                    s.Sx[p * S + si] -= 1;
                                    s.totSus[p] -= 1;
                                    s.I[p * H + h] += 1;
                                    s.totInf[p] += 1;
                                }
                                ginf += 1;
                                cB++;
                                desc = pack_event(EV_BIRTH, h, p, si, 0);
                            } else if (e == 1 || e == 2) {
                                // ---- Death / Sampling (:616-635)
                                int g = (int)pp[D.o_g + h];
                                if (lane == 0) {
                                    s.Sx[p * S + g] += 1;
                                    s.totSus[p] += 1;
                                    s.I[p * H + h] -= 1;
                                    s.totInf[p] -= 1;
                                }
                                ginf -= 1;
                                if (e == 1) cD++; else cS++;
                                desc = pack_event(e == 1 ? EV_DEATH : EV_SAMPLING, h, p, g, 0);
                            } else {
                                // ---- Mutation (:640-667)
                                int mi = small_choose([&](int i) { return pp[D.o_mu + h * U + i]; }, U, ev[3], rn);
                                if (mi < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                                const double *wv = pp + D.o_w + (h * U + mi) * 3;
                                int ds = small_choose([&](int i) { return wv[i]; }, 3, wv[0] + wv[1] + wv[2], rn);
                                if (ds < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                                int nh = mutate_hap(h, mi, ds, U);
                                if (lane == 0) {
                                    s.I[p * H + nh] += 1;
                                    s.I[p * H + h] -= 1;
                                }
                                cM++;
                                desc = pack_event(EV_MUTATION, h, p, nh, 0);
                            }
                            __syncwarp();
                            logged = true;
                            touched = p;
                            infect = true;
                        }
                    } else {
                        // ---- GenerateMigration (:672-694): rejection-sampled cross-deme infection
                        rn = (choose - Rt) / Rm;
                        int tp = warp_choose([&](int i) { return s.mp[i]; }, K, rn * Rm, rn);
                        if (tp < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                        int sp = warp_choose([&](int i) { return (double)s.totInf[i]; }, K,
                                             rn * (double)(ginf - s.totInf[tp]), rn, tp);
                        if (sp < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                        int h = warp_choose([&](int i) { return (double)s.I[sp * H + i]; }, H, rn * (double)s.totInf[sp], rn);
                        if (h < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                        int si = small_choose([&](int i) { return (double)s.Sx[tp * S + i]; }, S, (double)s.totSus[tp], rn);
                        if (si < 0) { errbits |= ERR_ZERO_WEIGHT; break; }
                        double p_accept = eff[sp * K + tp] * pp[D.o_b + h] * pp[D.o_sigT + si * H + h] / s.maxEBM[tp];
                        if (rn < p_accept) {
                            if (lane == 0) {
                                s.Sx[tp * S + si] -= 1;
                                s.totSus[tp] -= 1;
                                s.I[tp * H + h] += 1;
                                s.totInf[tp] += 1;
                            }
                            __syncwarp();
                            ginf += 1;
                            cGp++;
                            desc = pack_event(EV_MIGRATION, h, sp, si, tp);
                            logged = true;
                            touched = tp;
                            infect = true;
                        } else {
                            cGn++;
                            touched = tp;  // the reference still runs CheckLockdown on the target deme
                        }
                    }
                    if (logged) {
                        if (lane == 0) {
                            ev_time[evptr] = t;
                            ev_desc[evptr] = desc;
                        }
                        evptr++;
                        refresh_deme(D, pp, s, touched, infect);
                        refresh_migration(D, s, ginf);
                        double x = 0.0, y = 0.0;
                        for (int p = lane; p < K; p += 32) {
                            x += s.pr[p];
                            y += s.mp[p];
                        }
                        Rt = warp_sum(x);
                        Rm = warp_sum(y);
                    }
                    if (Rt == 0.0 || ginf == 0) break;
                    // ---- CheckLockdown(pi) (:412, :698-710)
                    int flips = 0;
                    if (lane == 0)
                        flips = check_lockdown(D, pp, touched, (long long)s.totInf[touched], s.cd, s.lock, t, &st.loc_n[r],
                                               loc_sp, loc_t, st.loc_cap, &st.err[r]);
                    flips = __shfl_sync(0xffffffffu, flips, 0);
                    if (flips) {
                        swaps += flips;
                        update_contact_rates(WarpGroup(), D, pp, s.cd, eff, s.c, s.maxEBM);
                        refresh_all(D, pp, s, ginf);
                        double x = 0.0, y = 0.0;
                        for (int p = lane; p < K; p += 32) {
                            x += s.pr[p];
                            y += s.mp[p];
                        }
                        Rt = warp_sum(x);
                        Rm = warp_sum(y);
                    }
                }
            }
            if (errbits) break;
            // ---- extinction retry (:414-418) and Restart (:714-738)
            if (evptr + st.ev_base[r] <= 100 && a.iterations > 100) {
                evptr = 0;
                leaps = 0;
                cB = cD = cS = cM = cI = cGp = cGn = 0;
                t = 0.0;
                for (int i = lane; i < K * H; i += 32) s.I[i] = (int)st.initI[(size_t)r * K * H + i];
                for (int i = lane; i < K * S; i += 32) s.Sx[i] = (int)st.initSx[(size_t)r * K * S + i];
                __syncwarp();
                refresh_all(D, pp, s, ginf);
                int flips = 0;
                if (lane == 0)
                    for (int p = 0; p < K; p++)
                        flips += check_lockdown(D, pp, p, (long long)s.totInf[p], s.cd, s.lock, t, &st.loc_n[r], loc_sp,
                                                loc_t, st.loc_cap, &st.err[r]);
                flips = __shfl_sync(0xffffffffu, flips, 0);
                if (flips) {
                    swaps += flips;
                    update_contact_rates(WarpGroup(), D, pp, s.cd, eff, s.c, s.maxEBM);
                    refresh_all(D, pp, s, ginf);
                }
                good_attempt = 0;
            } else {
                good_attempt = attempt + 1;
                break;
            }
        }

        // ---- commit
        __syncwarp();
        for (int i = lane; i < K * H; i += 32) st.I[(size_t)r * K * H + i] = s.I[i];
        for (int i = lane; i < K * S; i += 32) st.Sx[(size_t)r * K * S + i] = s.Sx[i];
        for (int i = lane; i < K; i += 32) {
            st.cd[(size_t)r * K + i] = s.cd[i];
            st.ceff[(size_t)r * K + i] = s.c[i];
            st.maxEBM[(size_t)r * K + i] = s.maxEBM[i];
            st.lock[(size_t)r * K + i] = s.lock[i];
        }
        if (lane == 0) {
            ctr[C_B] = cB; ctr[C_D] = cD; ctr[C_S] = cS; ctr[C_M] = cM; ctr[C_I] = cI;
            ctr[C_MIGP] = cGp; ctr[C_MIGN] = cGn;
            ctr[C_SWAP] += swaps;
            ctr[C_GOOD] = good_attempt;
            ctr[C_EVPTR] = evptr;
            ctr[C_LEAPS] = leaps;
            ctr[C_GINF] = ginf;
            st.time[r] = t;
            st.epoch[r] = epoch;
            if (errbits) st.err[r] |= errbits;
        }
        __syncwarp();
    }
}

// Deterministic parity tap: the rate hierarchy of replicate r's current state (UpdateAllRates, :279-351)
__global__ void rates_tap_kernel(DevState st, int r, double *ev, double *hp, double *popRate, double *migPop,
                                 double *totals) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Dims D = st.D;
    const int lane = threadIdx.x & 31;
    DirShared s;
    dir_carve(s, D, smem_raw);
    const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
    for (int i = lane; i < D.K * D.H; i += 32) s.I[i] = (int)st.I[(size_t)r * D.K * D.H + i];
    for (int i = lane; i < D.K * D.S; i += 32) s.Sx[i] = (int)st.Sx[(size_t)r * D.K * D.S + i];
    for (int i = lane; i < D.K; i += 32) {
        s.cd[i] = st.cd[(size_t)r * D.K + i];
        s.c[i] = st.ceff[(size_t)r * D.K + i];
        s.maxEBM[i] = st.maxEBM[(size_t)r * D.K + i];
    }
    __syncwarp();
    long long ginf = 0;
    refresh_all(D, pp, s, ginf);
    for (int i = lane; i < D.K * D.H; i += 32) {
        double e4[4];
        event_rates(D, pp, s, i / D.H, i % D.H, e4);
        for (int j = 0; j < 4; j++) ev[i * 4 + j] = e4[j];
        hp[i] = s.hp[i];
    }
    double x = 0.0, y = 0.0;
    for (int p = lane; p < D.K; p += 32) {
        popRate[p] = s.pr[p];
        migPop[p] = s.mp[p];
        x += s.pr[p];
        y += s.mp[p];
    }
    x = warp_sum(x);
    y = warp_sum(y);
    if (lane == 0) {
        totals[0] = x;
        totals[1] = y;
    }
}

cudaError_t launch_direct(const DevState &st, const SimArgs &a, cudaStream_t stream, int num_sms) {
    size_t wbytes = (dir_warp_bytes(st.D) + 15) & ~(size_t)15;
    int wpc = 16;
    while (wpc > 1 && wbytes * wpc > 220 * 1024) wpc >>= 1;
    size_t smem = wbytes * wpc;
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, direct_kernel, wpc * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int need = (st.R + wpc - 1) / wpc;
    int grid = num_sms * per_sm;
    if (grid > need) grid = need;
    direct_kernel<<<grid, wpc * 32, smem, stream>>>(st, a, wpc);
    return cudaGetLastError();
}

cudaError_t launch_rates_tap(const DevState &st, int r, double *ev, double *hp, double *popRate, double *migPop,
                             double *totals, cudaStream_t stream) {
    size_t smem = (dir_warp_bytes(st.D) + 15) & ~(size_t)15;
    cudaError_t e = cudaFuncSetAttribute(rates_tap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rates_tap_kernel<<<1, 32, smem, stream>>>(st, r, ev, hp, popRate, migPop, totals);
    return cudaGetLastError();
}

}  // namespace vg
