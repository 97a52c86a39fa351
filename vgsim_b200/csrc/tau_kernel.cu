// Replicate-batched tau-leaping (reference SimulatePopulation_tau, src/_BirthDeath.pyx:2293-2593).
//
// One CTA per replicate (grid-stride over replicates).  Compartment counts, drifts, per-leap deltas
// and the parameter point live in shared memory for the whole run; per leap the CTA
//   1. contracts the migration force of infection F[t,h] = sum_s eff[t,s] m[s,s] b[h] I[s,h]
//   2. assembles the net drifts and tau (ChooseTau, :2432-2450, incl. the float-epsilon quirk)
//   3. walks the P positional channels in chunks of 4: propensity -> lambda ->
//      Poisson draw from Philox(key=seed, ctr=(chunk, leap, retry|epoch)) -> shared-memory deltas,
//      and streams the int32 counts to the dense HBM log with 16-byte stores
//   4. checks feasibility (:2522-2528, with the source-deme book-keeping quirk Q8); on failure
//      halves tau and redraws the whole leap (:2316-2321)
//   5. applies the deltas, appends the MULTITYPE row, runs CheckLockdown for every deme.
// The same channel code backs the deterministic parity tap (propensity_kernel).
#include "common.cuh"
#include "rates.cuh"
#include "samplers.cuh"

namespace vg {

struct TauShared {
    // fp64
    double *b, *d, *sr, *q, *tmq, *sigT, *T, *sm, *cd, *c, *mdiag, *sizeD, *maxEBM, *dI, *dS, *F, *red;
    // int32
    int *I, *Sx, *chkI, *updI, *dSx, *g, *lock, *tot;
    int *flags;  // [0]=leap sampling count, [1]=bad flag, [2]=flip flag, [3]=overflow
};

__host__ __device__ inline size_t tau_smem_bytes(const Dims &D) {
    size_t nd = (size_t)D.H * 4 + (size_t)D.H * D.U * 3 + (size_t)D.S * D.H + (size_t)D.S * D.S + (size_t)D.K * 7 +
                (size_t)D.K * D.H * 2 + (size_t)D.K * D.S + 40;
    size_t ni = (size_t)D.K * D.H * 3 + (size_t)D.K * D.S * 2 + D.H + D.K * 2 + 8;
    return nd * 8 + ((ni + 1) & ~(size_t)1) * 4;
}

__device__ inline void carve(TauShared &s, const Dims &D, unsigned char *base) {
    double *p = reinterpret_cast<double *>(base);
    s.b = p; p += D.H;
    s.d = p; p += D.H;
    s.sr = p; p += D.H;
    s.tmq = p; p += D.H;
    s.q = p; p += D.H * D.U * 3;
    s.sigT = p; p += D.S * D.H;
    s.T = p; p += D.S * D.S;
    s.sm = p; p += D.K;
    s.cd = p; p += D.K;
    s.c = p; p += D.K;
    s.mdiag = p; p += D.K;
    s.sizeD = p; p += D.K;
    s.maxEBM = p; p += D.K;
    p += D.K;  // spare
    s.dI = p; p += D.K * D.H;
    s.F = p; p += D.K * D.H;
    s.dS = p; p += D.K * D.S;
    s.red = p; p += 40;
    int *q = reinterpret_cast<int *>(p);
    s.I = q; q += D.K * D.H;
    s.chkI = q; q += D.K * D.H;
    s.updI = q; q += D.K * D.H;
    s.Sx = q; q += D.K * D.S;
    s.dSx = q; q += D.K * D.S;
    s.g = q; q += D.H;
    s.lock = q; q += D.K;
    s.tot = q; q += D.K;
    s.flags = q;
}

// One reaction channel: positional index c -> propensity (per unit time), plus where its count goes.
struct Channel {
    int type;        // EV_* ; EV_MULTITYPE = padding (c >= P)
    int i_dec;       // I cell that loses n   (-1 none)
    int i_inc;       // I cell that gains n   (-1 none)
    int i_chk;       // I cell the feasibility check books the gain on (quirk Q8: source deme for migration)
    int s_dec;       // Sx cell that loses n  (-1 none)
    int s_inc;       // Sx cell that gains n  (-1 none)
    double prop;
};

__device__ __forceinline__ void decode_channel(int c, const Dims &D, const TauShared &s, const double *eff,
                                               Channel &ch) {
    const int K = D.K, H = D.H, S = D.S;
    ch.i_dec = ch.i_inc = ch.i_chk = ch.s_dec = ch.s_inc = -1;
    ch.prop = 0.0;
    if (c < D.NA) {
        // MIGRATION  [sp][tp != sp][s][h]  (:2366-2367)
        int row = c >> D.hshift, h = c & (H - 1);
        int pair = row / S, sn = row - pair * S;
        int sp = pair / (K - 1), tpp = pair - sp * (K - 1);
        int tp = tpp + (tpp >= sp ? 1 : 0);
        ch.type = EV_MIGRATION;
        ch.i_inc = tp * H + h;
        ch.i_chk = sp * H + h;
        ch.s_dec = tp * S + sn;
        int Ii = s.I[sp * H + h];
        if (Ii != 0)
            ch.prop = eff[tp * K + sp] * (double)s.Sx[tp * S + sn] * (double)Ii * s.b[h] * s.sigT[sn * H + h] * s.mdiag[sp];
        return;
    }
    if (c >= D.P) {
        ch.type = EV_MULTITYPE;
        return;
    }
    int c2 = c - D.NA;
    int p = c2 / D.PD, r = c2 - p * D.PD;
    if (r < D.SS1) {
        // SUSCCHANGE [p][ss][ts != ss]  (:2378)
        int ss = r / (S - 1), tsp = r - ss * (S - 1);
        int ts = tsp + (tsp >= ss ? 1 : 0);
        ch.type = EV_SUSCCHANGE;
        ch.s_dec = p * S + ss;
        ch.s_inc = p * S + ts;
        ch.prop = s.T[ss * S + ts] * (double)s.Sx[p * S + ss];
        return;
    }
    int r2 = r - D.SS1;
    int h = r2 / D.E, e = r2 - h * D.E;
    int cell = p * H + h;
    int Ii = s.I[cell];
    if (e == 0) {  // RECOVERY (:2386)
        ch.type = EV_DEATH;
        ch.i_dec = cell;
        ch.s_inc = p * S + s.g[h];
        ch.prop = s.d[h] * (double)Ii;
    } else if (e == 1) {  // SAMPLING (:2392)
        ch.type = EV_SAMPLING;
        ch.i_dec = cell;
        ch.s_inc = p * S + s.g[h];
        ch.prop = s.sr[h] * (double)Ii * s.sm[p];
    } else if (e < 2 + 3 * D.U) {  // MUTATION (:2400-2401)
        int uk = e - 2, u = uk / 3, k = uk - u * 3;
        ch.type = EV_MUTATION;
        ch.i_dec = cell;
        ch.i_inc = ch.i_chk = p * H + mutate_hap(h, u, k, D.U);
        ch.prop = s.q[h * D.U * 3 + uk] * (double)Ii;
    } else {  // TRANSMISSION (:2410-2414)
        int sn = e - 2 - 3 * D.U;
        ch.type = EV_BIRTH;
        ch.i_inc = ch.i_chk = cell;
        ch.s_dec = p * S + sn;
        if (Ii != 0) ch.prop = s.b[h] * s.sigT[sn * H + h] * s.c[p] * (double)s.Sx[p * S + sn] * (double)Ii;
    }
}

__device__ __forceinline__ double block_min(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[w] = v;
    __syncthreads();
    v = red[0];
    for (int i = 1; i < nw; i++) v = fmin(v, red[i]);
    return v;
}

// F, drifts and tau of the current shared-memory state (steps 1-2).  Returns tau (uniform).
__device__ double drifts_and_tau(const Dims &D, const TauShared &s, const double *eff) {
    const int K = D.K, H = D.H, S = D.S, U = D.U;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < K * H; i += nt) {
        int tp = i >> D.hshift, h = i & (H - 1);
        double acc = 0.0;
        for (int sp = 0; sp < K; sp++) {
            int Ii = s.I[sp * H + h];
            if (sp != tp && Ii != 0) acc += eff[tp * K + sp] * s.mdiag[sp] * (s.b[h] * (double)Ii);
        }
        s.F[i] = acc;
    }
    __syncthreads();
    double tmin = 1.0;
    const float eps = 0.03f;
    for (int i = tid; i < K * H; i += nt) {
        int p = i >> D.hshift, h = i & (H - 1);
        double Q = 0.0;
        for (int sn = 0; sn < S; sn++) Q += (double)s.Sx[p * S + sn] * s.sigT[sn * H + h];
        double Ii = (double)s.I[i];
        double v = Q * s.F[i] + s.c[p] * s.b[h] * Ii * Q - (s.d[h] + s.sr[h] * s.sm[p] + s.tmq[h]) * Ii;
        for (int u = 0; u < U; u++) {
            int sh = 2 * (U - u - 1);
            int hu = (h >> sh) & 3;
            for (int a = 0; a < 4; a++) {
                if (a == hu) continue;
                int src = h + ((a - hu) << sh);
                int k = hu - (hu > a ? 1 : 0);
                v += s.q[(src * U + u) * 3 + k] * (double)s.I[p * H + src];
            }
        }
        s.dI[i] = v;
        if (fabs(v) >= 1e-8) {
            double x = (double)(eps * (float)s.I[i]) / 2.0;  // float product, like the reference's generated C
            double t = (1.0 > x ? 1.0 : x) / fabs(v);
            tmin = fmin(tmin, t);
        }
    }
    for (int i = tid; i < K * S; i += nt) {
        int p = i / S, sn = i - p * S;
        double part = 0.0, rec = 0.0;
        for (int h = 0; h < H; h++) {
            double Ii = (double)s.I[p * H + h];
            part += s.sigT[sn * H + h] * (s.F[p * H + h] + s.c[p] * s.b[h] * Ii);
            if (s.g[h] == sn) rec += (s.d[h] + s.sr[h] * s.sm[p]) * Ii;
        }
        double v = -(double)s.Sx[i] * part + rec;
        for (int s2 = 0; s2 < S; s2++)
            if (s2 != sn) v += s.T[s2 * S + sn] * (double)s.Sx[p * S + s2] - s.T[sn * S + s2] * (double)s.Sx[i];
        s.dS[i] = v;
        if (fabs(v) >= 1e-8) {
            double x = (double)(eps * (float)s.Sx[i]) / 2.0;
            double t = (1.0 > x ? 1.0 : x) / fabs(v);
            tmin = fmin(tmin, t);
        }
    }
    return block_min(tmin, s.red);
}

// Load the parameter point and replicate state into shared memory.
__device__ void load_replicate(const DevState &st, int r, const Dims &D, TauShared &s, const double *pp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int K = D.K, H = D.H, S = D.S, U = D.U;
    for (int i = tid; i < H; i += nt) {
        s.b[i] = pp[D.o_b + i];
        s.d[i] = pp[D.o_d + i];
        s.sr[i] = pp[D.o_sr + i];
        s.tmq[i] = pp[D.o_tmq + i];
        s.g[i] = (int)pp[D.o_g + i];
    }
    for (int i = tid; i < H * U * 3; i += nt) s.q[i] = pp[D.o_q + i];
    for (int i = tid; i < S * H; i += nt) s.sigT[i] = pp[D.o_sigT + i];
    for (int i = tid; i < S * S; i += nt) s.T[i] = pp[D.o_T + i];
    for (int i = tid; i < K; i += nt) {
        s.sm[i] = pp[D.o_sm + i];
        s.mdiag[i] = pp[D.o_m + i * K + i];
        s.sizeD[i] = pp[D.o_size + i];
        s.cd[i] = st.cd[(size_t)r * K + i];
        s.c[i] = st.ceff[(size_t)r * K + i];
        s.maxEBM[i] = st.maxEBM[(size_t)r * K + i];
        s.lock[i] = st.lock[(size_t)r * K + i];
    }
    int ovf = 0;
    for (int i = tid; i < K * H; i += nt) {
        long long v = st.I[(size_t)r * K * H + i];
        if (v > 2147483647LL || v < 0) ovf = 1;
        s.I[i] = (int)v;
    }
    for (int i = tid; i < K * S; i += nt) {
        long long v = st.Sx[(size_t)r * K * S + i];
        if (v > 2147483647LL || v < 0) ovf = 1;
        s.Sx[i] = (int)v;
    }
    if (tid < 8) s.flags[tid] = 0;
    __syncthreads();
    if (ovf) atomicOr(&s.flags[3], 1);
    __syncthreads();
}

__device__ __forceinline__ long long block_sum_ll(long long v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    long long *r64 = reinterpret_cast<long long *>(red);
    int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) r64[w] = v;
    __syncthreads();
    v = 0;
    for (int i = 0; i < nw; i++) v += r64[i];
    __syncthreads();
    return v;
}

// per-deme infectious totals into s.tot[]; returns (uniformly) whether anyone is infectious
__device__ __forceinline__ int deme_totals(const Dims &D, const TauShared &s) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int any_inf = 0;
    for (int p = tid >> 5; p < D.K; p += nt >> 5) {
        int tot = 0;
        for (int h = tid & 31; h < D.H; h += 32) tot += s.I[p * D.H + h];
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (tot != 0) any_inf = 1;
        if ((tid & 31) == 0) s.tot[p] = tot;
    }
    return __syncthreads_or(any_inf);
}

// CheckLockdown for every deme (:2328-2329 / :449-450 / :736-737) by thread 0, then the rate refresh
__device__ __forceinline__ void lockdown_pass(const DevState &st, int r, const Dims &D, const TauShared &s,
                                              const double *pp, double *eff, double now) {
    if (threadIdx.x == 0) {
        int flips = 0;
        for (int p = 0; p < D.K; p++)
            flips += check_lockdown(D, pp, p, (long long)s.tot[p], s.cd, s.lock, now, &st.loc_n[r],
                                    st.loc_sp + (size_t)r * st.loc_cap, st.loc_t + (size_t)r * st.loc_cap, st.loc_cap,
                                    &st.err[r]);
        s.flags[2] = flips;
        s.flags[5] += flips;
    }
    __syncthreads();
    if (s.flags[2]) update_contact_rates(BlockGroup(), D, pp, s.cd, eff, s.c, s.maxEBM);
}

__global__ void __launch_bounds__(256) tau_kernel(DevState st, SimArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Dims D = st.D;
    const int K = D.K, H = D.H, S = D.S;
    const int tid = threadIdx.x, nt = blockDim.x;
    TauShared s;
    carve(s, D, smem_raw);

    for (int r = blockIdx.x; r < st.R; r += gridDim.x) {
        const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
        double *eff = st.eff + (size_t)r * K * K;
        long long *ctr = st.counters + (size_t)r * NCOUNT;
        const uint64_t seed = st.seeds[r];
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        __syncthreads();
        load_replicate(st, r, D, s, pp);
        if (s.flags[3]) {
            if (tid == 0) st.err[r] |= ERR_COUNT_OVERFLOW;
            continue;
        }

        // per-thread event-type tallies; reduced into the global counters at the end of the run
        long long accB = 0, accD = 0, accS = 0, accM = 0, accI = 0, accG = 0;
        long long base[6];
#pragma unroll
        for (int j = 0; j < 6; j++) base[j] = ctr[j];  // C_B..C_MIGP carried over from earlier calls
        long long sC = ctr[C_S];
        long long evptr = ctr[C_EVPTR], leaps = ctr[C_LEAPS];
        double t = st.time[r];
        unsigned epoch = st.epoch[r];
        long long good_attempt = ctr[C_GOOD];
        const long long ev_limit = evptr + a.iterations;  // events.ptr < events.size (:2312), intended capacity
        int *tau_counts = st.tau_counts + (size_t)r * st.leap_cap * D.Pp;
        double *tau_tt = st.tau_tt + (size_t)r * st.leap_cap * 2;
        double *ev_time = st.ev_time + (size_t)r * st.ev_cap;
        unsigned long long *ev_desc = st.ev_desc + (size_t)r * st.ev_cap;

        for (long long attempt = 0; attempt < a.attempts; attempt++) {
            epoch++;
            int any_inf = deme_totals(D, s);
            if (any_inf) {
                while (evptr < ev_limit && evptr < st.ev_cap && leaps < st.leap_cap &&
                       (a.sample_size == -1 || sC < a.sample_size) && (!a.has_time || t < (double)a.time)) {
                    double tau = drifts_and_tau(D, s, eff);
                    long long trB, trD, trS, trM, trI, trG;
                    int *row = tau_counts + (size_t)leaps * D.Pp;
                    // ---- draw all channels; halve tau and redraw on an infeasible leap (:2316-2321)
                    for (unsigned retry = 0;; retry++) {
                        for (int i = tid; i < K * H; i += nt) {
                            s.chkI[i] = 0;
                            s.updI[i] = 0;
                        }
                        for (int i = tid; i < K * S; i += nt) s.dSx[i] = 0;
                        __syncthreads();
                        trB = trD = trS = trM = trI = trG = 0;
                        PhiloxCtx ctx;
                        ctx.key = key;
                        ctx.c1 = (uint32_t)leaps;
                        ctx.c2 = (retry & 0xffu) | (epoch << 8);
                        for (int chunk = tid; chunk < D.Pp / 4; chunk += nt) {
                            ctx.c0 = (uint32_t)chunk;
                            int n4[4];
                            uint4 w = make_uint4(0, 0, 0, 0);
                            bool have_w = false;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                Channel ch;
                                decode_channel(chunk * 4 + j, D, s, eff, ch);
                                double lam = ch.prop * tau;
                                int n = 0;
                                if (lam > 0.0) {
                                    if (!have_w) {
                                        w = ctx.draw(0u);
                                        have_w = true;
                                    }
                                    n = (int)poisson_draw(lam, pick_word(w, j), ctx, j);
                                }
                                n4[j] = n;
                                if (n != 0) {
                                    if (ch.i_dec >= 0) {
                                        atomicSub(&s.chkI[ch.i_dec], n);
                                        atomicSub(&s.updI[ch.i_dec], n);
                                    }
                                    if (ch.i_inc >= 0) {
                                        atomicAdd(&s.updI[ch.i_inc], n);
                                        atomicAdd(&s.chkI[ch.i_chk], n);
                                    }
                                    if (ch.s_dec >= 0) atomicSub(&s.dSx[ch.s_dec], n);
                                    if (ch.s_inc >= 0) atomicAdd(&s.dSx[ch.s_inc], n);
                                    if (ch.type == EV_MIGRATION) trG += n;
                                    else if (ch.type == EV_BIRTH) trB += n;
                                    else if (ch.type == EV_DEATH) trD += n;
                                    else if (ch.type == EV_SAMPLING) trS += n;
                                    else if (ch.type == EV_MUTATION) trM += n;
                                    else trI += n;
                                }
                            }
                            reinterpret_cast<int4 *>(row)[chunk] = make_int4(n4[0], n4[1], n4[2], n4[3]);
                        }
                        __syncthreads();
                        // feasibility (:2522-2528)
                        int bad = 0;
                        for (int i = tid; i < K * H; i += nt) {
                            double v = (double)s.I[i] + (double)s.chkI[i];
                            if (v < 0.0 || v > s.sizeD[i >> D.hshift]) bad = 1;
                        }
                        for (int i = tid; i < K * S; i += nt) {
                            double v = (double)s.Sx[i] + (double)s.dSx[i];
                            if (v < 0.0 || v > s.sizeD[i / S]) bad = 1;
                        }
                        bad = __syncthreads_or(bad);
                        if (!bad) break;
                        tau *= 0.5;
                    }
                    // ---- apply (UpdateCompartmentCounts_tau, :2536-2593)
                    for (int i = tid; i < K * H; i += nt) s.I[i] += s.updI[i];
                    for (int i = tid; i < K * S; i += nt) s.Sx[i] += s.dSx[i];
                    accB += trB; accD += trD; accS += trS; accM += trM; accI += trI; accG += trG;
                    t += tau;
                    sC += block_sum_ll(trS, s.red);  // sCounter gates the loop (:2312), so it is kept exact per leap
                    if (tid == 0) {
                        tau_tt[leaps * 2] = t;
                        tau_tt[leaps * 2 + 1] = tau;
                        ev_time[evptr] = t;
                        ev_desc[evptr] = pack_multi((uint32_t)leaps);
                    }
                    leaps++;
                    evptr++;
                    // ---- extinction test and CheckLockdown for every deme (:2326-2329)
                    any_inf = deme_totals(D, s);
                    if (!any_inf) break;
                    lockdown_pass(st, r, D, s, pp, eff, t);
                }
            }
            // ---- extinction-retry (:2331-2335): <= 100 log rows with iterations > 100 => Restart (:714-738)
            if (evptr <= 100 && a.iterations > 100) {
                evptr = 0;
                leaps = 0;
                sC = 0;
                t = 0.0;
                accB = accD = accS = accM = accI = accG = 0;
#pragma unroll
                for (int j = 0; j < 6; j++) base[j] = 0;
                __syncthreads();
                for (int i = tid; i < K * H; i += nt) s.I[i] = (int)st.initI[(size_t)r * K * H + i];
                for (int i = tid; i < K * S; i += nt) s.Sx[i] = (int)st.initSx[(size_t)r * K * S + i];
                __syncthreads();
                deme_totals(D, s);
                lockdown_pass(st, r, D, s, pp, eff, t);
                good_attempt = 0;
                if (tid == 0) ctr[C_MIGN] = 0;
            } else {
                good_attempt = attempt + 1;
                break;
            }
        }

        // ---- commit the replicate back to HBM
        __syncthreads();
        for (int i = tid; i < K * H; i += nt) st.I[(size_t)r * K * H + i] = s.I[i];
        for (int i = tid; i < K * S; i += nt) st.Sx[(size_t)r * K * S + i] = s.Sx[i];
        for (int i = tid; i < K; i += nt) {
            st.cd[(size_t)r * K + i] = s.cd[i];
            st.ceff[(size_t)r * K + i] = s.c[i];
            st.maxEBM[(size_t)r * K + i] = s.maxEBM[i];
            st.lock[(size_t)r * K + i] = s.lock[i];
        }
        long long ginf = 0;
        for (int i = tid; i < K * H; i += nt) ginf += s.I[i];
        ginf = block_sum_ll(ginf, s.red);
        accB = block_sum_ll(accB, s.red);
        accD = block_sum_ll(accD, s.red);
        accM = block_sum_ll(accM, s.red);
        accI = block_sum_ll(accI, s.red);
        accG = block_sum_ll(accG, s.red);
        if (tid == 0) {
            ctr[C_B] = base[C_B] + accB;
            ctr[C_D] = base[C_D] + accD;
            ctr[C_S] = sC;
            ctr[C_M] = base[C_M] + accM;
            ctr[C_I] = base[C_I] + accI;
            ctr[C_MIGP] = base[C_MIGP] + accG;
            ctr[C_SWAP] += s.flags[5];
            ctr[C_GOOD] = good_attempt;
            ctr[C_EVPTR] = evptr;
            ctr[C_LEAPS] = leaps;
            ctr[C_GINF] = ginf;
            st.time[r] = t;
            st.epoch[r] = epoch;
        }
        __syncthreads();
    }
}

// Deterministic parity tap: propensities of the current state in positional order, drifts and tau.
__global__ void __launch_bounds__(256) propensity_kernel(DevState st, int r, double *out, double *dI, double *dS,
                                                         double *tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Dims D = st.D;
    TauShared s;
    carve(s, D, smem_raw);
    const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
    const double *eff = st.eff + (size_t)r * D.K * D.K;
    load_replicate(st, r, D, s, pp);
    double tau = drifts_and_tau(D, s, eff);
    for (int c = threadIdx.x; c < D.P; c += blockDim.x) {
        Channel ch;
        decode_channel(c, D, s, eff, ch);
        out[c] = ch.prop;
    }
    for (int i = threadIdx.x; i < D.K * D.H; i += blockDim.x) dI[i] = s.dI[i];
    for (int i = threadIdx.x; i < D.K * D.S; i += blockDim.x) dS[i] = s.dS[i];
    if (threadIdx.x == 0) *tau_out = tau;
}

// host launchers ---------------------------------------------------------------------------------
cudaError_t launch_tau(const DevState &st, const SimArgs &a, cudaStream_t stream, int num_sms) {
    size_t smem = tau_smem_bytes(st.D);
    cudaError_t e = cudaFuncSetAttribute(tau_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tau_kernel, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int grid = num_sms * per_sm;
    if (grid > st.R) grid = st.R;
    tau_kernel<<<grid, 256, smem, stream>>>(st, a);
    return cudaGetLastError();
}

cudaError_t launch_propensities(const DevState &st, int r, double *out, double *dI, double *dS, double *tau,
                                cudaStream_t stream) {
    size_t smem = tau_smem_bytes(st.D);
    cudaError_t e = cudaFuncSetAttribute(propensity_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    propensity_kernel<<<1, 256, smem, stream>>>(st, r, out, dI, dS, tau);
    return cudaGetLastError();
}

}  // namespace vg
