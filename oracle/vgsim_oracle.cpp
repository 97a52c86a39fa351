// TEST INFRASTRUCTURE ONLY — CPU oracle for the vgsim_b200 hot path.
//
// A sequential C++ restatement of the reference engine's algorithm (Genomics-HSE/VGsim,
// `cdef class BirthDeathModel`), written from the behavioural spec in SURVEY.md App. A and the
// reference files cited per function below (paths relative to /root/reference).  It exists so that
// tests/ (and bench.py's cpu_baseline / --impl reference legs, and __graft_entry__.smoke()) can
// check the CUDA path; the product path (vgsim_b200/) never imports, links or executes it.
//
// Parity status: PINNED.  With the same (seed, attempt) stream this oracle reproduces the
// out-of-tree build of the unmodified reference (oracle/build_ref.py -> oracle/_ref) bit for bit:
// event chains of the nine testing/check_simulator.py scenarios, tau-leap logs, parent arrays,
// mutation and migration tables (tests/golden/*, tests/test_oracle_vs_reference.py).  Seed-for-seed
// parity with *upstream binary wheels* is unpinned only in so far as the mc_lib seeding rule
// (np_random.h) could not be checked against upstream mc_lib offline.
//
// Arithmetic is kept in the reference's evaluation order (left-to-right products, sequential
// sums) and this file must be compiled with -ffp-contract=off so no FMA contraction changes bits.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "np_random.h"

using std::vector;
typedef int64_t i64;

namespace vgo {

enum { BIRTH = 0, DEATH = 1, SAMPLING = 2, MUTATION = 3, SUSCCHANGE = 4, MIGRATION = 5, MULTITYPE = 6 };  // src/events.pxi:2-8

struct EventLog {  // src/events.pxi:24-68
    i64 size = 0, ptr = 0;
    vector<double> t;
    vector<i64> type, hap, pop, nhap, npop;
    void resize(i64 n) {
        t.resize(n, 0.0);
        type.resize(n, 0);
        hap.resize(n, 0);
        pop.resize(n, 0);
        nhap.resize(n, 0);
        npop.resize(n, 0);
    }
    void create(i64 iterations) {  // CreateEvents, src/events.pxi:52-68
        if (ptr == 0) {
            size += iterations;
            resize(size);
        } else if (iterations + ptr - size > 0) {
            size = iterations + ptr;
            resize(size);
        }
    }
    void add(double time, i64 ty, i64 h, i64 p, i64 nh, i64 np) {
        if (ptr >= (i64)t.size()) resize(ptr + 1 + ptr / 2);
        t[ptr] = time;
        type[ptr] = ty;
        hap[ptr] = h;
        pop[ptr] = p;
        nhap[ptr] = nh;
        npop[ptr] = np;
        ptr++;
    }
};

struct MultiLog {  // src/events.pxi:105-152
    i64 size = 0, ptr = 0;
    bool overflow = false;  // reference would write out of bounds here
    vector<i64> num, type, hap, pop, nhap, npop;
    vector<double> t;
    void resize(i64 n) {
        num.resize(n, 0);
        t.resize(n, 0.0);
        type.resize(n, 0);
        hap.resize(n, 0);
        pop.resize(n, 0);
        nhap.resize(n, 0);
        npop.resize(n, 0);
    }
    void create(i64 n) {
        if (ptr == 0) {
            size = n;
        } else {
            size = n + ptr;
        }
        resize(size);
    }
    void add(i64 n, double time, i64 ty, i64 h, i64 p, i64 nh, i64 np) {
        if (ptr >= (i64)num.size()) {
            overflow = true;
            resize(ptr + 1 + ptr / 2);
        }
        num[ptr] = n;
        t[ptr] = time;
        type[ptr] = ty;
        hap[ptr] = h;
        pop[ptr] = p;
        nhap[ptr] = nh;
        npop[ptr] = np;
        ptr++;
    }
};

struct Model {
    int U, K, S, H;
    uint64_t user_seed;
    Pcg64 rng;
    bool first_simulation = false;
    int error = 0;  // 1 = zero weight sampled in choose() (reference: sys.exit(1), src/fast_choose.pxi:6-13)

    i64 bC = 0, dC = 0, sC = 0, mC = 0, iC = 0, swapLockdown = 0, migPlus = 0, migNonPlus = 0, globalInf = 0,
        good_attempt = 0;
    double currentTime = 0.0, totalRate = 0.0, totalMigrationRate = 0.0, rn = 0.0, tau_l = 0.01;

    EventLog ev;
    MultiLog mev;

    // parameters (src/_BirthDeath.pyx:157-199)
    vector<double> b, d, s, mRate, hapMutType, sigma, T, m, cd, cdBefore, cdAfter, startLD, endLD, sm;
    vector<i64> suscType, sizes, lockdownON;
    // state
    vector<i64> Sx, I, initSx, initI, totSus, totInf;
    // derived (src/_BirthDeath.pyx:150-156,182-199)
    vector<double> tm, tEv, hp, evr, shp, maxEBM, Tc, immPop, infPop, popRate, migPop, A, immSrc, eff;
    // tau (src/_BirthDeath.pyx:209-229)
    vector<double> pMigr, pSus, pRec, pSamp, pMut, pTr, dI, dS;
    vector<i64> eMigr, eSus, eRec, eSamp, eMut, eTr, deltaI, deltaS;
    // genealogy outputs
    vector<i64> tree, tree_pop;
    vector<double> times;
    vector<i64> mutNode, mutAS, mutDS, mutSite;
    vector<double> mutTime;
    vector<i64> migNode, migOld, migNew;
    vector<double> migTime;
    vector<i64> locState, locPop;
    vector<double> locTime;
    // injected uniform stream for the genealogy (parity mode of the CUDA kernel): when set,
    // genealogy uniforms are read from here instead of the PCG64 stream.
    const double *inj = nullptr;
    i64 inj_n = 0, inj_pos = 0;
    i64 clamped = 0;  // tau-log BIRTH records that asked for more coalescences than lineages allow

    Model(int sites, int K_, int S_, uint64_t seed) : U(sites), K(K_), S(S_), user_seed(seed) {
        H = 1;
        for (int i = 0; i < U; i++) H *= 4;
        seed_rndm_wrapper(rng, user_seed, 0);  // :74
        b.assign(H, 2.0);
        d.assign(H, 1.0);
        s.assign(H, 0.01);
        mRate.assign((size_t)H * std::max(U, 1), 0.01);
        hapMutType.assign((size_t)H * std::max(U, 1) * 3, 1.0);
        sigma.assign((size_t)H * S, 0.0);
        for (int h = 0; h < H; h++) sigma[(size_t)h * S] = 1.0;
        suscType.assign(H, 0);
        T.assign((size_t)S * S, 0.0);
        m.assign((size_t)K * K, 0.0);
        cd.assign(K, 1.0);
        cdBefore.assign(K, 1.0);
        cdAfter.assign(K, 0.0);
        startLD.assign(K, 1.0);
        endLD.assign(K, 1.0);
        sm.assign(K, 1.0);
        sizes.assign(K, 1000000);
        lockdownON.assign(K, 0);
        Sx.assign((size_t)K * S, 0);
        for (int p = 0; p < K; p++) Sx[(size_t)p * S] = 1000000;
        I.assign((size_t)K * H, 0);
        initSx.assign((size_t)K * S, 0);
        initI.assign((size_t)K * H, 0);
        totSus.assign(K, 1000000);
        totInf.assign(K, 0);
        tm.assign(H, 0);
        tEv.assign((size_t)K * H, 0);
        hp.assign((size_t)K * H, 0);
        evr.assign((size_t)K * H * 4, 0);
        shp.assign((size_t)K * H * S, 0);
        maxEBM.assign(K, 0);
        Tc.assign(S, 0);
        immPop.assign(K, 0);
        infPop.assign(K, 0);
        popRate.assign(K, 0);
        migPop.assign(K, 0);
        A.assign(K, 0);
        immSrc.assign((size_t)K * S, 0);
        eff.assign((size_t)K * K, 0);
        pMigr.assign((size_t)K * K * S * H, 0);
        pSus.assign((size_t)K * S * S, 0);
        pRec.assign((size_t)K * H, 0);
        pSamp.assign((size_t)K * H, 0);
        pMut.assign((size_t)K * H * std::max(U, 1) * 3, 0);
        pTr.assign((size_t)K * H * S, 0);
        dI.assign((size_t)K * H, 0);
        dS.assign((size_t)K * S, 0);
        eMigr.assign(pMigr.size(), 0);
        eSus.assign(pSus.size(), 0);
        eRec.assign(pRec.size(), 0);
        eSamp.assign(pSamp.size(), 0);
        eMut.assign(pMut.size(), 0);
        eTr.assign(pTr.size(), 0);
        deltaI.assign((size_t)K * H, 0);
        deltaS.assign((size_t)K * S, 0);
        tree.assign(1, 0);
        tree_pop.assign(1, 0);
    }

    i64 propNum() const {  // :2301
        return (i64)K * ((i64)(K - 1) * H * S + (i64)S * (S - 1) + (i64)H * (2 + U * 3 + S));
    }

    // ---- fastChoose / fastChoose_skip, src/fast_choose.pxi:18-52
    template <class W>
    i64 choose(const W *w, i64 n, W tw, double &r) {
        double x = tw * r;
        i64 i = 0;
        W total = w[0];
        while (total < x && i < n - 1) {
            i++;
            total += w[i];
        }
        if (w[i] == 0.0) error = 1;
        r = (x - (total - w[i])) / w[i];
        return i;
    }
    template <class W>
    i64 choose_skip(const W *w, i64 n, W tw, double &r, i64 skip) {
        double x = tw * r;
        i64 i = 0;
        if (skip == 0) i++;
        W total = w[i];
        while (total < x && i < n - 1) {
            i++;
            if (i != skip) total += w[i];
        }
        if (w[i] == 0.0) error = 1;
        r = (x - (total - w[i])) / w[i];
        return i;
    }

    // ---- compartment bookkeeping, :246-260
    void NewInfections(i64 p, i64 si, i64 h, i64 num = 1) {
        Sx[p * S + si] -= num;
        totSus[p] -= num;
        I[p * H + h] += num;
        totInf[p] += num;
        globalInf += num;
    }
    void NewRecoveries(i64 p, i64 si, i64 h, i64 num = 1) {
        Sx[p * S + si] += num;
        totSus[p] += num;
        I[p * H + h] -= num;
        totInf[p] -= num;
        globalInf -= num;
    }
    void FirstInfection() {  // :234-242
        if (globalInf == 0) {
            for (int sn = 0; sn < S; sn++) {
                if (Sx[sn] == 0) continue;
                NewInfections(0, sn, 0);
                return;
            }
        }
    }

    // ---- BirthRate, :382-392 (also refreshes susceptHapPopRate)
    double BirthRate(i64 p, i64 h) {
        double ps = 0.0;
        for (int sn = 0; sn < S; sn++) {
            double v = Sx[p * S + sn] * sigma[h * S + sn];
            shp[(p * H + h) * S + sn] = v;
            for (int q = 0; q < K; q++) ps += v * m[p * K + q] * m[p * K + q] * cd[q] / A[q];
        }
        return b[h] * ps;
    }

    // ---- UpdateAllRates, :279-351
    void UpdateAllRates() {
        for (int s1 = 0; s1 < S; s1++) {
            Tc[s1] = 0;
            for (int s2 = 0; s2 < S; s2++) Tc[s1] += T[s1 * S + s2];
        }
        for (int p1 = 0; p1 < K; p1++) {
            m[p1 * K + p1] = 1.0;
            A[p1] = 0.0;
            for (int p2 = 0; p2 < K; p2++) {
                if (p1 == p2) continue;
                m[p1 * K + p1] -= m[p1 * K + p2];
                A[p1] += m[p2 * K + p1] * sizes[p2];
            }
            A[p1] += m[p1 * K + p1] * sizes[p1];
        }
        totalRate = 0.0;
        for (int p = 0; p < K; p++) {
            infPop[p] = 0;
            immPop[p] = 0;
            popRate[p] = 0.;
        }
        for (int p = 0; p < K; p++) {
            for (int h = 0; h < H; h++) {
                tm[h] = 0;
                for (int u = 0; u < U; u++) tm[h] += mRate[h * U + u];
                double *e = &evr[((size_t)p * H + h) * 4];
                e[0] = BirthRate(p, h);
                e[1] = d[h];
                e[2] = s[h] * sm[p];
                e[3] = tm[h];
                tEv[p * H + h] = 0;
                for (int i = 0; i < 4; i++) tEv[p * H + h] += e[i];
                hp[p * H + h] = tEv[p * H + h] * I[p * H + h];
                infPop[p] += hp[p * H + h];
            }
            for (int sn = 0; sn < S; sn++) {
                immSrc[p * S + sn] = Tc[sn] * Sx[p * S + sn];
                immPop[p] += immSrc[p * S + sn];
            }
            popRate[p] = infPop[p] + immPop[p];
            totalRate += popRate[p];
        }
        vector<double> maxEff(K, 0.0);
        for (int p1 = 0; p1 < K; p1++)
            for (int p2 = 0; p2 < K; p2++) {
                if (p1 == p2) continue;
                double e = 0.0;
                for (int p3 = 0; p3 < K; p3++) e += m[p1 * K + p3] * m[p2 * K + p3] * cd[p3] / A[p3];
                eff[p1 * K + p2] = e;
                if (e > maxEff[p2]) maxEff[p2] = e;
            }
        double maxB = 0.0;
        for (int h = 0; h < H; h++)
            for (int sn = 0; sn < S; sn++)
                if (b[h] * sigma[h * S + sn] > maxB) maxB = b[h] * sigma[h * S + sn];
        totalMigrationRate = 0.0;
        for (int p = 0; p < K; p++) {
            maxEBM[p] = maxEff[p] * maxB;
            migPop[p] = maxEBM[p] * totSus[p] * (globalInf - totInf[p]);
            totalMigrationRate += migPop[p];
        }
    }

    // ---- UpdateRates, :516-546
    void UpdateRates(i64 p, bool infect, bool immune, bool migration) {
        if (infect) {
            infPop[p] = 0.0;
            for (int h = 0; h < H; h++) {
                double *e = &evr[((size_t)p * H + h) * 4];
                e[0] = BirthRate(p, h);
                double tmp = (e[0] + e[1] + e[2] + e[3]);
                tEv[p * H + h] = tmp;
                hp[p * H + h] = tEv[p * H + h] * I[p * H + h];
                infPop[p] += hp[p * H + h];
            }
        }
        if (immune) {
            immPop[p] = 0;
            for (int sn = 0; sn < S; sn++) immPop[p] += immSrc[p * S + sn];
        }
        if (infect || immune) {
            popRate[p] = infPop[p] + immPop[p];
            totalRate = 0.0;
            for (int q = 0; q < K; q++) totalRate += popRate[q];
        }
        if (migration) {
            totalMigrationRate = 0.0;
            for (int q = 0; q < K; q++) {
                migPop[q] = maxEBM[q] * totSus[q] * (globalInf - totInf[q]);
                totalMigrationRate += migPop[q];
            }
        }
    }

    // ---- CheckLockdown, :698-710
    void CheckLockdown(i64 p) {
        if (totInf[p] > startLD[p] * sizes[p] && lockdownON[p] == 0) {
            cd[p] = cdAfter[p];
            swapLockdown++;
            lockdownON[p] = 1;
            UpdateAllRates();
            locState.push_back(1);
            locPop.push_back(p);
            locTime.push_back(currentTime);
        }
        if (totInf[p] < endLD[p] * sizes[p] && lockdownON[p] == 1) {
            cd[p] = cdBefore[p];
            swapLockdown++;
            lockdownON[p] = 0;
            UpdateAllRates();
            locState.push_back(0);
            locPop.push_back(p);
            locTime.push_back(currentTime);
        }
    }

    // ---- Mutate, :2420-2427
    i64 Mutate(i64 h, i64 site, i64 DS) const {
        i64 digit4 = 1;
        for (int i = 0; i < U - site - 1; i++) digit4 *= 4;
        i64 AS = (h / digit4) % 4;
        if (DS >= AS) DS += 1;
        return h + (DS - AS) * digit4;
    }

    // ---- direct-method event handlers, :550-694
    void ImmunityTransition(i64 p) {
        i64 ssi = choose(&immSrc[p * S], S, immPop[p], rn);
        i64 tsi = choose(&T[ssi * S], S, Tc[ssi], rn);
        Sx[p * S + ssi] -= 1;
        Sx[p * S + tsi] += 1;
        immSrc[p * S + ssi] = Sx[p * S + ssi] * Tc[ssi];
        immSrc[p * S + tsi] = Sx[p * S + tsi] * Tc[tsi];
        UpdateRates(p, false, true, false);
        iC++;
        ev.add(currentTime, SUSCCHANGE, ssi, p, tsi, 0);
    }
    void Birth(i64 p, i64 h) {
        double ws = 0.0;
        for (int sn = 0; sn < S; sn++) ws += shp[(p * H + h) * S + sn];
        i64 si = choose(&shp[(p * H + h) * S], S, ws, rn);
        // recombination (:575-596) is out of scope: probability fixed at 0.0
        NewInfections(p, si, h);
        ev.add(currentTime, BIRTH, h, p, si, H);
        immSrc[p * S + si] = Tc[si] * Sx[p * S + si];
        UpdateRates(p, true, true, true);
        bC++;
    }
    void Death(i64 p, i64 h, bool add_event = true) {
        i64 st = suscType[h];
        NewRecoveries(p, st, h);
        immSrc[p * S + st] = Sx[p * S + st] * Tc[st];
        UpdateRates(p, true, true, true);
        if (add_event) {
            dC++;
            ev.add(currentTime, DEATH, h, p, st, 0);
        }
    }
    void Sampling(i64 p, i64 h) {
        Death(p, h, false);
        sC++;
        ev.add(currentTime, SAMPLING, h, p, suscType[h], 0);
    }
    void Mutation(i64 p, i64 h) {
        i64 mi = choose(&mRate[h * U], U, tm[h], rn);
        const double *w = &hapMutType[(h * U + mi) * 3];
        i64 DS = choose(w, 3, w[0] + w[1] + w[2], rn);
        i64 nh = Mutate(h, mi, DS);
        I[p * H + nh] += 1;
        I[p * H + h] -= 1;
        UpdateRates(p, true, false, false);
        mC++;
        ev.add(currentTime, MUTATION, h, p, nh, 0);
    }
    i64 GenerateMigration() {
        i64 tp = choose(migPop.data(), K, totalMigrationRate, rn);
        i64 sp = choose_skip(totInf.data(), K, globalInf - totInf[tp], rn, tp);
        i64 h = choose(&I[sp * H], H, totInf[sp], rn);
        i64 si = choose(&Sx[tp * S], S, totSus[tp], rn);
        double p_accept = eff[sp * K + tp] * b[h] * sigma[h * S + si] / maxEBM[tp];
        if (rn < p_accept) {
            NewInfections(tp, si, h);
            UpdateRates(tp, true, true, true);
            migPlus++;
            ev.add(currentTime, MIGRATION, h, sp, si, tp);
        } else {
            migNonPlus++;
        }
        return tp;
    }
    i64 GenerateEvent() {  // :483-512
        i64 p;
        rn = rng.next_double();
        double ch = rn * (totalRate + totalMigrationRate);
        if (totalRate > ch) {
            rn = ch / totalRate;
            p = choose(popRate.data(), K, totalRate, rn);
            ch = rn * popRate[p];
            if (immPop[p] > ch) {
                rn = ch / immPop[p];
                ImmunityTransition(p);
            } else {
                rn = (ch - immPop[p]) / infPop[p];
                i64 h = choose(&hp[p * H], H, infPop[p], rn);
                i64 e = choose(&evr[((size_t)p * H + h) * 4], 4, tEv[p * H + h], rn);
                if (e == BIRTH)
                    Birth(p, h);
                else if (e == DEATH)
                    Death(p, h);
                else if (e == SAMPLING)
                    Sampling(p, h);
                else
                    Mutation(p, h);
            }
        } else {
            rn = (ch - totalRate) / totalMigrationRate;
            p = GenerateMigration();
        }
        return p;
    }

    void Restart() {  // :714-738
        ev.ptr = 0;
        mev.ptr = 0;
        bC = dC = sC = mC = iC = migPlus = migNonPlus = 0;
        currentTime = 0.0;
        globalInf = 0;
        for (int p = 0; p < K; p++) {
            totSus[p] = 0;
            totInf[p] = 0;
            for (int sn = 0; sn < S; sn++) {
                Sx[p * S + sn] = initSx[p * S + sn];
                totSus[p] += initSx[p * S + sn];
            }
            for (int h = 0; h < H; h++) {
                I[p * H + h] = initI[p * H + h];
                totInf[p] += initI[p * H + h];
                globalInf += initI[p * H + h];
            }
        }
        for (int p = 0; p < K; p++) CheckLockdown(p);
        UpdateAllRates();
    }

    void PrepareParameters(i64 iterations) {  // :433-451
        ev.create(iterations);
        if (!first_simulation) {
            FirstInfection();
            globalInf = 0;
            for (int p = 0; p < K; p++) {
                totSus[p] = 0;
                for (int sn = 0; sn < S; sn++) {
                    initSx[p * S + sn] = Sx[p * S + sn];
                    totSus[p] += Sx[p * S + sn];
                }
                totInf[p] = 0;
                for (int h = 0; h < H; h++) {
                    initI[p * H + h] = I[p * H + h];
                    totInf[p] += I[p * H + h];
                    globalInf += I[p * H + h];
                }
            }
            first_simulation = true;
        }
        for (int p = 0; p < K; p++) CheckLockdown(p);
        UpdateAllRates();
    }

    // ---- SimulatePopulation, :396-429.  `time` is a C float (quirk Q1).
    void SimulateDirect(i64 iterations, i64 sample_size, float time, i64 attempts) {
        PrepareParameters(iterations);
        for (i64 i = 0; i < attempts; i++) {
            seed_rndm_wrapper(rng, user_seed, (uint32_t)i);
            if (totalRate + totalMigrationRate != 0.0 && globalInf != 0) {
                while (ev.ptr < ev.size && (sample_size == -1 || sC <= sample_size) &&
                       (time == -1 || currentTime < time)) {
                    double tau = -std::log(rng.next_double()) / (totalRate + totalMigrationRate);  // SampleTime :476-478
                    currentTime += tau;
                    i64 p = GenerateEvent();
                    if (error) return;
                    if (totalRate == 0.0 || globalInf == 0) break;
                    CheckLockdown(p);
                }
            }
            if (ev.ptr <= 100 && iterations > 100) {
                Restart();
            } else {
                good_attempt = i + 1;
                break;
            }
        }
    }

    // ---- Propensities, :2351-2417
    void Propensities() {
        std::fill(dI.begin(), dI.end(), 0.0);
        std::fill(dS.begin(), dS.end(), 0.0);
        for (int sp = 0; sp < K; sp++)
            for (int tp = 0; tp < K; tp++) {
                if (sp == tp) continue;
                for (int sn = 0; sn < S; sn++)
                    for (int h = 0; h < H; h++) {
                        double v = eff[tp * K + sp] * Sx[tp * S + sn] * I[sp * H + h] * b[h] * sigma[h * S + sn] *
                                   m[sp * K + sp];
                        pMigr[(((size_t)sp * K + tp) * S + sn) * H + h] = v;
                        dI[tp * H + h] += v;
                        dS[tp * S + sn] -= v;
                    }
            }
        for (int p = 0; p < K; p++) {
            for (int ss = 0; ss < S; ss++)
                for (int ts = 0; ts < S; ts++) {
                    if (ss == ts) continue;
                    double v = T[ss * S + ts] * Sx[p * S + ss];
                    pSus[((size_t)p * S + ss) * S + ts] = v;
                    dS[p * S + ts] += v;
                    dS[p * S + ss] -= v;
                }
            for (int h = 0; h < H; h++) {
                double v = d[h] * I[p * H + h];
                pRec[p * H + h] = v;
                dS[p * S + suscType[h]] += v;
                dI[p * H + h] -= v;
                v = s[h] * I[p * H + h] * sm[p];
                pSamp[p * H + h] = v;
                dS[p * S + suscType[h]] += v;
                dI[p * H + h] -= v;
                for (int u = 0; u < U; u++)
                    for (int i = 0; i < 3; i++) {
                        const double *w = &hapMutType[(h * U + u) * 3];
                        v = mRate[h * U + u] * w[i] / (w[0] + w[1] + w[2]) * I[p * H + h];
                        pMut[(((size_t)p * H + h) * U + u) * 3 + i] = v;
                        dI[p * H + Mutate(h, u, i)] += v;
                        dI[p * H + h] -= v;
                    }
            }
        }
        for (int tp = 0; tp < K; tp++)
            for (int h = 0; h < H; h++)
                for (int sn = 0; sn < S; sn++) {
                    double v = 0.0;
                    for (int sp = 0; sp < K; sp++)
                        v += b[h] * sigma[h * S + sn] * m[tp * K + sp] * m[tp * K + sp] * cd[sp] * Sx[tp * S + sn] *
                             I[tp * H + h] / A[sp];
                    pTr[((size_t)tp * H + h) * S + sn] = v;
                    dI[tp * H + h] += v;
                    dS[tp * S + sn] -= v;
                }
    }

    // ---- ChooseTau, :2432-2450.  epsilon is a C float and `epsilon * count` is a FLOAT product
    // (generated C: (double)(float_eps * int64) / 2.0) — reproduced on purpose.
    void ChooseTau() {
        const float epsilon = 0.03f;
        tau_l = 1.0;
        for (int p = 0; p < K; p++) {
            for (int h = 0; h < H; h++) {
                if (std::fabs(dI[p * H + h]) < 1e-8) continue;
                double x = ((double)(epsilon * I[p * H + h])) / 2.0;
                double tmp = (1.0 > x ? 1.0 : x) / std::fabs(dI[p * H + h]);
                if (tmp < tau_l) tau_l = tmp;
            }
            for (int sn = 0; sn < S; sn++) {
                if (std::fabs(dS[p * S + sn]) < 1e-8) continue;
                double x = ((double)(epsilon * Sx[p * S + sn])) / 2.0;
                double tmp = (1.0 > x ? 1.0 : x) / std::fabs(dS[p * S + sn]);
                if (tmp < tau_l) tau_l = tmp;
            }
        }
    }

    // ---- GenerateEvents_tau, :2454-2529 (note quirk Q8: migration arrivals booked on the SOURCE deme)
    bool GenerateEvents_tau() {
        std::fill(deltaS.begin(), deltaS.end(), 0);
        std::fill(deltaI.begin(), deltaI.end(), 0);
        for (int sp = 0; sp < K; sp++)
            for (int tp = 0; tp < K; tp++) {
                if (sp == tp) continue;
                for (int sn = 0; sn < S; sn++)
                    for (int h = 0; h < H; h++) {
                        size_t ix = (((size_t)sp * K + tp) * S + sn) * H + h;
                        i64 n = np_poisson(rng, pMigr[ix] * tau_l);
                        eMigr[ix] = n;
                        deltaI[sp * H + h] += n;
                        deltaS[tp * S + sn] -= n;
                    }
            }
        for (int p = 0; p < K; p++) {
            for (int ss = 0; ss < S; ss++)
                for (int ts = 0; ts < S; ts++) {
                    if (ss == ts) continue;
                    size_t ix = ((size_t)p * S + ss) * S + ts;
                    i64 n = np_poisson(rng, pSus[ix] * tau_l);
                    eSus[ix] = n;
                    deltaS[p * S + ts] += n;
                    deltaS[p * S + ss] -= n;
                }
            for (int h = 0; h < H; h++) {
                i64 n = np_poisson(rng, pRec[p * H + h] * tau_l);
                eRec[p * H + h] = n;
                deltaS[p * S + suscType[h]] += n;
                deltaI[p * H + h] -= n;
                n = np_poisson(rng, pSamp[p * H + h] * tau_l);
                eSamp[p * H + h] = n;
                deltaS[p * S + suscType[h]] += n;
                deltaI[p * H + h] -= n;
                for (int u = 0; u < U; u++)
                    for (int i = 0; i < 3; i++) {
                        size_t ix = (((size_t)p * H + h) * U + u) * 3 + i;
                        n = np_poisson(rng, pMut[ix] * tau_l);
                        eMut[ix] = n;
                        deltaI[p * H + Mutate(h, u, i)] += n;
                        deltaI[p * H + h] -= n;
                    }
                for (int sn = 0; sn < S; sn++) {
                    size_t ix = ((size_t)p * H + h) * S + sn;
                    n = np_poisson(rng, pTr[ix] * tau_l);
                    eTr[ix] = n;
                    deltaI[p * H + h] += n;
                    deltaS[p * S + sn] -= n;
                }
            }
        }
        for (int p = 0; p < K; p++) {
            for (int sn = 0; sn < S; sn++) {
                i64 v = deltaS[p * S + sn] + Sx[p * S + sn];
                if (v < 0 || v > sizes[p]) return false;
            }
            for (int h = 0; h < H; h++) {
                i64 v = deltaI[p * H + h] + I[p * H + h];
                if (v < 0 || v > sizes[p]) return false;
            }
        }
        return true;
    }

    // ---- UpdateCompartmentCounts_tau, :2536-2593
    void UpdateCompartmentCounts_tau() {
        for (int sp = 0; sp < K; sp++)
            for (int tp = 0; tp < K; tp++) {
                if (sp == tp) continue;
                for (int sn = 0; sn < S; sn++)
                    for (int h = 0; h < H; h++) {
                        i64 n = eMigr[(((size_t)sp * K + tp) * S + sn) * H + h];
                        NewInfections(tp, sn, h, n);
                        mev.add(n, currentTime, MIGRATION, h, sp, sn, tp);
                        migPlus += n;
                    }
            }
        for (int p = 0; p < K; p++) {
            for (int ss = 0; ss < S; ss++)
                for (int ts = 0; ts < S; ts++) {
                    if (ss == ts) continue;
                    i64 n = eSus[((size_t)p * S + ss) * S + ts];
                    Sx[p * S + ts] += n;
                    Sx[p * S + ss] -= n;
                    mev.add(n, currentTime, SUSCCHANGE, ss, p, ts, 0);
                    iC += n;
                }
            for (int h = 0; h < H; h++) {
                i64 n = eRec[p * H + h];
                NewRecoveries(p, suscType[h], h, n);
                mev.add(n, currentTime, DEATH, h, p, suscType[h], 0);
                dC += n;
                n = eSamp[p * H + h];
                NewRecoveries(p, suscType[h], h, n);
                mev.add(n, currentTime, SAMPLING, h, p, suscType[h], 0);
                sC += n;
                for (int u = 0; u < U; u++)
                    for (int i = 0; i < 3; i++) {
                        i64 nh = Mutate(h, u, i);
                        n = eMut[(((size_t)p * H + h) * U + u) * 3 + i];
                        I[p * H + nh] += n;
                        I[p * H + h] -= n;
                        mev.add(n, currentTime, MUTATION, h, p, nh, 0);
                        mC += n;
                    }
                for (int sn = 0; sn < S; sn++) {
                    n = eTr[((size_t)p * H + h) * S + sn];
                    NewInfections(p, sn, h, n);
                    mev.add(n, currentTime, BIRTH, h, p, sn, 0);
                    bC += n;
                }
            }
        }
    }

    // ---- SimulatePopulation_tau, :2293-2346
    void SimulateTau(i64 iterations, i64 sample_size, float time, i64 attempts) {
        PrepareParameters(iterations);
        i64 P = propNum();
        if (globalInf == 0) FirstInfection();
        mev.create(iterations * P);
        ev.create(iterations);  // second call: doubles ev.size on a fresh log (quirk Q3)
        UpdateAllRates();
        for (i64 i = 0; i < attempts; i++) {
            seed_rndm_wrapper(rng, user_seed, (uint32_t)i);
            if (totalRate + totalMigrationRate != 0.0 && globalInf != 0) {
                while (ev.ptr < ev.size && (sample_size == -1 || sC < sample_size) &&
                       (time == -1 || currentTime < time)) {
                    Propensities();
                    ChooseTau();
                    while (!GenerateEvents_tau()) tau_l /= 2;
                    currentTime += tau_l;
                    UpdateCompartmentCounts_tau();
                    ev.add(currentTime, MULTITYPE, mev.ptr - P, mev.ptr, 0, 0);
                    if (globalInf == 0) break;
                    for (int p = 0; p < K; p++) CheckLockdown(p);
                }
            }
            if (ev.ptr <= 100 && iterations > 100) {
                Restart();
            } else {
                good_attempt = i + 1;
                break;
            }
        }
    }

    // ---- genealogy, :743-1000
    double guniform() {
        if (inj) {
            if (inj_pos >= inj_n) {
                error = 2;
                return 0.5;
            }
            return inj[inj_pos++];
        }
        return rng.next_double();
    }
    void AddMutation(i64 node, i64 hap, i64 nhap, double time) {  // src/models.pxi:12-26
        i64 x = std::llabs(nhap - hap), site = 0;
        while (x >= 4) {
            x /= 4;
            site++;
        }
        i64 digit4 = 1;
        for (i64 i = 0; i < site; i++) digit4 *= 4;
        mutNode.push_back(node);
        mutDS.push_back((nhap / digit4) % 4);
        mutAS.push_back((hap / digit4) % 4);
        mutSite.push_back(site);
        mutTime.push_back(time);
    }
    void AddMigration(i64 node, double time, i64 oldp, i64 newp) {  // src/models.pxi:42-46
        migNode.push_back(node);
        migTime.push_back(time);
        migOld.push_back(oldp);
        migNew.push_back(newp);
    }

    int Genealogy(bool has_seed, uint64_t seed) {
        if (sC < 2) return 1;  // reference: sys.exit(0), :760-763
        if (has_seed) seed_rndm_wrapper(rng, seed, 0);
        i64 ptr = 0, n = 2 * sC - 1;
        tree.assign(n, 0);
        tree_pop.assign(n, 0);
        times.assign(n, 0.0);
        vector<vector<i64>> L((size_t)K * H), NL((size_t)K * H);
        std::fill(deltaI.begin(), deltaI.end(), 0);
        auto newnode = [&](i64 pop, double t) {
            tree[ptr] = -1;
            tree_pop[ptr] = pop;
            times[ptr] = t;
            return ptr++;
        };
        for (i64 e = ev.ptr - 1; e >= 0; e--) {
            double et = ev.t[e];
            i64 ty = ev.type[e], eh = ev.hap[e], ep = ev.pop[e], enh = ev.nhap[e], enp = ev.npop[e];
            if (ty == BIRTH) {
                vector<i64> &v = L[ep * H + eh];
                i64 lbs = (i64)v.size(), lbs_e = I[ep * H + eh];
                double p = ((double)lbs * ((double)lbs - 1.0)) / (double)lbs_e / ((double)lbs_e - 1.0);
                if (guniform() < p) {
                    i64 n1 = (i64)std::floor(lbs * guniform());
                    i64 n2 = (i64)std::floor((lbs - 1) * guniform());
                    if (n2 >= n1) n2 += 1;
                    i64 id1 = v[n1], id2 = v[n2], id3 = ptr;
                    v[n1] = id3;
                    v[n2] = v[lbs - 1];
                    v.pop_back();
                    tree[id1] = id3;
                    tree[id2] = id3;
                    newnode(ep, et);
                }
                I[ep * H + eh] -= 1;
            } else if (ty == DEATH) {
                I[ep * H + eh] += 1;
            } else if (ty == SAMPLING) {
                I[ep * H + eh] += 1;
                L[ep * H + eh].push_back(ptr);
                newnode(ep, et);
            } else if (ty == MUTATION) {
                vector<i64> &v = L[ep * H + enh];
                i64 lbs = (i64)v.size();
                double p = (double)lbs / (double)I[ep * H + enh];
                if (guniform() < p) {
                    i64 n1 = (i64)std::floor(lbs * guniform());
                    i64 id1 = v[n1];
                    v[n1] = v[lbs - 1];
                    v.pop_back();
                    L[ep * H + eh].push_back(id1);
                    AddMutation(id1, eh, enh, et);
                }
                I[ep * H + enh] -= 1;
                I[ep * H + eh] += 1;
            } else if (ty == SUSCCHANGE) {
            } else if (ty == MIGRATION) {
                vector<i64> &vt = L[enp * H + eh];
                i64 lbs = (i64)vt.size();
                double p = (double)lbs / (double)I[enp * H + eh];
                if (guniform() < p) {
                    i64 nt = (i64)std::floor(lbs * guniform());
                    vector<i64> &vs = L[ep * H + eh];
                    i64 lbss = (i64)vs.size();
                    double p1 = (double)lbss / (double)I[ep * H + eh];
                    if (guniform() < p1) {
                        i64 ns = (i64)std::floor(lbss * guniform());
                        i64 idt = vt[nt], ids = vs[ns], id3 = ptr;
                        vs[ns] = id3;
                        vt[nt] = vt[lbs - 1];
                        vt.pop_back();
                        tree[idt] = id3;
                        tree[ids] = id3;
                        newnode(ep, et);
                        AddMigration(idt, et, ep, enp);
                    } else {
                        vs.push_back(vt[nt]);
                        vt[nt] = vt[lbs - 1];
                        vt.pop_back();
                    }
                }
                I[enp * H + eh] -= 1;
            } else if (ty == MULTITYPE) {
                for (i64 me = eh; me < ep; me++) {
                    i64 num = mev.num[me];
                    double mt = mev.t[me];
                    i64 mty = mev.type[me], mh = mev.hap[me], mp = mev.pop[me], mnh = mev.nhap[me],
                        mnp = mev.npop[me];
                    // cells this record can touch (the reference sweeps all K*H cells after every record,
                    // :987-993; untouched cells have zero delta and no parked lineages, so sweeping only
                    // the touched ones in increasing (deme, haplotype) order is identical)
                    i64 touched[2] = {-1, -1};
                    if (mty == BIRTH) {
                        vector<i64> &v = L[mp * H + mh];
                        i64 lbs = (i64)v.size(), lbs_e = I[mp * H + mh];
                        i64 k = 0;
                        if (!(num == 0 || lbs == 0))
                            k = np_hypergeometric(rng, (i64)((lbs * (lbs - 1.0)) / 2.0),
                                                  (i64)(((lbs_e * (lbs_e - 1)) / 2) - ((lbs * (lbs - 1)) / 2)), num);
                        for (i64 i = 0; i < k; i++) {
                            if (lbs < 2) {  // reference runs into UB here; clamp and count
                                clamped++;
                                break;
                            }
                            i64 n1 = (i64)std::floor(lbs * guniform());
                            i64 n2 = (i64)std::floor((lbs - 1) * guniform());
                            if (n2 >= n1) n2 += 1;
                            i64 id1 = v[n1], id2 = v[n2], id3 = ptr;
                            NL[mp * H + mh].push_back(id3);
                            if (n1 == lbs - 1) {
                                v.pop_back();
                                v[n2] = v[lbs - 2];
                                v.pop_back();
                            } else if (n2 == lbs - 1) {
                                v.pop_back();
                                v[n1] = v[lbs - 2];
                                v.pop_back();
                            } else {
                                v[n1] = v[lbs - 1];
                                v.pop_back();
                                v[n2] = v[lbs - 2];
                                v.pop_back();
                            }
                            tree[id1] = id3;
                            tree[id2] = id3;
                            newnode(mp, mt);
                            lbs -= 2;
                        }
                        deltaI[mp * H + mh] -= num;
                        touched[0] = mp * H + mh;
                    } else if (mty == DEATH) {
                        deltaI[mp * H + mh] += num;
                        touched[0] = mp * H + mh;
                    } else if (mty == SAMPLING) {
                        deltaI[mp * H + mh] += num;
                        for (i64 i = 0; i < num; i++) {
                            NL[mp * H + mh].push_back(ptr);
                            newnode(mp, mt);
                        }
                        touched[0] = mp * H + mh;
                    } else if (mty == MUTATION) {
                        vector<i64> &v = L[mp * H + mnh];
                        i64 lbs = (i64)v.size();
                        i64 k = 0;
                        if (!(num == 0 || lbs == 0)) k = np_hypergeometric(rng, lbs, I[mp * H + mnh] - lbs, num);
                        for (i64 i = 0; i < k; i++) {
                            i64 n1 = (i64)std::floor(lbs * guniform());
                            i64 id1 = v[n1];
                            v[n1] = v[lbs - 1];
                            v.pop_back();
                            NL[mp * H + mh].push_back(id1);
                            AddMutation(id1, mh, mnh, mt);
                            lbs -= 1;
                        }
                        deltaI[mp * H + mnh] -= num;
                        deltaI[mp * H + mh] += num;
                        touched[0] = std::min(mp * H + mh, mp * H + mnh);
                        touched[1] = std::max(mp * H + mh, mp * H + mnh);
                    } else if (mty == SUSCCHANGE) {
                    } else if (mty == MIGRATION) {
                        vector<i64> &vt = L[mnp * H + mh];
                        i64 lbs = (i64)vt.size();
                        if (!(num == 0 || lbs == 0)) {
                            i64 k = np_hypergeometric(rng, lbs, I[mnp * H + mh] - lbs, num);
                            vector<i64> &vs = L[mp * H + mh];
                            i64 lbss = (i64)vs.size();
                            i64 k2 = 0;
                            if (!(k == 0 || lbss == 0)) k2 = np_hypergeometric(rng, lbss, I[mp * H + mh] - lbss, k);
                            for (i64 i = 0; i < k2; i++) {
                                i64 nt = (i64)std::floor(lbs * guniform());
                                i64 ns = (i64)std::floor(lbss * guniform());
                                i64 idt = vt[nt], ids = vs[ns], id3 = ptr;
                                vs[ns] = vs[vs.size() - 1];
                                vs.pop_back();
                                vt[nt] = vt[lbs - 1];
                                vt.pop_back();
                                NL[mp * H + mh].push_back(id3);
                                tree[idt] = id3;
                                tree[ids] = id3;
                                newnode(mp, mt);
                                AddMigration(idt, mt, mp, mnp);
                                lbss -= 1;
                                lbs -= 1;
                            }
                            for (i64 i = 0; i < k - k2; i++) {
                                i64 nt = (i64)std::floor(lbs * guniform());
                                NL[mp * H + mh].push_back(vt[nt]);
                                vt[nt] = vt[lbs - 1];
                                vt.pop_back();
                                lbs -= 1;
                            }
                        }
                        deltaI[mnp * H + mh] -= num;
                        touched[0] = std::min(mp * H + mh, mnp * H + mh);
                        touched[1] = std::max(mp * H + mh, mnp * H + mh);
                    } else {
                        return 2;
                    }
                    for (int c = 0; c < 2; c++) {
                        i64 cell = touched[c];
                        if (cell < 0) continue;
                        I[cell] += deltaI[cell];
                        deltaI[cell] = 0;
                        while (!NL[cell].empty()) {
                            L[cell].push_back(NL[cell].back());
                            NL[cell].pop_back();
                        }
                    }
                }
            } else {
                return 2;
            }
        }
        for (i64 i = 0; i < sC * 2 - 2; i++)
            if (tree_pop[tree[i]] != tree_pop[i]) AddMigration(i, times[i], tree_pop[tree[i]], tree_pop[i]);
        return 0;
    }
};

}  // namespace vgo

using vgo::Model;

extern "C" {

void *vgo_create(int sites, int K, int S, uint64_t seed) { return new Model(sites, K, S, seed); }
void vgo_destroy(void *h) { delete (Model *)h; }

#define CP(dst, src, n) \
    if (src) std::copy(src, src + (n), dst.begin());

// Parameter upload (any pointer may be NULL = keep current).  Layouts are the reference's C-order arrays.
void vgo_set_params(void *h, const double *b, const double *d, const double *s, const double *mRate,
                    const double *hapMutType, const double *sigma, const i64 *suscType, const double *T,
                    const double *m, const double *cd, const double *cdBefore, const double *cdAfter,
                    const double *startLD, const double *endLD, const double *sm, const i64 *sizes) {
    Model &M = *(Model *)h;
    CP(M.b, b, M.H);
    CP(M.d, d, M.H);
    CP(M.s, s, M.H);
    CP(M.mRate, mRate, (size_t)M.H * M.U);
    CP(M.hapMutType, hapMutType, (size_t)M.H * M.U * 3);
    CP(M.sigma, sigma, (size_t)M.H * M.S);
    CP(M.suscType, suscType, M.H);
    CP(M.T, T, (size_t)M.S * M.S);
    CP(M.m, m, (size_t)M.K * M.K);
    CP(M.cd, cd, M.K);
    CP(M.cdBefore, cdBefore, M.K);
    CP(M.cdAfter, cdAfter, M.K);
    CP(M.startLD, startLD, M.K);
    CP(M.endLD, endLD, M.K);
    CP(M.sm, sm, M.K);
    CP(M.sizes, sizes, M.K);
}
// Raw compartment upload (before the first simulate: like set_population_size/set_susceptible;
// afterwards: overwrites the live state and refreshes the totals).
void vgo_set_state(void *h, const i64 *Sx, const i64 *I) {
    Model &M = *(Model *)h;
    CP(M.Sx, Sx, (size_t)M.K * M.S);
    CP(M.I, I, (size_t)M.K * M.H);
    // totals always follow the arrays, so FirstInfection (:234-242) only fires when nobody is infected
    // (the reference cannot reach this path: its set_infectious always raises)
    M.globalInf = 0;
    for (int p = 0; p < M.K; p++) {
        M.totSus[p] = 0;
        M.totInf[p] = 0;
        for (int sn = 0; sn < M.S; sn++) M.totSus[p] += M.Sx[p * M.S + sn];
        for (int hh = 0; hh < M.H; hh++) M.totInf[p] += M.I[p * M.H + hh];
        M.globalInf += M.totInf[p];
    }
}
void vgo_get_state(void *h, i64 *Sx, i64 *I) {
    Model &M = *(Model *)h;
    std::copy(M.Sx.begin(), M.Sx.end(), Sx);
    std::copy(M.I.begin(), M.I.end(), I);
}
int vgo_simulate_direct(void *h, i64 iterations, i64 sample_size, float time, i64 attempts) {
    Model &M = *(Model *)h;
    M.SimulateDirect(iterations, sample_size, time, attempts);
    return M.error;
}
int vgo_simulate_tau(void *h, i64 iterations, i64 sample_size, float time, i64 attempts) {
    Model &M = *(Model *)h;
    M.SimulateTau(iterations, sample_size, time, attempts);
    return M.error;
}
// counters: b, d, s, m, i, migPlus, migNonPlus, swapLockdown, good_attempt, events.ptr, multievents.ptr, globalInfectious
void vgo_get_counters(void *h, i64 *out, double *current_time) {
    Model &M = *(Model *)h;
    i64 c[12] = {M.bC, M.dC, M.sC, M.mC, M.iC, M.migPlus, M.migNonPlus, M.swapLockdown, M.good_attempt, M.ev.ptr,
                 M.mev.ptr, M.globalInf};
    std::copy(c, c + 12, out);
    *current_time = M.currentTime;
}
i64 vgo_num_events(void *h) { return ((Model *)h)->ev.ptr; }
i64 vgo_num_multievents(void *h) { return ((Model *)h)->mev.ptr; }
i64 vgo_prop_num(void *h) { return ((Model *)h)->propNum(); }
// reference layout of export_chain_events (src/_BirthDeath.pyx:1849-1851): 6 x N, float64
void vgo_get_events(void *h, double *out6xN) {
    Model &M = *(Model *)h;
    i64 n = M.ev.ptr;
    for (i64 i = 0; i < n; i++) {
        out6xN[0 * n + i] = M.ev.t[i];
        out6xN[1 * n + i] = (double)M.ev.type[i];
        out6xN[2 * n + i] = (double)M.ev.hap[i];
        out6xN[3 * n + i] = (double)M.ev.pop[i];
        out6xN[4 * n + i] = (double)M.ev.nhap[i];
        out6xN[5 * n + i] = (double)M.ev.npop[i];
    }
}
// replace the event log (the fixed `set_chain_events`): 6 x N float64, plus sCounter recount
void vgo_set_events(void *h, const double *in6xN, i64 n) {
    Model &M = *(Model *)h;
    M.ev.size = n;
    M.ev.ptr = 0;
    M.ev.resize(n);
    M.sC = 0;
    for (i64 i = 0; i < n; i++) {
        M.ev.add(in6xN[i], (i64)in6xN[n + i], (i64)in6xN[2 * n + i], (i64)in6xN[3 * n + i], (i64)in6xN[4 * n + i],
                 (i64)in6xN[5 * n + i]);
        if ((i64)in6xN[n + i] == vgo::SAMPLING) M.sC++;
    }
}
void vgo_get_multievents(void *h, i64 *num, double *t, i64 *type, i64 *hap, i64 *pop, i64 *nhap, i64 *npop) {
    Model &M = *(Model *)h;
    i64 n = M.mev.ptr;
    std::copy(M.mev.num.begin(), M.mev.num.begin() + n, num);
    if (t) std::copy(M.mev.t.begin(), M.mev.t.begin() + n, t);
    if (type) std::copy(M.mev.type.begin(), M.mev.type.begin() + n, type);
    if (hap) std::copy(M.mev.hap.begin(), M.mev.hap.begin() + n, hap);
    if (pop) std::copy(M.mev.pop.begin(), M.mev.pop.begin() + n, pop);
    if (nhap) std::copy(M.mev.nhap.begin(), M.mev.nhap.begin() + n, nhap);
    if (npop) std::copy(M.mev.npop.begin(), M.mev.npop.begin() + n, npop);
}
// replace the multi-event counts and the MULTITYPE rows from a dense device log
// (counts[L][P] int32-as-int64, leap end times[L]); used to feed a CUDA tau log to the oracle genealogy.
void vgo_append_tau_log(void *h, const i64 *counts, const double *times, i64 L) {
    Model &M = *(Model *)h;
    i64 P = M.propNum();
    for (i64 l = 0; l < L; l++) {
        const i64 *c = counts + l * P;
        double t = times[l];
        i64 k = 0;
        for (int sp = 0; sp < M.K; sp++)
            for (int tp = 0; tp < M.K; tp++) {
                if (sp == tp) continue;
                for (int sn = 0; sn < M.S; sn++)
                    for (int hh = 0; hh < M.H; hh++) M.mev.add(c[k++], t, vgo::MIGRATION, hh, sp, sn, tp);
            }
        for (int p = 0; p < M.K; p++) {
            for (int ss = 0; ss < M.S; ss++)
                for (int ts = 0; ts < M.S; ts++)
                    if (ss != ts) M.mev.add(c[k++], t, vgo::SUSCCHANGE, ss, p, ts, 0);
            for (int hh = 0; hh < M.H; hh++) {
                M.mev.add(c[k++], t, vgo::DEATH, hh, p, M.suscType[hh], 0);
                M.sC += c[k];
                M.mev.add(c[k++], t, vgo::SAMPLING, hh, p, M.suscType[hh], 0);
                for (int u = 0; u < M.U; u++)
                    for (int i = 0; i < 3; i++) M.mev.add(c[k++], t, vgo::MUTATION, hh, p, M.Mutate(hh, u, i), 0);
                for (int sn = 0; sn < M.S; sn++) M.mev.add(c[k++], t, vgo::BIRTH, hh, p, sn, 0);
            }
        }
        M.ev.add(t, vgo::MULTITYPE, M.mev.ptr - P, M.mev.ptr, 0, 0);
    }
    M.mev.overflow = false;
}

// deterministic taps -------------------------------------------------------------------------
void vgo_update_all_rates(void *h) { ((Model *)h)->UpdateAllRates(); }
// PrintPropensities (src/_BirthDeath.pyx:2615-2649): UpdateAllRates + Propensities, positional channel order
void vgo_propensities(void *h, double *out, double *dI, double *dS, double *tau) {
    Model &M = *(Model *)h;
    M.UpdateAllRates();
    M.Propensities();
    M.ChooseTau();
    i64 k = 0;
    for (int sp = 0; sp < M.K; sp++)
        for (int tp = 0; tp < M.K; tp++) {
            if (sp == tp) continue;
            for (int sn = 0; sn < M.S; sn++)
                for (int hh = 0; hh < M.H; hh++) out[k++] = M.pMigr[(((size_t)sp * M.K + tp) * M.S + sn) * M.H + hh];
        }
    for (int p = 0; p < M.K; p++) {
        for (int ss = 0; ss < M.S; ss++)
            for (int ts = 0; ts < M.S; ts++)
                if (ss != ts) out[k++] = M.pSus[((size_t)p * M.S + ss) * M.S + ts];
        for (int hh = 0; hh < M.H; hh++) {
            out[k++] = M.pRec[p * M.H + hh];
            out[k++] = M.pSamp[p * M.H + hh];
            for (int u = 0; u < M.U; u++)
                for (int i = 0; i < 3; i++) out[k++] = M.pMut[(((size_t)p * M.H + hh) * M.U + u) * 3 + i];
            for (int sn = 0; sn < M.S; sn++) out[k++] = M.pTr[((size_t)p * M.H + hh) * M.S + sn];
        }
    }
    if (dI) std::copy(M.dI.begin(), M.dI.end(), dI);
    if (dS) std::copy(M.dS.begin(), M.dS.end(), dS);
    if (tau) *tau = M.tau_l;
}
// direct-method rate hierarchy after UpdateAllRates: A[K], eff[K*K], maxEBM[K], evr[K*H*4], hp[K*H],
// popRate[K], migPop[K], totals[2] = {totalRate, totalMigrationRate}
void vgo_rates(void *h, double *A, double *eff, double *maxEBM, double *evr, double *hp, double *popRate,
               double *migPop, double *totals) {
    Model &M = *(Model *)h;
    M.UpdateAllRates();
    if (A) std::copy(M.A.begin(), M.A.end(), A);
    if (eff) std::copy(M.eff.begin(), M.eff.end(), eff);
    if (maxEBM) std::copy(M.maxEBM.begin(), M.maxEBM.end(), maxEBM);
    if (evr) std::copy(M.evr.begin(), M.evr.end(), evr);
    if (hp) std::copy(M.hp.begin(), M.hp.end(), hp);
    if (popRate) std::copy(M.popRate.begin(), M.popRate.end(), popRate);
    if (migPop) std::copy(M.migPop.begin(), M.migPop.end(), migPop);
    if (totals) {
        totals[0] = M.totalRate;
        totals[1] = M.totalMigrationRate;
    }
}

// genealogy ------------------------------------------------------------------------------------
void vgo_inject_uniforms(void *h, const double *u, i64 n) {
    Model &M = *(Model *)h;
    M.inj = u;
    M.inj_n = n;
    M.inj_pos = 0;
}
i64 vgo_injected_used(void *h) { return ((Model *)h)->inj_pos; }
int vgo_genealogy(void *h, int has_seed, uint64_t seed) {
    Model &M = *(Model *)h;
    int r = M.Genealogy(has_seed != 0, seed);
    return r ? r : M.error;
}
i64 vgo_tree_size(void *h) { return (i64)((Model *)h)->tree.size(); }
void vgo_get_tree(void *h, i64 *tree, i64 *tree_pop, double *times) {
    Model &M = *(Model *)h;
    std::copy(M.tree.begin(), M.tree.end(), tree);
    if (tree_pop) std::copy(M.tree_pop.begin(), M.tree_pop.end(), tree_pop);
    if (times) std::copy(M.times.begin(), M.times.end(), times);
}
i64 vgo_num_mutations(void *h) { return (i64)((Model *)h)->mutNode.size(); }
void vgo_get_mutations(void *h, i64 *node, i64 *AS, i64 *DS, i64 *site, double *time) {
    Model &M = *(Model *)h;
    std::copy(M.mutNode.begin(), M.mutNode.end(), node);
    std::copy(M.mutAS.begin(), M.mutAS.end(), AS);
    std::copy(M.mutDS.begin(), M.mutDS.end(), DS);
    std::copy(M.mutSite.begin(), M.mutSite.end(), site);
    std::copy(M.mutTime.begin(), M.mutTime.end(), time);
}
i64 vgo_num_migrations(void *h) { return (i64)((Model *)h)->migNode.size(); }
void vgo_get_migrations(void *h, i64 *node, double *time, i64 *oldp, i64 *newp) {
    Model &M = *(Model *)h;
    std::copy(M.migNode.begin(), M.migNode.end(), node);
    std::copy(M.migTime.begin(), M.migTime.end(), time);
    std::copy(M.migOld.begin(), M.migOld.end(), oldp);
    std::copy(M.migNew.begin(), M.migNew.end(), newp);
}
i64 vgo_num_lockdowns(void *h) { return (i64)((Model *)h)->locState.size(); }
void vgo_get_lockdowns(void *h, i64 *state, i64 *pop, double *time) {
    Model &M = *(Model *)h;
    std::copy(M.locState.begin(), M.locState.end(), state);
    std::copy(M.locPop.begin(), M.locPop.end(), pop);
    std::copy(M.locTime.begin(), M.locTime.end(), time);
}
i64 vgo_clamped(void *h) { return ((Model *)h)->clamped; }
int vgo_multievents_overflowed(void *h) { return ((Model *)h)->mev.overflow ? 1 : 0; }

// RNG taps used by tests/test_oracle_rng.py to pin the restated numpy layer draw-for-draw -----
void *vgo_rng_create(uint64_t entropy, uint32_t num) {
    vgo::Pcg64 *g = new vgo::Pcg64();
    vgo::seed_rndm_wrapper(*g, entropy, num);
    return g;
}
void vgo_rng_destroy(void *g) { delete (vgo::Pcg64 *)g; }
void vgo_rng_doubles(void *g, double *out, i64 n) {
    for (i64 i = 0; i < n; i++) out[i] = ((vgo::Pcg64 *)g)->next_double();
}
void vgo_rng_raw(void *g, uint64_t *out, i64 n) {
    for (i64 i = 0; i < n; i++) out[i] = ((vgo::Pcg64 *)g)->next64();
}
void vgo_rng_poisson(void *g, const double *lam, i64 *out, i64 n) {
    for (i64 i = 0; i < n; i++) out[i] = vgo::np_poisson(*(vgo::Pcg64 *)g, lam[i]);
}
void vgo_rng_hypergeometric(void *g, const i64 *good, const i64 *bad, const i64 *sample, i64 *out, i64 n) {
    for (i64 i = 0; i < n; i++) out[i] = vgo::np_hypergeometric(*(vgo::Pcg64 *)g, good[i], bad[i], sample[i]);
}

}  // extern "C"
