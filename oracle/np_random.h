// TEST INFRASTRUCTURE ONLY (oracle).  Not linked into, imported by, or called from the product
// path (vgsim_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// Restatement of the third-party random layer the reference links against but does not vendor:
//   * mc_lib.rndm.RndmWrapper (mc_lib v0.4.1, reference pyproject.toml:9,33; call sites
//     src/_BirthDeath.pyx:11,74,403,766,2310): numpy PCG64 seeded by
//     SeedSequence(entropy, spawn_key=(num,)); uniform() == bitgen next_double.
//   * numpy.random C distributions (reference src/_BirthDeath.pyx:21): random_poisson (:2532) and
//     random_hypergeometric (:885,931,949,955), numpy 2.3 algorithms
//     (numpy/random/src/distributions/{distributions.c,random_hypergeometric.c,logfactorial.c}):
//     Poisson = multiplication method for lam < 10, Hoermann PTRS otherwise;
//     hypergeometric = urn simulation for small samples, Stadlober HRUA otherwise.
// Pinned in tests/test_oracle_rng.py draw-for-draw against numpy.random.Generator(PCG64(...)).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "logfact_table.h"

namespace vgo {

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state = 0, inc = 0;
    int has32 = 0;
    uint32_t buf32 = 0;

    static u128 mult() { return ((u128)0x2360ED051FC65DA4ULL << 64) | (u128)0x4385DF649FCCF645ULL; }
    void step() { state = state * mult() + inc; }
    // pcg_setseq_128_srandom_r
    void seed(u128 initstate, u128 initseq) {
        state = 0;
        inc = (initseq << 1) | 1;
        step();
        state += initstate;
        step();
        has32 = 0;
        buf32 = 0;
    }
    uint64_t next64() {
        step();
        uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        uint64_t x = hi ^ lo;
        unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((-rot) & 63));
    }
    uint32_t next32() {  // low half first, high half buffered (numpy pcg64_next32)
        if (has32) {
            has32 = 0;
            return buf32;
        }
        uint64_t n = next64();
        has32 = 1;
        buf32 = (uint32_t)(n >> 32);
        return (uint32_t)(n & 0xffffffffu);
    }
    double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
};

// numpy SeedSequence(entropy, spawn_key=(num,)).generate_state(4, uint64) -> PCG64 seed.
struct SeedSeq {
    uint32_t pool[4];
    static constexpr uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu,
                              MULT_B = 0x58f38dedu, MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    static uint32_t hashmix(uint32_t v, uint32_t &hc) {
        v ^= hc;
        hc *= MULT_A;
        v *= hc;
        v ^= v >> 16;
        return v;
    }
    static uint32_t mix(uint32_t x, uint32_t y) {
        uint32_t r = MIX_L * x - MIX_R * y;
        r ^= r >> 16;
        return r;
    }
    // entropy: non-negative integer < 2^64 ; spawn key: one non-negative integer < 2^32
    SeedSeq(uint64_t entropy, uint32_t spawn) {
        uint32_t ent[8];
        int n = 0;
        // _coerce_to_uint32_array(int): little-endian 32-bit words, [0] for zero
        if (entropy == 0) {
            ent[n++] = 0;
        } else {
            uint64_t e = entropy;
            while (e) {
                ent[n++] = (uint32_t)(e & 0xffffffffu);
                e >>= 32;
            }
        }
        while (n < 4) ent[n++] = 0;  // padded to pool size because a spawn key follows
        ent[n++] = spawn;
        uint32_t hc = INIT_A;
        for (int i = 0; i < 4; i++) pool[i] = hashmix(ent[i], hc);
        for (int s = 0; s < 4; s++)
            for (int d = 0; d < 4; d++)
                if (s != d) pool[d] = mix(pool[d], hashmix(pool[s], hc));
        for (int s = 4; s < n; s++)
            for (int d = 0; d < 4; d++) pool[d] = mix(pool[d], hashmix(ent[s], hc));
    }
    void generate64(uint64_t *out, int nwords) const {
        uint32_t hc = INIT_B;
        for (int i = 0; i < 2 * nwords; i++) {
            uint32_t v = pool[i & 3];
            v ^= hc;
            hc *= MULT_B;
            v *= hc;
            v ^= v >> 16;
            if (i & 1)
                out[i >> 1] |= (uint64_t)v << 32;
            else
                out[i >> 1] = v;
        }
    }
};

inline void seed_rndm_wrapper(Pcg64 &g, uint64_t entropy, uint32_t num) {
    SeedSeq ss(entropy, num);
    uint64_t w[4];
    ss.generate64(w, 4);
    u128 st = ((u128)w[0] << 64) | w[1];
    u128 ic = ((u128)w[2] << 64) | w[3];
    g.seed(st, ic);
}

// ---------------------------------------------------------------- numpy distributions
inline double np_loggam(double x) {
    static const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03, 7.936507936507937e-04,
                                 -5.952380952380952e-04, 8.417508417508418e-04, -1.917526917526918e-03,
                                 6.410256410256410e-03, -2.955065359477124e-02, 1.796443723688307e-01,
                                 -1.39243221690590e+00};
    if (x == 1.0 || x == 2.0) return 0.0;
    int64_t n = (x < 7.0) ? (int64_t)(7 - x) : 0;
    double x0 = x + n;
    double x2 = (1.0 / x0) * (1.0 / x0);
    const double lg2pi = 1.8378770664093453e+00;
    double gl0 = a[9];
    for (int k = 8; k >= 0; k--) {
        gl0 *= x2;
        gl0 += a[k];
    }
    double gl = gl0 / x0 + 0.5 * lg2pi + (x0 - 0.5) * std::log(x0) - x0;
    if (x < 7.0) {
        for (int64_t k = 1; k <= n; k++) {
            gl -= std::log(x0 - 1.0);
            x0 -= 1.0;
        }
    }
    return gl;
}

template <class G>
inline int64_t np_poisson(G &g, double lam) {
    if (lam >= 10) {
        double slam = std::sqrt(lam), loglam = std::log(lam);
        double b = 0.931 + 2.53 * slam;
        double a = -0.059 + 0.02483 * b;
        double invalpha = 1.1239 + 1.1328 / (b - 3.4);
        double vr = 0.9277 - 3.6224 / (b - 2);
        for (;;) {
            double U = g.next_double() - 0.5;
            double V = g.next_double();
            double us = 0.5 - std::fabs(U);
            int64_t k = (int64_t)std::floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) return k;
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if ((std::log(V) + std::log(invalpha) - std::log(a / (us * us) + b)) <=
                (-lam + k * loglam - np_loggam((double)(k + 1))))
                return k;
        }
    } else if (lam == 0) {
        return 0;
    } else {
        double enlam = std::exp(-lam);
        int64_t X = 0;
        double prod = 1.0;
        for (;;) {
            prod *= g.next_double();
            if (prod > enlam)
                X += 1;
            else
                return X;
        }
    }
}

inline double np_logfactorial(int64_t k) {
    const double halfln2pi = 0.9189385332046728;
    if (k < 126) return VG_LOGFACT[k];
    return (k + 0.5) * std::log((double)k) - k + (halfln2pi + (1.0 / k) * (1 / 12.0 - 1 / (360.0 * k * k)));
}

template <class G>
inline uint64_t np_interval(G &g, uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max, value;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    if (max <= 0xffffffffULL) {
        while ((value = (g.next32() & mask)) > max) {
        }
    } else {
        while ((value = (g.next64() & mask)) > max) {
        }
    }
    return value;
}

template <class G>
inline int64_t np_hypergeometric(G &g, int64_t good, int64_t bad, int64_t sample) {
    if (sample >= 10 && sample <= good + bad - 10) {
        const double D1 = 1.7155277699214135, D2 = 0.8989161620588988;
        int64_t popsize = good + bad;
        int64_t cs = sample < popsize - sample ? sample : popsize - sample;
        int64_t mn = good < bad ? good : bad, mx = good < bad ? bad : good;
        double p = ((double)mn) / popsize, q = ((double)mx) / popsize;
        double mu = cs * p;
        double a = mu + 0.5;
        double var = ((double)(popsize - cs) * cs * p * q / (popsize - 1));
        double c = std::sqrt(var + 0.5);
        double h = D1 * c + D2;
        int64_t m = (int64_t)std::floor((double)(cs + 1) * (mn + 1) / (popsize + 2));
        double gg = np_logfactorial(m) + np_logfactorial(mn - m) + np_logfactorial(cs - m) +
                    np_logfactorial(mx - cs + m);
        double b1 = (double)((cs < mn ? cs : mn) + 1), b2 = std::floor(a + 16 * c);
        double b = b1 < b2 ? b1 : b2;
        int64_t K;
        for (;;) {
            double U = g.next_double();
            double V = g.next_double();
            double X = a + h * (V - 0.5) / U;
            if (X < 0.0 || X >= b) continue;
            K = (int64_t)std::floor(X);
            double gp = np_logfactorial(K) + np_logfactorial(mn - K) + np_logfactorial(cs - K) +
                        np_logfactorial(mx - cs + K);
            double T = gg - gp;
            if ((U * (4.0 - U) - 3.0) <= T) break;
            if (U * (U - T) >= 1) continue;
            if (2.0 * std::log(U) <= T) break;
        }
        if (good > bad) K = cs - K;
        if (cs < sample) K = good - K;
        return K;
    }
    int64_t total = good + bad;
    int64_t cs = (sample > total / 2) ? total - sample : sample;
    int64_t rem_total = total, rem_good = good;
    while (cs > 0 && rem_good > 0 && rem_total > rem_good) {
        --rem_total;
        if ((int64_t)np_interval(g, (uint64_t)rem_total) < rem_good) --rem_good;
        --cs;
    }
    if (rem_total == rem_good) rem_good -= cs;
    return (sample > total / 2) ? rem_good : good - rem_good;
}

}  // namespace vgo
