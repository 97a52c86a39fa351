// Random source of the genealogy kernel + numpy-compatible hypergeometric sampler.
#pragma once
#include "common.cuh"
#include "logfact_table.cuh"

namespace vg {

struct GRng {
    int mode;  // 0 = Philox, 1 = injected doubles, 2 = injected raw 64-bit words
    const double *ud;
    const unsigned long long *uw;
    long long pos, end;
    uint2 key;
    unsigned long long ctr;
    uint4 buf;
    int have;
    int has32;
    uint32_t b32;
    int err;

    __device__ __forceinline__ unsigned long long next64() {
        if (mode == 2) {
            if (pos >= end) {
                err = 1;
                return 0x8000000000000000ull;
            }
            return uw[pos++];
        }
        if (mode == 0) {
            if (!have) {
                buf = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x47454e45u, 0u), key);
                ctr++;
                have = 1;
                return (unsigned long long)buf.x | ((unsigned long long)buf.y << 32);
            }
            have = 0;
            return (unsigned long long)buf.z | ((unsigned long long)buf.w << 32);
        }
        err = 1;  // a stream of doubles cannot supply raw words
        return 0x8000000000000000ull;
    }
    __device__ __forceinline__ double next_double() {
        if (mode == 1) {
            if (pos >= end) {
                err = 1;
                return 0.5;
            }
            return ud[pos++];
        }
        return (double)(next64() >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ uint32_t next32() {
        if (has32) {
            has32 = 0;
            return b32;
        }
        unsigned long long n = next64();
        has32 = 1;
        b32 = (uint32_t)(n >> 32);
        return (uint32_t)n;
    }
};

// ---- numpy-compatible hypergeometric (random_hypergeometric.c / logfactorial.c of numpy 2.x)
static __device__ __forceinline__ double logfactorial(long long k) {
    const double halfln2pi = 0.9189385332046728;
    if (k < 126) return LOGFACT[k];
    double dk = (double)k;
    return (dk + 0.5) * log(dk) - dk + (halfln2pi + (1.0 / dk) * (1 / 12.0 - 1 / (360.0 * dk * dk)));
}

static __device__ unsigned long long rng_interval(GRng &g, unsigned long long max) {
    if (max == 0) return 0;
    unsigned long long mask = max, value;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    if (max <= 0xffffffffull) {
        while ((value = (g.next32() & mask)) > max && !g.err) {
        }
    } else {
        while ((value = (g.next64() & mask)) > max && !g.err) {
        }
    }
    return value;
}

static __device__ long long hypergeometric(GRng &g, long long good, long long bad, long long sample) {
    if (sample >= 10 && sample <= good + bad - 10) {
        const double D1 = 1.7155277699214135, D2 = 0.8989161620588988;
        long long popsize = good + bad;
        long long cs = sample < popsize - sample ? sample : popsize - sample;
        long long mn = good < bad ? good : bad, mx = good < bad ? bad : good;
        double p = ((double)mn) / (double)popsize, q = ((double)mx) / (double)popsize;
        double mu = (double)cs * p;
        double a = mu + 0.5;
        double var = ((double)(popsize - cs) * (double)cs * p * q / (double)(popsize - 1));
        double c = sqrt(var + 0.5);
        double h = D1 * c + D2;
        long long m = (long long)floor((double)(cs + 1) * (double)(mn + 1) / (double)(popsize + 2));
        double gg = logfactorial(m) + logfactorial(mn - m) + logfactorial(cs - m) + logfactorial(mx - cs + m);
        double b1 = (double)((cs < mn ? cs : mn) + 1), b2 = floor(a + 16 * c);
        double b = b1 < b2 ? b1 : b2;
        long long Kk = 0;
        for (int it = 0; it < 100000 && !g.err; it++) {
            double U = g.next_double();
            double V = g.next_double();
            double X = a + h * (V - 0.5) / U;
            if (X < 0.0 || X >= b) continue;
            Kk = (long long)floor(X);
            double gp = logfactorial(Kk) + logfactorial(mn - Kk) + logfactorial(cs - Kk) + logfactorial(mx - cs + Kk);
            double T = gg - gp;
            if ((U * (4.0 - U) - 3.0) <= T) break;
            if (U * (U - T) >= 1) continue;
            if (2.0 * log(U) <= T) break;
        }
        if (good > bad) Kk = cs - Kk;
        if (cs < sample) Kk = good - Kk;
        return Kk;
    }
    long long total = good + bad;
    long long cs = (sample > total / 2) ? total - sample : sample;
    long long rem_total = total, rem_good = good;
    while (cs > 0 && rem_good > 0 && rem_total > rem_good && !g.err) {
        --rem_total;
        if ((long long)rng_interval(g, (unsigned long long)rem_total) < rem_good) --rem_good;
        --cs;
    }
    if (rem_total == rem_good) rem_good -= cs;
    return (sample > total / 2) ? rem_good : good - rem_good;
}

}  // namespace vg
