// Device samplers driven by counter-based Philox words.
//
// Poisson replaces numpy's random_poisson (reference src/_BirthDeath.pyx:2531-2532).  The reference
// draws from a sequential PCG64 stream (multiplication method for lam < 10, Hoermann PTRS above);
// here every draw is a pure function of (key, counter):
//   * lam == 0            -> 0, no random words consumed (same as numpy)
//   * 0 < lam < 10        -> exact inversion of one 53-bit uniform U (sequential search from 0).
//                            Fast path: if the top 32 bits of U already prove U < 1-lam <= exp(-lam)
//                            the draw is 0 without touching exp() or the low word.
//   * lam >= 10           -> PTRS transformed rejection (Hoermann 1993), 2 uniforms per trial,
//                            trial t uses Philox domain 2+t of the same channel counter.
#pragma once
#include "common.cuh"

namespace vg {

#ifndef VGSIM_TABLE_SPACE
#define VGSIM_TABLE_SPACE __constant__  // small, hit by a few distinct indices per warp: constant cache beats L1 (A/B 6.78 -> 6.66 ms)
#endif

struct PhiloxCtx {
    uint2 key;
    uint32_t c0, c1, c2;  // counter words x,y,z ; w = dom0 + kind * dstride
    uint32_t dom0 = 0, dstride = 1;  // tau kernel: dom0 = channel group within the cell, dstride = groups per cell
    __device__ __forceinline__ uint4 draw(uint32_t kind) const {
        return philox4x32_10(make_uint4(c0, c1, c2, dom0 + kind * dstride), key);
    }
};

// lane = which of the 4 words of a Philox block belongs to this draw
__device__ __forceinline__ uint32_t pick_word(const uint4 &w, int lane) {
    return lane == 0 ? w.x : lane == 1 ? w.y : lane == 2 ? w.z : w.w;
}

// log(k!) : exact table for k < 16, Stirling series beyond (next term 1/(1680 k^7) < 3e-12 at k = 16)
VGSIM_TABLE_SPACE const double LOGFACT16[16] = {
    0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458, 4.787491742782046, 6.579251212010101,
    8.525161361065415, 10.60460290274525, 12.801827480081469, 15.104412573075516, 17.502307845873887,
    19.987214495661885, 22.552163853123425, 25.19122118273868, 27.89927138384089};

__device__ __forceinline__ double log_factorial(long long k) {
    if (k < 16) return LOGFACT16[k];
    const double x = (double)k, r = 1.0 / x, r2 = r * r;
    return (x + 0.5) * log(x) - x + 0.9189385332046728 + r * (1.0 / 12.0 - r2 * (1.0 / 360.0 - r2 * (1.0 / 1260.0)));
}

// 1/n for the pmf recurrence of the inversion sampler (a table look-up instead of an fp64 division per step)
VGSIM_TABLE_SPACE const double RCP32[32] = {
    0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
    1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21, 1.0 / 22,
    1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31};

__device__ __forceinline__ long long poisson_ptrs_impl(double lam, const PhiloxCtx &ctx, int lane) {
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    double loglam = 0.0, invalpha = 0.0;  // only the (rare) full acceptance test needs them
    bool have = false;
    for (uint32_t t = 0; t < 64; t++) {
        // trial t: domain 2 + 4*t + lane gives 4 words = two 53-bit uniforms private to this channel
        uint4 w = ctx.draw(2u + 4u * t + (uint32_t)lane);
        double U = u53(w.x, w.y) - 0.5;
        double V = u53(w.z, w.w);
        double us = 0.5 - fabs(U);
        long long k = (long long)floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0 || (us < 0.013 && V > us)) continue;
        if (!have) {
            loglam = log(lam);
            invalpha = 1.1239 + 1.1328 / (b - 3.4);
            have = true;
        }
        // log(V) + log(invalpha) - log(a/us^2 + b) <= -lam + k log(lam) - log(k!)
        if (log(V * invalpha / (a / (us * us) + b)) <= -lam + (double)k * loglam - log_factorial(k)) return k;
    }
    return (long long)floor(lam + 0.5);  // unreachable in practice (acceptance ~0.9 per trial)
}
// Trial 0 of poisson_ptrs_impl with the quick acceptance test only: true and k when it accepts (the same k the complete
// sampler returns), false when the draw needs the slow test or another trial (the caller then runs the complete sampler,
// which repeats trial 0 and goes on).
__device__ __forceinline__ bool poisson_ptrs_quick(double lam, const PhiloxCtx &ctx, int lane, long long &k) {
    const double slam = sqrt(lam);
    const double b = 0.931 + 2.53 * slam;
    const double a = -0.059 + 0.02483 * b;
    const double vr = 0.9277 - 3.6224 / (b - 2.0);
    const uint4 w = ctx.draw(2u + (uint32_t)lane);
    const double U = u53(w.x, w.y) - 0.5;
    const double V = u53(w.z, w.w);
    const double us = 0.5 - fabs(U);
    k = (long long)floor((2.0 * a / us + b) * U + lam + 0.43);
    return us >= 0.07 && V <= vr;
}
// out-of-line copy for call sites that sit inside divergent code (team kernel, per-channel parity paths); the warp
// kernel's drain calls the inline body once, with its lanes converged
static __device__ __noinline__ long long poisson_ptrs(double lam, const PhiloxCtx ctx, int lane) {
    return poisson_ptrs_impl(lam, ctx, lane);
}

// Inversion by sequential search.  U = (hi + f) / 2^32 with f in [0,1) supplied by the low word of a second Philox
// block: the walk is first done with f = 0, and if (hi + 1) / 2^32 lands in the same pmf cell the low word cannot
// change the answer and is never drawn (it decides in ~n * 2^-32 of the calls); the result is the same either way.
__device__ __forceinline__ long long poisson_inversion_impl(double lam, uint32_t hi, const PhiloxCtx &ctx, int lane) {
    const double Ulo = (double)hi * (1.0 / 4294967296.0), Uhi = ((double)hi + 1.0) * (1.0 / 4294967296.0);
    double p = exp(-lam), cdf = p;
    long long n = 0;
    while (Ulo >= cdf && n < 256) {
        n++;
        p *= lam * (n < 32 ? RCP32[n] : 1.0 / (double)n);  // lam < 10: the walk almost never passes 32
        cdf += p;
    }
    if (Uhi <= cdf || n >= 256) return n;
    const uint4 w = ctx.draw(1u);
    const double U = u53(hi, pick_word(w, lane));
    while (U >= cdf && n < 256) {
        n++;
        p *= lam * (n < 32 ? RCP32[n] : 1.0 / (double)n);
        cdf += p;
    }
    return n;
}
static __device__ __noinline__ long long poisson_inversion(double lam, uint32_t hi, const PhiloxCtx ctx, int lane) {
    return poisson_inversion_impl(lam, hi, ctx, lane);
}

// `hi`: this channel's word of the chunk's domain-0 Philox block.
__device__ __forceinline__ long long poisson_draw(double lam, uint32_t hi, const PhiloxCtx &ctx, int lane) {
    if (lam < 10.0) {
        // U = (hi + f)/2^32 with f in [0,1).  (hi+1)/2^32 <= 1 - lam  =>  U < 1-lam <= exp(-lam)  =>  0
        if ((double)hi + 1.0 <= (1.0 - lam) * 4294967296.0) return 0;
        return poisson_inversion(lam, hi, ctx, lane);
    }
    return poisson_ptrs(lam, ctx, lane);
}

}  // namespace vg
