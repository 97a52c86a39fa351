// Derived constants that depend on the LIVE contact density (UpdateAllRates, reference
// src/_BirthDeath.pyx:279-351, the part that changes on a lockdown flip) and CheckLockdown (:698-710).
// Written group-cooperatively: a "group" is a whole CTA (tau kernel) or one warp (direct kernel).
#pragma once
#include "common.cuh"

namespace vg {

struct BlockGroup {
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct WarpGroup {
    __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
    __device__ __forceinline__ int size() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};

// c[p]   = sum_r m[p,r]^2 cd[r]/A[r]                 (inner factor of BirthRate, :388-390)
// eff[p,q] = sum_r m[p,r] m[q,r] cd[r]/A[r], p != q  (:328-338)
// maxEBM[q] = max_{p!=q} eff[p,q] * max_{h,s} b[h] sigma[h,s]   (:337-348)
// cd: live contact density [K]; eff_g: [K*K] (global); c_out, maxEBM_out: [K] (shared or global)
template <class G>
__device__ void update_contact_rates(const G &g, const Dims &D, const double *__restrict__ pp,
                                     const double *cd, double *eff_g, double *c_out, double *maxEBM_out) {
    const int K = D.K;
    const double *m = pp + D.o_m, *A = pp + D.o_A;
    for (int i = g.tid(); i < K * K; i += g.size()) {
        int p = i / K, q = i - p * K;
        double acc = 0.0;
        if (p == q) {
            for (int r = 0; r < K; r++) acc += m[p * K + r] * m[p * K + r] * cd[r] / A[r];
            c_out[p] = acc;
            eff_g[i] = 0.0;
        } else {
            for (int r = 0; r < K; r++) acc += m[p * K + r] * m[q * K + r] * cd[r] / A[r];
            eff_g[i] = acc;
        }
    }
    g.sync();
    __threadfence_block();
    const double maxB = pp[D.o_maxB];
    for (int q = g.tid(); q < K; q += g.size()) {
        double mx = 0.0;
        for (int p = 0; p < K; p++)
            if (p != q) {
                double e = eff_g[p * K + q];
                if (e > mx) mx = e;
            }
        maxEBM_out[q] = mx * maxB;
    }
    g.sync();
}

// CheckLockdown for deme p (called by ONE thread of the group): returns the number of contact-density
// flips (the caller must run update_contact_rates if non-zero).  totInf is the deme's infectious total.
__device__ inline int check_lockdown(const Dims &D, const double *__restrict__ pp, int p, long long totInf,
                                     double *cd, int *lock, double now, int *loc_n, int *loc_sp, double *loc_t,
                                     int loc_cap, int *err) {
    int flips = 0;
    if ((double)totInf > pp[D.o_startN + p] && lock[p] == 0) {
        cd[p] = pp[D.o_cdA + p];
        lock[p] = 1;
        flips++;
        int n = *loc_n;
        if (n < loc_cap) {
            loc_sp[n] = 1 | (p << 1);
            loc_t[n] = now;
            *loc_n = n + 1;
        } else
            *err |= ERR_LOCKDOWN_OVERFLOW;
    }
    if ((double)totInf < pp[D.o_endN + p] && lock[p] == 1) {
        cd[p] = pp[D.o_cdB + p];
        lock[p] = 0;
        flips++;
        int n = *loc_n;
        if (n < loc_cap) {
            loc_sp[n] = 0 | (p << 1);
            loc_t[n] = now;
            *loc_n = n + 1;
        } else
            *err |= ERR_LOCKDOWN_OVERFLOW;
    }
    return flips;
}

}  // namespace vg
