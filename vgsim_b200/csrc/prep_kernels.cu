// PrepareParameters (reference src/_BirthDeath.pyx:433-451): FirstInfection (:234-242), snapshot of the
// initial state for Restart (:438-448), CheckLockdown for every deme and the contact-density dependent
// part of UpdateAllRates (:279-351), per replicate.
#include "common.cuh"
#include "handle.h"
#include "rates.cuh"

namespace vg {

__global__ void prepare_kernel(DevState st, int first, int tau_mode) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.R) return;
    const Dims D = st.D;
    long long *I = st.I + (size_t)r * D.K * D.H;
    long long *Sx = st.Sx + (size_t)r * D.K * D.S;
    long long ginf = 0;
    for (int i = 0; i < D.K * D.H; i++) ginf += I[i];
    auto first_infection = [&]() {
        for (int sn = 0; sn < D.S; sn++) {
            if (Sx[sn] == 0) continue;
            Sx[sn] -= 1;
            I[0] += 1;
            ginf += 1;
            return;
        }
    };
    if (first) {
        if (ginf == 0) first_infection();
        long long *iI = st.initI + (size_t)r * D.K * D.H;
        long long *iS = st.initSx + (size_t)r * D.K * D.S;
        for (int i = 0; i < D.K * D.H; i++) iI[i] = I[i];
        for (int i = 0; i < D.K * D.S; i++) iS[i] = Sx[i];
    }
    if (tau_mode && ginf == 0) first_infection();  // :2302-2303
    st.counters[(size_t)r * NCOUNT + C_GINF] = ginf;
}

// one CTA per replicate: CheckLockdown for every deme, then refresh c[], eff[][], maxEBM[]
__global__ void __launch_bounds__(128) refresh_kernel(DevState st) {
    const Dims D = st.D;
    __shared__ int flips;
    for (int r = blockIdx.x; r < st.R; r += gridDim.x) {
        const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
        double *cd = st.cd + (size_t)r * D.K;
        int *lock = st.lock + (size_t)r * D.K;
        if (threadIdx.x == 0) {
            int f = 0;
            const long long *I = st.I + (size_t)r * D.K * D.H;
            for (int p = 0; p < D.K; p++) {
                long long tot = 0;
                for (int h = 0; h < D.H; h++) tot += I[p * D.H + h];
                f += check_lockdown(D, pp, p, tot, cd, lock, st.time[r], &st.loc_n[r], st.loc_sp + (size_t)r * st.loc_cap,
                                    st.loc_t + (size_t)r * st.loc_cap, st.loc_cap, &st.err[r]);
            }
            st.counters[(size_t)r * NCOUNT + C_SWAP] += f;
            flips = f;
        }
        __syncthreads();
        update_contact_rates(BlockGroup(), D, pp, cd, st.eff + (size_t)r * D.K * D.K, st.ceff + (size_t)r * D.K,
                             st.maxEBM + (size_t)r * D.K);
        __syncthreads();
    }
}

cudaError_t launch_prepare(const DevState &st, int first, int tau_mode, cudaStream_t stream) {
    int nb = (st.R + 127) / 128;
    prepare_kernel<<<nb, 128, 0, stream>>>(st, first, tau_mode);
    return cudaGetLastError();
}

cudaError_t launch_refresh(const DevState &st, cudaStream_t stream) {
    int nb = st.R < 4096 ? st.R : 4096;
    refresh_kernel<<<nb, 128, 0, stream>>>(st);
    return cudaGetLastError();
}

}  // namespace vg
