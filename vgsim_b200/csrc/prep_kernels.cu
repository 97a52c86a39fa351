// PrepareParameters (reference src/_BirthDeath.pyx:433-451): FirstInfection (:234-242), snapshot of the
// initial state for Restart (:438-448), CheckLockdown for every deme and the contact-density dependent
// part of UpdateAllRates (:279-351), per replicate.
#include "common.cuh"
#include "handle.h"
#include "rates.cuh"

namespace vg {

// one warp per replicate: globalInfectious, FirstInfection (:234-242; tau mode :2302-2303), snapshot for Restart
__global__ void __launch_bounds__(256) prepare_kernel(DevState st, int first, int tau_mode) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= st.R) return;
    const Dims &D = st.D;
    const int KH = D.K * D.H, KS = D.K * D.S;
    long long *I = st.I + (size_t)r * KH;
    long long *Sx = st.Sx + (size_t)r * KS;
    long long ginf = 0;
    for (int i = lane; i < KH; i += 32) ginf += I[i];
    for (int o = 16; o > 0; o >>= 1) ginf += __shfl_xor_sync(0xffffffffu, ginf, o);
    // FirstInfection: one host of deme 0, haplotype 0, taken from the first non-empty susceptibility group
    if (ginf == 0 && (first || tau_mode == 1)) {  // tau_mode 2: a later block of the same call, no new seed case
        if (lane == 0)
            for (int sn = 0; sn < D.S; sn++)
                if (Sx[sn] != 0) {
                    Sx[sn] -= 1;
                    I[0] += 1;
                    ginf = 1;
                    break;
                }
        ginf = __shfl_sync(0xffffffffu, ginf, 0);
        __syncwarp();
    }
    if (first) {
        long long *iI = st.initI + (size_t)r * KH;
        long long *iS = st.initSx + (size_t)r * KS;
        for (int i = lane; i < KH; i += 32) iI[i] = I[i];
        for (int i = lane; i < KS; i += 32) iS[i] = Sx[i];
    }
    if (lane == 0) st.counters[(size_t)r * NCOUNT + C_GINF] = ginf;
}

// one CTA per replicate: CheckLockdown for every deme, then refresh c[], eff[][], maxEBM[]
__global__ void __launch_bounds__(128) refresh_kernel(DevState st) {
    const Dims &D = st.D;
    __shared__ int flips;
    extern __shared__ long long tot_s[];  // [K] per-deme infectious totals
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = blockIdx.x; r < st.R; r += gridDim.x) {
        const double *pp = st.params + (size_t)st.rep_pp[r] * D.blob;
        double *cd = st.cd + (size_t)r * D.K;
        int *lock = st.lock + (size_t)r * D.K;
        const long long *I = st.I + (size_t)r * D.K * D.H;
        for (int p = wib; p < D.K; p += nw) {  // a warp per deme, coalesced
            long long t = 0;
            for (int h = lane; h < D.H; h += 32) t += I[p * D.H + h];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) tot_s[p] = t;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int f = 0;
            for (int p = 0; p < D.K; p++)
                f += check_lockdown(D, pp, p, tot_s[p], cd, lock, st.time[r], &st.loc_n[r], st.loc_sp + (size_t)r * st.loc_cap,
                                    st.loc_t + (size_t)r * st.loc_cap, st.loc_cap, &st.err[r]);
            if (f) st.counters[(size_t)r * NCOUNT + C_SWAP] += f;
            flips = f;
        }
        __syncthreads();
        update_contact_rates(BlockGroup(), D, pp, cd, st.eff + (size_t)r * D.K * D.K, st.ceff + (size_t)r * D.K,
                             st.maxEBM + (size_t)r * D.K);
        __syncthreads();
    }
}

cudaError_t launch_prepare(const DevState &st, int first, int tau_mode, cudaStream_t stream) {
    int nb = (st.R * 32 + 255) / 256;
    prepare_kernel<<<nb, 256, 0, stream>>>(st, first, tau_mode);
    return cudaGetLastError();
}

cudaError_t launch_refresh(const DevState &st, cudaStream_t stream) {
    int nb = st.R < 4096 ? st.R : 4096;
    refresh_kernel<<<nb, 128, (size_t)st.D.K * 8, stream>>>(st);
    return cudaGetLastError();
}

}  // namespace vg
