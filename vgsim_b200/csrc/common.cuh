// Shared device/host definitions for the vgsim_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vg {

enum { EV_BIRTH = 0, EV_DEATH = 1, EV_SAMPLING = 2, EV_MUTATION = 3, EV_SUSCCHANGE = 4, EV_MIGRATION = 5, EV_MULTITYPE = 6 };

// counters per replicate (include/vgsim_b200.h VGSIM_NCOUNTERS)
enum { C_B = 0, C_D, C_S, C_M, C_I, C_MIGP, C_MIGN, C_SWAP, C_GOOD, C_EVPTR, C_LEAPS, C_GINF, NCOUNT };

// sticky per-replicate error bits
enum {
    ERR_ZERO_WEIGHT = 1,      // categorical draw hit a zero weight (reference: sys.exit in fastChoose)
    ERR_LOCKDOWN_OVERFLOW = 2,
    ERR_ARENA = 4,            // genealogy lineage arena exhausted
    ERR_STREAM = 8,           // injected uniform stream exhausted
    ERR_CLAMPED = 16,         // tau-log BIRTH record asked for more coalescences than lineages allow
    ERR_COUNT_OVERFLOW = 32,  // compartment count does not fit int32
    ERR_BADLOG = 64,
    ERR_TAU_STUCK = 128,      // tau leap infeasible after 80 halvings (state itself violates the bounds)
    ERR_SIDE_TABLE = 256,     // genealogy mutation / migration table too small (the host retries with larger tables)
};

// ---------------------------------------------------------------------------------------------
// Model dimensions and the layout of one parameter point inside a flat fp64 blob.
struct Dims {
    int K, H, S, U;
    int E;        // per-haplotype channels in a deme block: 2 + 3U + S
    int SS1;      // S*(S-1)
    int PD;       // per-deme block: SS1 + H*E
    int NA;       // migration section: K*(K-1)*S*H
    int P;        // total channels
    int Pp;       // P rounded up to a multiple of 8 (row stride of the dense log, int32 units: rows are 32-byte aligned)
    int hshift;   // log2(H)
    // q = n / d for the runtime divisors of the channel decode (logrec.cuh): q = umulhi(n, m) with m = ceil(2^32 / d) is n / d
    // or n / d + 1 for every n < 2^32 (the excess n * (m * d - 2^32) / (d * 2^32) is below 1), one compare fixes it up
    unsigned mgS, mgK1, mgPD, mgS1, mgE;
    // parameter blob offsets (in doubles)
    int o_b, o_d, o_sr, o_q, o_tmq, o_sigT, o_T, o_Tc, o_m, o_A, o_sm, o_cdB, o_cdA, o_startN, o_endN,
        o_size, o_g, o_mu, o_w, o_maxB, o_cd0, blob;
};

__host__ __device__ inline Dims make_dims(int U, int K, int S) {
    Dims D;
    D.K = K; D.S = S; D.U = U;
    int H = 1;
    for (int i = 0; i < U; i++) H *= 4;
    D.H = H;
    D.hshift = 2 * U;
    D.E = 2 + 3 * U + S;
    D.SS1 = S * (S - 1);
    D.PD = D.SS1 + H * D.E;
    D.NA = K * (K - 1) * S * H;
    D.P = D.NA + K * D.PD;
    D.Pp = (D.P + 7) & ~7;
    {
        auto magic = [](int d) { return d > 1 ? (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d) : 0u; };
        D.mgS = magic(S); D.mgK1 = magic(K - 1); D.mgPD = magic(D.PD); D.mgS1 = magic(S - 1); D.mgE = magic(D.E);
    }
    int o = 0;
    D.o_b = o; o += H;
    D.o_d = o; o += H;
    D.o_sr = o; o += H;
    D.o_q = o; o += H * U * 3;      // mutation channel rate per infected: mRate*w_k/(w0+w1+w2)
    D.o_tmq = o; o += H;            // sum_u,k q
    D.o_sigT = o; o += S * H;       // susceptibility transposed [s][h]
    D.o_T = o; o += S * S;
    D.o_Tc = o; o += S;
    D.o_m = o; o += K * K;          // migration matrix with diagonal filled in
    D.o_A = o; o += K;              // actual sizes
    D.o_sm = o; o += K;
    D.o_cdB = o; o += K;
    D.o_cdA = o; o += K;
    D.o_startN = o; o += K;         // startLD*size
    D.o_endN = o; o += K;           // endLD*size
    D.o_size = o; o += K;           // sizes as fp64
    D.o_g = o; o += H;              // suscType as fp64
    D.o_mu = o; o += H * (U > 0 ? U : 1);      // mRate (direct method: site choice)
    D.o_w = o; o += H * (U > 0 ? U : 1) * 3;   // hapMutType weights (direct method: allele choice)
    D.o_maxB = o; o += 1;           // max_h,s b*sigma
    D.o_cd0 = o; o += K;            // contact density as uploaded (what vgsim_reset restores)
    D.blob = (o + 1) & ~1;
    return D;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: word i of draw (ctr,key) is a pure function.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    // 53-bit uniform in [0,1): hi supplies the top 32 bits, lo the next 21
    return (double)(((uint64_t)hi << 21) | (uint64_t)(lo >> 11)) * (1.0 / 9007199254740992.0);
}

// Packed 64-bit event descriptor: type(3) | hap(18) | pop(12) | nhap(18) | npop(12)
__host__ __device__ inline uint64_t pack_event(int type, int hap, int pop, int nhap, int npop) {
    return (uint64_t)type | ((uint64_t)hap << 3) | ((uint64_t)pop << 21) | ((uint64_t)nhap << 33) |
           ((uint64_t)npop << 51);
}
__host__ __device__ inline void unpack_event(uint64_t d, int &type, int &hap, int &pop, int &nhap, int &npop) {
    type = (int)(d & 7);
    hap = (int)((d >> 3) & 0x3FFFF);
    pop = (int)((d >> 21) & 0xFFF);
    nhap = (int)((d >> 33) & 0x3FFFF);
    npop = (int)((d >> 51) & 0xFFF);
}
__host__ __device__ inline uint64_t pack_multi(uint32_t leap) { return (uint64_t)EV_MULTITYPE | ((uint64_t)leap << 3); }
__host__ __device__ inline uint32_t unpack_multi(uint64_t d) { return (uint32_t)(d >> 3); }

// Mutate (reference src/_BirthDeath.pyx:2420-2427): haplotype after putting the k-th OTHER allele at site u
__host__ __device__ inline int mutate_hap(int h, int u, int k, int U) {
    int sh = 2 * (U - u - 1);
    int as = (h >> sh) & 3;
    int ds = k + (k >= as ? 1 : 0);
    return h + ((ds - as) << sh);
}

// ---------------------------------------------------------------------------------------------
// Device-resident state of a handle (all pointers are device pointers).
// n / d through the precomputed multiplier (d >= 1; d == 1 has no multiplier)
__host__ __device__ __forceinline__ int magic_div(int n, unsigned m, int d) {
    if (d <= 1) return n;
#ifdef __CUDA_ARCH__
    int q = (int)__umulhi((unsigned)n, m);
#else
    int q = (int)(((unsigned long long)(unsigned)n * m) >> 32);
#endif
    return q * d > n ? q - 1 : q;
}

struct DevState {
    Dims D;
    int R;
    int n_pp;
    const double *params;      // [n_pp][D.blob]
    const int *rep_pp;         // [R]
    const uint64_t *seeds;     // [R]
    long long *I;              // [R][K*H]
    long long *Sx;             // [R][K*S]
    long long *initI, *initSx; // snapshot taken by the first simulate call (Restart target)
    double *cd;                // [R][K] live contact density
    int *lock;                 // [R][K] lockdownON
    double *eff;               // [R][K*K] effective migration (depends on live contact density)
    double *ceff;              // [R][K]   c[p] = sum_r m[p,r]^2 cd[r]/A[r]
    double *maxEBM;            // [R][K]
    double *time;              // [R]
    double *rate_tot;          // [R][2] the direct kernel's incrementally maintained totalRate / totalMigrationRate at its exit (parity tap)
    long long *counters;       // [R][NCOUNT]
    unsigned *epoch;           // [R] Philox stream epoch (bumped per attempt)
    int *err;                  // [R] sticky error bits
    // event log
    long long ev_cap;
    double *ev_time;           // [R][ev_cap]
    unsigned long long *ev_desc;
    // dense tau log
    long long leap_cap;        // leaps per replicate that tau_tt / sp_off can hold (dense + archived)
    long long dense_cap;       // dense rows per replicate
    int *tau_counts;           // [R][dense_cap][Pp]: leap L of replicate r is row L - dense_base[r]
    double *tau_tt;            // [R][leap_cap][2]  (time after the leap, tau)
    // sparse archive of the dense rows (vgsim_archive_tau_log): the non-zero counts of leap L < dense_base[r], in
    // ascending channel order, are sp_ent[r * sp_cap + sp_off[r][L] .. sp_off[r][L + 1])
    long long *dense_base;     // [R] leaps already archived (0: everything is dense)
    int2 *sp_ent;              // [R][sp_cap] (channel, count)
    long long sp_cap;
    int *sp_off;               // [R][leap_cap + 1]
    int *sp_n;                 // [R] entries in use
    // lockdown records
    int loc_cap;
    int *loc_n;                // [R]
    int *loc_sp;               // [R][loc_cap]  state | pop<<1
    double *loc_t;             // [R][loc_cap]
    long long *ev_base;        // [R] log rows dropped by vgsim_recycle_log (the <= 100-row Restart test counts them)
    int first_simulation;      // 0 until the first simulate call snapshotted the initial state
};

struct SimArgs {
    long long iterations;
    long long sample_size;
    float time;     // C float like the reference (quirk Q1)
    int has_time;
    long long attempts;
    int cont;       // 1: a later block of one SimulatePopulation_tau call (vgsim_simulate_tau_blocks) -- the first attempt
                    // keeps the epoch of the block before, so that the blocks draw what the single call would have drawn
};

}  // namespace vg
