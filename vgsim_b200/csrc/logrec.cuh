// Positional tau-log channel index -> multi-event record, shared by the kernels that READ the dense log (genealogy
// replay, epidemic curves).  The dense row stores only counts; channel identity is the position (SURVEY App. A.4 =
// the loop order of GenerateEvents_tau, reference src/_BirthDeath.pyx:2454-2532), so the reference's multiEvents
// fields (src/events.pxi:105-152) are recomputed here instead of being read from HBM.
#pragma once
#include "common.cuh"

namespace vg {

// decode a tau-log channel index into a multi-event record (type, hap, pop, nhap, npop)
__device__ __forceinline__ void decode_record(int c, const Dims &D, const double *__restrict__ pp, int &type, int &hap,
                                              int &pop, int &nhap, int &npop) {
    const int K = D.K, H = D.H, S = D.S;
    if (c < D.NA) {
        int row = c >> D.hshift;
        hap = c & (H - 1);
        int pair = magic_div(row, D.mgS, S);
        nhap = row - pair * S;
        pop = magic_div(pair, D.mgK1, K - 1);
        int tpp = pair - pop * (K - 1);
        npop = tpp + (tpp >= pop ? 1 : 0);
        type = EV_MIGRATION;
        return;
    }
    int c2 = c - D.NA;
    pop = magic_div(c2, D.mgPD, D.PD);
    int r = c2 - pop * D.PD;
    npop = 0;
    if (r < D.SS1) {
        hap = magic_div(r, D.mgS1, S - 1);
        int tsp = r - hap * (S - 1);
        nhap = tsp + (tsp >= hap ? 1 : 0);
        type = EV_SUSCCHANGE;
        return;
    }
    int r2 = r - D.SS1;
    hap = magic_div(r2, D.mgE, D.E);
    int e = r2 - hap * D.E;
    if (e == 0) {
        type = EV_DEATH;
        nhap = (int)pp[D.o_g + hap];
    } else if (e == 1) {
        type = EV_SAMPLING;
        nhap = (int)pp[D.o_g + hap];
    } else if (e < 2 + 3 * D.U) {
        int uk = e - 2, u = uk / 3, k = uk - u * 3;
        type = EV_MUTATION;
        nhap = mutate_hap(hap, u, k, D.U);
    } else {
        type = EV_BIRTH;
        nhap = e - 2 - 3 * D.U;
    }
}

}  // namespace vg
