// Host-side handle behind the C ABI (include/vgsim_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace vg {

struct HostParams {  // host copy of one parameter point in the reference's layouts
    std::vector<double> b, d, s, mRate, hapMutType, sigma, T, m, cd, cdBefore, cdAfter, startLD, endLD, sm;
    std::vector<long long> suscType, sizes;
    bool uploaded = false;
};

struct GenealogyBuffers {
    // per replicate outputs; node capacity = 2*sC-1
    long long *node_off = nullptr;     // [R+1] prefix offsets into the node arrays
    int *parent = nullptr, *pop = nullptr;
    double *time = nullptr;
    long long total_nodes = 0;
    // side tables
    long long *mut_off = nullptr, *mig_off = nullptr;  // [R+1] capacity offsets
    int *mut_n = nullptr, *mig_n = nullptr;            // [R] used
    int *mut_node = nullptr, *mut_hap = nullptr;       // hap | nhap packed separately
    int *mut_nhap = nullptr;
    double *mut_time = nullptr;
    int *mig_node = nullptr, *mig_old = nullptr, *mig_new = nullptr;
    double *mig_time = nullptr;
    // lineage arena
    long long *arena_off = nullptr;  // [R+1]
    int *arena = nullptr;
    int *cell_hdr = nullptr;         // [R][K*H][3] (offset, size, cap)
    int *n_nodes = nullptr;          // [R] nodes actually created
    int *scratch = nullptr;          // one int per node: newLineages scratch of the replay, then tree-shape scratch
    bool valid = false;
    std::vector<long long> h_node_off, h_mut_off, h_mig_off;
};

struct Handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = 0;
    Dims D;
    int R = 0, n_pp = 0;
    DevState st;
    std::vector<HostParams> hp;
    long long ev_bound = 0, leap_bound = 0;  // host upper bounds of log rows / leaps per replicate
    long long dense_bound = 0;               // ... of leaps whose dense row has not been archived
    long long launches = 0;
    int tau_variant = 0;  // 0 = aggregated small groups (product), 1 = one Poisson draw per channel (parity tap)
    bool state_set = false;
    bool async_host = false;  // vgsim_set_async: host-buffer copies are enqueued without blocking the calling thread
    GenealogyBuffers gen;
    double *summaries = nullptr;  // [R][VGSIM_NSUMMARY]
    int *work = nullptr;           // [4] atomic work counters of the kernels that hand out replicates dynamically
    int *tau_order = nullptr;      // [2R] scratch of the tau kernel's size-sorted schedule (weights, order)
    int *arch_cnt = nullptr, *arch_need = nullptr;  // scratch of vgsim_archive_tau_log: non-zero counts per dense row / per replicate
    size_t arch_cnt_cap = 0;
    std::vector<int> rep_pp_host;  // host copy of the replicate -> parameter point map
    // event pairs that bracket the hot kernels on the handle's stream: a ring, so that a pipelined driver can read the
    // duration of launch `id` after the fact (vgsim_kernel_ms) instead of blocking on every launch
    static const int NTIMER = 64;
    cudaEvent_t ev_ring0[NTIMER] = {}, ev_ring1[NTIMER] = {};
    long long timer_id = -1;  // id of the last hot-kernel launch
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // the pair of the last launch
    bool ev_valid = false;
    std::vector<void *> allocs;
};

// kernels' host launchers (defined in the .cu files)
cudaError_t launch_tau(const DevState &st, const SimArgs &a, cudaStream_t stream, int num_sms, int variant, int uniform_pp,
                       int *order_buf);
cudaError_t tau_phase_cycles(unsigned long long *out16, int reset);
cudaError_t tau_cta_end(unsigned long long *out1024, int reset);
cudaError_t launch_propensities(const DevState &st, int r, double *out, double *dI, double *dS, double *tau,
                                cudaStream_t stream);
cudaError_t launch_archive_count(const DevState &st, int *cnt, int *need, cudaStream_t stream);
cudaError_t launch_archive_write(const DevState &st, const int *cnt, const int *need, cudaStream_t stream);
cudaError_t launch_prepare(const DevState &st, int first, int tau_mode, cudaStream_t stream);
cudaError_t launch_refresh(const DevState &st, cudaStream_t stream);
cudaError_t launch_direct(const DevState &st, const SimArgs &a, cudaStream_t stream, int num_sms, int uniform_pp, int *work);
cudaError_t launch_curves(const DevState &st, int rep_first, int rep_count, int step_num, long long *inf, long long *sus,
                          long long *removed, long long *sampled, double *time_points, int *last_point,
                          cudaStream_t stream, int num_sms);
cudaError_t launch_rates_tap(const DevState &st, int r, double *ev, double *hp, double *popRate, double *migPop,
                             double *totals, cudaStream_t stream);

}  // namespace vg
