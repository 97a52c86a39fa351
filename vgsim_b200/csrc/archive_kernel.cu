// Sparse archive of the dense tau log (vgsim_archive_tau_log).
//
// The dense row of a leap -- one int32 per positional channel, 4P bytes (105 KB at the T3 shape, 1.97 MB at the world
// shape) -- is what the tau kernels write and what the roofline is quoted on.  ~99 % of it is zeros, so a long run keeps
// only a block of leaps dense and moves finished blocks into an archive of (channel, count) pairs, 8 bytes per non-zero
// count, in ascending channel order (the order fixes which random numbers the genealogy replay consumes, so the
// archive must reproduce it).  Two passes over the block's rows, both HBM-read bound:
//   count: non-zero counts per (replicate, leap); a per-replicate scan turns them into the leap offsets sp_off;
//   write: every row is compacted to its offset (ballot-based ordered compaction, one int4 per lane and step).
// One warp per (replicate, row), grid-stride over all rows of the block, so that a batch of 32 replicates fills the GPU as
// well as one of 4,096 (per-replicate CTAs read 150 GB/s at the world shape: 32 CTAs cannot keep HBM busy).
#include "common.cuh"
#include "handle.h"

namespace vg {

__global__ void __launch_bounds__(256) archive_count_kernel(const DevState st, int *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int n16 = st.D.Pp >> 2;
    const long long rows = (long long)st.R * st.dense_cap, nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < rows; w += nwarps) {
        const int r = (int)(w / st.dense_cap), l = (int)(w % st.dense_cap);
        if (l >= (int)(st.counters[(size_t)r * NCOUNT + C_LEAPS] - st.dense_base[r])) continue;
        const int4 *row = reinterpret_cast<const int4 *>(st.tau_counts + (size_t)w * st.D.Pp);
        int c = 0;
#pragma unroll 8
        for (int j = lane; j < n16; j += 32) {
            const int4 v = __ldcs(row + j);
            c += (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) cnt[w] = c;
    }
}

// one warp per replicate: exclusive scan of the block's counts -> sp_off[r][base + l], total -> need[r]
__global__ void archive_scan_kernel(const DevState st, const int *__restrict__ cnt, int *__restrict__ need) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= st.R) return;
    const long long base = st.dense_base[r];
    const int nl = (int)(st.counters[(size_t)r * NCOUNT + C_LEAPS] - base);
    int *off = st.sp_off + (size_t)r * (st.leap_cap + 1);
    int run = st.sp_n[r];
    for (int l0 = 0; l0 < nl; l0 += 32) {
        const int l = l0 + lane;
        const int c = l < nl ? cnt[(size_t)r * st.dense_cap + l] : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (l < nl) off[base + l] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        off[base + nl] = run;
        need[r] = run - st.sp_n[r];
    }
}

__global__ void __launch_bounds__(256) archive_write_kernel(const DevState st) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int n16 = st.D.Pp >> 2;
    const long long rows = (long long)st.R * st.dense_cap, nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < rows; w += nwarps) {
        const int r = (int)(w / st.dense_cap), l = (int)(w % st.dense_cap);
        const long long base = st.dense_base[r];
        if (l >= (int)(st.counters[(size_t)r * NCOUNT + C_LEAPS] - base)) continue;
        const int *off = st.sp_off + (size_t)r * (st.leap_cap + 1);
        int2 *ent = st.sp_ent + (size_t)r * st.sp_cap;
        const int4 *row = reinterpret_cast<const int4 *>(st.tau_counts + (size_t)w * st.D.Pp);
        int pos0 = off[base + l];
        // four steps' loads (2 KB per warp) are issued before the first is compacted
        for (int j0 = 0; j0 < n16; j0 += 128) {
            int4 v4[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u * 32 + lane;
                v4[u] = j < n16 ? __ldcs(row + j) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int4 v = v4[u];
                const int cl = (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
                const unsigned b0 = __ballot_sync(0xffffffffu, cl & 1), b1 = __ballot_sync(0xffffffffu, cl & 2),
                               b2 = __ballot_sync(0xffffffffu, cl & 4);
                if ((b0 | b1 | b2) == 0u) continue;
                int pos = pos0 + __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
                const int c0 = (j0 + u * 32 + lane) * 4;
                if (v.x != 0) ent[pos++] = make_int2(c0, v.x);
                if (v.y != 0) ent[pos++] = make_int2(c0 + 1, v.y);
                if (v.z != 0) ent[pos++] = make_int2(c0 + 2, v.z);
                if (v.w != 0) ent[pos++] = make_int2(c0 + 3, v.w);
                pos0 += __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
            }
        }
    }
}

// after the rows are in the archive: the archive's fill level and the number of archived leaps per replicate
__global__ void archive_commit_kernel(const DevState st, const int *__restrict__ need) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= st.R) return;
    st.sp_n[r] += need[r];
    st.dense_base[r] = st.counters[(size_t)r * NCOUNT + C_LEAPS];
}

cudaError_t launch_archive_count(const DevState &st, int *cnt, int *need, cudaStream_t stream) {
    const long long want = ((long long)st.R * st.dense_cap + 7) / 8;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    archive_count_kernel<<<grid, 256, 0, stream>>>(st, cnt);
    archive_scan_kernel<<<(st.R * 32 + 127) / 128, 128, 0, stream>>>(st, cnt, need);
    return cudaGetLastError();
}

cudaError_t launch_archive_write(const DevState &st, const int *cnt, const int *need, cudaStream_t stream) {
    (void)cnt;
    const long long want = ((long long)st.R * st.dense_cap + 7) / 8;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    archive_write_kernel<<<grid, 256, 0, stream>>>(st);
    archive_commit_kernel<<<(st.R + 255) / 256, 256, 0, stream>>>(st, need);
    return cudaGetLastError();
}

}  // namespace vg
