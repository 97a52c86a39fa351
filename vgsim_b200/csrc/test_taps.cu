// Test taps for the device samplers (include/vgsim_b200.h: vgsim_test_poisson).
#include <cuda_runtime.h>
#include <string>
#include "../../include/vgsim_b200.h"
#include "common.cuh"
#include "samplers.cuh"

namespace vg {
__global__ void poisson_tap_kernel(const double *lam, long long n, uint64_t seed, long long *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // same addressing as the tau kernel: 4 draws share one Philox block (chunk = i/4, lane = i%4)
    PhiloxCtx ctx;
    ctx.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    ctx.c0 = (uint32_t)(i >> 2);
    ctx.c1 = (uint32_t)(i >> 34);
    ctx.c2 = 0x5eedu;
    double l = lam[i];
    long long v = 0;
    if (l > 0.0) {
        uint4 w = ctx.draw(0u);
        v = poisson_draw(l, pick_word(w, (int)(i & 3)), ctx, (int)(i & 3));
    }
    out[i] = v;
}
}  // namespace vg

extern "C" int vgsim_test_poisson(const double *lam, int64_t n, uint64_t seed, int64_t *out) {
    double *dl = nullptr;
    long long *dout = nullptr;
    if (cudaMalloc(&dl, n * 8) != cudaSuccess || cudaMalloc(&dout, n * 8) != cudaSuccess) return 1;
    cudaMemcpy(dl, lam, n * 8, cudaMemcpyHostToDevice);
    vg::poisson_tap_kernel<<<(unsigned)((n + 255) / 256), 256>>>(dl, n, seed, dout);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out, dout, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(dl);
    cudaFree(dout);
    return e == cudaSuccess ? 0 : 1;
}
