// placeholder until the batched direct kernel lands
#include "common.cuh"
#include "handle.h"
namespace vg {
cudaError_t launch_direct(const DevState &, const SimArgs &, cudaStream_t, int) { return cudaErrorNotSupported; }
cudaError_t launch_rates_tap(const DevState &, int, double *, double *, double *, double *, double *, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace vg
