// placeholder until the genealogy kernel lands
#include <cuda_runtime.h>
#include <string>
#include "../../include/vgsim_b200.h"
#include "common.cuh"
#include "handle.h"
struct vgsim_handle_s : public vg::Handle {};
extern "C" {
int vgsim_genealogy(vgsim_handle, const uint64_t *, const double *, const int64_t *, int) { return 1; }
int64_t vgsim_tree_size(vgsim_handle, int) { return 0; }
int vgsim_get_tree(vgsim_handle, int, int64_t *, int64_t *, double *) { return 1; }
int64_t vgsim_num_mutations(vgsim_handle, int) { return 0; }
int vgsim_get_mutations(vgsim_handle, int, int64_t *, int64_t *, int64_t *, int64_t *, double *) { return 1; }
int64_t vgsim_num_migrations(vgsim_handle, int) { return 0; }
int vgsim_get_migrations(vgsim_handle, int, int64_t *, double *, int64_t *, int64_t *) { return 1; }
int vgsim_summaries(vgsim_handle, double *) { return 1; }
int vgsim_summaries_dev(vgsim_handle, void **) { return 1; }
int vgsim_set_event_log(vgsim_handle, int, const double *, int64_t, const int64_t *) { return 1; }
int vgsim_test_hypergeometric(const int64_t *, const int64_t *, const int64_t *, int64_t, const uint64_t *, int64_t,
                              int64_t *, int64_t *) { return 1; }
}
