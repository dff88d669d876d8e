// Tiled deposit (bucket by tile -> shared-memory accumulation -> coalesced flush).
// Placeholder until the tiled kernels land: reports "not supported" so that pyl_deposit
// routes every request to the atomic kernel.
#include "common.cuh"

namespace pyl {

bool deposit_tiled_supported(int, int64_t, int, int) { return false; }
size_t deposit_tiled_workspace(int, int64_t, int, int, int) { return 0; }
int deposit_tiled(int, const float *, float *, const float *, int64_t, int, int, float, int, void *,
                  size_t, cudaStream_t) {
    set_last_error("tiled deposit not built");
    return PYL_ERR_ARG;
}

}  // namespace pyl
