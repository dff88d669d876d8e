// C entry points of the mass-assignment stage: argument checks + algorithm selection.
#include "common.cuh"

namespace pyl {
int deposit_atomic(int mas, const float *pos, float *number, const float *W, int64_t particles,
                   int dims, int axes, float BoxSize, bool slab, int x_origin, int x_planes,
                   int64_t *dropped, cudaStream_t stream, float plane_mult = 0.0f, const unsigned *n_dev = nullptr);
size_t deposit_tiled_workspace(int mas, int64_t particles, int dims, int axes, int mode, int x_own);
int deposit_tiled(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                  float BoxSize, int x_origin, int x_own, int x_planes, int64_t *dropped, void *ws,
                  cudaStream_t stream, const unsigned *n_dev = nullptr);
bool deposit_tiled_supported(int mas, int64_t particles, int dims, int axes, int x_own);
bool deposit_sorted_supported(int64_t particles);
size_t deposit_sorted_workspace(int64_t particles, int dims, int axes);
int deposit_sorted(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                   int axes, float BoxSize, void *ws, cudaStream_t stream);
int stencil_base_plane(int mas, const float *pos, int64_t particles, int dims, float BoxSize, int32_t *plane,
                       cudaStream_t stream);
}  // namespace pyl

using namespace pyl;

static int resolve_mode(int mas, int64_t particles, int dims, int axes, int mode) {
    if (mode == PYL_MODE_ATOMIC) return mode;
    if (mode == PYL_MODE_DETERMINISTIC) return mode;      // sorted-segment kernel: any dims, 2D or 3D
    const bool ok = deposit_tiled_supported(mas, particles, dims, axes, -1);
    if (mode == PYL_MODE_AUTO) return ok ? PYL_MODE_TILED : PYL_MODE_ATOMIC;
    return ok ? mode : PYL_MODE_ATOMIC;   // TILED / DETERMINISTIC requested but not applicable
}

namespace pyl {
// MAS_c.c semantics for planes (one add per cell): used by the host-pointer entry points
int deposit_plane_c_core(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                         float BoxSize, cudaStream_t stream) {
    return deposit_atomic(mas, pos, number, W, particles, dims, 2, BoxSize, false, 0, dims, nullptr, stream, 1.0f);
}
}  // namespace pyl

extern "C" {

size_t pyl_deposit_workspace_bytes(int mas, int64_t particles, int dims, int axes, int mode) {
    if (mas < PYL_MAS_NGP || mas > PYL_MAS_PCS || particles <= 0 || dims <= 0) return 0;
    const int m = resolve_mode(mas, particles, dims, axes, mode);
    if (m == PYL_MODE_ATOMIC) return 0;
    if (m == PYL_MODE_DETERMINISTIC)
        return deposit_sorted_supported(particles) ? deposit_sorted_workspace(particles, dims, axes) : 0;
    return deposit_tiled_workspace(mas, particles, dims, axes, m, -1);
}

int pyl_deposit(int mas, const float *pos, float *number, const float *W, int64_t particles,
                int dims, int axes, float BoxSize, int mode, void *ws, size_t ws_bytes,
                pyl_stream_t stream) {
    PYL_REQUIRE(mas >= PYL_MAS_NGP && mas <= PYL_MAS_PCS, "pyl_deposit: unknown scheme");
    PYL_REQUIRE(axes == 2 || axes == 3, "pyl_deposit: axes must be 2 or 3");
    PYL_REQUIRE(dims > 0, "pyl_deposit: dims must be positive");
    PYL_REQUIRE(particles >= 0, "pyl_deposit: negative particle count");
    PYL_REQUIRE(BoxSize > 0.0f, "pyl_deposit: BoxSize must be positive");
    PYL_REQUIRE(mode >= PYL_MODE_AUTO && mode <= PYL_MODE_DETERMINISTIC, "pyl_deposit: unknown mode");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(pos != nullptr && number != nullptr, "pyl_deposit: NULL pos/number");
    const int m = resolve_mode(mas, particles, dims, axes, mode);
    if (m == PYL_MODE_ATOMIC)
        return deposit_atomic(mas, pos, number, W, particles, dims, axes, BoxSize, false, 0, dims,
                              nullptr, as_stream(stream));
    if (m == PYL_MODE_DETERMINISTIC)
        PYL_REQUIRE(deposit_sorted_supported(particles), "pyl_deposit: deterministic mode takes < 2^32 particles per call");
    const size_t need = m == PYL_MODE_DETERMINISTIC ? deposit_sorted_workspace(particles, dims, axes)
                                                    : deposit_tiled_workspace(mas, particles, dims, axes, m, -1);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("pyl_deposit: workspace of %zu bytes required, %zu given", need, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    if (m == PYL_MODE_DETERMINISTIC)
        return deposit_sorted(mas, pos, number, W, particles, dims, axes, BoxSize, ws, as_stream(stream));
    return deposit_tiled(mas, pos, number, W, particles, dims, BoxSize, 0, -1, -1, nullptr, ws,
                         as_stream(stream));
}

int pyl_stencil_base_plane(int mas, const float *pos, int64_t particles, int dims, float BoxSize,
                           int32_t *plane, pyl_stream_t stream) {
    PYL_REQUIRE(mas >= PYL_MAS_NGP && mas <= PYL_MAS_PCS, "pyl_stencil_base_plane: unknown scheme");
    PYL_REQUIRE(dims > 0 && BoxSize > 0.0f && particles >= 0, "pyl_stencil_base_plane: bad sizes");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(pos != nullptr && plane != nullptr, "pyl_stencil_base_plane: NULL pointer");
    return stencil_base_plane(mas, pos, particles, dims, BoxSize, plane, as_stream(stream));
}

size_t pyl_deposit_slab_workspace_bytes(int mas, int64_t particles, int dims, int x_own) {
    if (mas < PYL_MAS_NGP || mas > PYL_MAS_PCS || particles <= 0 || dims <= 0 || x_own <= 0) return 0;
    if (!deposit_tiled_supported(mas, particles, dims, 3, x_own)) return 0;
    return deposit_tiled_workspace(mas, particles, dims, 3, PYL_MODE_TILED, x_own);
}

static int deposit_slab_impl(int mas, const float *pos, float *number, const float *W, int64_t particles,
                             const unsigned *n_dev, int dims, float BoxSize, int x_origin, int x_own, int x_planes,
                             int64_t *dropped, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    PYL_REQUIRE(mas >= PYL_MAS_NGP && mas <= PYL_MAS_PCS, "pyl_deposit_slab: unknown scheme");
    PYL_REQUIRE(dims > 0 && x_planes > 0 && x_planes <= dims, "pyl_deposit_slab: bad plane window");
    PYL_REQUIRE(x_own > 0 && x_own <= x_planes, "pyl_deposit_slab: x_own must be in 1..x_planes");
    PYL_REQUIRE(x_origin >= 0 && x_origin < dims, "pyl_deposit_slab: x_origin outside [0,dims)");
    PYL_REQUIRE(particles >= 0 && BoxSize > 0.0f, "pyl_deposit_slab: bad particles/BoxSize");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(pos != nullptr && number != nullptr, "pyl_deposit_slab: NULL pos/number");
    if (x_origin == 0 && x_own == dims && x_planes == dims) {
        // a one-rank "slab": the window is the whole periodic grid
        const size_t whole = pyl_deposit_workspace_bytes(mas, particles, dims, 3, PYL_MODE_TILED);
        if (whole > 0 && ws != nullptr && ws_bytes >= whole)
            return deposit_tiled(mas, pos, number, W, particles, dims, BoxSize, 0, -1, -1, dropped, ws,
                                 as_stream(stream), n_dev);
    }
    const size_t need = pyl_deposit_slab_workspace_bytes(mas, particles, dims, x_own);
    if (need > 0 && ws != nullptr && ws_bytes >= need && x_planes < dims)
        return deposit_tiled(mas, pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dropped,
                             ws, as_stream(stream), n_dev);
    // no (or too small a) workspace, sparse input, or a window spanning the whole grid: atomic kernel
    return deposit_atomic(mas, pos, number, W, particles, dims, 3, BoxSize, true, x_origin,
                          x_planes, dropped, as_stream(stream), 0.0f, n_dev);
}

int pyl_deposit_slab(int mas, const float *pos, float *number, const float *W,
                     int64_t particles, int dims, float BoxSize, int x_origin, int x_own, int x_planes,
                     int64_t *dropped, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    return deposit_slab_impl(mas, pos, number, W, particles, nullptr, dims, BoxSize, x_origin, x_own, x_planes,
                             dropped, ws, ws_bytes, stream);
}

int pyl_deposit_slab_counted(int mas, const float *pos, float *number, const float *W, int64_t capacity,
                             const uint32_t *count, int dims, float BoxSize, int x_origin, int x_own, int x_planes,
                             int64_t *dropped, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    PYL_REQUIRE(count != nullptr, "pyl_deposit_slab_counted: NULL count");
    return deposit_slab_impl(mas, pos, number, W, capacity, count, dims, BoxSize, x_origin, x_own, x_planes,
                             dropped, ws, ws_bytes, stream);
}

}  // extern "C"
