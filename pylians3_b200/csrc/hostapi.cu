// Host-pointer entry points with the reference's exact native signature.
//
// MAS_c.h:3-10 declares  void NGP|CIC|TSC|PCS(FLOAT *pos, FLOAT *number, FLOAT *W,
// long particles, int dims, int axes, FLOAT BoxSize, int threads)  on HOST arrays; these are
// the same argument lists (return value: PYL_* status instead of void) so MAS_c.pxd can bind
// them without touching MAS_library.pyx:1305-1389.  Each call stages pos / W / number through
// a grow-only device arena owned by this file (the only place in the library that allocates),
// runs pyl_deposit on a private stream and copies `number` back -- `number` is accumulated,
// like the reference (+= semantics).
#include <mutex>

#include "common.cuh"

namespace pyl {

int deposit_plane_c_core(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                         float BoxSize, cudaStream_t stream);

struct Arena {
    void *ptr = nullptr;
    size_t bytes = 0;
    cudaStream_t stream = nullptr;
    int device = -1;
};
static Arena g_arena;
static std::mutex g_arena_mu;

static int arena_reserve(size_t bytes) {
    int dev = 0;
    PYL_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_arena.device != dev && g_arena.ptr != nullptr) {
        cudaFree(g_arena.ptr);
        g_arena.ptr = nullptr; g_arena.bytes = 0;
    }
    if (g_arena.stream == nullptr || g_arena.device != dev) {
        if (g_arena.stream) cudaStreamDestroy(g_arena.stream);
        PYL_CUDA_CHECK(cudaStreamCreateWithFlags(&g_arena.stream, cudaStreamNonBlocking));
    }
    g_arena.device = dev;
    if (bytes > g_arena.bytes) {
        if (g_arena.ptr) PYL_CUDA_CHECK(cudaFree(g_arena.ptr));
        g_arena.ptr = nullptr; g_arena.bytes = 0;
        PYL_CUDA_CHECK(cudaMalloc(&g_arena.ptr, bytes));
        g_arena.bytes = bytes;
    }
    return PYL_OK;
}

static int deposit_host(int mas, float *pos, float *number, float *W, long particles, int dims,
                        int axes, float BoxSize) {
    PYL_REQUIRE(axes == 2 || axes == 3, "host deposit: axes must be 2 or 3");
    PYL_REQUIRE(dims > 0 && particles >= 0, "host deposit: bad dims/particles");
    PYL_REQUIRE(number != nullptr && (pos != nullptr || particles == 0), "host deposit: NULL pointer");
    if (particles == 0) return PYL_OK;
    std::lock_guard<std::mutex> lock(g_arena_mu);

    size_t cells = (size_t)dims * dims * (axes == 3 ? dims : 1);
    const size_t b_pos = align_up((size_t)particles * axes * sizeof(float), 256);
    const size_t b_w = W ? align_up((size_t)particles * sizeof(float), 256) : 0;
    const size_t b_grid = align_up(cells * sizeof(float), 256);
    const size_t b_ws = align_up(pyl_deposit_workspace_bytes(mas, particles, dims, axes, PYL_MODE_AUTO), 256);
    int st = arena_reserve(b_pos + b_w + b_grid + b_ws);
    if (st != PYL_OK) return st;

    char *base = reinterpret_cast<char *>(g_arena.ptr);
    float *d_pos = reinterpret_cast<float *>(base);
    float *d_w = W ? reinterpret_cast<float *>(base + b_pos) : nullptr;
    float *d_grid = reinterpret_cast<float *>(base + b_pos + b_w);
    void *d_ws = b_ws ? (void *)(base + b_pos + b_w + b_grid) : nullptr;
    cudaStream_t s = g_arena.stream;

    PYL_CUDA_CHECK(cudaMemcpyAsync(d_pos, pos, (size_t)particles * axes * sizeof(float), cudaMemcpyHostToDevice, s));
    if (W) PYL_CUDA_CHECK(cudaMemcpyAsync(d_w, W, (size_t)particles * sizeof(float), cudaMemcpyHostToDevice, s));
    PYL_CUDA_CHECK(cudaMemcpyAsync(d_grid, number, cells * sizeof(float), cudaMemcpyHostToDevice, s));
    if (axes == 2)   // MAS_c.c deposits a plane with n_max = 1: ONE add per cell (unlike MA's Cython 2D path)
        st = deposit_plane_c_core(mas, d_pos, d_grid, d_w, particles, dims, BoxSize, s);
    else
        st = pyl_deposit(mas, d_pos, d_grid, d_w, particles, dims, axes, BoxSize, PYL_MODE_AUTO, d_ws, b_ws,
                         reinterpret_cast<pyl_stream_t>(s));
    if (st != PYL_OK) return st;
    PYL_CUDA_CHECK(cudaMemcpyAsync(number, d_grid, cells * sizeof(float), cudaMemcpyDeviceToHost, s));
    PYL_CUDA_CHECK(cudaStreamSynchronize(s));
    return PYL_OK;
}

}  // namespace pyl

using namespace pyl;

extern "C" {

int pyl_NGP(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads;
    return deposit_host(PYL_MAS_NGP, pos, number, W, particles, dims, axes, BoxSize);
}
int pyl_CIC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads;
    return deposit_host(PYL_MAS_CIC, pos, number, W, particles, dims, axes, BoxSize);
}
int pyl_TSC(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads;
    return deposit_host(PYL_MAS_TSC, pos, number, W, particles, dims, axes, BoxSize);
}
int pyl_PCS(float *pos, float *number, float *W, long particles, int dims, int axes, float BoxSize, int threads) {
    (void)threads;
    return deposit_host(PYL_MAS_PCS, pos, number, W, particles, dims, axes, BoxSize);
}

int pyl_host_arena_release(void) {
    std::lock_guard<std::mutex> lock(g_arena_mu);
    if (g_arena.ptr) cudaFree(g_arena.ptr);
    if (g_arena.stream) cudaStreamDestroy(g_arena.stream);
    g_arena = Arena();
    return PYL_OK;
}

}  // extern "C"
