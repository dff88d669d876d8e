// K3: deterministic deposit (PYL_MODE_DETERMINISTIC) -- sort by cell, fixed-order segmented sums.
//
// Replaces the same reference loops as deposit_atomic.cu (MAS_library.pyx:142-166, 288-292, 388-404,
// 481-497 and W variants) with a summation order that does not depend on scheduling: two runs on the
// same input give bit-identical grids (the atomic and tiled kernels only promise the 1e-5 per-cell
// tolerance, because float32 addition is not associative and their order of adds varies).
//
//   1. sorted_key    key = linear index of the particle's FIRST stencil cell (reference arithmetic,
//                    stencil.cuh), value = particle index
//   2. cub::DeviceRadixSort::SortPairs on the significant key bits only.  LSD radix sort is stable, so
//                    equal cells keep ascending particle index: the reference's own visiting order.
//   3. sorted_gather record[p] = (dist.xyz, W) of the p-th particle in cell order (16-byte records, so the
//                    passes below stream linearly)
//   4. one pass per stencil offset (l,m,n), S^axes launches in fixed order.  In a pass every particle
//                    contributes to cell(key) + (l,m,n): particles of one cell form one contiguous run and
//                    runs of different cells hit different targets (a translation of the torus is a
//                    bijection), so the run's first thread owns the target cell: it loads it, adds the
//                    run's contributions one by one in particle order -- exactly the chain of float32
//                    adds the reference performs for that offset -- and stores it.  No atomics.
// Weights are the reference's own expressions (axis_stencil<MAS>), products in its order.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "stencil.cuh"

namespace pyl {

template <int MAS, int AXES>
__global__ void __launch_bounds__(256) sorted_key_kernel(const float *__restrict__ pos, int64_t particles,
                                                         int dims, float inv_cell_size,
                                                         unsigned long long *__restrict__ keys,
                                                         unsigned *__restrict__ idx) {
    constexpr int S = StencilWidth<MAS>::value;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < particles; i += stride) {
        unsigned long long key = 0;
#pragma unroll
        for (int a = 0; a < AXES; a++) {
            int c[S];
            float w[S];
            axis_stencil<MAS>(cell_coordinate(__ldg(pos + i * AXES + a), inv_cell_size), dims, c, w);
            key = key * (unsigned long long)dims + (unsigned long long)c[0];
        }
        keys[i] = key;
        idx[i] = (unsigned)i;
    }
}

template <int AXES>
__global__ void __launch_bounds__(256) sorted_gather_kernel(const float *__restrict__ pos,
                                                            const float *__restrict__ W,
                                                            const unsigned *__restrict__ idx, int64_t particles,
                                                            float inv_cell_size, float4 *__restrict__ rec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < particles; p += stride) {
        const int64_t i = idx[p];
        float4 r;
        r.x = cell_coordinate(__ldg(pos + i * AXES), inv_cell_size);
        r.y = cell_coordinate(__ldg(pos + i * AXES + 1), inv_cell_size);
        r.z = AXES == 3 ? cell_coordinate(__ldg(pos + i * AXES + 2), inv_cell_size) : 0.0f;
        r.w = W != nullptr ? __ldg(W + i) : 1.0f;
        rec[p] = r;
    }
}

// w[j] / c[j] for a run-time j without spilling the arrays to local memory
template <int S, typename T>
__device__ __forceinline__ T pick(const T *a, int j) {
    T v = a[0];
#pragma unroll
    for (int i = 1; i < S; i++) v = (j == i) ? a[i] : v;
    return v;
}

template <int MAS, int AXES, bool WEIGHTED>
__global__ void __launch_bounds__(256) sorted_pass_kernel(const unsigned long long *__restrict__ keys,
                                                          const float4 *__restrict__ rec, int64_t particles,
                                                          float *__restrict__ number, int dims, int l, int m,
                                                          int n) {
    constexpr int S = StencilWidth<MAS>::value;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= particles) return;
    const unsigned long long key = keys[p];
    if (p > 0 && keys[p - 1] == key) return;          // not the first particle of its cell

    int c[3][S];
    float w[3][S];
    float4 r = rec[p];
    axis_stencil<MAS>(r.x, dims, c[0], w[0]);
    axis_stencil<MAS>(r.y, dims, c[1], w[1]);
    if (AXES == 3) axis_stencil<MAS>(r.z, dims, c[2], w[2]);
    const int cx = pick<S>(c[0], l), cy = pick<S>(c[1], m);
    float *cell = AXES == 3 ? number + ((int64_t)cx * dims + cy) * dims + pick<S>(c[2], n)
                            : number + (int64_t)cx * dims + cy;
    float acc = *cell;
    int64_t q = p;
    for (;;) {
        float v = __fmul_rn(pick<S>(w[0], l), pick<S>(w[1], m));
        if (AXES == 3) v = __fmul_rn(v, pick<S>(w[2], n));
        if (WEIGHTED) v = __fmul_rn(v, r.w);
        // planes: the reference's loop visits the pinned third axis S times (MAS_library.pyx:138-139)
        if (AXES == 2) v = v * (float)S;
        acc = __fadd_rn(acc, v);
        if (++q >= particles || keys[q] != key) break;
        r = rec[q];
        axis_stencil<MAS>(r.x, dims, c[0], w[0]);
        axis_stencil<MAS>(r.y, dims, c[1], w[1]);
        if (AXES == 3) axis_stencil<MAS>(r.z, dims, c[2], w[2]);
    }
    *cell = acc;
}

struct SortedWorkspace {
    unsigned long long *keys_in, *keys_out;
    unsigned *idx_in, *idx_out;
    float4 *rec;
    void *cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static int key_bits(int dims, int axes) {
    double cells = 1.0;
    for (int a = 0; a < axes; a++) cells *= (double)dims;
    int bits = 1;
    while (bits < 64 && (double)(1ull << bits) < cells) bits++;
    return bits;
}

static SortedWorkspace carve_sorted(void *ws, int64_t particles, int dims, int axes) {
    SortedWorkspace w;
    char *base = reinterpret_cast<char *>(ws);
    size_t off = 0;
    const size_t n = (size_t)particles;
    w.rec = reinterpret_cast<float4 *>(base + off);                  off += align_up(n * 16, 256);
    w.keys_in = reinterpret_cast<unsigned long long *>(base + off);  off += align_up(n * 8, 256);
    w.keys_out = reinterpret_cast<unsigned long long *>(base + off); off += align_up(n * 8, 256);
    w.idx_in = reinterpret_cast<unsigned *>(base + off);             off += align_up(n * 4, 256);
    w.idx_out = reinterpret_cast<unsigned *>(base + off);            off += align_up(n * 4, 256);
    w.cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, (unsigned long long *)nullptr,
                                    (unsigned long long *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr,
                                    (int64_t)particles, 0, key_bits(dims, axes));
    w.cub_tmp = base + off;
    off += align_up(w.cub_bytes, 256);
    w.total = off;
    return w;
}

bool deposit_sorted_supported(int64_t particles) { return particles > 0 && particles < ((int64_t)1 << 32); }

size_t deposit_sorted_workspace(int64_t particles, int dims, int axes) {
    return carve_sorted(nullptr, particles, dims, axes).total;
}

template <int MAS, int AXES>
static int run_sorted(const float *pos, float *number, const float *W, int64_t particles, int dims,
                      float BoxSize, void *ws, cudaStream_t stream) {
    constexpr int S = StencilWidth<MAS>::value;
    SortedWorkspace w = carve_sorted(ws, particles, dims, AXES);
    const float inv = (float)dims / BoxSize;          // float32 division, MAS_library.pyx:135
    int64_t blocks = (particles + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    const int gs = (int)(blocks > cap ? cap : blocks);

    sorted_key_kernel<MAS, AXES><<<gs, 256, 0, stream>>>(pos, particles, dims, inv, w.keys_in, w.idx_in);
    PYL_LAUNCH_CHECK();
    PYL_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, w.cub_bytes, w.keys_in, w.keys_out, w.idx_in,
                                                   w.idx_out, (int64_t)particles, 0, key_bits(dims, AXES),
                                                   stream));
    sorted_gather_kernel<AXES><<<gs, 256, 0, stream>>>(pos, W, w.idx_out, particles, inv, w.rec);
    PYL_LAUNCH_CHECK();
    for (int l = 0; l < S; l++)
        for (int m = 0; m < S; m++)
            for (int n = 0; n < (AXES == 3 ? S : 1); n++) {
                if (W)
                    sorted_pass_kernel<MAS, AXES, true><<<(unsigned)blocks, 256, 0, stream>>>(
                        w.keys_out, w.rec, particles, number, dims, l, m, n);
                else
                    sorted_pass_kernel<MAS, AXES, false><<<(unsigned)blocks, 256, 0, stream>>>(
                        w.keys_out, w.rec, particles, number, dims, l, m, n);
                PYL_LAUNCH_CHECK();
            }
    return PYL_OK;
}

int deposit_sorted(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                   int axes, float BoxSize, void *ws, cudaStream_t stream) {
#define PYL_SORTED_CASE(M)                                                                          \
    case M:                                                                                         \
        return axes == 3 ? run_sorted<M, 3>(pos, number, W, particles, dims, BoxSize, ws, stream)    \
                         : run_sorted<M, 2>(pos, number, W, particles, dims, BoxSize, ws, stream);
    switch (mas) {
        PYL_SORTED_CASE(PYL_MAS_NGP)
        PYL_SORTED_CASE(PYL_MAS_CIC)
        PYL_SORTED_CASE(PYL_MAS_TSC)
        PYL_SORTED_CASE(PYL_MAS_PCS)
    }
#undef PYL_SORTED_CASE
    set_last_error("deposit_sorted: unknown scheme %d", mas);
    return PYL_ERR_ARG;
}

}  // namespace pyl
