// K2f: particle-parallel deposit with one red.global.add.f32 per stencil cell.
//
// Replaces the reference's serial loops MAS_library.pyx:142-166 (CIC), :288-292 (NGP),
// :388-404 (TSC), :481-497 (PCS) and their W variants, and the OpenMP C core MAS_c.c:8-266.
// It is the first kernel of the path (minimum slice), the fallback of the tiled kernel for
// sparse inputs, and the only kernel used for slab deposits with ghost planes.
//
// Layout: pos is [particles][AXES] float32.  Each thread takes FOUR consecutive particles so
// that positions arrive as aligned float4 loads (3 x float4 = 4 particles in 3D, 2 x float4
// in 2D) and weights as one float4; a scalar path covers unaligned bases and the tail.
#include "common.cuh"
#include "deposit_point.cuh"
#include "stencil.cuh"

namespace pyl {

template <int MAS, int AXES, bool WEIGHTED, bool SLAB>
__global__ void __launch_bounds__(256)
deposit_atomic_kernel(const float *__restrict__ pos, const float *__restrict__ W,
                      float *__restrict__ number, int64_t particles, int dims,
                      float inv_cell_size, SlabWindow win, unsigned long long *dropped_out,
                      int vec_ok, const unsigned *__restrict__ n_dev) {
    if (n_dev != nullptr) particles = min(particles, (int64_t)__ldg(n_dev));     // count known on the device only
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long dropped = 0;
    const int64_t groups = vec_ok ? (particles >> 2) : 0;

    for (int64_t g = tid; g < groups; g += stride) {
        float p[4 * AXES];
        const float4 *src = reinterpret_cast<const float4 *>(pos + g * 4 * AXES);
#pragma unroll
        for (int q = 0; q < AXES; q++) {
            const float4 v = __ldg(src + q);
            p[4 * q + 0] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
        }
        float wv[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        if (WEIGHTED) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(W + g * 4));
            wv[0] = v.x; wv[1] = v.y; wv[2] = v.z; wv[3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 4; q++)
            deposit_one<MAS, AXES, WEIGHTED, SLAB>(p + q * AXES, wv[q], number, dims,
                                                   inv_cell_size, win, dropped);
    }
    // scalar tail (or everything, when the base pointers are not 16-byte aligned)
    for (int64_t i = (groups << 2) + tid; i < particles; i += stride) {
        float p[AXES];
#pragma unroll
        for (int a = 0; a < AXES; a++) p[a] = __ldg(pos + i * AXES + a);
        const float wp = WEIGHTED ? __ldg(W + i) : 1.0f;
        deposit_one<MAS, AXES, WEIGHTED, SLAB>(p, wp, number, dims, inv_cell_size, win, dropped);
    }
    if (SLAB && dropped_out != nullptr && dropped != 0) atomicAdd(dropped_out, dropped);
}

template <int MAS, int AXES, bool WEIGHTED, bool SLAB>
static int launch_atomic(const float *pos, const float *W, float *number, int64_t particles,
                         int dims, float inv_cell_size, SlabWindow win,
                         unsigned long long *dropped, cudaStream_t stream, const unsigned *n_dev) {
    const int vec_ok = ((reinterpret_cast<uintptr_t>(pos) & 15) == 0) &&
                       (!WEIGHTED || (reinterpret_cast<uintptr_t>(W) & 15) == 0);
    int64_t blocks = ((particles + 3) / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;    // grid-stride beyond 32 CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    deposit_atomic_kernel<MAS, AXES, WEIGHTED, SLAB><<<(int)blocks, 256, 0, stream>>>(
        pos, W, number, particles, dims, inv_cell_size, win, dropped, vec_ok, n_dev);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

template <int MAS>
static int dispatch_atomic(const float *pos, const float *W, float *number, int64_t particles,
                           int dims, int axes, float inv, bool slab, SlabWindow win,
                           unsigned long long *dropped, cudaStream_t s, const unsigned *n_dev) {
    if (slab) {
        return W ? launch_atomic<MAS, 3, true, true>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev)
                 : launch_atomic<MAS, 3, false, true>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev);
    }
    if (axes == 3) {
        return W ? launch_atomic<MAS, 3, true, false>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev)
                 : launch_atomic<MAS, 3, false, false>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev);
    }
    return W ? launch_atomic<MAS, 2, true, false>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev)
             : launch_atomic<MAS, 2, false, false>(pos, W, number, particles, dims, inv, win, dropped, s, n_dev);
}

// first stencil cell along x (wrapped), same arithmetic as axis_stencil<MAS>
template <int MAS>
__global__ void __launch_bounds__(256) base_plane_kernel(const float *__restrict__ pos, int64_t particles,
                                                         int dims, float inv_cell_size,
                                                         int32_t *__restrict__ plane) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < particles; i += stride) {
        int idx[StencilWidth<MAS>::value];
        float w[StencilWidth<MAS>::value];
        axis_stencil<MAS>(cell_coordinate(__ldg(pos + i * 3), inv_cell_size), dims, idx, w);
        plane[i] = idx[0];
    }
}

int stencil_base_plane(int mas, const float *pos, int64_t particles, int dims, float BoxSize,
                       int32_t *plane, cudaStream_t stream) {
    const float inv = (float)dims / BoxSize;
    int64_t blocks = (particles + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    switch (mas) {
        case PYL_MAS_NGP: base_plane_kernel<PYL_MAS_NGP><<<(int)blocks, 256, 0, stream>>>(pos, particles, dims, inv, plane); break;
        case PYL_MAS_CIC: base_plane_kernel<PYL_MAS_CIC><<<(int)blocks, 256, 0, stream>>>(pos, particles, dims, inv, plane); break;
        case PYL_MAS_TSC: base_plane_kernel<PYL_MAS_TSC><<<(int)blocks, 256, 0, stream>>>(pos, particles, dims, inv, plane); break;
        case PYL_MAS_PCS: base_plane_kernel<PYL_MAS_PCS><<<(int)blocks, 256, 0, stream>>>(pos, particles, dims, inv, plane); break;
        default: set_last_error("stencil_base_plane: unknown scheme %d", mas); return PYL_ERR_ARG;
    }
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

// shared with deposit.cu
int deposit_atomic(int mas, const float *pos, float *number, const float *W, int64_t particles,
                   int dims, int axes, float BoxSize, bool slab, int x_origin, int x_planes,
                   int64_t *dropped, cudaStream_t stream, float plane_mult, const unsigned *n_dev) {
    // inv_cell_size = dims/BoxSize evaluated in float32 like `cdef float inv_cell_size`
    // (MAS_library.pyx:135); IEEE division on the host, identical to the CPU's.
    const float inv = (float)dims / BoxSize;
    SlabWindow win{x_origin, x_planes, plane_mult};
    unsigned long long *dr = reinterpret_cast<unsigned long long *>(dropped);
    switch (mas) {
        case PYL_MAS_NGP: return dispatch_atomic<PYL_MAS_NGP>(pos, W, number, particles, dims, axes, inv, slab, win, dr, stream, n_dev);
        case PYL_MAS_CIC: return dispatch_atomic<PYL_MAS_CIC>(pos, W, number, particles, dims, axes, inv, slab, win, dr, stream, n_dev);
        case PYL_MAS_TSC: return dispatch_atomic<PYL_MAS_TSC>(pos, W, number, particles, dims, axes, inv, slab, win, dr, stream, n_dev);
        case PYL_MAS_PCS: return dispatch_atomic<PYL_MAS_PCS>(pos, W, number, particles, dims, axes, inv, slab, win, dr, stream, n_dev);
    }
    set_last_error("deposit: unknown mass-assignment scheme %d", mas);
    return PYL_ERR_ARG;
}

}  // namespace pyl
