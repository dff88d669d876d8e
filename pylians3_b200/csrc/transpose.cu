// K8: slab-FFT transpose as ONE kernel over peer memory (NVLink 5 / NVSwitch), replacing "pack into a send
// buffer + NCCL all-to-all".
//
// After the batched 2D r2c over (y,z) every rank holds (nx_local, N, nz) complex64 for its own x-planes; the
// 1D transform along x needs, on rank r, ALL x for the ky rows that rank owns (mirrored pairs, see
// pyl_pk_bin_mirrored).  Each rank therefore writes row (ix, ky, :) straight into the receive buffer of the
// rank owning ky, at [x0 + ix][stored row of ky][:] -- plain 8-byte stores to peer pointers (P2P mappings of
// symmetric allocations), coalesced along kz.  No staging copy, no separate collective: the NVLink traffic is
// issued by the same kernel that reads the FFT output.  The caller brackets the kernel with two device-side
// barriers (receive buffers free / all rows landed).
//
// The reference has no counterpart (single process, FFTW on the whole grid, Pk_library.pyx:117-130).
#include "common.cuh"

namespace pyl {

#ifndef PYL_TR_THREADS
#define PYL_TR_THREADS 256
#endif
#ifndef PYL_TR_V
#define PYL_TR_V 2           // 1: 8-byte stores; 2: 16-byte stores after aligning the destination row
#endif
#ifndef PYL_TR_CTAS_PER_SM
#define PYL_TR_CTAS_PER_SM 2
#endif
constexpr int TR_THREADS = PYL_TR_THREADS;
constexpr int TR_MAX_RANKS = 16;

struct TransposeArgs {
    const float2 *src;                 // (nx, N, nz)
    float2 *peer[TR_MAX_RANKS];        // receive buffer of every rank: (N, nky[r], nz)
    int nky[TR_MAX_RANKS];
    const int *ky_owner;               // [N] rank owning global row ky
    const int *ky_row;                 // [N] stored row index of ky on its owner
    const int *ky_order;               // [N] or NULL: row handled by CTA column b (a permutation of 0..N-1)
    int N, nz, nx, x0;
    int ky_major;                      // receive buffers are (nky[r], N, nz) instead of (N, nky[r], nz)
};

__global__ void __launch_bounds__(TR_THREADS) transpose_scatter_kernel(const TransposeArgs A) {
    // A SMALL persistent grid (PYL_TR_CTAS_PER_SM CTAs per SM, grid-stride over the rows): the kernel is bound by the
    // NVLink, which a quarter of the thread slots saturates, and it runs NEXT TO the 2D-FFT kernels of the following
    // batch.  With one CTA per row (65 536 per batch at 4096^3) whichever kernel was scheduled first filled every SM
    // and the two streams took turns: 2D FFTs + transposes cost their SUM (102 of 38 + 66 ms), not their maximum.
    // Row order: CTAs are scheduled in blockIdx order.  With ky = column every rank would sweep the owners in the same
    // order at the same time -- all senders on one or two receivers' NVLink ingress, the other links idle (the
    // all-to-all incast).  ky_order interleaves the owners row by row, starting with a different owner on every
    // rank, so the rows in flight at any time go to all ranks.
    const int64_t rows = (int64_t)A.N * A.nx;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int col = (int)(row % A.N), ix = (int)(row / A.N);
    const int ky = A.ky_order != nullptr ? __ldg(A.ky_order + col) : col;
    const int r = __ldg(A.ky_owner + ky), j = __ldg(A.ky_row + ky);
    const float2 *src = A.src + ((int64_t)ix * A.N + ky) * A.nz;
    float2 *dst = A.ky_major ? A.peer[r] + ((int64_t)j * A.N + (A.x0 + ix)) * A.nz
                             : A.peer[r] + ((int64_t)(A.x0 + ix) * A.nky[r] + j) * A.nz;
#if PYL_TR_V == 1
    // four loads in flight per thread before the first peer store: a store over NVLink costs microseconds of latency,
    // and with one element per iteration the row (16 KB at 4096^3) was latency-bound
    for (int kz0 = threadIdx.x; kz0 < A.nz; kz0 += 4 * TR_THREADS) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int kz = kz0 + u * TR_THREADS;
            if (kz < A.nz) v[u] = __ldg(src + kz);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int kz = kz0 + u * TR_THREADS;
            if (kz < A.nz) dst[kz] = v[u];
        }
    }
#else
    // 16-byte peer stores: rows are nz = dims/2+1 complex64 long (odd for even dims), so a row starts on an 8-byte
    // boundary only; one leading element aligns the DESTINATION, the body moves pairs (a warp stores 512 contiguous
    // bytes per instruction), one trailing element may remain.  Four pairs per thread are loaded before the first
    // store leaves: a store over NVLink costs microseconds of latency.
    const int head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1u);
    if (head && threadIdx.x == 0) dst[0] = __ldg(src);
    const int npair = (A.nz - head) >> 1;
    const float2 *s2 = src + head;
    float4 *d4 = reinterpret_cast<float4 *>(dst + head);
    const bool src16 = (reinterpret_cast<uintptr_t>(s2) & 15u) == 0;
    for (int p0 = threadIdx.x; p0 < npair; p0 += 4 * TR_THREADS) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * TR_THREADS;
            if (p < npair) {
                if (src16) {
                    v[u] = __ldg(reinterpret_cast<const float4 *>(s2) + p);
                } else {
                    const float2 a = __ldg(s2 + 2 * p), b = __ldg(s2 + 2 * p + 1);
                    v[u] = make_float4(a.x, a.y, b.x, b.y);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int p = p0 + u * TR_THREADS;
            if (p < npair) d4[p] = v[u];
        }
    }
    if (((A.nz - head) & 1) && threadIdx.x == 32) dst[A.nz - 1] = __ldg(src + A.nz - 1);
#endif
    }
}

}  // namespace pyl

using namespace pyl;

static int transpose_scatter(const float *slab_k, void *const *peer_recv, const int *nky_of_rank,
                             const int *ky_owner, const int *ky_row, const int *ky_order, int dims, int nx, int x0,
                             int nranks, int ky_major, pyl_stream_t stream) {
    PYL_REQUIRE(nranks >= 1 && nranks <= TR_MAX_RANKS, "pyl_transpose_scatter: 1..16 ranks");
    PYL_REQUIRE(dims > 0 && nx >= 0 && x0 >= 0 && x0 + nx <= dims, "pyl_transpose_scatter: bad plane range");
    if (nx == 0) return PYL_OK;
    PYL_REQUIRE(slab_k != nullptr && peer_recv != nullptr && nky_of_rank != nullptr && ky_owner != nullptr &&
                    ky_row != nullptr, "pyl_transpose_scatter: NULL pointer");
    PYL_REQUIRE(nx <= 65535, "pyl_transpose_scatter: more than 65535 local planes");
    TransposeArgs A;
    A.src = reinterpret_cast<const float2 *>(slab_k);
    for (int r = 0; r < nranks; r++) {
        PYL_REQUIRE(peer_recv[r] != nullptr, "pyl_transpose_scatter: NULL peer buffer");
        A.peer[r] = reinterpret_cast<float2 *>(peer_recv[r]);
        A.nky[r] = nky_of_rank[r];
    }
    A.ky_owner = ky_owner; A.ky_row = ky_row; A.ky_order = ky_order;
    A.N = dims; A.nz = dims / 2 + 1; A.nx = nx; A.x0 = x0; A.ky_major = ky_major;
    int64_t ctas = (int64_t)sm_count() * PYL_TR_CTAS_PER_SM;
    if (ctas > (int64_t)dims * nx) ctas = (int64_t)dims * nx;
    transpose_scatter_kernel<<<(unsigned)ctas, TR_THREADS, 0, as_stream(stream)>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

extern "C" int pyl_transpose_scatter(const float *slab_k, void *const *peer_recv, const int *nky_of_rank,
                                     const int *ky_owner, const int *ky_row, const int *ky_order, int dims, int nx,
                                     int x0, int nranks, pyl_stream_t stream) {
    return transpose_scatter(slab_k, peer_recv, nky_of_rank, ky_owner, ky_row, ky_order, dims, nx, x0, nranks, 0,
                             stream);
}

// Same, into receive buffers laid out (nky[r], N, nz): the x axis in the MIDDLE.  The 1D transforms along x then
// stride by one row (nz elements) inside one (N, nz) plane per ky instead of by nky*nz elements across the whole
// buffer -- at 4096^3 over 8 ranks the latter touches a different 2 MB page for every x (TLB-bound: 83 ms for the
// x transforms, profiles/r2_config5_pieces.md).
extern "C" int pyl_transpose_scatter_kymajor(const float *slab_k, void *const *peer_recv, const int *nky_of_rank,
                                             const int *ky_owner, const int *ky_row, const int *ky_order, int dims,
                                             int nx, int x0, int nranks, pyl_stream_t stream) {
    return transpose_scatter(slab_k, peer_recv, nky_of_rank, ky_owner, ky_row, ky_order, dims, nx, x0, nranks, 1,
                             stream);
}
