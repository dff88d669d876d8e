// K9: particle routing to the x-slab owners as ONE kernel over peer memory (NVLink 5 / NVSwitch).
//
// The reference has no counterpart (single process, whole grid in host RAM).  On P GPUs every
// particle belongs to the rank that owns the x-plane of its FIRST stencil cell (the rule of pyl_stencil_base_plane /
// pyl_deposit_slab).  Instead of "sort by owner, exchange counts, all-to-all" (three passes, two host
// synchronisations), each CTA ranks 4096 particles by owner in shared memory, reserves one run per destination with
// ONE system-scope atomic on that rank's cursor (a word of its symmetric receive block, reached through its P2P
// mapping), and writes the run with consecutive lanes straight into the destination's receive arrays -- plain
// stores to peer pointers.  The count a rank received stays on the device (its own cursor): pyl_deposit_slab_counted
// reads it there, so the step has no host round trip.
//
// Receive block of a rank (same layout on every rank, symmetric allocation):
//   [0, 256)                       header: word 0 = cursor (particles received so far)
//   [256, 256 + 12*cap)            positions, AoS float32 (x, y, z) like the reference's `pos`
//   [.. , .. + 4*cap)              weights (float32), 256-byte aligned start; present even when unused
// The caller brackets the kernel with two barriers of the symmetric-memory group: cursors zeroed everywhere before
// the first store, every store landed before the deposit reads the block.
#include "common.cuh"
#include "stencil.cuh"

namespace pyl {

constexpr int RT_THREADS = 512;
constexpr int RT_PER = 8;
constexpr int RT_CHUNK = RT_THREADS * RT_PER;        // particles per CTA
constexpr int RT_MAX_RANKS = 16;

struct RouteArgs {
    const float *pos;
    const float *W;
    int64_t n;
    int dims, nranks;
    float inv_cell_size;
    int x_end[RT_MAX_RANKS];            // first plane NOT owned by rank r (planes are split in ascending order)
    unsigned char *peer[RT_MAX_RANKS];  // receive block of every rank
    unsigned cap;                       // particles a receive block can hold
    size_t w_off;                       // byte offset of the weights inside a block
    unsigned long long *lost;           // particles that found a receive block full (must stay 0)
};

__host__ __device__ inline size_t route_w_offset(unsigned cap) { return (256 + (size_t)cap * 12 + 255) / 256 * 256; }

template <int MAS, bool WEIGHTED>
__global__ void __launch_bounds__(RT_THREADS, 2) route_scatter_kernel(const RouteArgs A) {
    extern __shared__ __align__(16) unsigned char route_smem[];
    float4 *stage = reinterpret_cast<float4 *>(route_smem);                      // the chunk in destination order
    unsigned char *sdst = route_smem + (size_t)RT_CHUNK * 16;                   // destination of a staged particle
    __shared__ unsigned cnt[RT_MAX_RANKS], start[RT_MAX_RANKS], goff[RT_MAX_RANKS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t first = (int64_t)blockIdx.x * RT_CHUNK;
    const int n_in = (int)min((int64_t)RT_CHUNK, A.n - first);
    if (tid < RT_MAX_RANKS) cnt[tid] = 0u;
    __syncthreads();

    float p[RT_PER][3], w[RT_PER];
#pragma unroll
    for (int q = 0; q < RT_PER; q++) {                 // all loads first (see partition_kernel)
        const int i = q * RT_THREADS + tid;
        p[q][0] = p[q][1] = p[q][2] = 0.0f;
        w[q] = 1.0f;
        if (i < n_in) {
            const float *src = A.pos + (first + i) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) p[q][k] = __ldg(src + k);
            if (WEIGHTED) w[q] = __ldg(A.W + first + i);
        }
    }
    unsigned dr[RT_PER];                               // (destination << 16) | rank within the destination
#pragma unroll
    for (int q = 0; q < RT_PER; q++) {
        const int i = q * RT_THREADS + tid;
        int dst = -1;
        if (i < n_in) {
            const int plane = wrap_index(axis_base<MAS>(cell_coordinate(p[q][0], A.inv_cell_size)), A.dims);
            dst = 0;
#pragma unroll 1
            for (int r = 0; r < A.nranks - 1; r++) dst += plane >= A.x_end[r] ? 1 : 0;
        }
        // whole warps usually agree on the destination (particles mostly stay where they are): one atomic per
        // group of agreeing lanes, at most twice, then per-lane atomics
        unsigned rank = 0;
        bool todo = dst >= 0;
#pragma unroll 1
        for (int it = 0; it < 2; it++) {
            const unsigned pending = __ballot_sync(0xffffffffu, todo);
            if (pending == 0) break;
            const int leader = __ffs(pending) - 1;
            const int ld = __shfl_sync(0xffffffffu, dst, leader);
            const unsigned m = __ballot_sync(0xffffffffu, todo && dst == ld);
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(cnt + ld, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (todo && dst == ld) {
                rank = base + __popc(m & ((1u << lane) - 1u));
                todo = false;
            }
        }
        if (todo) rank = atomicAdd(cnt + dst, 1u);
        dr[q] = dst >= 0 ? (((unsigned)dst << 16) | rank) : 0xffffffffu;
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive scan over the destinations and one reservation per non-empty destination: a system-scope
        // atomic on the cursor word of that rank's receive block
        const unsigned c = tid < A.nranks ? cnt[tid] : 0u;
        unsigned incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (tid < A.nranks) {
            start[tid] = incl - c;
            unsigned got = 0;
            if (c > 0) got = atomicAdd_system(reinterpret_cast<unsigned *>(A.peer[tid]), c);
            goff[tid] = got - (incl - c);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RT_PER; q++) {
        if (dr[q] != 0xffffffffu) {
            const unsigned d = dr[q] >> 16;
            const unsigned at = start[d] + (dr[q] & 0xffffu);
            stage[at] = make_float4(p[q][0], p[q][1], p[q][2], w[q]);
            sdst[at] = (unsigned char)d;
        }
    }
    __syncthreads();
    unsigned long long lost = 0;
    for (int j = tid; j < n_in; j += RT_THREADS) {
        const unsigned d = sdst[j];
        const unsigned slot = goff[d] + (unsigned)j;
        const float4 v = stage[j];
        if (slot < A.cap) {
            float *dp = reinterpret_cast<float *>(A.peer[d] + 256) + (size_t)slot * 3;
            dp[0] = v.x; dp[1] = v.y; dp[2] = v.z;
            if (WEIGHTED) reinterpret_cast<float *>(A.peer[d] + A.w_off)[slot] = v.w;
        } else {
            lost++;
        }
    }
    if (lost != 0) atomicAdd(A.lost, lost);
}

constexpr size_t RT_SMEM = (size_t)RT_CHUNK * 17;

template <int MAS, bool WEIGHTED>
static int launch_route_w(const RouteArgs &A, cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        PYL_CUDA_CHECK(cudaFuncSetAttribute(route_scatter_kernel<MAS, WEIGHTED>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM));
        attr_done = true;
    }
    const unsigned blocks = (unsigned)((A.n + RT_CHUNK - 1) / RT_CHUNK);
    route_scatter_kernel<MAS, WEIGHTED><<<blocks, RT_THREADS, RT_SMEM, stream>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

template <int MAS>
static int launch_route(const RouteArgs &A, cudaStream_t stream) {
    return A.W ? launch_route_w<MAS, true>(A, stream) : launch_route_w<MAS, false>(A, stream);
}

}  // namespace pyl

using namespace pyl;

extern "C" {

size_t pyl_route_block_bytes(int64_t capacity) {
    if (capacity <= 0 || capacity >= ((int64_t)1 << 32)) return 0;
    return route_w_offset((unsigned)capacity) + (size_t)capacity * 4;
}

int pyl_route_scatter(int mas, const float *pos, const float *W, int64_t particles, int dims, float BoxSize,
                      int nranks, const int *x_offsets, void *const *peer_blocks, int64_t capacity, int64_t *lost,
                      pyl_stream_t stream) {
    PYL_REQUIRE(mas >= PYL_MAS_NGP && mas <= PYL_MAS_PCS, "pyl_route_scatter: unknown scheme");
    PYL_REQUIRE(nranks >= 1 && nranks <= RT_MAX_RANKS, "pyl_route_scatter: 1..16 ranks");
    PYL_REQUIRE(dims > 0 && BoxSize > 0.0f && particles >= 0, "pyl_route_scatter: bad sizes");
    PYL_REQUIRE(capacity > 0 && capacity < ((int64_t)1 << 32), "pyl_route_scatter: capacity must fit 32 bits");
    PYL_REQUIRE(x_offsets != nullptr && peer_blocks != nullptr && lost != nullptr, "pyl_route_scatter: NULL pointer");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(pos != nullptr, "pyl_route_scatter: NULL pos");
    RouteArgs A;
    A.pos = pos; A.W = W; A.n = particles; A.dims = dims; A.nranks = nranks;
    A.inv_cell_size = (float)dims / BoxSize;            // float32 division, MAS_library.pyx:135
    for (int r = 0; r < nranks; r++) {
        PYL_REQUIRE(peer_blocks[r] != nullptr, "pyl_route_scatter: NULL peer block");
        PYL_REQUIRE(x_offsets[r + 1] >= x_offsets[r], "pyl_route_scatter: plane offsets must ascend");
        A.x_end[r] = x_offsets[r + 1];
        A.peer[r] = reinterpret_cast<unsigned char *>(peer_blocks[r]);
    }
    A.cap = (unsigned)capacity;
    A.w_off = route_w_offset(A.cap);
    A.lost = reinterpret_cast<unsigned long long *>(lost);
    cudaStream_t s = as_stream(stream);
    switch (mas) {
        case PYL_MAS_NGP: return launch_route<PYL_MAS_NGP>(A, s);
        case PYL_MAS_CIC: return launch_route<PYL_MAS_CIC>(A, s);
        case PYL_MAS_TSC: return launch_route<PYL_MAS_TSC>(A, s);
        default: return launch_route<PYL_MAS_PCS>(A, s);
    }
}

}  // extern "C"
