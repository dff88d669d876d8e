// Tiled deposit, round 2: two-level particle partition with shared-memory ranking and coalesced run writes, then
// one CTA per grid tile that accumulates in shared memory WITHOUT atomics and flushes with bulk reductions.
//
// Replaces the same reference loops as deposit_atomic.cu (MAS_library.pyx:142-166, 288-292, 388-404,
// 481-497 and W variants).
//
// Why this shape (measured on B200, scratch/ubench.cu, profiles/r2_ubench.md):
//   * the particle-parallel red.global kernel moves 21x the algorithmic bytes through DRAM (1.8% of roofline);
//   * a one-level scatter with one returning global atomic and one scattered 16-byte store per particle is bound
//     by the L2 request rate (2.8 ms for 134 M records); runs of >= 8 records written by consecutive lanes reach
//     bandwidth (0.45 ms).  Runs need few bins per pass, hence TWO passes of <= 2048 bins each;
//   * same-address returning atomics serialise (~20-100 ns each): the per-run cursors of the first pass are
//     replicated (one sub-segment per replica) so that a cursor sees only 1/REPL of the CTAs;
//   * shared-memory integer atomics are cheap (3.6 cycles per warp instruction), match.any is not (64 cycles), float
//     shared atomics are a CAS loop (13 cycles): ranking uses ATOMS, duplicates are found with shuffles on the
//     cell-sorted order, accumulation is plain conflict-free read-modify-write;
//   * cp.reduce.async.bulk (UBLKRED) adds a 128-byte row of the tile to the grid at twice the rate of coalesced
//     red.global.
//
// Pipeline (all on the caller's stream, scratch in the caller's workspace, no host synchronisation):
//   1. tile_count    histogram of a 1-in-8 SAMPLE of the particles over tiles (tile = 8 x 16 x 32 cells).
//   2. tile_caps + exclusive scan: per-tile bucket capacity = 1.125 x estimate + 4 sigma + 32 (multiple of 4) and
//                    the bucket start offsets; tile_setup derives the super-tile segments (a super-tile = 2^k
//                    consecutive tiles, k chosen so that both passes have about sqrt(ntiles) bins).
//   3. partition<1>  pos (AoS) -> buf1 (SoA planes x,y,z[,W] of cell coordinates dist = fl32(pos*inv)), grouped by
//                    super-tile.  A CTA ranks 4096 particles by bin in shared memory (ATOMS), reserves one run per
//                    non-empty bin with ONE global atomic, and writes the runs with consecutive lanes.
//   4. partition<2>  buf1 -> buf2 grouped by tile, same kernel body, bins = tiles of one super-tile.
//   5. tile_deposit  one CTA per tile: bucket chunks arrive by cp.async.bulk + mbarrier (UBLKCP); counting sort by
//                    cell inside shared memory; each warp owns four z-planes of the tile with PRIVATE accumulators
//                    (own planes + stencil halo), so a particle is visited once, does all S^3 plain
//                    read-modify-writes, and no barrier is needed inside the accumulation; the overlapping private
//                    planes are summed and the tile (+ halo) is added to the grid with cp.reduce.async.bulk.
//   A particle that finds its segment or bucket full (capacities come from a sample) is deposited on the spot with
//   red.global -- correctness never depends on the estimate.
// Weights are the reference's own arithmetic (stencil.cuh: float64 where the reference promotes, then rounded),
// identical to the atomic kernel's; products are formed as (wx*W)*wy*wz and added with a fused multiply-add.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "deposit_point.cuh"
#include "stencil.cuh"

namespace pyl {

constexpr int TX = 8, TY = 16, TZ = 32;         // tile extent in cells
constexpr int TILE_CELLS = TX * TY * TZ;        // 4096
#ifndef PYL_P2_GROUP
#define PYL_P2_GROUP 4
#endif
#ifndef PYL_TNT
#define PYL_TNT 256
#endif
constexpr int TNT = PYL_TNT;                    // threads per tile CTA
constexpr int SAMPLE = 8;                       // tile_count looks at one particle group in SAMPLE
#ifndef PYL_PT
#define PYL_PT 512
#endif
#ifndef PYL_PPER
#define PYL_PPER 8
#endif
constexpr int PT = PYL_PT;                      // threads per partition CTA
constexpr int PPER = PYL_PPER;                  // particles per thread
constexpr int PP = PT * PPER;                   // 4096 particles per partition CTA
constexpr int MAX_BINS = 2048;                  // bins per partition pass
constexpr int MAX_REPL = 64;                    // cursor replicas of the first pass
constexpr int DIRECT_SLOTS = 1024;              // a CTA of pass 1 goes straight to the tile buckets if the bounding box of
                                                // its particles' tiles holds at most this many tiles
constexpr int OUT_ROW = 36;                     // floats per flushed row (32 + halo, 16-byte multiple)

struct TileGeom {
    int dims;
    int ntx, nty, ntz;
    unsigned ntiles;
    float inv_cell_size;
    // x window of the destination buffer (multi-GPU slabs).  Whole grid: x_origin = 0, x_own = x_planes =
    // dims and planes wrap periodically.  Slab: the buffer holds global planes x_origin .. x_origin +
    // x_planes - 1; a particle belongs here iff its first stencil plane is one of the first x_own planes
    // (the other x_planes - x_own planes are the upward ghost planes its stencil may reach).
    int x_origin, x_own, x_planes;
    int tps_shift;        // a super-tile = 1 << tps_shift consecutive tiles
    unsigned nsuper;      // number of super-tiles
    unsigned repl;        // cursor replicas (sub-segments) per super-tile in pass 1
};

constexpr unsigned NO_TILE = 0xffffffffu;

// periodic wrap for the partition kernels: branch-free single fold (inputs inside the documented domain
// 0 <= pos <= BoxSize need no more); anything still outside is sent through the true modulo
__device__ __forceinline__ int wrap_once(int i, int dims) {
    i += i < 0 ? dims : 0;
    i -= i >= dims ? dims : 0;
    return i;
}

template <int MAS>
__device__ __forceinline__ unsigned tile_and_local(const float d[3], const TileGeom &g, int local[3],
                                                   unsigned *tcoord = nullptr) {
    static_assert(TX == 8 && TY == 16 && TZ == 32, "shifts below");
    int wb[3];
#pragma unroll
    for (int a = 0; a < 3; a++) wb[a] = wrap_once(axis_base<MAS>(d[a]), g.dims);
    if (__builtin_expect(((unsigned)wb[0] >= (unsigned)g.dims) | ((unsigned)wb[1] >= (unsigned)g.dims) |
                         ((unsigned)wb[2] >= (unsigned)g.dims), 0)) {
#pragma unroll
        for (int a = 0; a < 3; a++) wb[a] = wrap_index_slow(wb[a], g.dims);
    }
    wb[0] -= g.x_origin;                // plane index inside the destination window
    wb[0] += wb[0] < 0 ? g.dims : 0;
    const bool mine = wb[0] < g.x_own;
    const unsigned t0 = (unsigned)wb[0] >> 3, t1 = (unsigned)wb[1] >> 4, t2 = (unsigned)wb[2] >> 5;
    local[0] = wb[0] & (TX - 1);
    local[1] = wb[1] & (TY - 1);
    local[2] = wb[2] & (TZ - 1);
    if (tcoord != nullptr) *tcoord = (t0 << 20) | (t1 << 10) | t2;       // tile coordinates, 10 bits each
    return mine ? (t0 * g.nty + t1) * g.ntz + t2 : NO_TILE;
}

// ---- 1. sampled histogram --------------------------------------------------------------------------
template <int MAS>
__global__ void __launch_bounds__(256) tile_count_kernel(const float *__restrict__ pos, int64_t particles,
                                                         TileGeom g, unsigned *__restrict__ counts, int vec_ok,
                                                         const unsigned *__restrict__ n_dev) {
    if (n_dev != nullptr) particles = min(particles, (int64_t)__ldg(n_dev));     // count known on the device only
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t groups = vec_ok ? (particles >> 2) : 0;
    int local[3];
    // one group of 4 particles out of every SAMPLE consecutive groups, at a pseudo-random place: a FIXED stride aliases
    // with lattice-ordered input (a stride of 32 particles is one tile length at Np = N^3: every sampled particle then
    // sits at the same offset inside its tile, the estimate follows the displacement across the tile boundary
    // instead of the density, and ~0.5% of the particles overflowed their buckets)
    for (int64_t sg = tid; sg * SAMPLE < groups; sg += stride) {
        unsigned h = (unsigned)sg * 0x9E3779B1u;
        h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
        int64_t grp = sg * SAMPLE + (h % SAMPLE);
        if (grp >= groups) grp = sg * SAMPLE;
        float p[12];
        const float4 *src = reinterpret_cast<const float4 *>(pos + grp * 12);
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float4 v = __ldg(src + q);
            p[4 * q] = v.x; p[4 * q + 1] = v.y; p[4 * q + 2] = v.z; p[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float d[3];
#pragma unroll
            for (int a = 0; a < 3; a++) d[a] = cell_coordinate(p[3 * q + a], g.inv_cell_size);
            const unsigned t = tile_and_local<MAS>(d, g, local);
            if (t != NO_TILE) atomicAdd(counts + t, 1u);
        }
    }
    // unaligned input / tail: every SAMPLE-th particle
    for (int64_t si = tid; (groups << 2) + si * SAMPLE < particles; si += stride) {
        unsigned h = (unsigned)si * 0x9E3779B1u;
        h ^= h >> 15; h *= 0x85EBCA77u; h ^= h >> 13;
        int64_t i = (groups << 2) + si * SAMPLE + (h % SAMPLE);
        if (i >= particles) i = (groups << 2) + si * SAMPLE;
        float d[3];
#pragma unroll
        for (int a = 0; a < 3; a++) d[a] = cell_coordinate(__ldg(pos + i * 3 + a), g.inv_cell_size);
        const unsigned t = tile_and_local<MAS>(d, g, local);
        if (t != NO_TILE) atomicAdd(counts + t, 1u);
    }
}

// ---- 2. capacities from the sampled counts (in place; then scanned into bucket starts) ---------------
__host__ __device__ __forceinline__ unsigned tile_capacity(unsigned sampled) {
    // 1.125 x estimate + 4 sigma (sigma of the estimate = SAMPLE*sqrt(sampled)) + 32, rounded up to a multiple of
    // 4 slots so that every bucket starts on a 16-byte boundary of its plane (bulk copies)
    const unsigned est = sampled * SAMPLE;
    const unsigned sig = (unsigned)(4.0f * SAMPLE * sqrtf((float)sampled)) + 1u;
    return (est + (est >> 3) + sig + 32u + 3u) & ~3u;
}

__global__ void tile_caps_kernel(unsigned *__restrict__ counts, unsigned ntiles) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) counts[i] = tile_capacity(counts[i]);
    else if (i == ntiles) counts[i] = 0u;
}

// segment of super-tile s, replica r, inside the bucket array: [seg_begin, seg_begin + rep_cap)
struct Segment {
    unsigned begin, cap;
};
__device__ __forceinline__ Segment super_segment(const unsigned *__restrict__ starts, const TileGeom &g, unsigned s,
                                                 unsigned r) {
    const unsigned t0 = s << g.tps_shift;
    unsigned t1 = (s + 1) << g.tps_shift;
    if (t1 > g.ntiles) t1 = g.ntiles;
    const unsigned b = __ldg(starts + t0), e = __ldg(starts + t1);
    Segment sg;
    sg.cap = ((e - b) / g.repl) & ~3u;
    sg.begin = b + r * sg.cap;
    return sg;
}

// cur2[t] = bucket start of tile t; cur1[r][s] = start of sub-segment (s, r) of buf1
__global__ void __launch_bounds__(1024) tile_setup_kernel(const unsigned *__restrict__ starts, TileGeom g,
                                                          unsigned *__restrict__ cur1, unsigned *__restrict__ cur2) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g.ntiles) cur2[i] = starts[i];
    if (i < g.nsuper) {
        const Segment sg = super_segment(starts, g, i, 0);
        for (unsigned r = 0; r < g.repl; r++) cur1[r * g.nsuper + i] = sg.begin + r * sg.cap;
    }
}

// ---- 3./4. partition passes ---------------------------------------------------------------------------------
// bucket full (capacity came from a sample): deposit the particle directly.  Out of line: it is rare.
template <int MAS, bool WEIGHTED>
__device__ __noinline__ void overflow_deposit(float d0, float d1, float d2, float wp, TileGeom g,
                                              float *__restrict__ number, unsigned long long *dropped) {
    const float d[3] = {d0, d1, d2};
    unsigned long long dr = 0;
    if (g.x_planes == g.dims)
        deposit_dist<MAS, 3, WEIGHTED, false>(d, wp, number, g.dims, SlabWindow{0, g.dims, 0.0f}, dr);
    else
        deposit_dist<MAS, 3, WEIGHTED, true>(d, wp, number, g.dims, SlabWindow{g.x_origin, g.x_planes, 0.0f}, dr);
    *dropped += dr;
}

struct PartArgs {
    const float *pos;            // pass 1 input (AoS) ...
    const float *W;
    int64_t particles;
    const float4 *in;            // pass 2 input (buf1)
    float4 *out;                 // output records (dist.x, dist.y, dist.z, W): buf1 in pass 1, buf2 in pass 2
    float4 *out2;                // pass 1: buf2, written directly by CTAs whose particles touch few tiles
    const unsigned *starts;      // ntiles + 1 bucket starts
    unsigned *cur1;              // [repl][nsuper]
    unsigned *cur2;              // [ntiles]
    float *number;               // the grid (overflow path only)
    unsigned long long *dropped;
    double *wsum;                // sum of |W| over the particles of this call (pass 1, weighted only)
    const unsigned *n_dev;       // optional: the particle count lives on the device (routed particles); then
                                 // `particles` is the capacity of the arrays
};

// dynamic shared memory of a partition CTA: PP staged records + three words per bin + the bin of every staged record
static size_t part_smem_bytes(int nb) {
    if (nb < DIRECT_SLOTS) nb = DIRECT_SLOTS;
    return (size_t)PP * 16 + 3 * (size_t)((nb + 3) & ~3) * 4 + (size_t)PP * 2 + 64 * 4;
}

template <int MAS, bool WEIGHTED, int LEVEL>
__global__ void __launch_bounds__(PT, 1024 / PT) partition_kernel(PartArgs a, TileGeom g) {
    constexpr int S = StencilWidth<MAS>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int nb = LEVEL == 1 ? (int)g.nsuper : (1 << g.tps_shift);
    const int nb4 = ((nb < DIRECT_SLOTS ? DIRECT_SLOTS : nb) + 3) & ~3;
    float4 *stage = reinterpret_cast<float4 *>(smem_raw);                      // PP records in bin order
    unsigned *cnt = reinterpret_cast<unsigned *>(stage + PP);                  // nb: counts -> bin starts
    unsigned *goff = cnt + nb4;                                                // global slot of sorted index 0 of a bin
    unsigned *lim = goff + nb4;                                                // first slot past the bin's segment
    unsigned short *sbin = reinterpret_cast<unsigned short *>(lim + nb4);       // PP: bin of a staged record
    unsigned *misc = reinterpret_cast<unsigned *>(sbin + PP);                  // [0..15] warp partials, [16..] work item

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- which items -------------------------------------------------------------------------------------
    // Pass 2 is persistent: 2 CTAs per SM walk the (super-tile, replica) sub-segments of buf1 and take the PP-record
    // chunks of the non-empty ones.  (One CTA per possible chunk was ~380 000 CTAs at 1024^3, 1.8 ms of work
    // lookups even when pass 1 had sent everything straight to the tiles.)  Segment order: PYL_P2_GROUP replicas of
    // one super-tile, then the next super-tile -- CTAs running at the same time reserve runs on DIFFERENT bucket
    // cursors (same-address atomics serialise) while the runs written at the same time still fall into a limited
    // set of buckets (L2 write combining).  Warp 0 keeps the walk state; 32 candidate segments per look-up round.
    unsigned seg_next = blockIdx.x, w_super = 0, w_at = 0, w_fill = 0;
    for (;;) {
    int64_t first;              // index of this CTA's first item (particle index / slot of buf1)
    int n_in;                   // items of this CTA
    unsigned super = 0, repl = 0;
    if (LEVEL == 1) {
        first = (int64_t)blockIdx.x * PP;
        const int64_t n_all = a.n_dev != nullptr ? min(a.particles, (int64_t)__ldg(a.n_dev)) : a.particles;
        if (first >= n_all) return;
        n_in = (int)min((int64_t)PP, n_all - first);
        repl = blockIdx.x % g.repl;
    } else {
        if (warp == 0) {
            if (w_at >= w_fill) {
                const unsigned G = min((unsigned)PYL_P2_GROUP, g.repl);
                const unsigned nseg = (g.repl + G - 1) / G * G * g.nsuper;
                while (seg_next < nseg) {
                    const unsigned my = seg_next + lane * gridDim.x;
                    unsigned mb = 0, mf = 0, ms = 0;
                    if (my < nseg) {
                        const unsigned s = (my / G) % g.nsuper, r = my / (G * g.nsuper) * G + my % G;
                        if (r < g.repl) {
                            const Segment sg = super_segment(a.starts, g, s, r);
                            mf = min(a.cur1[r * g.nsuper + s], sg.begin + sg.cap);
                            mb = sg.begin;
                            ms = s;
                        }
                    }
                    const unsigned hit = __ballot_sync(0xffffffffu, mf > mb);
                    if (hit != 0u) {
                        const int l = __ffs(hit) - 1;
                        w_super = __shfl_sync(0xffffffffu, ms, l);
                        w_at = __shfl_sync(0xffffffffu, mb, l);
                        w_fill = __shfl_sync(0xffffffffu, mf, l);
                        seg_next += (unsigned)(l + 1) * gridDim.x;
                        break;
                    }
                    seg_next += 32u * gridDim.x;
                }
            }
            if (lane == 0) {
                misc[16] = w_super;
                misc[17] = w_at;
                misc[18] = w_at < w_fill ? min((unsigned)PP, w_fill - w_at) : 0u;
            }
            if (w_at < w_fill) w_at += PP;
        }
        __syncthreads();
        n_in = (int)misc[18];
        if (n_in == 0) return;                      // the walk is over
        super = misc[16];
        first = misc[17];
    }


    // ---- load, bin, rank ------------------------------------------------------------------------------------
    float d[PPER][3], wv[PPER];
    unsigned br[PPER];                              // (bin << 12) | rank within the bin; 0xffffffff = no item
    unsigned long long dropped = 0;
    // all loads first: the shared-memory atomics of the ranking below are ordering points for the compiler, and a
    // load issued after one of them would expose a full DRAM latency per particle instead of one per CTA
#pragma unroll
    for (int q = 0; q < PPER; q++) {
        const int i = q * PT + tid;
        wv[q] = 1.0f;
        d[q][0] = d[q][1] = d[q][2] = 0.0f;
        if (i < n_in) {
            if (LEVEL == 1) {
                const float *p = a.pos + (first + i) * 3;
#pragma unroll
                for (int k = 0; k < 3; k++) d[q][k] = __ldg(p + k);
                if (WEIGHTED) wv[q] = __ldg(a.W + first + i);
            } else {
                const float4 v = __ldg(a.in + first + i);
                d[q][0] = v.x; d[q][1] = v.y; d[q][2] = v.z;
                wv[q] = v.w;
            }
        }
    }
    // tile of every particle
    unsigned tl[PPER], tc[PPER];
#pragma unroll
    for (int q = 0; q < PPER; q++) {
        const int i = q * PT + tid;
        tl[q] = NO_TILE;
        tc[q] = 0u;
        if (i < n_in) {
            if (LEVEL == 1) {
#pragma unroll
                for (int k = 0; k < 3; k++) d[q][k] = cell_coordinate(d[q][k], g.inv_cell_size);
            }
            int local[3];
            tl[q] = tile_and_local<MAS>(d[q], g, local, LEVEL == 1 ? &tc[q] : nullptr);
            if (tl[q] == NO_TILE) dropped += S * S * S;  // not routed to this slab: nothing of it is deposited here
        }
    }
    // ---- pass 1, spatially ordered input (lattice order, Peano-Hilbert order: snapshots usually are): the 4096
    // consecutive particles of a CTA sit in a small box of tiles, so their runs are long enough to go STRAIGHT to
    // the tile buckets (buf2) and pass 2 has nothing left to do for them.  The CTA takes the bounding box of its
    // particles' tile coordinates (relative to its first particle, periodic) and uses the position inside the box
    // as the bin; a box of more than DIRECT_SLOTS tiles (random order: always) keeps the super-tile bins.
    bool direct = false;
    int box_lo[3] = {0, 0, 0}, box_n[3] = {1, 1, 1}, ref[3] = {0, 0, 0};
    const int nt[3] = {g.ntx, g.nty, g.ntz};
    for (int i = tid; i < nb4; i += PT) cnt[i] = 0u;
    if (LEVEL == 1 && tid < 6) misc[24 + tid] = tid < 3 ? 0x7fffffffu : 0x80000000u;   // bounding box: 3 minima, 3 maxima (int)
    if (LEVEL == 1 && tid == 0) misc[30] = tc[0];           // reference tile of the box: the first particle's (any tile does)
    __syncthreads();
    bool ordered = false;
    if (LEVEL == 1 && g.ntx < 1024 && g.nty < 1024 && g.ntz < 1024) {
        // cheap vote first: in half of the warps at least three lanes share a tile (random order: none does).  The
        // largest group counts, not lane 0's: in lattice order lane 0 sits ON a tile boundary, and a coherent flow of
        // a cell or two towards lower z leaves it alone in the previous tile (38 % of the CTAs of BASELINE config 3
        // failed the lane-0 vote and took both passes)
        const int agree = __reduce_max_sync(0xffffffffu, __popc(__match_any_sync(0xffffffffu, tl[0])));
        ordered = __syncthreads_count(lane == 0 && agree >= 3) >= PT / 32 / 2;
    }
    if (ordered) {
        const unsigned rc = misc[30];
        ref[0] = (int)(rc >> 20); ref[1] = (int)((rc >> 10) & 1023u); ref[2] = (int)(rc & 1023u);
        int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
#pragma unroll
        for (int q = 0; q < PPER; q++) {
            if (tl[q] != NO_TILE) {
                const int c[3] = {(int)(tc[q] >> 20), (int)((tc[q] >> 10) & 1023u), (int)(tc[q] & 1023u)};
                unsigned packed = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    int r = c[k] - ref[k];                  // periodic distance from the reference tile
                    if (2 * r >= nt[k]) r -= nt[k];
                    else if (2 * r < -nt[k]) r += nt[k];
                    lo[k] = min(lo[k], r);
                    hi[k] = max(hi[k], r);
                    packed = (packed << 10) | (unsigned)(r + 512);
                }
                tc[q] = packed;                             // relative coordinates, biased by 512
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
            hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicMin(reinterpret_cast<int *>(misc) + 24 + k, lo[k]);
                atomicMax(reinterpret_cast<int *>(misc) + 27 + k, hi[k]);
            }
        }
        __syncthreads();
        long long vol = 1;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            box_lo[k] = reinterpret_cast<int *>(misc)[24 + k];
            box_n[k] = reinterpret_cast<int *>(misc)[27 + k] - box_lo[k] + 1;
            if (box_n[k] < 1 || box_n[k] > nt[k]) box_n[k] = DIRECT_SLOTS + 1;      // empty CTA / wraps onto itself
            vol *= box_n[k];
        }
        direct = vol <= DIRECT_SLOTS;
    }
    // tile of bin b of the box (direct mode)
    auto box_tile = [&](int b) -> unsigned {
        const int rz = b % box_n[2], ry = (b / box_n[2]) % box_n[1], rx = b / (box_n[2] * box_n[1]);
        int c[3] = {ref[0] + box_lo[0] + rx, ref[1] + box_lo[1] + ry, ref[2] + box_lo[2] + rz};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (c[k] < 0) c[k] += nt[k];
            else if (c[k] >= nt[k]) c[k] -= nt[k];
        }
        return ((unsigned)c[0] * g.nty + c[1]) * g.ntz + c[2];
    };
    if (direct) nb = box_n[0] * box_n[1] * box_n[2];
#pragma unroll
    for (int q = 0; q < PPER; q++) {
        int bin = -1;
        if (tl[q] != NO_TILE) {
            if (LEVEL == 2) {
                bin = (int)(tl[q] - (super << g.tps_shift));
            } else if (!direct) {
                bin = (int)(tl[q] >> g.tps_shift);
            } else {
                const int r0 = (int)(tc[q] >> 20) - 512 - box_lo[0], r1 = (int)((tc[q] >> 10) & 1023u) - 512 - box_lo[1],
                          r2 = (int)(tc[q] & 1023u) - 512 - box_lo[2];
                bin = (r0 * box_n[1] + r1) * box_n[2] + r2;
            }
        }
        // rank within the bin.  Ordered inputs (lattice / Peano-Hilbert order) send whole warps to one bin, and
        // same-address shared atomics serialise: when neighbouring lanes agree, the lanes of the leading bin share
        // one atomic (ballot + leader), at most twice, before the per-lane atomics.
        unsigned rank = 0;
        bool todo = bin >= 0;
        const int nbin = __shfl_down_sync(0xffffffffu, bin, 1);
        if (__popc(__ballot_sync(0xffffffffu, todo && nbin == bin && lane < 31)) >= 8) {
#pragma unroll 1
            for (int it = 0; it < 2; it++) {
                const unsigned pending = __ballot_sync(0xffffffffu, todo);
                if (pending == 0) break;
                const int leader = __ffs(pending) - 1;
                const int lb = __shfl_sync(0xffffffffu, bin, leader);
                const unsigned m = __ballot_sync(0xffffffffu, todo && bin == lb);
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(cnt + lb, (unsigned)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (todo && bin == lb) {
                    rank = base + __popc(m & ((1u << lane) - 1u));
                    todo = false;
                }
            }
        }
        if (todo) rank = atomicAdd(cnt + bin, 1u);
        br[q] = bin >= 0 ? (((unsigned)bin << 12) | rank) : 0xffffffffu;
    }
    __syncthreads();

    // ---- exclusive scan of the bin counts; one global reservation per non-empty bin -------------------------------
    constexpr int EPT = MAX_BINS / PT;          // 4 bins per thread
    unsigned c[EPT], got[EPT], end[EPT], run0;
    {
        unsigned sum = 0;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int b = tid * EPT + i;
            c[i] = b < nb ? cnt[b] : 0u;
            sum += c[i];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) misc[warp] = incl;
        __syncthreads();
        unsigned run = incl - sum;
#pragma unroll
        for (int w = 0; w < PT / 32; w++) run += (w < warp) ? misc[w] : 0u;
        run0 = run;
        // the reservations of this thread's bins are issued back to back and consumed only AFTER the records have
        // been staged: returning global atomics take microseconds under load, and the staging below needs nothing
        // but the CTA-local bin starts
        unsigned btile[EPT];
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int b = tid * EPT + i;
            got[i] = end[i] = btile[i] = 0u;
            if (b < nb && c[i] > 0) {
                if (LEVEL == 1 && direct) {
                    btile[i] = box_tile(b);
                    end[i] = __ldg(a.starts + btile[i] + 1);
                } else if (LEVEL == 1) {
                    const Segment sg = super_segment(a.starts, g, (unsigned)b, repl);
                    end[i] = sg.begin + sg.cap;
                } else {
                    end[i] = __ldg(a.starts + (super << g.tps_shift) + b + 1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int b = tid * EPT + i;
            if (b < nb && c[i] > 0) {
                if (LEVEL == 1 && direct) got[i] = atomicAdd(a.cur2 + btile[i], c[i]);
                else if (LEVEL == 1) got[i] = atomicAdd(a.cur1 + repl * g.nsuper + b, c[i]);
                else got[i] = atomicAdd(a.cur2 + (super << g.tps_shift) + b, c[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int b = tid * EPT + i;
            if (b < nb) cnt[b] = run;
            run += c[i];
        }
        if (tid == PT - 1) misc[20] = run;         // items staged by this CTA
    }
    __syncthreads();

    // ---- stage in bin order -------------------------------------------------------------------------------------
    float wabs = 0.0f;
#pragma unroll
    for (int q = 0; q < PPER; q++) {
        if (br[q] != 0xffffffffu) {
            const unsigned bin = br[q] >> 12;
            const unsigned dst = cnt[bin] + (br[q] & 0xfffu);
            stage[dst] = make_float4(d[q][0], d[q][1], d[q][2], wv[q]);
            sbin[dst] = (unsigned short)bin;
            if (WEIGHTED && LEVEL == 1) wabs += fabsf(wv[q]);
        }
    }
    if (WEIGHTED && LEVEL == 1) {
        // scale of the fixed-point accumulators of the tile kernel: sum of |W| of the particles deposited here
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wabs += __shfl_xor_sync(0xffffffffu, wabs, o);
        if (lane == 0 && wabs != 0.0f) atomicAdd(a.wsum, (double)wabs);
    }
    {
        // now the reservations: global slot of sorted index 0 of every bin, and the end of its segment
        unsigned run = run0;
#pragma unroll
        for (int i = 0; i < EPT; i++) {
            const int b = tid * EPT + i;
            if (b < nb && c[i] > 0) {
                goff[b] = got[i] - run;
                lim[b] = end[i];
            }
            run += c[i];
        }
    }
    __syncthreads();

    // ---- runs out: consecutive lanes write consecutive slots of a bin -------------------------------------------------
    const int n_out = (int)misc[20];
    for (int j = tid; j < n_out; j += PT) {
        const unsigned b = sbin[j];
        const unsigned slot = goff[b] + (unsigned)j;
        const float4 v = stage[j];
        if (slot < lim[b]) (LEVEL == 1 && direct ? a.out2 : a.out)[slot] = v;
        else overflow_deposit<MAS, WEIGHTED>(v.x, v.y, v.z, v.w, g, a.number, &dropped);
    }
    if (a.dropped != nullptr && dropped != 0) atomicAdd(a.dropped, dropped);
    if (LEVEL == 1) return;
    __syncthreads();                                // the next chunk reuses the staging area and the bin tables
    }
}

// ---- 5. per-tile deposit ------------------------------------------------------------------------------------------
// Accumulators are 64-bit FIXED-POINT sums held as two 32-bit words per cell (tile + stencil halo) and fed with
// native 32-bit shared-memory atomics (ATOMS.ADD: 3.6 cycles per warp instruction, the price of a plain
// read-modify-write; 64-bit and float shared atomics are CAS loops).  A contribution c = fl32(wx*wy*wz*W) becomes
// q = rint(c * 2^k) with 2^k ~ 2^26 / mean|W|: one returning atomic on the low word, the carry (and the high word of
// a large |q|) goes to the high word -- rarely.  Integer addition is associative, so the tile's sums do not depend on
// the order the particles arrive in (bit-reproducible), and they are exact to 2^-26 of a typical contribution before
// the single rounding to float32 when the tile is added to the grid.  Unweighted NGP stays bit-exact (q = 2^26).
template <int MAS>
struct TileAcc {
    static constexpr int S = MAS + 1;
    static constexpr int AX = TX + S - 1, AY = TY + S - 1;
    static constexpr int ROWS = AX * AY;
    static constexpr int CELLS = ROWS * OUT_ROW;               // OUT_ROW floats per row: rows are flushed in place
    static constexpr size_t bytes = (size_t)CELLS * 8 + 16;
    static_assert(TZ + S - 1 <= OUT_ROW, "row too short");
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MAS, bool WEIGHTED>
__global__ void __launch_bounds__(TNT)
tile_deposit_kernel(const float4 *__restrict__ bucket, const unsigned *__restrict__ starts,
                    const unsigned *__restrict__ cur2, float *__restrict__ number, TileGeom g,
                    const double *__restrict__ wsum, double particles, const unsigned *__restrict__ n_dev,
                    int bulk_ok) {
    using TA = TileAcc<MAS>;
    constexpr int S = TA::S, AY = TA::AY;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *lo = reinterpret_cast<unsigned *>(smem_raw);            // CELLS low words; later the float rows to flush
    unsigned *hi = lo + TA::CELLS;                                    // CELLS high words

    const unsigned tile = blockIdx.x;
    const unsigned begin = __ldg(starts + tile);
    const unsigned end = min(__ldg(cur2 + tile), __ldg(starts + tile + 1));      // overflow went the direct way
    if (begin >= end) return;                                   // empty tile: nothing to add

    const int tid = threadIdx.x;
    const int tz = tile % g.ntz, ty = (tile / g.ntz) % g.nty, tx = tile / (g.ntz * g.nty);
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;

    // power-of-two scale: a typical contribution lands near 2^26 (exact scaling; see the header comment)
    int kexp = 26;
    if (WEIGHTED) {
        if (n_dev != nullptr) particles = fmin(particles, (double)__ldg(n_dev));
        const double mean = __ldg(wsum) / fmax(particles, 1.0);
        int e = 0;
        if (mean > 0.0 && mean < 1e300) frexp(mean, &e);         // mean = f * 2^e, f in [0.5, 1)
        kexp = 26 - e;
        kexp = kexp > 120 ? 120 : (kexp < -120 ? -120 : kexp);
    }
    const float scale = __int_as_float((127 + kexp) << 23);
    const double inv_scale = __longlong_as_double((long long)(1023 - kexp) << 52);

    for (int i = tid; i < TA::CELLS / 2; i += TNT) reinterpret_cast<uint4 *>(lo)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    // the next record is requested before the current one is accumulated (the shared-memory atomics below are
    // ordering points: without this every iteration would wait for its own load)
    float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
    if (begin + tid < end) nxt = __ldg(bucket + begin + tid);
    for (unsigned i = begin + tid; i < end; i += TNT) {
        const float4 v = nxt;
        if (i + TNT < end) nxt = __ldg(bucket + i + TNT);
        const float dd[3] = {v.x, v.y, v.z};
        int lc[3];
        float w[3][S];
#pragma unroll
        for (int k = 0; k < 3; k++) axis_weights<MAS>(dd[k], axis_base<MAS>(dd[k]), w[k]);
        tile_and_local<MAS>(dd, g, lc);
        const float ws = WEIGHTED ? v.w * scale : scale;
#pragma unroll
        for (int l = 0; l < S; l++) w[0][l] *= ws;
        const int cell = (lc[0] * AY + lc[1]) * OUT_ROW + lc[2];
        // one x-plane of the stencil at a time: its S*S low-word atomics are issued back to back (no branch between
        // them, so their latencies overlap); the carries and the rare non-zero high words follow under ONE branch
#pragma unroll
        for (int l = 0; l < S; l++) {
            unsigned old[S][S], ql[S][S];
            int qh[S][S];
#pragma unroll
            for (int m = 0; m < S; m++) {
                const float wxy = w[0][l] * w[1][m];
#pragma unroll
                for (int n = 0; n < S; n++) {
                    const long long q = __float2ll_rn(wxy * w[2][n]);
                    ql[m][n] = (unsigned)q;
                    qh[m][n] = (int)(q >> 32);
                }
            }
#pragma unroll
            for (int m = 0; m < S; m++) {
#pragma unroll
                for (int n = 0; n < S; n++) old[m][n] = atomicAdd(lo + cell + (l * AY + m) * OUT_ROW + n, ql[m][n]);
            }
            int any = 0;
#pragma unroll
            for (int m = 0; m < S; m++) {
#pragma unroll
                for (int n = 0; n < S; n++) {
                    qh[m][n] += (old[m][n] + ql[m][n]) < old[m][n] ? 1 : 0;
                    any |= qh[m][n];
                }
            }
            if (any != 0) {
#pragma unroll
                for (int m = 0; m < S; m++) {
#pragma unroll
                    for (int n = 0; n < S; n++)
                        if (qh[m][n] != 0) atomicAdd(hi + cell + (l * AY + m) * OUT_ROW + n, (unsigned)qh[m][n]);
                }
            }
        }
    }
    __syncthreads();

    // ---- fixed point -> float32, in place: lo[] becomes the rows to add to the grid ------------------------------------
    float *outst = reinterpret_cast<float *>(lo);
    for (int i = tid; i < TA::CELLS; i += TNT) {
        const long long q = (long long)(((unsigned long long)hi[i] << 32) | lo[i]);
        outst[i] = (float)((double)q * inv_scale);               // exact scaling, one rounding
    }
    const int dims = g.dims;
    const int lane = tid & 31, warp = tid >> 5;
    if (bulk_ok) {
        // make the generic-proxy writes of the rows visible to the bulk-copy engine
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid < TA::ROWS) {
            const int r = tid;
            const int ax = r / AY, ay = r - ax * AY;
            int gx = ox + ax, gy = oy + ay;                    // gx: plane inside the destination buffer
            if (gx >= g.x_planes) gx -= g.x_planes;            // only when the buffer is the whole periodic grid
            if (gy >= dims) gy -= dims;
            const uint4 *rowv = reinterpret_cast<const uint4 *>(outst + r * OUT_ROW);
            unsigned nz = 0;
#pragma unroll
            for (int k = 0; k < OUT_ROW / 4; k++) {
                const uint4 t = rowv[k];
                nz |= (t.x | t.y | t.z | t.w) << 1;             // (the shift drops the sign of a -0.0)
            }
            // rows beyond a partial tile's extent and untouched rows carry only zeros: skipped
            if (nz != 0 && gx < g.x_planes && gy < dims) {
                float *row = number + ((int64_t)gx * dims + gy) * dims;
                const unsigned src = smem_u32(outst + r * OUT_ROW);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(row + oz), "r"(src), "r"(TZ * 4) : "memory");
                if (S > 1) {
                    int gz = oz + TZ;
                    if (gz >= dims) gz -= dims;
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 ::"l"(row + gz), "r"(src + TZ * 4), "r"(16) : "memory");
                }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        for (int r = warp; r < TA::ROWS; r += TNT / 32) {
            const int ax = r / AY, ay = r - ax * AY;
            int gx = ox + ax, gy = oy + ay;
            if (gx >= g.x_planes) gx -= g.x_planes;
            if (gy >= dims) gy -= dims;
            if (gx >= g.x_planes || gy >= dims) continue;       // beyond a partial tile: zeros only
            float *row = number + ((int64_t)gx * dims + gy) * dims;
            for (int az = lane; az < TZ + S - 1; az += 32) {
                const float val = outst[r * OUT_ROW + az];
                int gz = oz + az;
                if (gz >= dims) gz -= dims;
                if (val != 0.0f && gz < dims) atomicAdd(row + gz, val);
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------
static TileGeom make_geom(int dims, float BoxSize, int64_t particles, int x_origin = 0, int x_own = -1,
                          int x_planes = -1) {
    TileGeom g;
    g.dims = dims;
    g.x_origin = x_origin;
    g.x_own = x_own < 0 ? dims : x_own;
    g.x_planes = x_planes < 0 ? dims : x_planes;
    g.ntx = (g.x_own + TX - 1) / TX;
    g.nty = (dims + TY - 1) / TY;
    g.ntz = (dims + TZ - 1) / TZ;
    g.ntiles = (unsigned)g.ntx * g.nty * g.ntz;
    g.inv_cell_size = (float)dims / BoxSize;      // float32 division, MAS_library.pyx:135
    // both passes get about sqrt(ntiles) bins (a power of two tiles per super-tile)
    int sh = 0;
    while (((uint64_t)1 << (2 * sh)) < g.ntiles) sh++;
    while ((g.ntiles + ((1u << sh) - 1)) >> sh > (unsigned)MAX_BINS) sh++;
    g.tps_shift = sh;
    g.nsuper = (g.ntiles + ((1u << sh) - 1)) >> sh;
    // one cursor replica per 64 first-pass CTAs: a cursor then serves few enough same-address atomics, and every
    // replica still receives an even share of the particles
    const int64_t ctas = (particles + PP - 1) / PP;
    int64_t r = ctas / 64;
    g.repl = (unsigned)(r < 1 ? 1 : (r > MAX_REPL ? MAX_REPL : r));
    return g;
}

static size_t scan_temp_bytes(unsigned n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned *)nullptr, (unsigned *)nullptr, (int)n);
    return bytes;
}

// upper bound of sum_t tile_capacity(c_t) given sum_t c_t <= particles/SAMPLE + 4 (Cauchy-Schwarz on
// the sqrt term); the partition kernels can therefore never write past the bucket planes
static size_t bucket_slots_bound(int64_t particles, unsigned ntiles) {
    const double sampled = (double)particles / SAMPLE + 8.0;
    const double bound = 1.125 * SAMPLE * sampled + (4.0 * SAMPLE + 1.0) * sqrt((double)ntiles * sampled) +
                         38.0 * (double)ntiles;
    return ((size_t)bound + 1024 + 3) & ~(size_t)3;
}

bool deposit_tiled_supported(int mas, int64_t particles, int dims, int axes, int x_own) {
    (void)mas;
    if (axes != 3 || dims < 64) return false;                       // halo wrap assumes dims >> stencil
    const TileGeom g = make_geom(dims, 1.0f, particles, 0, x_own, x_own);
    if ((uint64_t)g.ntx * g.nty * g.ntz > (uint64_t)MAX_BINS * MAX_BINS) return false;
    if (g.tps_shift > 11) return false;                             // more than MAX_BINS tiles per super-tile
    // 32-bit slots; a cursor may run past its bucket by the number of overflowing particles
    if (bucket_slots_bound(particles, g.ntiles) + (size_t)particles >= ((size_t)1 << 32)) return false;
    return particles >= (int64_t)g.ntiles * 64;                     // sparse inputs: per-tile overhead loses
}

struct TiledWorkspace {
    float4 *buf1, *buf2;          // `slots` records (dist.x, dist.y, dist.z, W) each
    size_t slots;
    double *wsum;                 // sum of |W| of this call's particles
    unsigned *starts;             // ntiles + 1 : sampled counts -> capacities -> exclusive scan
    unsigned *cur2;               // ntiles : next free slot of a tile's bucket
    unsigned *cur1;               // MAX_REPL x nsuper : next free slot of a super-tile sub-segment
    void *scan_tmp;
    size_t scan_bytes;
    size_t total;
};

static TiledWorkspace carve(void *ws, int64_t particles, const TileGeom &g) {
    TiledWorkspace w;
    char *base = reinterpret_cast<char *>(ws);
    size_t off = 0;
    w.slots = bucket_slots_bound(particles, g.ntiles);
    w.buf1 = reinterpret_cast<float4 *>(base + off);
    off += align_up(w.slots * 16, 256);
    w.buf2 = reinterpret_cast<float4 *>(base + off);
    off += align_up(w.slots * 16, 256);
    w.wsum = reinterpret_cast<double *>(base + off);
    off += 256;
    w.starts = reinterpret_cast<unsigned *>(base + off);
    off += align_up(((size_t)g.ntiles + 1) * 4, 256);
    w.cur2 = reinterpret_cast<unsigned *>(base + off);
    off += align_up((size_t)g.ntiles * 4, 256);
    w.cur1 = reinterpret_cast<unsigned *>(base + off);
    off += align_up((size_t)MAX_REPL * g.nsuper * 4, 256);
    w.scan_tmp = base + off;
    w.scan_bytes = scan_temp_bytes(g.ntiles + 1);
    off += align_up(w.scan_bytes, 256);
    w.total = off;
    return w;
}

size_t deposit_tiled_workspace(int mas, int64_t particles, int dims, int axes, int mode, int x_own) {
    (void)mas; (void)axes; (void)mode;
    const TileGeom g = make_geom(dims, 1.0f, particles, 0, x_own, x_own);
    return carve(nullptr, particles, g).total;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    PYL_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PYL_OK;
}

template <int MAS, bool WEIGHTED>
static int run_tiled_w(const float *pos, float *number, const float *W, int64_t particles, const TileGeom &g,
                       const TiledWorkspace &w, unsigned long long *dropped, const unsigned *n_dev,
                       cudaStream_t stream) {
    static bool attr_done = false;
    if (!attr_done) {
        int st = set_smem(partition_kernel<MAS, WEIGHTED, 1>, part_smem_bytes(MAX_BINS));
        if (st == PYL_OK) st = set_smem(partition_kernel<MAS, WEIGHTED, 2>, part_smem_bytes(MAX_BINS));
        if (st == PYL_OK) st = set_smem(tile_deposit_kernel<MAS, WEIGHTED>, TileAcc<MAS>::bytes);
        if (st != PYL_OK) return st;
        attr_done = true;
    }
    PartArgs a;
    a.pos = pos; a.W = W; a.particles = particles; a.in = w.buf1; a.out = w.buf1; a.out2 = w.buf2;
    a.starts = w.starts; a.cur1 = w.cur1; a.cur2 = w.cur2; a.number = number; a.dropped = dropped;
    a.wsum = w.wsum;
    a.n_dev = n_dev;
    if (WEIGHTED) PYL_CUDA_CHECK(cudaMemsetAsync(w.wsum, 0, 8, stream));
    const unsigned ctas1 = (unsigned)((particles + PP - 1) / PP);
    partition_kernel<MAS, WEIGHTED, 1><<<ctas1, PT, part_smem_bytes((int)g.nsuper), stream>>>(a, g);
    PYL_LAUNCH_CHECK();
    a.out = w.buf2;
    const unsigned ctas2 = 2u * (unsigned)sm_count();          // persistent: two CTAs per SM walk the segments
    partition_kernel<MAS, WEIGHTED, 2><<<ctas2, PT, part_smem_bytes(1 << g.tps_shift), stream>>>(a, g);
    PYL_LAUNCH_CHECK();
    // bulk reductions need 16-byte aligned 128-byte rows: whole tiles along z and an aligned grid
    const int bulk_ok = (g.dims % TZ == 0) && ((reinterpret_cast<uintptr_t>(number) & 15) == 0);
    tile_deposit_kernel<MAS, WEIGHTED><<<g.ntiles, TNT, TileAcc<MAS>::bytes, stream>>>(
        w.buf2, w.starts, w.cur2, number, g, w.wsum, (double)particles, n_dev, bulk_ok);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

template <int MAS>
static int run_tiled(const float *pos, float *number, const float *W, int64_t particles, int dims,
                     float BoxSize, int x_origin, int x_own, int x_planes, unsigned long long *dropped,
                     void *ws, const unsigned *n_dev, cudaStream_t stream) {
    const TileGeom g = make_geom(dims, BoxSize, particles, x_origin, x_own, x_planes);
    const TiledWorkspace w = carve(ws, particles, g);

    const int vec_ok = ((reinterpret_cast<uintptr_t>(pos) & 15) == 0) &&
                       (W == nullptr || (reinterpret_cast<uintptr_t>(W) & 15) == 0);
    int64_t blocks = ((particles + 3) / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int64_t cblocks = (blocks + SAMPLE - 1) / SAMPLE;
    if (cblocks < 1) cblocks = 1;

    PYL_CUDA_CHECK(cudaMemsetAsync(w.starts, 0, ((size_t)g.ntiles + 1) * 4, stream));
    tile_count_kernel<MAS><<<(int)cblocks, 256, 0, stream>>>(pos, particles, g, w.starts, vec_ok, n_dev);
    PYL_LAUNCH_CHECK();
    tile_caps_kernel<<<(g.ntiles + 1 + 255) / 256, 256, 0, stream>>>(w.starts, g.ntiles);
    PYL_LAUNCH_CHECK();
    PYL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.scan_tmp, const_cast<size_t &>(w.scan_bytes), w.starts, w.starts,
                                                 (int)(g.ntiles + 1), stream));
    tile_setup_kernel<<<(g.ntiles + 1023) / 1024, 1024, 0, stream>>>(w.starts, g, w.cur1, w.cur2);
    PYL_LAUNCH_CHECK();
    if (W) return run_tiled_w<MAS, true>(pos, number, W, particles, g, w, dropped, n_dev, stream);
    return run_tiled_w<MAS, false>(pos, number, W, particles, g, w, dropped, n_dev, stream);
}

// x_own < 0: whole periodic grid.  Otherwise the slab window of pyl_deposit_slab.
int deposit_tiled(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                  float BoxSize, int x_origin, int x_own, int x_planes, int64_t *dropped, void *ws,
                  cudaStream_t stream, const unsigned *n_dev) {
    unsigned long long *dr = reinterpret_cast<unsigned long long *>(dropped);
    switch (mas) {
        case PYL_MAS_NGP: return run_tiled<PYL_MAS_NGP>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, n_dev, stream);
        case PYL_MAS_CIC: return run_tiled<PYL_MAS_CIC>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, n_dev, stream);
        case PYL_MAS_TSC: return run_tiled<PYL_MAS_TSC>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, n_dev, stream);
        case PYL_MAS_PCS: return run_tiled<PYL_MAS_PCS>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, n_dev, stream);
    }
    set_last_error("deposit_tiled: unknown scheme %d", mas);
    return PYL_ERR_ARG;
}

}  // namespace pyl
