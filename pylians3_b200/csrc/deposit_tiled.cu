// K0/K1/K2: tiled deposit -- bucket particles by grid tile, accumulate every tile in shared memory
// WITHOUT atomics, flush each tile once with coalesced reductions.
//
// Replaces the same reference loops as deposit_atomic.cu (MAS_library.pyx:142-166, 288-292, 388-404,
// 481-497 and W variants).  Why: measured on B200 (profiles/r1_deposit_atomic.md) the particle-parallel
// red.global kernel moves 21x the algorithmic bytes through DRAM (every stencil row of a randomly
// placed particle is a 32-byte sector read-modify-write) and runs at 1.8% of the HBM roofline.
//
// Pipeline (all on the caller's stream, scratch in the caller's workspace):
//   1. tile_count    histogram of a 1-in-8 SAMPLE of the particles over tiles (tile = 8 x 16 x 32 cells):
//                    cell coordinate dist = fl32(pos*inv) per axis, stencil base cell -> tile id.
//   2. tile_caps + exclusive scan (cub::DeviceScan): per-tile bucket capacity = 1.125 x estimate +
//                    4 sigma of the sampling noise + 32, and the bucket start offsets.
//   3. tile_scatter  ONE full pass over pos: one 64-bit atomicAdd on the tile's packed cursor returns the slot
//                    and the bucket end; writes (dist.xyz, W) as one aligned float4 into the tile's bucket.  A particle that finds its bucket full (rare:
//                    beyond 4 sigma) is deposited on the spot with red.global -- correctness never
//                    depends on the estimate.
//   4. tile_deposit  one CTA per tile, looping over the bucket in chunks of 1024 particles:
//        a. counting sort of the chunk by local cell (x,y,z) inside shared memory: one packed-u16
//           shared atomic per particle for the rank, block scan, scatter into a sorted float4 array
//           (stencil fractions + W) -- so every (x,y) row of the tile is contiguous and z-ordered;
//        b. each warp OWNS target x-planes of the shared accumulator (tile + stencil halo).  For a
//           target plane X it walks the source rows x = X - l (l = 0..S-1), 32 particles at a time;
//           lanes of equal cell form contiguous runs (sorted), so a segmented warp-shuffle scan
//           leaves each run's sum in its head lane, and head lanes do plain shared-memory
//           read-modify-writes.  Target planes are exclusive to a warp and a warp's instructions are
//           ordered, hence no shared atomics and no block barriers inside the stencil loops;
//        c. after the last chunk the accumulator (tile + halo) is added to the grid with coalesced
//           red.global.add.f32 (halo cells are shared with neighbouring tiles).
// Weights: fractions are taken against the same unwrapped base cell the reference uses; CIC is
// operation-identical to the reference, TSC/PCS evaluate the same polynomials in float32 (<= 3 ulp from
// the reference's float64-then-rounded values, far inside the 1e-5 per-cell tolerance).  Unweighted
// NGP stays bit-exact (sums of 1.0f).
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "deposit_point.cuh"
#include "stencil.cuh"

namespace pyl {

constexpr int TX = 8, TY = 16, TZ = 32;         // tile extent in cells
constexpr int TILE_CELLS = TX * TY * TZ;        // 4096
constexpr int TNT = 256;                        // threads per tile CTA
constexpr int CHUNK = 1024;                     // particles sorted per pass
constexpr int PER = CHUNK / TNT;                // 4 per thread
constexpr int SAMPLE = 8;                       // tile_count looks at one particle group in SAMPLE

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
};

constexpr unsigned NO_TILE = 0xffffffffu;

// unwrapped base cell of the stencil (first of the S cells) and the fraction the weights depend on
template <int MAS>
__device__ __forceinline__ int stencil_base(float dist, float &frac) {
    int b;
    if (MAS == PYL_MAS_NGP) {
        b = __double2int_rz(__dadd_rn((double)dist, 0.5));
        frac = 0.0f;
    } else if (MAS == PYL_MAS_CIC) {
        b = __float2int_rz(dist);
        frac = __fsub_rn(dist, (float)b);                 // u in [0,1)
    } else if (MAS == PYL_MAS_TSC) {
        b = __double2int_rd(__dadd_rn((double)dist, -1.5)) + 1;
        frac = __fsub_rn(dist, (float)b);                 // r in [0.5,1.5)
    } else {
        b = __double2int_rd(__dadd_rn((double)dist, -2.0)) + 1;
        frac = __fsub_rn(dist, (float)b);                 // 1+u in [1,2)
    }
    return b;
}

// all S weights of one axis from the fraction (see header comment)
template <int MAS>
__device__ __forceinline__ void stencil_weights(float frac, float *w) {
    if (MAS == PYL_MAS_NGP) {
        w[0] = 1.0f;
    } else if (MAS == PYL_MAS_CIC) {
        w[0] = __fsub_rn(1.0f, frac);
        w[1] = frac;
    } else if (MAS == PYL_MAS_TSC) {
        // diffs: r, |r-1|, 2-r   (MAS_library.pyx:394-398)
        const float a = 1.5f - frac;
        const float c = frac - 0.5f;
        const float d1 = frac - 1.0f;
        w[0] = 0.5f * a * a;
        w[1] = 0.75f - d1 * d1;
        w[2] = 0.5f * c * c;
    } else {
        // diffs: 1+u, u, 1-u, 2-u   (MAS_library.pyx:487-491)
        const float u = frac - 1.0f;
        const float v = 1.0f - u;
        const float sixth = 1.0f / 6.0f;
        w[0] = v * v * v * sixth;
        w[1] = (4.0f - 6.0f * u * u + 3.0f * u * u * u) * sixth;
        w[2] = (4.0f - 6.0f * v * v + 3.0f * v * v * v) * sixth;
        w[3] = u * u * u * sixth;
    }
}

template <int MAS>
__device__ __forceinline__ unsigned tile_and_local(const float d[3], const TileGeom &g, int local[3],
                                                   float frac[3]) {
    int t[3];
    const int T[3] = {TX, TY, TZ};
    bool mine = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const int b = stencil_base<MAS>(d[a], frac[a]);
        int wb = wrap_index(b, g.dims);
        if (a == 0) {                       // plane index inside the destination window
            wb -= g.x_origin;
            if (wb < 0) wb += g.dims;
            mine = wb < g.x_own;
        }
        t[a] = wb / T[a];
        local[a] = wb - t[a] * T[a];
    }
    return mine ? ((unsigned)t[0] * g.nty + t[1]) * g.ntz + t[2] : NO_TILE;
}

// ---- 1. sampled histogram --------------------------------------------------------------------------
template <int MAS>
__global__ void __launch_bounds__(256) tile_count_kernel(const float *__restrict__ pos, int64_t particles,
                                                         TileGeom g, unsigned *__restrict__ counts, int vec_ok) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t groups = vec_ok ? (particles >> 2) : 0;
    int local[3];
    float frac[3];
    // every SAMPLE-th group of 4 particles
    for (int64_t sg = tid; sg * SAMPLE < groups; sg += stride) {
        const int64_t grp = sg * SAMPLE;
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
            const unsigned t = tile_and_local<MAS>(d, g, local, frac);
            if (t != NO_TILE) atomicAdd(counts + t, 1u);
        }
    }
    // unaligned input / tail: every SAMPLE-th particle
    for (int64_t si = tid; (groups << 2) + si * SAMPLE < particles; si += stride) {
        const int64_t i = (groups << 2) + si * SAMPLE;
        float d[3];
#pragma unroll
        for (int a = 0; a < 3; a++) d[a] = cell_coordinate(__ldg(pos + i * 3 + a), g.inv_cell_size);
        const unsigned t = tile_and_local<MAS>(d, g, local, frac);
        if (t != NO_TILE) atomicAdd(counts + t, 1u);
    }
}

// ---- 2. capacities from the sampled counts (in place; then scanned into bucket starts) ---------------
__host__ __device__ __forceinline__ unsigned tile_capacity(unsigned sampled) {
    // 1.125 x estimate + 4 sigma (sigma of the estimate = SAMPLE*sqrt(sampled)) + 32
    const unsigned est = sampled * SAMPLE;
    const unsigned sig = (unsigned)(4.0f * SAMPLE * sqrtf((float)sampled)) + 1u;
    return est + (est >> 3) + sig + 32u;
}

__global__ void tile_caps_kernel(unsigned *__restrict__ counts, unsigned ntiles) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) counts[i] = tile_capacity(counts[i]);
    else if (i == ntiles) counts[i] = 0u;
}

// cursor[t] = (bucket end << 32) | next free slot: ONE 64-bit atomicAdd in the scatter kernel returns both the
// slot and the capacity limit (measured: the two extra loads of `starts` per particle were 25% of the scatter
// kernel's L2 requests, and the kernel is bound by the L2 request rate)
__global__ void tile_cursor_kernel(const unsigned *__restrict__ starts, unsigned long long *__restrict__ cursor,
                                   unsigned ntiles) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) cursor[i] = ((unsigned long long)starts[i + 1] << 32) | starts[i];
}

// ---- 3. scatter into buckets (single full pass) --------------------------------------------------------
// bucket full (capacity came from a sample): deposit the particle directly.  Out of line: it is rare, and
// inlined four times it made every atomicAdd below wait for the previous particle's branch.
template <int MAS, bool WEIGHTED>
__device__ __noinline__ void scatter_overflow(float d0, float d1, float d2, float wp, TileGeom g,
                                              float *__restrict__ number, unsigned long long *dropped) {
    const float d[3] = {d0, d1, d2};
    unsigned long long dr = 0;
    if (g.x_planes == g.dims)
        deposit_dist<MAS, 3, WEIGHTED, false>(d, wp, number, g.dims, SlabWindow{0, g.dims, 0.0f}, dr);
    else
        deposit_dist<MAS, 3, WEIGHTED, true>(d, wp, number, g.dims, SlabWindow{g.x_origin, g.x_planes, 0.0f}, dr);
    *dropped += dr;
}

// One particle per thread, scalar loads: measured on B200 (scratch/s1_bench.cu) this simplest form beats the
// float4 / 4-particles-per-thread form (2.85 vs 3.25 ms at 512^3) -- the kernel is bound by the rate of
// returning atomics (1.5 ms for 134 M of them) plus the scattered 16-byte stores, and more independent
// threads keep more of both in flight.  A two-level shared-memory-staged partition was prototyped in the
// same file and lost (3.2 ms).
// Lanes of a warp that go to the same tile share ONE atomic (match_any + leader): snapshot-ordered inputs
// (lattice order, Peano-Hilbert order) send whole warps to one tile, and same-address returning atomics
// serialise in L2 (measured 9.7 ms instead of 2.8 ms on a Zel'dovich-displaced lattice without this).
template <int MAS, bool WEIGHTED>
__global__ void __launch_bounds__(512) tile_scatter_kernel(const float *__restrict__ pos,
                                                           const float *__restrict__ W, int64_t particles,
                                                           TileGeom g, unsigned long long *__restrict__ cursor,
                                                           float4 *__restrict__ bucket,
                                                           float *__restrict__ number,
                                                           unsigned long long *__restrict__ dropped_out) {
    constexpr int S = StencilWidth<MAS>::value;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < particles;
    const int lane = threadIdx.x & 31;
    float d[3] = {0.f, 0.f, 0.f};
    float wp = 1.0f;
    unsigned t = NO_TILE;
    if (live) {
#pragma unroll
        for (int a = 0; a < 3; a++) d[a] = cell_coordinate(__ldg(pos + i * 3 + a), g.inv_cell_size);
        if (WEIGHTED) wp = __ldg(W + i);
        int local[3];
        float frac[3];
        t = tile_and_local<MAS>(d, g, local, frac);
    }
    // match_any only when neighbouring lanes agree (ordered input); for scattered input it would cost ~8%
    unsigned peers = 1u << lane;
    if (__any_sync(0xffffffffu, __shfl_down_sync(0xffffffffu, t, 1) == t && lane < 31))
        peers = __match_any_sync(0xffffffffu, t);
    const int leader = __ffs(peers) - 1;
    const unsigned rank = __popc(peers & ((1u << lane) - 1u));
    unsigned long long cur = 0ull;
    if (lane == leader && t != NO_TILE) cur = atomicAdd(cursor + t, (unsigned long long)__popc(peers));
    cur = __shfl_sync(0xffffffffu, cur, leader);
    const unsigned slot = (unsigned)cur + rank, end = (unsigned)(cur >> 32);
    unsigned long long dropped = 0;
    if (t != NO_TILE) {
        if (slot < end) bucket[slot] = make_float4(d[0], d[1], d[2], wp);
        else scatter_overflow<MAS, WEIGHTED>(d[0], d[1], d[2], wp, g, number, &dropped);
    } else if (live) {
        dropped = S * S * S;            // not routed to this slab: nothing of it is deposited here
    }
    if (dropped_out != nullptr && dropped != 0) atomicAdd(dropped_out, dropped);
}

// ---- 4. per-tile deposit ----------------------------------------------------------------------------
template <int MAS>
struct TileSmem {
    static constexpr int S = MAS + 1;
    static constexpr int AX = TX + S - 1, AY = TY + S - 1, AZ = TZ + S - 1;
    static constexpr int ACC = AX * AY * AZ;
    static constexpr int CNT_WORDS = TILE_CELLS / 2 + 1;      // packed u16 pairs + one end marker
    static constexpr size_t bytes =
        (size_t)ACC * 4 + (size_t)CNT_WORDS * 4 + (size_t)CHUNK * 16 + (size_t)CHUNK * 2 + 64;
};

__device__ __forceinline__ unsigned off16(const unsigned *cnt, int key) {
    const unsigned w = cnt[key >> 1];
    return (key & 1) ? (w >> 16) : (w & 0xffffu);
}

template <int MAS>
__global__ void __launch_bounds__(TNT, 4)
tile_deposit_kernel(const float4 *__restrict__ bucket, const unsigned long long *__restrict__ cursor,
                    float *__restrict__ number, TileGeom g) {
    using SM = TileSmem<MAS>;
    constexpr int S = SM::S, AY = SM::AY, AZ = SM::AZ, AX = SM::AX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sorted = reinterpret_cast<float4 *>(smem_raw);                       // CHUNK float4
    float *acc = reinterpret_cast<float *>(smem_raw + (size_t)CHUNK * 16);        // ACC floats
    unsigned *cnt = reinterpret_cast<unsigned *>(acc + SM::ACC);                  // CNT_WORDS
    unsigned short *sorted_yz = reinterpret_cast<unsigned short *>(cnt + SM::CNT_WORDS);   // CHUNK: y*TZ+z
    unsigned *warp_part = reinterpret_cast<unsigned *>(sorted_yz + CHUNK);        // 8 words

    const unsigned tile = blockIdx.x;
    // cursor = (bucket end << 32) | (bucket begin + particles that asked for a slot); the next tile's bucket
    // begins where this one ends, so this tile's begin is the previous tile's end
    const unsigned long long cur = cursor[tile];
    const unsigned begin = tile == 0 ? 0u : (unsigned)(cursor[tile - 1] >> 32);
    const unsigned end = min((unsigned)cur, (unsigned)(cur >> 32));            // overflow went the direct way
    if (begin == end) return;                                 // empty tile: nothing to add

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tz = tile % g.ntz, ty = (tile / g.ntz) % g.nty, tx = tile / (g.ntz * g.nty);
    const int ox = tx * TX, oy = ty * TY, oz = tz * TZ;

    for (int i = tid; i < SM::ACC; i += TNT) acc[i] = 0.0f;

    // the next chunk's particles are fetched while the current chunk is accumulated (the exposed latency of
    // this load was 18% of the kernel's stall samples)
    float4 nxt[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) {
        const unsigned i = begin + tid + q * TNT;
        nxt[q] = i < end ? __ldg(bucket + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (unsigned c0 = begin; c0 < end; c0 += CHUNK) {
        const int n = (int)min((unsigned)CHUNK, end - c0);
        for (int i = tid; i < SM::CNT_WORDS; i += TNT) cnt[i] = 0u;
        __syncthreads();

        // ---- a. rank every particle within its cell ------------------------------------------------
        // sort key: source plane, then the shared-memory BANK of the particle's first target cell, then y.
        // Lanes of one accumulation batch sit nb sorted slots apart (below), i.e. in different banks most
        // of the time: the read-modify-writes of a batch then need ~2 wavefronts instead of ~5.
        float4 part[PER];
        int key[PER], yzq[PER];
        unsigned rank[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = tid + q * TNT;
            key[q] = -1;
            if (i < n) {
                const float4 v = nxt[q];
                const float d[3] = {v.x, v.y, v.z};
                float fr[3];
                int lc[3];
                const int org[3] = {ox, oy, oz};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int b = stencil_base<MAS>(d[a], fr[a]);
                    int wb = wrap_index(b, g.dims);
                    if (a == 0) { wb -= g.x_origin; if (wb < 0) wb += g.dims; }
                    lc[a] = wb - org[a];
                }
                const int bank = (lc[1] * (S - 1) + lc[2]) & (TZ - 1);       // (y*AZ + z) mod 32, AZ = 32+S-1
                key[q] = (lc[0] * TZ + bank) * TY + lc[1];
                yzq[q] = lc[1] * TZ + lc[2];
                part[q] = make_float4(fr[0], fr[1], fr[2], v.w);
                const unsigned old = atomicAdd(cnt + (key[q] >> 1), (key[q] & 1) ? 0x10000u : 1u);
                rank[q] = (key[q] & 1) ? (old >> 16) : (old & 0xffffu);
            }
        }
        __syncthreads();

        // ---- exclusive scan of the 8192 packed counters (32 per thread) -------------------------------
        {
            constexpr int WPT = TILE_CELLS / 2 / TNT;          // 16 words per thread
            unsigned words[WPT];
            unsigned sum = 0;
#pragma unroll
            for (int i = 0; i < WPT; i++) {
                words[i] = cnt[tid * WPT + i];
                sum += (words[i] & 0xffffu) + (words[i] >> 16);
            }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_part[warp] = incl;
            __syncthreads();
            unsigned base = 0;
#pragma unroll
            for (int w = 0; w < TNT / 32; w++) base += (w < warp) ? warp_part[w] : 0u;
            unsigned run = base + incl - sum;
#pragma unroll
            for (int i = 0; i < WPT; i++) {
                const unsigned lo = words[i] & 0xffffu, hi = words[i] >> 16;
                cnt[tid * WPT + i] = run | ((run + lo) << 16);
                run += lo + hi;
            }
            if (tid == TNT - 1) cnt[TILE_CELLS / 2] = run;     // == n : end marker for the last row
        }
        __syncthreads();

        // ---- scatter into cell order ------------------------------------------------------------------
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (key[q] >= 0) {
                const unsigned dst = off16(cnt, key[q]) + rank[q];
                sorted[dst] = part[q];
                sorted_yz[dst] = (unsigned short)yzq[q];
            }
        }
        __syncthreads();

        if (c0 + CHUNK < end) {
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const unsigned i = c0 + CHUNK + tid + q * TNT;
                nxt[q] = i < end ? __ldg(bucket + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        // ---- b. stencil accumulation: warp `warp` owns target planes X = warp, warp+8 -------------------
        // (with TX = 8 and 8 warps every warp gets exactly S source-plane visits per chunk: balanced)
        for (int X = warp; X < AX; X += TNT / 32) {
            float *plane = acc + X * AY * AZ;
#pragma unroll 1
            for (int l = 0; l < S; l++) {
                const int x = X - l;                           // source plane
                if (x < 0 || x >= TX) continue;
                // the particles of source plane x are contiguous and ordered by (y,z).  Lane i takes the slots
                // q0 + i*nb + j: lanes of one batch are nb sorted slots apart, so two of them share a cell only
                // when a run of equal cells is longer than nb -- at ~1 particle per cell almost never, and every
                // lane then owns a distinct cell of plane X (plain read-modify-writes, no shuffles).  nb is made
                // odd so that the 16-byte reads of `sorted` at stride nb stay free of bank conflicts.
                const int q0 = (int)off16(cnt, x * TY * TZ);
                const int q1 = (int)off16(cnt, (x + 1) * TY * TZ);
                const int nb = q1 > q0 ? (((q1 - q0 + 31) >> 5) | 1) : 0;
#pragma unroll 1
                for (int j = 0; j < nb; j++) {
                    const int p = q0 + lane * nb + j;
                    const bool valid = p < q1;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    int yz = 0x4000 + lane;                    // invalid lanes: singleton segments
                    if (valid) { v = sorted[p]; yz = sorted_yz[p]; }
                    const int y = (yz >> 5) & (TY - 1), lz = yz & (TZ - 1);
                    float wx[S], wy[S], wz[S];
                    stencil_weights<MAS>(v.x, wx);
                    stencil_weights<MAS>(v.y, wy);
                    stencil_weights<MAS>(v.z, wz);
                    float wxl = wx[0];
#pragma unroll
                    for (int jj = 1; jj < S; jj++) wxl = (l == jj) ? wx[jj] : wxl;
                    const float wxw = wxl * v.w;

                    // sorted slots are monotone in the lane: equal cells are adjacent lanes
                    const int yz_up = __shfl_down_sync(0xffffffffu, yz, 1);
                    const bool dup = __any_sync(0xffffffffu, lane < 31 && yz_up == yz);
                    unsigned peers = 1u << lane;
                    bool head = valid;
                    int run_max = 1;
                    if (dup) {
                        peers = __match_any_sync(0xffffffffu, yz);
                        head = valid && (lane == __ffs(peers) - 1);
                        run_max = __reduce_max_sync(0xffffffffu, __popc(peers));
                    }
                    float *cellp = plane + y * AZ + lz;

                    if (run_max == 1) {
                        // every lane owns a distinct cell of the target plane: plain read-modify-writes
#pragma unroll
                        for (int m = 0; m < S; m++) {
                            const float wxy = wxw * wy[m];
#pragma unroll
                            for (int nn = 0; nn < S; nn++) {
                                if (valid) cellp[m * AZ + nn] += wxy * wz[nn];
                                __syncwarp();
                            }
                        }
                    } else {
                        // segmented suffix scan over each run; its head lane adds the run's sum
                        const unsigned above = peers >> lane;          // bit d: lane+d is in my run
                        const int nsteps = 32 - __clz(run_max - 1);
#pragma unroll
                        for (int m = 0; m < S; m++) {
                            const float wxy = wxw * wy[m];
#pragma unroll
                            for (int nn = 0; nn < S; nn++) {
                                float val = wxy * wz[nn];
                                for (int k = 0; k < nsteps; k++) {
                                    const float o = __shfl_down_sync(0xffffffffu, val, 1 << k);
                                    if (above & (1u << (1 << k))) val += o;
                                }
                                if (head) cellp[m * AZ + nn] += val;
                                __syncwarp();
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- c. flush tile + halo into the grid: a warp per (x,y) row, lanes along z (coalesced reductions) -----
    const int dims = g.dims;
    for (int r = warp; r < AX * AY; r += TNT / 32) {
        const int ax = r / AY, ay = r - ax * AY;
        int gx = ox + ax, gy = oy + ay;                        // gx: plane inside the destination buffer
        if (gx >= g.x_planes) gx -= g.x_planes;                // only when the buffer is the whole periodic grid
        if (gy >= dims) gy -= dims;
        float *row = number + ((int64_t)gx * dims + gy) * dims;
        for (int az = lane; az < AZ; az += 32) {
            const float val = acc[r * AZ + az];
            int gz = oz + az;
            if (gz >= dims) gz -= dims;
            if (val != 0.0f) atomicAdd(row + gz, val);
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------
static TileGeom make_geom(int dims, float BoxSize, int x_origin = 0, int x_own = -1, int x_planes = -1) {
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
    return g;
}

static size_t scan_temp_bytes(unsigned n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (unsigned *)nullptr, (unsigned *)nullptr, (int)n);
    return bytes;
}

// upper bound of sum_t tile_capacity(c_t) given sum_t c_t <= particles/SAMPLE + 4 (Cauchy-Schwarz on
// the sqrt term); the scatter kernel can therefore never write past the bucket array
static size_t bucket_slots_bound(int64_t particles, unsigned ntiles) {
    const double sampled = (double)particles / SAMPLE + 8.0;
    const double bound = 1.125 * SAMPLE * sampled + (4.0 * SAMPLE + 1.0) * sqrt((double)ntiles * sampled) +
                         34.0 * (double)ntiles;
    return (size_t)bound + 1024;
}

bool deposit_tiled_supported(int mas, int64_t particles, int dims, int axes, int x_own) {
    (void)mas;
    if (axes != 3 || dims < 64) return false;                       // halo wrap assumes dims >> stencil
    const TileGeom g = make_geom(dims, 1.0f, 0, x_own, x_own);
    if ((int64_t)g.ntiles > ((int64_t)1 << 30)) return false;
    // 32-bit slots; the cursor's low word may run past its bucket by the number of overflowing particles
    if (bucket_slots_bound(particles, g.ntiles) + (size_t)particles >= ((size_t)1 << 32)) return false;
    return particles >= (int64_t)g.ntiles * 64;                     // sparse inputs: per-tile overhead loses
}

struct TiledWorkspace {
    float4 *bucket;
    unsigned *starts;   // ntiles + 1 : sampled counts -> capacities -> exclusive scan
    unsigned long long *cursor;   // ntiles : (bucket end << 32) | next free slot
    void *scan_tmp;
    size_t scan_bytes;
    size_t total;
};

static TiledWorkspace carve(void *ws, int64_t particles, unsigned ntiles) {
    TiledWorkspace w;
    char *base = reinterpret_cast<char *>(ws);
    size_t off = 0;
    w.bucket = reinterpret_cast<float4 *>(base + off);
    off += align_up(bucket_slots_bound(particles, ntiles) * 16, 256);
    w.starts = reinterpret_cast<unsigned *>(base + off);
    off += align_up(((size_t)ntiles + 1) * 4, 256);
    w.cursor = reinterpret_cast<unsigned long long *>(base + off);
    off += align_up((size_t)ntiles * 8, 256);
    w.scan_tmp = base + off;
    w.scan_bytes = scan_temp_bytes(ntiles + 1);
    off += align_up(w.scan_bytes, 256);
    w.total = off;
    return w;
}

size_t deposit_tiled_workspace(int mas, int64_t particles, int dims, int axes, int mode, int x_own) {
    (void)mas; (void)axes; (void)mode;
    const TileGeom g = make_geom(dims, 1.0f, 0, x_own, x_own);
    return carve(nullptr, particles, g.ntiles).total;
}

template <int MAS>
static int run_tiled(const float *pos, float *number, const float *W, int64_t particles, int dims,
                     float BoxSize, int x_origin, int x_own, int x_planes, unsigned long long *dropped,
                     void *ws, cudaStream_t stream) {
    const TileGeom g = make_geom(dims, BoxSize, x_origin, x_own, x_planes);
    const TiledWorkspace w = carve(ws, particles, g.ntiles);

    const int vec_ok = ((reinterpret_cast<uintptr_t>(pos) & 15) == 0) &&
                       (W == nullptr || (reinterpret_cast<uintptr_t>(W) & 15) == 0);
    int64_t blocks = ((particles + 3) / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    int64_t cblocks = (blocks + SAMPLE - 1) / SAMPLE;
    if (cblocks < 1) cblocks = 1;

    PYL_CUDA_CHECK(cudaMemsetAsync(w.starts, 0, ((size_t)g.ntiles + 1) * 4, stream));
    tile_count_kernel<MAS><<<(int)cblocks, 256, 0, stream>>>(pos, particles, g, w.starts, vec_ok);
    PYL_LAUNCH_CHECK();
    tile_caps_kernel<<<(g.ntiles + 1 + 255) / 256, 256, 0, stream>>>(w.starts, g.ntiles);
    PYL_LAUNCH_CHECK();
    PYL_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(w.scan_tmp, const_cast<size_t &>(w.scan_bytes), w.starts, w.starts,
                                                 (int)(g.ntiles + 1), stream));
    tile_cursor_kernel<<<(g.ntiles + 255) / 256, 256, 0, stream>>>(w.starts, w.cursor, g.ntiles);
    PYL_LAUNCH_CHECK();
    const unsigned sblocks = (unsigned)((particles + 511) / 512);
    if (W)
        tile_scatter_kernel<MAS, true><<<sblocks, 512, 0, stream>>>(pos, W, particles, g, w.cursor, w.bucket, number,
                                                                    dropped);
    else
        tile_scatter_kernel<MAS, false><<<sblocks, 512, 0, stream>>>(pos, W, particles, g, w.cursor, w.bucket, number,
                                                                     dropped);
    PYL_LAUNCH_CHECK();

    static bool attr_done[4] = {false, false, false, false};
    if (!attr_done[MAS]) {
        PYL_CUDA_CHECK(cudaFuncSetAttribute(tile_deposit_kernel<MAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)TileSmem<MAS>::bytes));
        attr_done[MAS] = true;
    }
    tile_deposit_kernel<MAS><<<g.ntiles, TNT, TileSmem<MAS>::bytes, stream>>>(w.bucket, w.cursor, number, g);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

// x_own < 0: whole periodic grid.  Otherwise the slab window of pyl_deposit_slab.
int deposit_tiled(int mas, const float *pos, float *number, const float *W, int64_t particles, int dims,
                  float BoxSize, int x_origin, int x_own, int x_planes, int64_t *dropped, void *ws,
                  cudaStream_t stream) {
    unsigned long long *dr = reinterpret_cast<unsigned long long *>(dropped);
    switch (mas) {
        case PYL_MAS_NGP: return run_tiled<PYL_MAS_NGP>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, stream);
        case PYL_MAS_CIC: return run_tiled<PYL_MAS_CIC>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, stream);
        case PYL_MAS_TSC: return run_tiled<PYL_MAS_TSC>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, stream);
        case PYL_MAS_PCS: return run_tiled<PYL_MAS_PCS>(pos, number, W, particles, dims, BoxSize, x_origin, x_own, x_planes, dr, ws, stream);
    }
    set_last_error("deposit_tiled: unknown scheme %d", mas);
    return PYL_ERR_ARG;
}

}  // namespace pyl
