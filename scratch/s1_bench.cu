// Prototype benchmark: two-level shared-memory-staged partition of particles into tile buckets
// (x-slab first, then (ty,tz) inside the slab) versus the one-level atomic-cursor scatter.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o s1_bench s1_bench.cu && ./s1_bench 512
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int TX = 8, TY = 16, TZ = 32;

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__global__ void gen_kernel(float *pos, int64_t n, float box) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 3) return;
    pos[i] = (hash32((unsigned)i * 2654435761u + 12345u) >> 8) * (1.0f / 16777216.0f) * box;
}

struct Geom { int dims, ntx, nty, ntz; float inv; };

__device__ __forceinline__ void cell_of(const float d[3], int dims, int c[3]) {
#pragma unroll
    for (int a = 0; a < 3; a++) { int b = (int)d[a]; if (b >= dims) b -= dims; c[a] = b; }
}

// ---- level 1: pos -> slab-ordered records ----------------------------------------------------------
template <int CHUNK, int NT>
__global__ void __launch_bounds__(NT) level1_kernel(const float *__restrict__ pos, int64_t n, Geom g,
                                                    const unsigned *__restrict__ slab_start,
                                                    unsigned *__restrict__ slab_cursor, float4 *__restrict__ out) {
    constexpr int PER = CHUNK / NT;
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *stage = reinterpret_cast<float4 *>(smem);                 // CHUNK
    unsigned short *sbin = reinterpret_cast<unsigned short *>(stage + CHUNK);   // CHUNK
    unsigned *cnt = reinterpret_cast<unsigned *>(sbin + CHUNK);       // nbins
    unsigned *bstart = cnt + 512;                                     // nbins
    unsigned *gbase = bstart + 512;                                   // nbins
    const int nbins = g.ntx;
    const int tid = threadIdx.x;
    const int64_t nchunks = (n + CHUNK - 1) / CHUNK;
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const int64_t base = ch * CHUNK;
        const int m = (int)min((int64_t)CHUNK, n - base);
        for (int i = tid; i < nbins; i += NT) cnt[i] = 0;
        __syncthreads();
        float4 rec[PER];
        int bin[PER];
        unsigned rank[PER];
        // CHUNK*3 floats, read as float4 (CHUNK % 4 == 0, base*3*4 bytes is 16-aligned since CHUNK%4==0)
        const float *src = pos + base * 3;
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = tid + q * NT;
            bin[q] = -1;
            if (i < m) {
                const float x = __ldg(src + i * 3), y = __ldg(src + i * 3 + 1), z = __ldg(src + i * 3 + 2);
                const float d[3] = {x * g.inv, y * g.inv, z * g.inv};
                int c[3];
                cell_of(d, g.dims, c);
                bin[q] = c[0] / TX;
                rec[q] = make_float4(d[0], d[1], d[2], 1.0f);
                rank[q] = atomicAdd(cnt + bin[q], 1u);
            }
        }
        __syncthreads();
        // scan (nbins <= 512 <= 2*NT)
        if (tid < 32) {
            unsigned run = 0;
            for (int b0 = 0; b0 < nbins; b0 += 32) {
                const int b = b0 + tid;
                const unsigned c = b < nbins ? cnt[b] : 0;
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
                if (b < nbins) {
                    bstart[b] = run + incl - c;
                    gbase[b] = c ? slab_start[b] + atomicAdd(slab_cursor + b, c) : 0;
                }
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (bin[q] >= 0) {
                const unsigned s = bstart[bin[q]] + rank[q];
                stage[s] = rec[q];
                sbin[s] = (unsigned short)bin[q];
            }
        }
        __syncthreads();
        for (int s = tid; s < m; s += NT) {
            const int b = sbin[s];
            out[gbase[b] + (s - bstart[b])] = stage[s];
        }
        __syncthreads();
    }
}

// ---- level 2: slab-ordered records -> tile buckets -------------------------------------------------
template <int CHUNK, int NT>
__global__ void __launch_bounds__(NT) level2_kernel(const float4 *__restrict__ in, Geom g,
                                                    const unsigned *__restrict__ slab_start,
                                                    const unsigned *__restrict__ slab_cursor,
                                                    const unsigned *__restrict__ tile_start,
                                                    unsigned *__restrict__ tile_cursor, float4 *__restrict__ out) {
    constexpr int PER = CHUNK / NT;
    extern __shared__ __align__(16) unsigned char smem[];
    float4 *stage = reinterpret_cast<float4 *>(smem);
    unsigned short *sbin = reinterpret_cast<unsigned short *>(stage + CHUNK);
    unsigned *cnt = reinterpret_cast<unsigned *>(sbin + CHUNK);
    unsigned *bstart = cnt + 512;
    unsigned *gbase = bstart + 512;
    const int slab = blockIdx.y;
    const int nbins = g.nty * g.ntz;
    const int tid = threadIdx.x;
    const unsigned s0 = slab_start[slab], cntslab = slab_cursor[slab];
    const unsigned nchunks = (cntslab + CHUNK - 1) / CHUNK;
    for (unsigned ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const unsigned base = s0 + ch * CHUNK;
        const int m = (int)min((unsigned)CHUNK, cntslab - ch * CHUNK);
        for (int i = tid; i < nbins; i += NT) cnt[i] = 0;
        __syncthreads();
        float4 rec[PER];
        int bin[PER];
        unsigned rank[PER];
#pragma unroll
        for (int q = 0; q < PER; q++) {
            const int i = tid + q * NT;
            bin[q] = -1;
            if (i < m) {
                rec[q] = __ldg(in + base + i);
                const float d[3] = {rec[q].x, rec[q].y, rec[q].z};
                int c[3];
                cell_of(d, g.dims, c);
                bin[q] = (c[1] / TY) * g.ntz + c[2] / TZ;
                rank[q] = atomicAdd(cnt + bin[q], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {
            unsigned run = 0;
            for (int b0 = 0; b0 < nbins; b0 += 32) {
                const int b = b0 + tid;
                const unsigned c = b < nbins ? cnt[b] : 0;
                unsigned incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
                if (b < nbins) {
                    bstart[b] = run + incl - c;
                    const unsigned tile = (unsigned)slab * nbins + b;
                    gbase[b] = c ? tile_start[tile] + atomicAdd(tile_cursor + tile, c) : 0;
                }
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; q++) {
            if (bin[q] >= 0) {
                const unsigned s = bstart[bin[q]] + rank[q];
                stage[s] = rec[q];
                sbin[s] = (unsigned short)bin[q];
            }
        }
        __syncthreads();
        for (int s = tid; s < m; s += NT) {
            const int b = sbin[s];
            out[gbase[b] + (s - bstart[b])] = stage[s];
        }
        __syncthreads();
    }
}

// ---- reference: exact tile histogram; one-level scatter ----------------------------------------------
__global__ void hist_kernel(const float *__restrict__ pos, int64_t n, Geom g, unsigned *__restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d[3] = {pos[i * 3] * g.inv, pos[i * 3 + 1] * g.inv, pos[i * 3 + 2] * g.inv};
    int c[3];
    cell_of(d, g.dims, c);
    atomicAdd(counts + ((unsigned)(c[0] / TX) * g.nty + c[1] / TY) * g.ntz + c[2] / TZ, 1u);
}

__global__ void onelevel_kernel(const float *__restrict__ pos, int64_t n, Geom g, const unsigned *__restrict__ tile_start,
                                unsigned *__restrict__ tile_cursor, float4 *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d[3] = {pos[i * 3] * g.inv, pos[i * 3 + 1] * g.inv, pos[i * 3 + 2] * g.inv};
    int c[3];
    cell_of(d, g.dims, c);
    const unsigned t = ((unsigned)(c[0] / TX) * g.nty + c[1] / TY) * g.ntz + c[2] / TZ;
    out[tile_start[t] + atomicAdd(tile_cursor + t, 1u)] = make_float4(d[0], d[1], d[2], 1.0f);
}

template <int MODE>   // 0 full(packed 64-bit cursor) 1 atomic only 2 store only 3 load only 4 atomic32+store no start
__global__ void variant_kernel(const float *__restrict__ pos, int64_t n, Geom g, unsigned long long *__restrict__ cur64,
                               unsigned *__restrict__ cur32, float4 *__restrict__ out, unsigned *sink) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d[3] = {__ldg(pos + i * 3) * g.inv, __ldg(pos + i * 3 + 1) * g.inv, __ldg(pos + i * 3 + 2) * g.inv};
    int c[3];
    cell_of(d, g.dims, c);
    const unsigned t = ((unsigned)(c[0] / TX) * g.nty + c[1] / TY) * g.ntz + c[2] / TZ;
    if (MODE == 0) {
        const unsigned long long v = atomicAdd(cur64 + t, 1ull);
        if ((unsigned)v < (unsigned)(v >> 32)) out[(unsigned)v] = make_float4(d[0], d[1], d[2], 1.0f);
    } else if (MODE == 1) {
        const unsigned v = atomicAdd(cur32 + t, 1u);
        if (v == 0xffffffffu) sink[0] = t;
    } else if (MODE == 2) {
        const unsigned slot = hash32((unsigned)i) % (unsigned)n;
        out[slot] = make_float4(d[0], d[1], d[2], 1.0f);
    } else if (MODE == 3) {
        if (t == 0xffffffffu) sink[0] = t;
    } else if (MODE == 4) {
        const unsigned v = atomicAdd(cur32 + t, 1u);
        out[(size_t)t * 8192 % (size_t)n + (v & 8191)] = make_float4(d[0], d[1], d[2], 1.0f);
    } else if (MODE == 5) {   // RED (no return) + store to hashed slot
        atomicAdd(cur32 + t, 1u);
        const unsigned slot = hash32((unsigned)i) % (unsigned)n;
        out[slot] = make_float4(d[0], d[1], d[2], 1.0f);
    }
}

__global__ void init64_kernel(const unsigned *start, unsigned long long *cur, unsigned nt) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nt) cur[t] = ((unsigned long long)start[t + 1] << 32) | start[t];
}

// per-tile order-independent checksum of a bucket array
__global__ void checksum_kernel(const float4 *__restrict__ b, const unsigned *__restrict__ tile_start,
                                const unsigned *__restrict__ counts, unsigned ntiles, unsigned long long *sum) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    unsigned long long s = 0;
    for (unsigned i = 0; i < counts[t]; i++) {
        const float4 v = b[tile_start[t] + i];
        s += (unsigned long long)__float_as_uint(v.x) * 3 + (unsigned long long)__float_as_uint(v.y) * 5 + __float_as_uint(v.z);
    }
    sum[t] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 512;
    const int64_t n = (int64_t)N * N * N;
    Geom g{N, N / TX, N / TY, N / TZ, (float)N / 1000.0f};
    const unsigned ntiles = g.ntx * g.nty * g.ntz;
    float *pos; CK(cudaMalloc(&pos, n * 12));
    gen_kernel<<<(unsigned)((n * 3 + 255) / 256), 256>>>(pos, n, 1000.0f);
    unsigned *counts, *tile_start, *tile_cursor, *slab_start, *slab_cursor;
    CK(cudaMalloc(&counts, ntiles * 4)); CK(cudaMalloc(&tile_start, (ntiles + 1) * 4)); CK(cudaMalloc(&tile_cursor, ntiles * 4));
    CK(cudaMalloc(&slab_start, (g.ntx + 1) * 4)); CK(cudaMalloc(&slab_cursor, g.ntx * 4));
    CK(cudaMemset(counts, 0, ntiles * 4));
    hist_kernel<<<(unsigned)((n + 255) / 256), 256>>>(pos, n, g, counts);
    // exact starts on the host (prototype)
    unsigned *h = (unsigned *)malloc((ntiles + 1) * 4), *hs = (unsigned *)malloc((ntiles + 1) * 4);
    CK(cudaMemcpy(h, counts, ntiles * 4, cudaMemcpyDeviceToHost));
    unsigned run = 0;
    for (unsigned t = 0; t < ntiles; t++) { hs[t] = run; run += h[t]; }
    hs[ntiles] = run;
    CK(cudaMemcpy(tile_start, hs, (ntiles + 1) * 4, cudaMemcpyHostToDevice));
    unsigned *hslab = (unsigned *)malloc((g.ntx + 1) * 4);
    for (int s = 0; s <= g.ntx; s++) hslab[s] = hs[(unsigned)s * g.nty * g.ntz];
    CK(cudaMemcpy(slab_start, hslab, (g.ntx + 1) * 4, cudaMemcpyHostToDevice));
    float4 *ba, *bb, *bc;
    CK(cudaMalloc(&ba, n * 16)); CK(cudaMalloc(&bb, n * 16)); CK(cudaMalloc(&bc, n * 16));
    unsigned long long *ck1, *ck2;
    CK(cudaMalloc(&ck1, ntiles * 8)); CK(cudaMalloc(&ck2, ntiles * 8));

    float t1 = time_ms([&] {
        cudaMemsetAsync(tile_cursor, 0, ntiles * 4);
        onelevel_kernel<<<(unsigned)((n + 255) / 256), 256>>>(pos, n, g, tile_start, tile_cursor, bc);
    });
    printf("one-level atomic scatter: %.3f ms\n", t1);
    checksum_kernel<<<(ntiles + 127) / 128, 128>>>(bc, tile_start, counts, ntiles, ck1);

    unsigned long long *cur64; CK(cudaMalloc(&cur64, ntiles * 8));
    unsigned *sink; CK(cudaMalloc(&sink, 4));
    for (int bs : {128, 256, 512, 1024}) {
        const unsigned gb = (unsigned)((n + bs - 1) / bs);
        float m0 = time_ms([&] { init64_kernel<<<(ntiles + 255) / 256, 256>>>(tile_start, cur64, ntiles);
                                 variant_kernel<0><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        float m1 = time_ms([&] { cudaMemsetAsync(tile_cursor, 0, ntiles * 4); variant_kernel<1><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        float m2 = time_ms([&] { variant_kernel<2><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        float m3 = time_ms([&] { variant_kernel<3><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        float m4 = time_ms([&] { cudaMemsetAsync(tile_cursor, 0, ntiles * 4); variant_kernel<4><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        float m5 = time_ms([&] { cudaMemsetAsync(tile_cursor, 0, ntiles * 4); variant_kernel<5><<<gb, bs>>>(pos, n, g, cur64, tile_cursor, bc, sink); });
        printf("block %4d: full64 %.3f  atomic-only %.3f  store-only(random) %.3f  load-only %.3f  atomic32+store %.3f  red+randomstore %.3f ms\n", bs, m0, m1, m2, m3, m4, m5);
    }

    auto run2 = [&](auto l1, auto l2, int chunk, int nt, const char *name) {
        const size_t sm = (size_t)chunk * 18 + 3 * 512 * 4;
        CK(cudaFuncSetAttribute(l1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaFuncSetAttribute(l2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int occ1 = 0, occ2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, l1, nt, sm);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, l2, nt, sm);
        float a = time_ms([&] {
            cudaMemsetAsync(slab_cursor, 0, g.ntx * 4);
            l1<<<148 * occ1, nt, sm>>>(pos, n, g, slab_start, slab_cursor, ba);
        });
        const unsigned per_slab = (unsigned)((n / g.ntx) * 1.2 / chunk) + 1;
        float b = time_ms([&] {
            cudaMemsetAsync(tile_cursor, 0, ntiles * 4);
            l2<<<dim3(per_slab, g.ntx), nt, sm>>>(ba, g, slab_start, slab_cursor, tile_start, tile_cursor, bb);
        });
        checksum_kernel<<<(ntiles + 127) / 128, 128>>>(bb, tile_start, counts, ntiles, ck2);
        unsigned long long *x = (unsigned long long *)malloc(ntiles * 8), *y = (unsigned long long *)malloc(ntiles * 8);
        CK(cudaMemcpy(x, ck1, ntiles * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(y, ck2, ntiles * 8, cudaMemcpyDeviceToHost));
        unsigned bad = 0;
        for (unsigned t = 0; t < ntiles; t++) bad += x[t] != y[t];
        printf("%s: level1 %.3f ms (occ %d)  level2 %.3f ms (occ %d)  total %.3f ms  mismatching tiles %u\n", name, a, occ1,
               b, occ2, a + b, bad);
        free(x); free(y);
    };
    run2(level1_kernel<4096, 512>, level2_kernel<4096, 512>, 4096, 512, "chunk 4096 / 512 thr");
    return 0;
}
