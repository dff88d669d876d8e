// Micro-benchmarks behind the round-2 deposit design (B200).  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o scratch/ubench scratch/ubench.cu ; run: scratch/ubench
// Each line: what, time, rate.  Rates decide: shared-memory ranking by ATOMS vs match_any, per-particle global
// atomics vs per-run atomics, red.global flush vs cp.reduce.async.bulk flush, run length needed for scattered writes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ---- 1. shared-memory ops: ITER ops per thread on a table of TBL words --------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) smem_ops(unsigned *out, int iters, int tbl_mask) {
    extern __shared__ unsigned tbl[];
    for (int i = threadIdx.x; i <= tbl_mask; i += blockDim.x) tbl[i] = 0;
    __syncthreads();
    unsigned h = hash32(blockIdx.x * 256 + threadIdx.x);
    unsigned acc = 0;
    float *ftbl = reinterpret_cast<float *>(tbl);
    for (int i = 0; i < iters; i++) {
        h = h * 1664525u + 1013904223u;
        unsigned a = (h >> 8) & tbl_mask;
        if (KIND == 0) acc += atomicAdd(&tbl[a], 1u);                       // ATOMS.ADD returning
        else if (KIND == 1) atomicAdd(&ftbl[a], 1.0f);                      // float: CAS loop?
        else if (KIND == 2) { acc += __popc(__match_any_sync(0xffffffffu, a)); }
        else if (KIND == 3) { float v = ftbl[a]; ftbl[a] = v + 1.0f; }      // plain RMW, random banks
        else if (KIND == 4) { unsigned b = (a & ~31u) | (threadIdx.x & 31); float v = ftbl[b]; ftbl[b] = v + 1.0f; }  // conflict-free RMW
        else if (KIND == 5) atomicAdd(&tbl[a], 1u);                         // ATOMS non-returning (RED.shared)
        else if (KIND == 6) { unsigned b = (a & ~31u) | (threadIdx.x & 31); acc += atomicAdd(&tbl[b], 1u); }  // ATOMS, one lane per bank
    }
    if (acc == 0xdeadbeef) out[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && tbl[0] == 0xdeadbeef) out[1] = 1;
}

// ---- 2. global atomics ------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(512) gatom(unsigned long long *cur, unsigned *out, int64_t n, unsigned mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned a = hash32((unsigned)i) & mask;
    if (KIND == 0) { unsigned long long r = atomicAdd(cur + a, 1ull); if (r == 0xdeadbeefdeadull) out[0] = 1; }
    else if (KIND == 1) { unsigned r = atomicAdd(reinterpret_cast<unsigned *>(cur) + a, 1u); if (r == 0xdeadbeef) out[0] = 1; }
    else if (KIND == 2) atomicAdd(reinterpret_cast<unsigned *>(cur) + a, 1u);      // RED
    else if (KIND == 3) atomicAdd(reinterpret_cast<float *>(cur) + a, 1.0f);       // RED.F32
}

// ---- 3. scattered run writes: every warp writes runs of R float4 at pseudo-random bucket cursors ---------
// emulates the partition output: per run one returning atomic on one of `nb` cursors, then R x 16 B contiguous
template <int R>
__global__ void __launch_bounds__(256) run_writes(float4 *buf, unsigned *cursor, unsigned nb_mask, unsigned cap, int64_t nruns) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    constexpr int RPW = 32 / R;              // runs per warp instruction
    const int64_t run = warp * RPW + lane / R;
    if (run >= nruns) return;
    unsigned b = hash32((unsigned)run) & nb_mask;
    unsigned slot = 0;
    if (lane % R == 0) slot = atomicAdd(cursor + b, (unsigned)R);
    slot = __shfl_sync(0xffffffffu, slot, lane - lane % R);
    slot = slot % cap;
    buf[(int64_t)b * cap + slot + lane % R] = make_float4(1.f, 2.f, 3.f, (float)run);
}

// ---- 4. flush: red.global rows vs cp.reduce.async.bulk rows ----------------------------------------------
// grid of N^3 floats; each CTA owns a tile of 8 x 16 x 32 cells and adds AX*AY rows of 32 floats (+halo ignored)
__global__ void __launch_bounds__(256) flush_red(float *grid, int N) {
    __shared__ float acc[9 * 17 * 36];
    for (int i = threadIdx.x; i < 9 * 17 * 36; i += 256) acc[i] = 1.0f;
    __syncthreads();
    const int ntz = N / 32, nty = N / 16;
    const int tile = blockIdx.x;
    const int tz = tile % ntz, ty = (tile / ntz) % nty, tx = tile / (ntz * nty);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < 9 * 17; r += 8) {
        int ax = r / 17, ay = r % 17;
        int gx = (tx * 8 + ax) % N, gy = (ty * 16 + ay) % N;
        float *row = grid + ((int64_t)gx * N + gy) * N + tz * 32;
        atomicAdd(row + lane, acc[r * 36 + lane]);
        if (lane < 1) { int gz = (tz * 32 + 32 + lane) % N; atomicAdd(grid + ((int64_t)gx * N + gy) * N + gz, acc[r * 36 + 32 + lane]); }
    }
}

__global__ void __launch_bounds__(256) flush_bulk(float *grid, int N) {
    __shared__ __align__(128) float acc[9 * 17 * 36];
    for (int i = threadIdx.x; i < 9 * 17 * 36; i += 256) acc[i] = 1.0f;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int ntz = N / 32, nty = N / 16;
    const int tile = blockIdx.x;
    const int tz = tile % ntz, ty = (tile / ntz) % nty, tx = tile / (ntz * nty);
    for (int r = threadIdx.x; r < 9 * 17; r += 256) {
        int ax = r / 17, ay = r % 17;
        int gx = (tx * 8 + ax) % N, gy = (ty * 16 + ay) % N;
        float *row = grid + ((int64_t)gx * N + gy) * N + tz * 32;
        unsigned s = (unsigned)__cvta_generic_to_shared(acc + r * 36);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(row), "r"(s), "r"(128) : "memory");
        int gz = (tz * 32 + 32) % N;
        float *row2 = grid + ((int64_t)gx * N + gy) * N + gz;
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(row2), "r"(s + 128), "r"(16) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- 5. bulk load of a chunk (global -> shared) with an mbarrier, then a sum: load path check ---------------
__global__ void __launch_bounds__(256) bulk_load(const float *src, float *out, int chunk_floats, int chunks_per_cta) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = reinterpret_cast<float *>(sm);
    __shared__ __align__(8) unsigned long long bar[2];
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    const unsigned bytes = chunk_floats * 4;
    const float *base = src + (int64_t)blockIdx.x * chunks_per_cta * chunk_floats;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(buf)), "l"(base), "r"(bytes), "r"(bar0) : "memory");
    }
    for (int c = 0; c < chunks_per_cta; c++) {
        const int s = c & 1;
        if (threadIdx.x == 0 && c + 1 < chunks_per_cta) {
            const unsigned b = bar0 + 8 * (s ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(buf + (s ^ 1) * chunk_floats)), "l"(base + (int64_t)(c + 1) * chunk_floats), "r"(bytes), "r"(b) : "memory");
        }
        const unsigned phase = (c >> 1) & 1;
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar0 + 8 * s), "r"(phase) : "memory");
        }
        const float *b = buf + s * chunk_floats;
        for (int i = threadIdx.x; i < chunk_floats; i += 256) acc += b[i];
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_load(const float *src, float *out, int chunk_floats, int chunks_per_cta) {
    float acc = 0.f;
    const float4 *base = reinterpret_cast<const float4 *>(src + (int64_t)blockIdx.x * chunks_per_cta * chunk_floats);
    const int n4 = chunks_per_cta * chunk_floats / 4;
    for (int i = threadIdx.x; i < n4; i += 256) { float4 v = __ldg(base + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
static float timeit(F f, int reps = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate * 1e-6;
    printf("device %s, %d SMs, %.3f GHz\n", p.name, sms, ghz);
    unsigned *out; CK(cudaMalloc(&out, 64));

    {   // shared-memory ops: 8 CTAs/SM x 256 threads, table 4096 words
        const int iters = 4096, blocks = sms * 8;
        const char *names[] = {"ATOMS.ADD u32 returning, random", "atomicAdd float smem, random", "match_any", "plain RMW random banks", "plain RMW conflict-free", "ATOMS.ADD u32 no return", "ATOMS.ADD u32 returning, conflict-free"};
        float ms[7];
        ms[6] = timeit([&] { smem_ops<6><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[0] = timeit([&] { smem_ops<0><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[1] = timeit([&] { smem_ops<1><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[2] = timeit([&] { smem_ops<2><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[3] = timeit([&] { smem_ops<3><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[4] = timeit([&] { smem_ops<4><<<blocks, 256, 16384>>>(out, iters, 4095); });
        ms[5] = timeit([&] { smem_ops<5><<<blocks, 256, 16384>>>(out, iters, 4095); });
        for (int k = 0; k < 7; k++) {
            const double warp_ops_per_sm = (double)iters * 8 * 8;      // warps per SM x iters
            printf("smem %-34s %8.3f ms  %.2f cycles per warp-op per SM\n", names[k], ms[k], ms[k] * 1e-3 * ghz * 1e9 / warp_ops_per_sm);
        }
    }
    {   // global atomics, 134M ops over 32768 / 262144 addresses
        const int64_t n = 134217728;
        unsigned long long *cur; CK(cudaMalloc(&cur, 8 << 20));
        CK(cudaMemset(cur, 0, 8 << 20));
        const unsigned masks[] = {32767u, 262143u, 127u};
        for (unsigned m : masks) {
            float a = timeit([&] { gatom<0><<<(unsigned)(n / 512), 512>>>(cur, out, n, m); });
            float b = timeit([&] { gatom<1><<<(unsigned)(n / 512), 512>>>(cur, out, n, m); });
            float c = timeit([&] { gatom<2><<<(unsigned)(n / 512), 512>>>(cur, out, n, m); });
            float d = timeit([&] { gatom<3><<<(unsigned)(n / 512), 512>>>(cur, out, n, m); });
            printf("global atomics over %6u addresses, 134M ops: u64 ret %.3f ms | u32 ret %.3f ms | u32 red %.3f ms | f32 red %.3f ms\n", m + 1, a, b, c, d);
        }
        CK(cudaFree(cur));
    }
    {   // scattered runs: 134M records of 16 B in runs of R into nb buckets
        const int64_t nrec = 134217728;
        const unsigned nbs[] = {32768u, 512u};
        for (unsigned nb : nbs) {
            const unsigned cap = (unsigned)(nrec / nb);
            float4 *buf; CK(cudaMalloc(&buf, (size_t)nrec * 16 + 4096));
            unsigned *cursor; CK(cudaMalloc(&cursor, nb * 4));
            CK(cudaMemset(cursor, 0, nb * 4));
            float t1 = timeit([&] { run_writes<1><<<(unsigned)(nrec / 256), 256>>>(buf, cursor, nb - 1, cap, nrec); });
            float t2 = timeit([&] { run_writes<2><<<(unsigned)(nrec / 256), 256>>>(buf, cursor, nb - 1, cap, nrec / 2); });
            float t4 = timeit([&] { run_writes<4><<<(unsigned)(nrec / 256), 256>>>(buf, cursor, nb - 1, cap, nrec / 4); });
            float t8 = timeit([&] { run_writes<8><<<(unsigned)(nrec / 256), 256>>>(buf, cursor, nb - 1, cap, nrec / 8); });
            float t16 = timeit([&] { run_writes<16><<<(unsigned)(nrec / 256), 256>>>(buf, cursor, nb - 1, cap, nrec / 16); });
            printf("run writes, %u buckets, 2.1 GB: R=1 %.3f | R=2 %.3f | R=4 %.3f | R=8 %.3f | R=16 %.3f ms\n", nb, t1, t2, t4, t8, t16);
            CK(cudaFree(buf)); CK(cudaFree(cursor));
        }
    }
    {   // flush of a 512^3 grid by tiles
        const int N = 512;
        float *grid; CK(cudaMalloc(&grid, (size_t)N * N * N * 4));
        CK(cudaMemset(grid, 0, (size_t)N * N * N * 4));
        const unsigned tiles = (N / 8) * (N / 16) * (N / 32);
        float a = timeit([&] { flush_red<<<tiles, 256>>>(grid, N); });
        float b = timeit([&] { flush_bulk<<<tiles, 256>>>(grid, N); });
        float h[4];
        CK(cudaMemcpy(h, grid + (size_t)N * N * 17 + N * 5 + 33, 16, cudaMemcpyDeviceToHost));
        printf("flush 512^3 by 8x16x32 tiles (+1 halo): red.global %.3f ms | cp.reduce.async.bulk %.3f ms   (cell value %.1f, expect equal counts from both: %s)\n", a, b, h[0], "see value");
        CK(cudaFree(grid));
    }
    {   // chunked bulk load vs ldg: 2.1 GB
        const int chunk_floats = 4096, chunks = 64;
        const int blocks = 2048;
        float *src; CK(cudaMalloc(&src, (size_t)blocks * chunks * chunk_floats * 4));
        CK(cudaMemset(src, 0, (size_t)blocks * chunks * chunk_floats * 4));
        CK(cudaFuncSetAttribute(bulk_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * chunk_floats * 4));
        float a = timeit([&] { bulk_load<<<blocks, 256, 2 * chunk_floats * 4>>>(src, reinterpret_cast<float *>(out), chunk_floats, chunks); });
        float b = timeit([&] { ldg_load<<<blocks, 256>>>(src, reinterpret_cast<float *>(out), chunk_floats, chunks); });
        const double gb = (double)blocks * chunks * chunk_floats * 4 / 1e9;
        printf("stream %.2f GB through 16 KB chunks: cp.async.bulk+mbarrier %.3f ms (%.0f GB/s) | ldg float4 %.3f ms (%.0f GB/s)\n", gb, a, gb / a * 1e3, b, gb / b * 1e3);
        CK(cudaFree(src));
    }
    return 0;
}
