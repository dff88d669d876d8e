// Error reporting, version string and small elementwise / reduction utilities.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace pyl {

static thread_local char g_err[512] = "";

unsigned long long launches();

void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// ---- elementwise helpers (grid-stride, float4 body, scalar head/tail) -----------------
__global__ void __launch_bounds__(256) affine_kernel(float *__restrict__ x, int64_t n, float a,
                                                     float b) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // x is 16-byte aligned whenever it comes from a device allocator; guard anyway
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int64_t n4 = n >> 2;
        float4 *x4 = reinterpret_cast<float4 *>(x);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 v = x4[i];
            v.x = __fmaf_rn(v.x, a, b); v.y = __fmaf_rn(v.y, a, b);
            v.z = __fmaf_rn(v.z, a, b); v.w = __fmaf_rn(v.w, a, b);
            x4[i] = v;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) x[i] = __fmaf_rn(x[i], a, b);
    } else {
        for (int64_t i = tid; i < n; i += stride) x[i] = __fmaf_rn(x[i], a, b);
    }
}

// x[i] = x[i] / d : true IEEE division, what numpy's `number2 /= 3.0` does on a float32 plane
__global__ void __launch_bounds__(256) divide_kernel(float *__restrict__ x, int64_t n, float d) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        x[i] = __fdiv_rn(x[i], d);
}

// x[i] = float(double(x[i]) / mean) - 1  with mean = sum[0]/count: what the reference's callers do
// in NumPy, `delta /= np.mean(delta, dtype=np.float64); delta -= 1.0` (docs/source/construction.rst:50,
// Pk_snapshot.py:88) -- the division happens in float64, the subtraction in float32.  The float64
// quotient is formed as x * (1/mean): it differs from x/mean by at most one float64 ulp, which survives
// the rounding to float32 in about one element per 2^28.
__device__ __forceinline__ float overdensity_of(float x, double rmean) {
    return __fsub_rn(__double2float_rn(__dmul_rn((double)x, rmean)), 1.0f);
}

__global__ void __launch_bounds__(256) overdensity_kernel(float *__restrict__ x, int64_t n,
                                                          const double *__restrict__ sum, double count) {
    const double rmean = count / sum[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int64_t n4 = n >> 2;
        float4 *x4 = reinterpret_cast<float4 *>(x);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 v = x4[i];
            v.x = overdensity_of(v.x, rmean); v.y = overdensity_of(v.y, rmean);
            v.z = overdensity_of(v.z, rmean); v.w = overdensity_of(v.w, rmean);
            x4[i] = v;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) x[i] = overdensity_of(x[i], rmean);
    } else {
        for (int64_t i = tid; i < n; i += stride) x[i] = overdensity_of(x[i], rmean);
    }
}

__global__ void __launch_bounds__(256) add_kernel(float *__restrict__ out,
                                                  const float *__restrict__ in, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(in)) & 15) == 0) {
        const int64_t n4 = n >> 2;
        float4 *o4 = reinterpret_cast<float4 *>(out);
        const float4 *i4 = reinterpret_cast<const float4 *>(in);
        for (int64_t i = tid; i < n4; i += stride) {
            float4 a = o4[i];
            const float4 b = i4[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            o4[i] = a;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) out[i] += in[i];
    } else {
        for (int64_t i = tid; i < n; i += stride) out[i] += in[i];
    }
}

__global__ void __launch_bounds__(256) sum_f64_kernel(const float *__restrict__ x, int64_t n,
                                                      double *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int64_t n4 = n >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        for (int64_t i = tid; i < n4; i += stride) {
            const float4 v = x4[i];
            acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) acc += (double)x[i];
    } else {
        for (int64_t i = tid; i < n; i += stride) acc += (double)x[i];
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) part[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        acc = (lane < (blockDim.x >> 5)) ? part[lane] : 0.0;
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) atomicAdd(out, acc);
    }
}

// x[i] = -c, c = float32(numerator/cells) read from DEVICE memory; c_out[0] = c as float64 (field.prebias_)
__global__ void __launch_bounds__(256) fill_negative_kernel(float *__restrict__ x, int64_t n,
                                                            const double *__restrict__ numerator, double cells,
                                                            double *__restrict__ c_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float c = (float)(numerator[0] / cells);
    if (tid == 0) c_out[0] = (double)c;
    const float v = -c;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int64_t n4 = n >> 2;
        float4 *x4 = reinterpret_cast<float4 *>(x);
        const float4 v4 = make_float4(v, v, v, v);
        for (int64_t i = tid; i < n4; i += stride) x4[i] = v4;
        for (int64_t i = (n4 << 2) + tid; i < n; i += stride) x[i] = v;
    } else {
        for (int64_t i = tid; i < n; i += stride) x[i] = v;
    }
}

static int ew_grid(int64_t n) {
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace pyl

using namespace pyl;

extern "C" {

const char *pyl_error_string(int status) {
    switch (status) {
        case PYL_OK: return "ok";
        case PYL_ERR_ARG: return "invalid argument";
        case PYL_ERR_CUDA: return "CUDA error";
        case PYL_ERR_CUFFT: return "cuFFT error";
        case PYL_ERR_WORKSPACE: return "workspace missing or too small";
        case PYL_ERR_DOMAIN: return "particle stencil outside the local slab";
        default: return "unknown status";
    }
}

const char *pyl_last_error(void) { return g_err; }

const char *pyl_version(void) { return "pyl_b200 0.1.0 sm_100a"; }

unsigned long long pyl_kernel_launches(void) { return pyl::launches(); }

int pyl_affine_inplace(float *x, int64_t n, float a, float b, pyl_stream_t stream) {
    PYL_REQUIRE(x != nullptr || n == 0, "pyl_affine_inplace: x is NULL");
    if (n <= 0) return PYL_OK;
    affine_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, n, a, b);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_scale_inplace(float *x, int64_t n, float factor, pyl_stream_t stream) {
    return pyl_affine_inplace(x, n, factor, 0.0f, stream);
}

int pyl_divide_inplace(float *x, int64_t n, float divisor, pyl_stream_t stream) {
    PYL_REQUIRE(x != nullptr || n == 0, "pyl_divide_inplace: x is NULL");
    if (n <= 0) return PYL_OK;
    divide_kernel<<<ew_grid(n * 4), 256, 0, as_stream(stream)>>>(x, n, divisor);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_overdensity_inplace(float *x, int64_t n, const double *sum, double count, pyl_stream_t stream) {
    PYL_REQUIRE(x != nullptr || n == 0, "pyl_overdensity_inplace: x is NULL");
    PYL_REQUIRE(sum != nullptr && count > 0.0, "pyl_overdensity_inplace: bad sum/count");
    if (n <= 0) return PYL_OK;
    overdensity_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, n, sum, count);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_fill_negative(float *x, int64_t n, const double *numerator, double cells, double *c_out,
                      pyl_stream_t stream) {
    PYL_REQUIRE(x != nullptr || n == 0, "pyl_fill_negative: x is NULL");
    PYL_REQUIRE(numerator != nullptr && c_out != nullptr && cells > 0.0, "pyl_fill_negative: bad numerator/cells");
    fill_negative_kernel<<<ew_grid(n > 0 ? n : 1), 256, 0, as_stream(stream)>>>(x, n > 0 ? n : 0, numerator, cells, c_out);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_add_inplace(float *out, const float *in, int64_t n, pyl_stream_t stream) {
    PYL_REQUIRE((out != nullptr && in != nullptr) || n == 0, "pyl_add_inplace: NULL pointer");
    if (n <= 0) return PYL_OK;
    add_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(out, in, n);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_sum_f64(const float *x, int64_t n, double *out, pyl_stream_t stream) {
    PYL_REQUIRE(out != nullptr, "pyl_sum_f64: out is NULL");
    PYL_REQUIRE(x != nullptr || n == 0, "pyl_sum_f64: x is NULL");
    PYL_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(double), as_stream(stream)));
    if (n <= 0) return PYL_OK;
    sum_f64_kernel<<<ew_grid(n), 256, 0, as_stream(stream)>>>(x, n, out);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

}  // extern "C"
