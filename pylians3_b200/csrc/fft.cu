// FFT stage: cuFFT plans behind the C ABI (library call, not a hand-written kernel).
// Replaces FFT3Dr_f (library/Pk_library/Pk_library.pyx:117-130): unnormalised forward r2c,
// float32 (N,N,N) -> complex64 (N,N,N/2+1), C order, input left untouched.
// Plans are created with auto-allocation OFF so that cuFFT's work area comes from the
// caller's workspace (torch's caching allocator owns all memory).
#include <cufft.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace pyl {

enum PlanKind { PLAN_R2C_3D = 0, PLAN_R2C_YZ = 1, PLAN_C2C_X = 2, PLAN_C2R_3D = 3, PLAN_C2R_2D = 4, PLAN_C2C_XMID = 5 };

struct Plan {
    cufftHandle handle = 0;
    size_t work = 0;
};

using PlanKey = std::tuple<int, int, int, int, int>;   // device, kind, dims, batch, nky (x plans: row stride)
// slab transforms run in batches so that cuFFT's work area (which grows with the batch) stays bounded at
// 4096^3-class grids: at most this many (y,z) planes / x columns per cufftExec call
// (smaller batches were measured and lose: 4096^3 over 8 ranks, x transforms 83 ms at 65536 columns per call, 95 ms at
// 2049; 2D transforms 38 ms at 16 planes per call, 44 ms at 1 -- profiles/r2_config5_pieces.md)
static int yz_batch(int) { return 16; }
static int x_batch(int) { return 1 << 16; }
static std::map<PlanKey, Plan> g_plans;
static std::mutex g_plans_mu;

static const char *cufft_name(cufftResult r) {
    switch (r) {
        case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
        case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
        case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
        case CUFFT_INVALID_TYPE: return "CUFFT_INVALID_TYPE";
        case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
        case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
        case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
        case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
        case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
        case CUFFT_UNALIGNED_DATA: return "CUFFT_UNALIGNED_DATA";
        case CUFFT_INVALID_DEVICE: return "CUFFT_INVALID_DEVICE";
        case CUFFT_NOT_SUPPORTED: return "CUFFT_NOT_SUPPORTED";
        default: return "CUFFT_<other>";
    }
}

#define PYL_CUFFT_CHECK(expr)                                                             \
    do {                                                                                  \
        cufftResult _r = (expr);                                                          \
        if (_r != CUFFT_SUCCESS) {                                                        \
            pyl::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                                pyl::cufft_name(_r));                                     \
            return PYL_ERR_CUFFT;                                                         \
        }                                                                                 \
    } while (0)

static int get_plan(PlanKind kind, int dims, int extent, Plan *out, int nky = 0) {
    int dev = 0;
    PYL_CUDA_CHECK(cudaGetDevice(&dev));
    const PlanKey key(dev, (int)kind, dims, extent, nky);
    std::lock_guard<std::mutex> lock(g_plans_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *out = it->second; return PYL_OK; }

    Plan p;
    PYL_CUFFT_CHECK(cufftCreate(&p.handle));
    cufftResult r = cufftSetAutoAllocation(p.handle, 0);
    const long long N = dims, nz = dims / 2 + 1;
    if (r == CUFFT_SUCCESS) {
        if (kind == PLAN_R2C_3D) {
            long long n[3] = {N, N, N};
            r = cufftMakePlanMany64(p.handle, 3, n, nullptr, 1, N * N * N, nullptr, 1, N * N * nz,
                                    CUFFT_R2C, 1, &p.work);
        } else if (kind == PLAN_C2R_3D) {
            long long n[3] = {N, N, N};
            r = cufftMakePlanMany64(p.handle, 3, n, nullptr, 1, N * N * nz, nullptr, 1, N * N * N,
                                    CUFFT_C2R, 1, &p.work);
        } else if (kind == PLAN_C2R_2D) {
            long long n[2] = {N, N};
            r = cufftMakePlanMany64(p.handle, 2, n, nullptr, 1, N * nz, nullptr, 1, N * N, CUFFT_C2R, 1, &p.work);
        } else if (kind == PLAN_R2C_YZ) {
            long long n[2] = {N, N};
            r = cufftMakePlanMany64(p.handle, 2, n, nullptr, 1, N * N, nullptr, 1, N * nz, CUFFT_R2C,
                                    (long long)extent, &p.work);
        } else if (kind == PLAN_C2C_XMID) {
            // in-place 1D transforms along the FIRST axis of one (N, nz) plane: element stride nz, nz columns
            long long n[1] = {N};
            long long embed[1] = {N};
            r = cufftMakePlanMany64(p.handle, 1, n, embed, nz, 1, embed, nz, 1, CUFFT_C2C, nz, &p.work);
        } else {
            // in-place 1D transforms along x of a (N, nky, nz) array: element stride nky*nz, `extent` columns
            long long n[1] = {N};
            long long embed[1] = {N};
            const long long stride = (long long)nky * nz;
            r = cufftMakePlanMany64(p.handle, 1, n, embed, stride, 1, embed, stride, 1, CUFFT_C2C,
                                    (long long)extent, &p.work);
        }
    }
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(p.handle);
        set_last_error("cuFFT plan (kind %d, dims %d, extent %d) failed: %s", (int)kind, dims,
                       extent, cufft_name(r));
        return PYL_ERR_CUFFT;
    }
    g_plans[key] = p;
    *out = p;
    return PYL_OK;
}

static int bind(const Plan &p, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (p.work > 0 && (ws == nullptr || ws_bytes < p.work)) {
        set_last_error("FFT workspace of %zu bytes required, %zu given", p.work, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    PYL_CUFFT_CHECK(cufftSetStream(p.handle, stream));
    if (p.work > 0) PYL_CUFFT_CHECK(cufftSetWorkArea(p.handle, ws));
    return PYL_OK;
}

}  // namespace pyl

using namespace pyl;

extern "C" {

size_t pyl_fft_r2c_workspace_bytes(int dims) {
    if (dims <= 0) return 0;
    Plan p;
    if (get_plan(PLAN_R2C_3D, dims, 0, &p) != PYL_OK) return (size_t)-1;
    return p.work;
}

int pyl_fft_r2c(const float *delta, float *delta_k, int dims, void *ws, size_t ws_bytes,
                pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0, "pyl_fft_r2c: dims must be positive");
    PYL_REQUIRE(delta != nullptr && delta_k != nullptr, "pyl_fft_r2c: NULL pointer");
    PYL_REQUIRE((void *)delta != (void *)delta_k, "pyl_fft_r2c: transform is out of place");
    Plan p;
    int st = get_plan(PLAN_R2C_3D, dims, 0, &p);
    if (st != PYL_OK) return st;
    st = bind(p, ws, ws_bytes, as_stream(stream));
    if (st != PYL_OK) return st;
    PYL_CUFFT_CHECK(cufftExecR2C(p.handle, const_cast<cufftReal *>(delta),
                                 reinterpret_cast<cufftComplex *>(delta_k)));
    return PYL_OK;
}

size_t pyl_fft_c2r_workspace_bytes(int dims) {
    if (dims <= 0) return 0;
    Plan p;
    if (get_plan(PLAN_C2R_3D, dims, 0, &p) != PYL_OK) return (size_t)-1;
    return p.work;
}

int pyl_fft_c2r(float *delta_k, float *delta, int dims, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0, "pyl_fft_c2r: dims must be positive");
    PYL_REQUIRE(delta != nullptr && delta_k != nullptr, "pyl_fft_c2r: NULL pointer");
    PYL_REQUIRE((void *)delta != (void *)delta_k, "pyl_fft_c2r: transform is out of place");
    Plan p;
    int st = get_plan(PLAN_C2R_3D, dims, 0, &p);
    if (st != PYL_OK) return st;
    st = bind(p, ws, ws_bytes, as_stream(stream));
    if (st != PYL_OK) return st;
    PYL_CUFFT_CHECK(cufftExecC2R(p.handle, reinterpret_cast<cufftComplex *>(delta_k), delta));
    return PYL_OK;
}

size_t pyl_fft2d_c2r_workspace_bytes(int dims) {
    if (dims <= 0) return 0;
    Plan p;
    if (get_plan(PLAN_C2R_2D, dims, 0, &p) != PYL_OK) return (size_t)-1;
    return p.work;
}

int pyl_fft2d_c2r(float *image_k, float *image, int dims, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0, "pyl_fft2d_c2r: dims must be positive");
    PYL_REQUIRE(image != nullptr && image_k != nullptr, "pyl_fft2d_c2r: NULL pointer");
    PYL_REQUIRE((void *)image != (void *)image_k, "pyl_fft2d_c2r: transform is out of place");
    Plan p;
    int st = get_plan(PLAN_C2R_2D, dims, 0, &p);
    if (st != PYL_OK) return st;
    st = bind(p, ws, ws_bytes, as_stream(stream));
    if (st != PYL_OK) return st;
    PYL_CUFFT_CHECK(cufftExecC2R(p.handle, reinterpret_cast<cufftComplex *>(image_k), image));
    return PYL_OK;
}

size_t pyl_fft_slab_workspace_bytes(int dims, int nx, int nky) {
    if (dims <= 0) return 0;
    size_t need = 0;
    Plan p;
    if (nx > 0) {
        const int YZ_BATCH = yz_batch(dims);
        const int full = nx < YZ_BATCH ? nx : YZ_BATCH, rest = nx % full;
        if (get_plan(PLAN_R2C_YZ, dims, full, &p) != PYL_OK) return (size_t)-1;
        need = p.work;
        if (rest > 0) {
            if (get_plan(PLAN_R2C_YZ, dims, rest, &p) != PYL_OK) return (size_t)-1;
            if (p.work > need) need = p.work;
        }
    }
    if (nky > 0) {
        const long long cols = (long long)nky * (dims / 2 + 1);
        const int X_BATCH = x_batch(dims);
        const int full = cols < X_BATCH ? (int)cols : X_BATCH, rest = (int)(cols % full);
        if (get_plan(PLAN_C2C_X, dims, full, &p, nky) != PYL_OK) return (size_t)-1;
        if (p.work > need) need = p.work;
        if (rest > 0) {
            if (get_plan(PLAN_C2C_X, dims, rest, &p, nky) != PYL_OK) return (size_t)-1;
            if (p.work > need) need = p.work;
        }
    }
    return need;
}

int pyl_fft_slab_yz(const float *slab, float *slab_k, int dims, int nx, void *ws, size_t ws_bytes,
                    pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && nx >= 0, "pyl_fft_slab_yz: bad sizes");
    if (nx == 0) return PYL_OK;
    PYL_REQUIRE(slab != nullptr && slab_k != nullptr, "pyl_fft_slab_yz: NULL pointer");
    const long long plane_in = (long long)dims * dims, plane_out = (long long)dims * (dims / 2 + 1);
    const int YZ_BATCH = yz_batch(dims);
    for (int x = 0; x < nx; x += YZ_BATCH) {
        const int b = nx - x < YZ_BATCH ? nx - x : YZ_BATCH;
        Plan p;
        int st = get_plan(PLAN_R2C_YZ, dims, b, &p);
        if (st != PYL_OK) return st;
        st = bind(p, ws, ws_bytes, as_stream(stream));
        if (st != PYL_OK) return st;
        PYL_CUFFT_CHECK(cufftExecR2C(p.handle, const_cast<cufftReal *>(slab) + x * plane_in,
                                     reinterpret_cast<cufftComplex *>(slab_k) + x * plane_out));
    }
    return PYL_OK;
}

int pyl_fft_slab_x(float *cols_k, int dims, int nky, void *ws, size_t ws_bytes,
                   pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && nky >= 0, "pyl_fft_slab_x: bad sizes");
    if (nky == 0) return PYL_OK;
    PYL_REQUIRE(cols_k != nullptr, "pyl_fft_slab_x: NULL pointer");
    const long long cols = (long long)nky * (dims / 2 + 1);
    cufftComplex *c = reinterpret_cast<cufftComplex *>(cols_k);
    const int X_BATCH = x_batch(dims);
    for (long long j = 0; j < cols; j += X_BATCH) {
        const int b = cols - j < X_BATCH ? (int)(cols - j) : X_BATCH;
        Plan p;
        int st = get_plan(PLAN_C2C_X, dims, b, &p, nky);
        if (st != PYL_OK) return st;
        st = bind(p, ws, ws_bytes, as_stream(stream));
        if (st != PYL_OK) return st;
        PYL_CUFFT_CHECK(cufftExecC2C(p.handle, c + j, c + j, CUFFT_FORWARD));
    }
    return PYL_OK;
}

size_t pyl_fft_slab_x_kymajor_workspace_bytes(int dims) {
    if (dims <= 0) return 0;
    Plan p;
    if (get_plan(PLAN_C2C_XMID, dims, 0, &p) != PYL_OK) return (size_t)-1;
    return p.work;
}

int pyl_fft_slab_x_kymajor(float *cols_k, int dims, int nky, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && nky >= 0, "pyl_fft_slab_x_kymajor: bad sizes");
    if (nky == 0) return PYL_OK;
    PYL_REQUIRE(cols_k != nullptr, "pyl_fft_slab_x_kymajor: NULL pointer");
    Plan p;
    int st = get_plan(PLAN_C2C_XMID, dims, 0, &p);
    if (st != PYL_OK) return st;
    st = bind(p, ws, ws_bytes, as_stream(stream));
    if (st != PYL_OK) return st;
    cufftComplex *c = reinterpret_cast<cufftComplex *>(cols_k);
    const long long plane = (long long)dims * (dims / 2 + 1);
    for (int j = 0; j < nky; j++) PYL_CUFFT_CHECK(cufftExecC2C(p.handle, c + j * plane, c + j * plane, CUFFT_FORWARD));
    return PYL_OK;
}

int pyl_fft_clear_plans(void) {
    int dev = 0;
    PYL_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plans_mu);
    for (auto it = g_plans.begin(); it != g_plans.end();) {
        if (std::get<0>(it->first) == dev) {
            cufftDestroy(it->second.handle);
            it = g_plans.erase(it);
        } else {
            ++it;
        }
    }
    return PYL_OK;
}

}  // extern "C"
