// Shell estimators and half-spectrum passes: CUDA wrappers around the bodies of shell_body.cuh.
//
// Replaces the serial loops of library/Pk_library/Pk_library.pyx: Pk_plane :470-499, XPk_plane :1151-1200,
// Pk_theta :1273-1316, XPk_dv :1386-1432, XPk_vv :1515-1568, expected_Pk :2004-2037, the real-space binning of
// Xi / XXi :2233-2267 / :2378-2412, the mode loops of correct_MAS :1909-1929 and Xi / XXi :2198-2218 / :2335-2362,
// smoothing_library.pyx:227-232 / :255-258 (field_k * filter_k) and the filter loops of FT_filter / FT_filter_2D
// (smoothing_library.pyx:37-114, :141-203).
//
// Accumulation: each thread keeps the sums of its current |k| bin in registers and issues red.global.add.f64 /
// .u64 only on a bin change, into one of SHELL_NREP replicas of the (small) accumulator block chosen by
// blockIdx; a second tiny kernel folds the replicas.  Counts are uint64, everything else float64.
#include "common.cuh"
#include "shell_body.cuh"

namespace pyl {

constexpr int SHELL_BLOCK = 128;
constexpr int SHELL_NREP = 8;

struct DeviceSink {
    unsigned long long *base;
    __device__ __forceinline__ void add(long long w, double v) { atomicAdd(reinterpret_cast<double *>(base + w), v); }
    __device__ __forceinline__ void count(long long w, unsigned long long c) { atomicAdd(base + w, c); }
};

template <int KIND>
__global__ void __launch_bounds__(SHELL_BLOCK) shell_bin_kernel(const ShellArgs A, unsigned long long *rep,
                                                                long long rep_words) {
    const long long t = (long long)blockIdx.x * SHELL_BLOCK + threadIdx.x;
    if (t >= A.T) return;
    DeviceSink sink{rep + (long long)(blockIdx.x % SHELL_NREP) * rep_words};
    shell_thread<KIND>(A, t, (int)blockIdx.y, sink);
}

// out[w] = sum over replicas; words [n3, 2*n3) are uint64 counts, the rest float64
__global__ void shell_fold_kernel(unsigned long long *out, const unsigned long long *rep, long long words, int n3) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    if (w >= n3 && w < 2LL * n3) {
        unsigned long long s = 0;
        for (int r = 0; r < SHELL_NREP; r++) s += rep[(long long)r * words + w];
        out[w] = s;
    } else {
        double s = 0.0;
        for (int r = 0; r < SHELL_NREP; r++) s += __longlong_as_double((long long)rep[(long long)r * words + w]);
        out[w] = (unsigned long long)__double_as_longlong(s);
    }
}

__global__ void shell_window_kernel(double *tab, int m1, int N, int p0, int p1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m1) return;
    tab[i] = shell_window(i, N, p0);
    tab[m1 + i] = shell_window(i, N, p1);
}

template <int OP>
__global__ void __launch_bounds__(256) mode_kernel(const ModeArgs A) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < A.total; e += stride) mode_element<OP>(A, e);
}

// a[e] *= b[e], complex64 (smoothing_library.pyx:227-232)
__global__ void __launch_bounds__(256) cmul_kernel(float2 *__restrict__ a, const float2 *__restrict__ b, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const float2 x = a[e], y = __ldg(b + e);
        float2 r;
        r.x = __fsub_rn(__fmul_rn(x.x, y.x), __fmul_rn(x.y, y.y));
        r.y = __fadd_rn(__fmul_rn(x.x, y.y), __fmul_rn(x.y, y.x));
        a[e] = r;
    }
}

// v[i] *= (1 + d[i]) : the momentum field (1+delta)*V of XPk_dv / XPk_vv (Pk_library.pyx:1367, :1491-1492),
// float32 like NumPy's in-place `Vx *= (1.0 + delta)`
__global__ void __launch_bounds__(256) mul_one_plus_kernel(float *__restrict__ v, const float *__restrict__ d, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        v[i] = __fmul_rn(v[i], __fadd_rn(1.0f, __ldg(d + i)));
}

template <int KIND>
__global__ void __launch_bounds__(256) filter_kernel(const FilterArgs A) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < A.total; e += stride) filter_element<KIND>(A, e);
}

// x[i] = float(double(x[i]) / d[0]) : `field[i,j,l] = field[i,j,l]/normalization` with a double normalisation
// (smoothing_library.pyx:110-114)
__global__ void __launch_bounds__(256) divide_by_f64_kernel(float *__restrict__ x, long long n, const double *__restrict__ d) {
    const double div = d[0];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        x[i] = __double2float_rn(__ddiv_rn((double)x[i], div));
}

static int shell_bins(int kind, int dims) {
    const double m = (double)(dims / 2);
    const bool plane = (kind == SK_PLANE || kind == SK_XPLANE);
    return (int)sqrt(plane ? 2.0 * m * m : 3.0 * m * m) + 1;     // kmax + 1 (frequencies / frequencies_2D)
}

static size_t shell_ws_bytes(int kind, int dims) {
    const size_t tab = align_up((size_t)2 * (dims / 2 + 1) * sizeof(double), 256);
    const size_t words = (size_t)(2 + shell_nvals(kind)) * shell_bins(kind, dims);
    return tab + (size_t)SHELL_NREP * words * 8;
}

static unsigned grid_for(long long n, int block) {
    long long b = (n + block - 1) / block;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}


// ---- XXi_projected (Pk_library.pyx:2684-2789): the 2D forms of the mode pass and of the real-space binning -------
// mode pass: every stored mode of the (grid, grid/2+1) half-spectra is deconvolved (no duplicate-mode rule in the
// reference's 2D loop, :2735-2758) and a_k <- (re_a*re_b + im_a*im_b, 0), products and sum in float32
__global__ void __launch_bounds__(256) modes_power_2d_kernel(float2 *__restrict__ a, const float2 *__restrict__ b,
                                                             const double *__restrict__ wa, const double *__restrict__ wb,
                                                             int N, int m) {
    const int m1 = m + 1;
    const long long total = (long long)N * m1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int kxx = (int)(t / m1), kyy = (int)(t - (long long)kxx * m1);
        const int ax = kxx > m ? N - kxx : kxx;                   // the window is even in k
        const float fa = __double2float_rn(wa[ax] * wa[kyy]);    // cdef float MAS_factor = corr_x * corr_y (:2745-2748)
        const float fb = __double2float_rn(wb[ax] * wb[kyy]);
        const float2 va = a[t], vb = b[t];
        const float r1 = __fmul_rn(va.x, fa), i1 = __fmul_rn(va.y, fa);
        const float r2 = __fmul_rn(vb.x, fb), i2 = __fmul_rn(vb.y, fb);
        a[t] = make_float2(__fadd_rn(__fmul_rn(r1, r2), __fmul_rn(i1, i2)), 0.0f);
    }
}

// binning of every cell of the (grid, grid) image by int(sqrt(kx^2 + ky^2)) (:2776-2787): out = [sum k | count | sum xi]
// per bin, float64 / uint64; a block keeps private sums in shared memory when the bins fit
__global__ void __launch_bounds__(256) radial_bin_2d_kernel(const float *__restrict__ img, int N, int m, int nb,
                                                            float scale, unsigned long long *__restrict__ out,
                                                            int use_smem) {
    extern __shared__ unsigned long long rb_smem[];
    if (use_smem) {
        for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) rb_smem[i] = 0ull;
        __syncthreads();
    }
    unsigned long long *acc = use_smem ? rb_smem : out;
    const long long total = (long long)N * N;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int kxx = (int)(t / N), kyy = (int)(t - (long long)kxx * N);
        const int kx = kxx > m ? kxx - N : kxx, ky = kyy > m ? kyy - N : kyy;
        const double k = sqrt((double)(kx * kx + ky * ky));
        const int bin = (int)k;
        if (bin >= nb) continue;
        const float v = __fmul_rn(__ldg(img + t), scale);
        atomicAdd(reinterpret_cast<double *>(acc + bin), k);
        atomicAdd(acc + nb + bin, 1ull);
        atomicAdd(reinterpret_cast<double *>(acc + 2 * nb + bin), (double)v);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) {
            const unsigned long long w = rb_smem[i];
            if (w == 0ull) continue;
            if (i >= nb && i < 2 * nb) atomicAdd(out + i, w);
            else atomicAdd(reinterpret_cast<double *>(out + i), __longlong_as_double((long long)w));
        }
    }
}

}  // namespace pyl

using namespace pyl;

extern "C" {

int pyl_shell_layout(int kind, int dims, int *bins, int *values) {
    PYL_REQUIRE(kind >= 0 && kind < SK_COUNT && dims > 0, "pyl_shell_layout: bad kind/dims");
    if (bins) *bins = shell_bins(kind, dims);
    if (values) *values = shell_nvals(kind);
    return PYL_OK;
}

size_t pyl_shell_bin_workspace_bytes(int kind, int dims) {
    if (kind < 0 || kind >= SK_COUNT || dims <= 0) return 0;
    return shell_ws_bytes(kind, dims);
}

int pyl_shell_bin(int kind, const float *const *fields, int nfields, const int *mas_index, int dims, int axis,
                  float scale, const pyl_shell_table_t *table, void *out, void *ws, size_t ws_bytes,
                  pyl_stream_t stream) {
    PYL_REQUIRE(kind >= 0 && kind < SK_COUNT, "pyl_shell_bin: unknown kind");
    PYL_REQUIRE(dims > 0 && dims <= 8192, "pyl_shell_bin: dims must be in 1..8192");
    PYL_REQUIRE(nfields == shell_nfields(kind), "pyl_shell_bin: wrong number of fields for this kind");
    PYL_REQUIRE(out != nullptr, "pyl_shell_bin: NULL output");
    PYL_REQUIRE(axis >= 0 && axis <= 2, "pyl_shell_bin: axis must be 0, 1 or 2");
    for (int f = 0; f < nfields; f++) PYL_REQUIRE(fields != nullptr && fields[f] != nullptr, "pyl_shell_bin: NULL field");
    int p0 = 0, p1 = 0;
    if (kind != SK_XI && kind != SK_EXPECTED) {
        PYL_REQUIRE(mas_index != nullptr, "pyl_shell_bin: NULL mas_index");
        p0 = mas_index[0];
        p1 = (kind == SK_XPLANE) ? mas_index[1] : p0;
        PYL_REQUIRE(p0 >= 0 && p0 <= 4 && p1 >= 0 && p1 <= 4, "pyl_shell_bin: mas_index must be 0..4");
    }
    if (kind == SK_EXPECTED)
        PYL_REQUIRE(table != nullptr && table->k != nullptr && table->P != nullptr && table->n >= 2 && table->deltak > 0.0,
                    "pyl_shell_bin: expected_Pk needs an interpolation table");
    const size_t need = shell_ws_bytes(kind, dims);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("pyl_shell_bin: workspace of %zu bytes required, %zu given", need, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    const int m = dims / 2, m1 = m + 1;
    double *tab = reinterpret_cast<double *>(ws);
    const size_t tab_bytes = align_up((size_t)2 * m1 * sizeof(double), 256);
    unsigned long long *rep = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(ws) + tab_bytes);

    ShellArgs A;
    memset(&A, 0, sizeof(A));
    for (int f = 0; f < nfields; f++) A.f[f] = fields[f];
    A.win[0] = tab; A.win[1] = tab + m1;
    A.N = dims; A.m = m; A.even = (dims % 2 == 0);
    const bool plane = (kind == SK_PLANE || kind == SK_XPLANE);
    A.nx = plane ? 1 : dims;
    A.nzs = (kind == SK_XI) ? dims : m1;
    A.hermitian = (kind == SK_XI) ? 0 : 1;
    A.axis = axis;
    A.n3 = shell_bins(kind, dims);
    A.scale = scale;
    if (kind == SK_EXPECTED) {
        A.tab_k = table->k; A.tab_P = table->P; A.tab_n = table->n;
        A.kF = table->kF; A.log10_kmin = table->log10_kmin; A.deltak = table->deltak;
    }
    shell_geometry(A, sm_count());
    const long long words = (long long)(2 + shell_nvals(kind)) * A.n3;

    PYL_CUDA_CHECK(cudaMemsetAsync(rep, 0, (size_t)SHELL_NREP * words * 8, s));
    shell_window_kernel<<<(m1 + 127) / 128, 128, 0, s>>>(tab, m1, dims, p0, p1);
    PYL_LAUNCH_CHECK();
    dim3 grid((unsigned)((A.T + SHELL_BLOCK - 1) / SHELL_BLOCK), (unsigned)A.nseg);
    switch (kind) {
        case SK_THETA: shell_bin_kernel<SK_THETA><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        case SK_DV: shell_bin_kernel<SK_DV><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        case SK_VV: shell_bin_kernel<SK_VV><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        case SK_EXPECTED: shell_bin_kernel<SK_EXPECTED><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        case SK_PLANE: shell_bin_kernel<SK_PLANE><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        case SK_XPLANE: shell_bin_kernel<SK_XPLANE><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
        default: shell_bin_kernel<SK_XI><<<grid, SHELL_BLOCK, 0, s>>>(A, rep, words); break;
    }
    PYL_LAUNCH_CHECK();
    shell_fold_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(reinterpret_cast<unsigned long long *>(out), rep,
                                                                      words, A.n3);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

size_t pyl_modes_workspace_bytes(int dims) {
    if (dims <= 0) return 0;
    return align_up((size_t)2 * (dims / 2 + 1) * sizeof(double), 256);
}

static int modes_entry(int op, float *a_k, const float *b_k, int dims, int mas_a, int mas_b, void *ws, size_t ws_bytes,
                       pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && a_k != nullptr, "pyl_modes_*: bad dims or NULL field");
    PYL_REQUIRE(mas_a >= 0 && mas_a <= 4 && mas_b >= 0 && mas_b <= 4, "pyl_modes_*: mas_index must be 0..4");
    const size_t need = pyl_modes_workspace_bytes(dims);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("pyl_modes_*: workspace of %zu bytes required, %zu given", need, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    const int m1 = dims / 2 + 1;
    double *tab = reinterpret_cast<double *>(ws);
    shell_window_kernel<<<(m1 + 127) / 128, 128, 0, s>>>(tab, m1, dims, mas_a, mas_b);
    PYL_LAUNCH_CHECK();
    ModeArgs A;
    A.a = reinterpret_cast<float2 *>(a_k);
    A.b = reinterpret_cast<const float2 *>(b_k);
    A.win[0] = tab; A.win[1] = tab + m1;
    A.N = dims; A.m = dims / 2; A.even = (dims % 2 == 0);
    A.total = (long long)dims * dims * m1;
    const unsigned g = grid_for(A.total, 256);
    if (op == MO_DECONVOLVE) mode_kernel<MO_DECONVOLVE><<<g, 256, 0, s>>>(A);
    else mode_kernel<MO_POWER><<<g, 256, 0, s>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_modes_deconvolve(float *delta_k, int dims, int mas_index, void *ws, size_t ws_bytes, pyl_stream_t stream) {
    return modes_entry(MO_DECONVOLVE, delta_k, nullptr, dims, mas_index, mas_index, ws, ws_bytes, stream);
}

int pyl_modes_power(float *a_k, const float *b_k, int dims, int mas_a, int mas_b, void *ws, size_t ws_bytes,
                    pyl_stream_t stream) {
    return modes_entry(MO_POWER, a_k, b_k, dims, mas_a, b_k ? mas_b : mas_a, ws, ws_bytes, stream);
}

int pyl_cmul_inplace(float *a_k, const float *b_k, int64_t n_complex, pyl_stream_t stream) {
    PYL_REQUIRE(n_complex >= 0, "pyl_cmul_inplace: negative size");
    if (n_complex == 0) return PYL_OK;
    PYL_REQUIRE(a_k != nullptr && b_k != nullptr, "pyl_cmul_inplace: NULL pointer");
    cmul_kernel<<<grid_for(n_complex, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<float2 *>(a_k),
                                                                         reinterpret_cast<const float2 *>(b_k), n_complex);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_filter_fill(int kind, void *out, int dims, int axes, float R2, float kF, float kmin, float kmax,
                    pyl_stream_t stream) {
    PYL_REQUIRE(kind >= FK_TOPHAT && kind <= FK_TOPHAT_K, "pyl_filter_fill: kind must be 0 (Top-Hat), 1 (Gaussian) or 2 (Top-Hat-k)");
    PYL_REQUIRE(out != nullptr && dims > 0 && (axes == 2 || axes == 3), "pyl_filter_fill: bad output, dims or axes");
    FilterArgs A;
    A.real = reinterpret_cast<float *>(out);
    A.cplx = reinterpret_cast<float2 *>(out);
    A.N = dims; A.m = dims / 2; A.axes = axes;
    A.R2 = R2; A.kF = kF; A.kmin = kmin; A.kmax = kmax;
    const long long last = (kind == FK_TOPHAT_K) ? dims / 2 + 1 : dims;
    A.total = (axes == 3 ? (long long)dims * dims : (long long)dims) * last;
    const unsigned g = grid_for(A.total, 256);
    cudaStream_t s = as_stream(stream);
    if (kind == FK_TOPHAT) filter_kernel<FK_TOPHAT><<<g, 256, 0, s>>>(A);
    else if (kind == FK_GAUSSIAN) filter_kernel<FK_GAUSSIAN><<<g, 256, 0, s>>>(A);
    else filter_kernel<FK_TOPHAT_K><<<g, 256, 0, s>>>(A);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_divide_by_f64(float *x, int64_t n, const double *divisor, pyl_stream_t stream) {
    PYL_REQUIRE(n >= 0, "pyl_divide_by_f64: negative size");
    if (n == 0) return PYL_OK;
    PYL_REQUIRE(x != nullptr && divisor != nullptr, "pyl_divide_by_f64: NULL pointer");
    divide_by_f64_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, n, divisor);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_mul_one_plus(float *v, const float *delta, int64_t n, pyl_stream_t stream) {
    PYL_REQUIRE(n >= 0, "pyl_mul_one_plus: negative size");
    if (n == 0) return PYL_OK;
    PYL_REQUIRE(v != nullptr && delta != nullptr, "pyl_mul_one_plus: NULL pointer");
    mul_one_plus_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(v, delta, n);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}


int pyl_modes_power_2d(float *a_k, const float *b_k, int dims, int mas_a, int mas_b, void *ws, size_t ws_bytes,
                       pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && a_k != nullptr && b_k != nullptr, "pyl_modes_power_2d: bad dims or NULL field");
    PYL_REQUIRE(mas_a >= 0 && mas_a <= 4 && mas_b >= 0 && mas_b <= 4, "pyl_modes_power_2d: mas_index must be 0..4");
    const size_t need = pyl_modes_workspace_bytes(dims);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("pyl_modes_power_2d: workspace of %zu bytes required, %zu given", need, ws_bytes);
        return PYL_ERR_WORKSPACE;
    }
    cudaStream_t s = as_stream(stream);
    const int m = dims / 2, m1 = m + 1;
    double *tab = reinterpret_cast<double *>(ws);
    shell_window_kernel<<<(m1 + 127) / 128, 128, 0, s>>>(tab, m1, dims, mas_a, mas_b);
    PYL_LAUNCH_CHECK();
    const long long total = (long long)dims * m1;
    modes_power_2d_kernel<<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<float2 *>(a_k),
                                                               reinterpret_cast<const float2 *>(b_k), tab, tab + m1, dims, m);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

int pyl_radial_bin_2d(const float *image, int dims, float scale, void *out, pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && image != nullptr && out != nullptr, "pyl_radial_bin_2d: bad dims or NULL pointer");
    const int nb = (int)((dims / 2) * sqrt(2.0)) + 1;                  // kmax + 1, kmax = int((grid//2)*sqrt(2)) (:2722)
    cudaStream_t s = as_stream(stream);
    PYL_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)3 * nb * 8, s));
    const size_t smem = (size_t)3 * nb * 8;
    const int use_smem = smem <= 40 * 1024;
    long long blocks = ((long long)dims * dims + 256 * 16 - 1) / (256 * 16);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    radial_bin_2d_kernel<<<(unsigned)blocks, 256, use_smem ? smem : 0, s>>>(
        image, dims, dims / 2, nb, scale, reinterpret_cast<unsigned long long *>(out), use_smem);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

}  // extern "C"
