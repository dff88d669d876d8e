// Per-thread bodies of the "shell" kernels: estimators of Pk_library.pyx that reduce every mode (or every
// real-space cell) of a grid into bins of |k| (one bin per fundamental frequency), plus the elementwise passes
// over a half-spectrum that precede an inverse transform.
//
//   SK_THETA     Pk_theta    Pk_library.pyx:1273-1316   fields Vx,Vy,Vz (complex64)           1 value / bin
//   SK_DV        XPk_dv      :1386-1432                  fields delta,Vx,Vy,Vz                 3 values
//   SK_VV        XPk_vv      :1515-1568                  fields Vx1,Vy1,Vz1,Vx2,Vy2,Vz2        3 values
//   SK_EXPECTED  expected_Pk :2004-2037                  no field, interpolation table         1 value
//   SK_PLANE     Pk_plane    :470-499                    one (grid, grid/2+1) image            1 value
//   SK_XPLANE    XPk_plane   :1151-1200                  two images, two windows               3 values
//   SK_XI        Xi / XXi    :2233-2267 / :2378-2412     one REAL (dims,dims,dims) grid        3 values (l=0,2,4)
//
// Everything here is `__host__ __device__`: the CUDA kernels in pk_shell.cu are thin wrappers that map
// (blockIdx, threadIdx) to (t, seg) and hand the body a sink that issues red.global; tests/harness/shell_host.cpp
// runs the SAME bodies serially on the CPU with a plain-add sink, so that the decomposition of the grid into
// threads and segments, the Hermitian-duplicate rule, the index arithmetic and every per-mode functional are
// checked against the CPU restatement of the reference in the GPU-less authoring container.  That harness is test infrastructure only.
//
// Decomposition: thread t owns (|kx|, |kz|) -- lanes run along the contiguous last axis, so a warp's loads are
// 256-byte row segments -- and walks |ky| over its segment of [0, N/2], handling the sign combinations that share
// |k| together (see shell_thread).  k^2 grows monotonically along the walk, so the sums of the current bin live
// in registers and reach memory only when the bin changes.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define PYL_HD __host__ __device__ __forceinline__
#else
#define PYL_HD inline
struct float2 { float x, y; };
#endif

namespace pyl {

enum ShellKind { SK_THETA = 0, SK_DV = 1, SK_VV = 2, SK_EXPECTED = 3, SK_PLANE = 4, SK_XPLANE = 5, SK_XI = 6, SK_COUNT = 7 };

PYL_HD int shell_nvals(int kind) { return (kind == SK_THETA || kind == SK_EXPECTED || kind == SK_PLANE) ? 1 : 3; }
PYL_HD int shell_nfields(int kind) {
    return kind == SK_THETA ? 3 : kind == SK_DV ? 4 : kind == SK_VV ? 6 : kind == SK_EXPECTED ? 0 : kind == SK_XPLANE ? 2 : 1;
}

struct ShellArgs {
    const void *f[6];        // fields (float2 half-spectra, or one float grid for SK_XI)
    const double *win[2];    // (x/sin x)^p per |k| index, [m+1]; win[1] is the second image's window (SK_XPLANE)
    int N, m, even;
    int nx;                  // stored rows of the slowest axis: N (cube) or 1 (image: kx = 0)
    int nzs;                 // stored extent of the last axis: m+1 (half-spectrum) or N (real grid)
    int hermitian;           // apply the duplicate-mode rule of Pk_library.pyx:324-327 / :480-481
    int axis;                // line of sight (SK_XI)
    int seg_len, nseg;
    long long T;             // threads per segment = nx * nzs
    int n3;                  // bins = kmax + 1
    float scale;             // SK_XI: value = grid * scale (the 1/N^3 of the normalised inverse transform)
    // SK_EXPECTED
    const float *tab_k, *tab_P;
    int tab_n;
    float kF;
    double log10_kmin, deltak;
};

// word offsets inside one accumulator block of (2 + NV) * n3 eight-byte words
PYL_HD long long shell_off_ksum(const ShellArgs &, int bin) { return bin; }
PYL_HD long long shell_off_count(const ShellArgs &A, int bin) { return (long long)A.n3 + bin; }
PYL_HD long long shell_off_val(const ShellArgs &A, int j, int bin) { return (long long)(2 + j) * A.n3 + bin; }

PYL_HD float shell_fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);          // a rounded float32 product, never contracted
#else
    return a * b;
#endif
}
PYL_HD float shell_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
template <class T>
PYL_HD T shell_ld(const T *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// MAS window (x/sin x)^p at x = pi*i/N  (Pk_library.pyx:83-84; even in k, so tabulated per |k|)
PYL_HD double shell_window(int i, int N, int p) {
    if (i == 0 || p == 0) return 1.0;
    const double x = (M_PI / (double)N) * (double)i;
    return pow(x / sin(x), (double)p);
}

// what one walk step needs from memory
template <int KIND>
struct ShellLoad {
    float2 c[(KIND == SK_THETA) ? 3 : (KIND == SK_DV) ? 4 : (KIND == SK_VV) ? 6 : (KIND == SK_XPLANE) ? 2 : 1];
    float r;
};

template <int KIND>
PYL_HD void shell_fetch(const ShellArgs &A, long long e, ShellLoad<KIND> &L) {
    if (KIND == SK_XI) {
        L.r = shell_ld(reinterpret_cast<const float *>(A.f[0]) + e);
    } else if (KIND != SK_EXPECTED) {
        constexpr int NF = (KIND == SK_THETA) ? 3 : (KIND == SK_DV) ? 4 : (KIND == SK_VV) ? 6 : (KIND == SK_XPLANE) ? 2 : 1;
#pragma unroll
        for (int c = 0; c < NF; c++) L.c[c] = shell_ld(reinterpret_cast<const float2 *>(A.f[c]) + e);
    }
}

// k . V of three deconvolved complex components, formed in float32 like the reference's C `int * float`
// products (Pk_library.pyx:1303-1309)
PYL_HD void shell_kdotv(const float2 *V, float fac, int kx, int ky, int kz, float &dre, float &dim) {
    const float fx = (float)kx, fy = (float)ky, fz = (float)kz;
    dre = shell_fma(fz, shell_fmul(V[2].x, fac), shell_fma(fy, shell_fmul(V[1].x, fac), shell_fmul(fx, shell_fmul(V[0].x, fac))));
    dim = shell_fma(fz, shell_fmul(V[2].y, fac), shell_fma(fy, shell_fmul(V[1].y, fac), shell_fmul(fx, shell_fmul(V[0].y, fac))));
}

// The per-mode functional of one loaded mode/cell with signed wavenumbers (kx, ky, kz); adds into v[].
template <int KIND>
PYL_HD void shell_mode(const ShellArgs &A, const ShellLoad<KIND> &L, int kx, int ky, int kz, float fac, float fac1,
                       double k, int k2, double *v) {
    if (KIND == SK_XI) {
        const int kpar = (A.axis == 0) ? kx : (A.axis == 1 ? ky : kz);
        const double mu = (k2 == 0) ? 0.0 : (double)kpar / k;
        const double mu2 = mu * mu;
        const double x = (double)shell_fmul(L.r, A.scale);
        v[0] += x;
        v[1] += (x * (3.0 * mu2 - 1.0) / 2.0);
        v[2] += (x * (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0);
    } else if (KIND == SK_PLANE) {
        const double re = (double)shell_fmul(L.c[0].x, fac), im = (double)shell_fmul(L.c[0].y, fac);
        v[0] += re * re + im * im;
    } else if (KIND == SK_XPLANE) {
        const double r0 = (double)shell_fmul(L.c[0].x, fac), i0 = (double)shell_fmul(L.c[0].y, fac);
        const double r1 = (double)shell_fmul(L.c[1].x, fac1), i1 = (double)shell_fmul(L.c[1].y, fac1);
        v[0] += r0 * r0 + i0 * i0;
        v[1] += r1 * r1 + i1 * i1;
        v[2] += r0 * r1 + i0 * i1;
    } else if (KIND == SK_THETA) {
        float dre, dim;
        shell_kdotv(&L.c[0], fac, kx, ky, kz, dre, dim);
        const double re = -(double)dim, im = (double)dre;          // theta = i k.V (:1303-1309)
        v[0] += re * re + im * im;
    } else if (KIND == SK_DV || KIND == SK_VV) {
        double r1, i1, r2, i2;
        float dre, dim;
        if (KIND == SK_DV) {
            r1 = (double)shell_fmul(L.c[0].x, fac); i1 = (double)shell_fmul(L.c[0].y, fac);
            shell_kdotv(&L.c[1], fac, kx, ky, kz, dre, dim);
            r2 = (double)dim; i2 = -(double)dre;                   // :1419-1425
        } else {
            shell_kdotv(&L.c[0], fac, kx, ky, kz, dre, dim);
            r1 = (double)dim; i1 = -(double)dre;                   // :1549-1554
            shell_kdotv(&L.c[KIND == SK_VV ? 3 : 0], fac, kx, ky, kz, dre, dim);
            r2 = (double)dim; i2 = -(double)dre;                   // :1556-1561
        }
        v[0] += r1 * r1 + i1 * i1;
        v[1] += r2 * r2 + i2 * i2;
        v[2] += r1 * r2 + i1 * i2;
    }
}

// The body.  Sink: add(word, double) and count(word, unsigned long long) accumulate into the block the caller chose.
//
// FOLD: thread t owns (|kx|, |kz|) = (a, iz) and walks s = |ky| over its segment of [0, m].  The sign combinations
// (+-kx, +-ky[, +-kz for the real grid of SK_XI]) that exist as separate stored elements share |k|, hence the bin,
// the (even) window factor and mu^2: they are loaded together, their functionals (which see the SIGNED wavenumbers)
// are summed into the same registers, and the duplicate-mode rule is applied per mode, so counts stay exact.
// k^2 = a^2 + iz^2 + s^2 grows monotonically along the walk: a bin is flushed once per thread and the flush is
// shared by 4 (8) modes -- the red.global rate, not the loads, bounded the unfolded form of this kernel
// (measured: Xi binning 2.56 ms at 512^3 unfolded).
template <int KIND, class Sink>
PYL_HD void shell_thread(const ShellArgs &A, long long t, int seg, Sink &sink) {
    constexpr int NV = (KIND == SK_THETA || KIND == SK_EXPECTED || KIND == SK_PLANE) ? 1 : 3;
    constexpr int NZ = (KIND == SK_XI) ? 2 : 1;
    const int N = A.N, m = A.m, nzh = A.m + 1;
    const bool even = A.even != 0;
    const int a = (A.nx == 1) ? 0 : (int)(t / nzh);
    const int iz = (int)(t - (long long)(t / nzh) * nzh);
    const int s0 = seg * A.seg_len;
    const int s1 = (s0 + A.seg_len < nzh) ? s0 + A.seg_len : nzh;
    if (s0 >= s1) return;

    const bool has_nx = (A.nx != 1) && a != 0 && !(even && a == m);        // -a is a separate stored row
    const bool has_nz = (NZ == 2) && iz != 0 && !(even && iz == m);
    const bool zspecial = A.hermitian && (iz == 0 || (iz == m && even));
    const bool xself = (a == 0) || (a == m && even);
    // duplicate-mode rule (Pk_library.pyx:324-327): on the special planes drop kx < 0, and ky < 0 where kx is 0 / Nyquist
    const bool use_nx = has_nx && !zspecial;
    const bool drop_ny = zspecial && xself;

    double cx0 = 1.0, cz0 = 1.0, cx1 = 1.0, cz1 = 1.0;
    if (KIND != SK_XI && KIND != SK_EXPECTED) {
        cx0 = A.win[0][a]; cz0 = A.win[0][iz];
        if (KIND == SK_XPLANE) { cx1 = A.win[1][a]; cz1 = A.win[1][iz]; }
    }
    const int base2 = a * a + iz * iz;
    const int ixs[2] = {a, N - a};
    const int izs[2] = {iz, N - iz};
    const int kxs[2] = {a, -a};
    const int kzs[2] = {iz, -iz};
    const int nxv = use_nx ? 2 : 1, nzv = has_nz ? 2 : 1;

    int cur = -1;
    unsigned int cnt = 0;
    double ksum = 0.0, v[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = 0.0;

    auto flush = [&]() {
        if (cnt != 0) {
            sink.count(shell_off_count(A, cur), cnt);
            sink.add(shell_off_ksum(A, cur), ksum);
#pragma unroll
            for (int j = 0; j < NV; j++) sink.add(shell_off_val(A, j, cur), v[j]);
        }
        cnt = 0; ksum = 0.0;
#pragma unroll
        for (int j = 0; j < NV; j++) v[j] = 0.0;
    };

    for (int s = s0; s < s1; s++) {
        const bool use_ny = s != 0 && !(even && s == m) && !drop_ny;
        const int nyv = use_ny ? 2 : 1;
        const int iys[2] = {s, N - s};
        // all loads of the step first (independent, in flight together), then the arithmetic
        ShellLoad<KIND> L[2][2][NZ];
        if (KIND != SK_EXPECTED) {
#pragma unroll
            for (int jx = 0; jx < 2; jx++)
#pragma unroll
                for (int jy = 0; jy < 2; jy++)
#pragma unroll
                    for (int jz = 0; jz < NZ; jz++)
                        if (jx < nxv && jy < nyv && jz < nzv)
                            shell_fetch<KIND>(A, ((long long)ixs[jx] * N + iys[jy]) * A.nzs + izs[jz], L[jx][jy][jz]);
        }
        const int mult = nxv * nyv * nzv;
        const int k2 = base2 + s * s;
        const double k = sqrt((double)k2);
        int kidx = (int)k;
        if (KIND == SK_EXPECTED) {
            // `cdef float k` in the reference (Pk_library.pyx:1961,2021-2028); the DC mode is skipped (:2025)
            float kf = (float)k;
            kidx = (int)kf;
            if (kf != 0.0f) {
                if (kidx != cur) { flush(); cur = kidx; }
                kf = shell_fmul(kf, A.kF);
                int i = (int)((log10((double)kf) - A.log10_kmin) / A.deltak);
                i = i < 0 ? 0 : (i > A.tab_n - 2 ? A.tab_n - 2 : i);     // the reference reads out of bounds here
                const float k0 = A.tab_k[i], k1 = A.tab_k[i + 1], p0 = A.tab_P[i], p1 = A.tab_P[i + 1];
                const float Pi = shell_fmul((p1 - p0) / (k1 - k0), kf - k0) + p0;
                cnt += mult; ksum += (double)mult * (double)kf; v[0] += (double)mult * (double)Pi;
            }
            continue;
        }
        if (kidx != cur) { flush(); cur = kidx; }
        cnt += mult; ksum += (double)mult * k;
        // window factor: product in float64 in the reference's order, rounded to float32 (:351, :488)
        float fac = 1.0f, fac1 = 1.0f;
        if (KIND != SK_XI) {
            fac = (float)(cx0 * A.win[0][s] * cz0);
            if (KIND == SK_XPLANE) fac1 = (float)(cx1 * A.win[1][s] * cz1);
        }
#pragma unroll
        for (int jx = 0; jx < 2; jx++)
#pragma unroll
            for (int jy = 0; jy < 2; jy++)
#pragma unroll
                for (int jz = 0; jz < NZ; jz++)
                    if (jx < nxv && jy < nyv && jz < nzv)
                        shell_mode<KIND>(A, L[jx][jy][jz], kxs[jx], jy ? -s : s, kzs[jz], fac, fac1, k, k2, v);
    }
    flush();
}

// launch geometry shared by the device launcher and the host harness
PYL_HD void shell_geometry(ShellArgs &A, int sms) {
    const int walk = A.m + 1;                            // |ky| = 0..m
    A.T = (long long)(A.nx == 1 ? 1 : A.m + 1) * (A.m + 1);
    const long long warps_per_seg = (A.T + 31) / 32;
    const long long want_warps = (long long)sms * 64;
    long long nseg = (want_warps + warps_per_seg - 1) / warps_per_seg;
    const long long max_seg = (walk + 7) / 8;            // at least 8 walk steps per segment
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    A.seg_len = (int)((walk + nseg - 1) / nseg);
    if (A.seg_len < 1) A.seg_len = 1;
    A.nseg = (walk + A.seg_len - 1) / A.seg_len;
}

// ---- elementwise passes over a (N, N, N/2+1) half-spectrum ---------------------------------------------------
enum ModeOp { MO_DECONVOLVE = 0, MO_POWER = 1 };

struct ModeArgs {
    float2 *a;               // in/out
    const float2 *b;         // MO_POWER: second field or nullptr (auto)
    const double *win[2];
    int N, m, even;
    long long total;         // N*N*(m+1)
};

// MO_DECONVOLVE -- correct_MAS, Pk_library.pyx:1909-1929.  The reference multiplies the modes it counts as
// independent by the float32 window w and leaves their Hermitian duplicates on the kz = 0 / Nyquist planes
// untouched, then runs a c2r transform on the (now non-Hermitian) planes.  A c2r transform that goes
// "c2c over x,y then c2r over z" (FFTW, pocketfft) uses only the Hermitian part of those planes, i.e. both
// members of a duplicate pair effectively carry (1 + w)/2.  That is applied here explicitly, so the planes stay
// Hermitian and the result does not depend on how cuFFT treats inconsistent input.  Self-conjugate modes get w.
// MO_POWER -- Xi / XXi, :2198-2218 / :2335-2362: every stored mode becomes (re1*re2 + im1*im2, 0) of the
// deconvolved values, all in float32 (`cdef float real, imag`).
template <int OP>
PYL_HD void mode_element(const ModeArgs &A, long long e) {
    const int N = A.N, m = A.m, nz = m + 1;
    const bool even = A.even != 0;
    const int iz = (int)(e % nz);
    const long long q = e / nz;
    const int iy = (int)(q % N), ix = (int)(q / N);
    const int kx = ix > m ? ix - N : ix, ky = iy > m ? iy - N : iy;
    const int ax = kx < 0 ? -kx : kx, ay = ky < 0 ? -ky : ky;
    const float w0 = (float)(A.win[0][ax] * A.win[0][ay] * A.win[0][iz]);
    float2 v = A.a[e];
    if (OP == MO_DECONVOLVE) {
        float f = w0;
        const bool zspecial = (iz == 0) || (iz == m && even);
        if (zspecial) {
            const bool selfconj = (kx == 0 || (kx == m && even)) && (ky == 0 || (ky == m && even));
            if (!selfconj) f = shell_fmul(0.5f, 1.0f + w0);
        }
        v.x = shell_fmul(v.x, f); v.y = shell_fmul(v.y, f);
    } else {
        const float r1 = shell_fmul(v.x, w0), i1 = shell_fmul(v.y, w0);
        float r2 = r1, i2 = i1;
        if (A.b != nullptr) {
            const float w1 = (float)(A.win[1][ax] * A.win[1][ay] * A.win[1][iz]);
            const float2 u = shell_ld(A.b + e);
            r2 = shell_fmul(u.x, w1); i2 = shell_fmul(u.y, w1);
        }
        v.x = shell_fma(i1, i2, shell_fmul(r1, r2));
        v.y = 0.0f;
    }
    A.a[e] = v;
}

// ---- smoothing filters placed on a grid (smoothing_library.pyx:37-100 / :141-191) --------------------------------
enum FilterKind { FK_TOPHAT = 0, FK_GAUSSIAN = 1, FK_TOPHAT_K = 2 };

struct FilterArgs {
    float *real;             // FK_TOPHAT / FK_GAUSSIAN: (N,)*axes float32
    float2 *cplx;            // FK_TOPHAT_K: (N,)*(axes-1) x (N/2+1) complex64
    int N, m, axes;
    float R2, kF, kmin, kmax;       // `cdef float` in the reference
    long long total;
};

// element e of the filter grid.  d2 is an int, `d2 <= R2` compares it as a float, the Gaussian is a double exp
// stored as float, k = kF*sqrt(d2) is a float (the reference's C types).
template <int KIND>
PYL_HD void filter_element(const FilterArgs &A, long long e) {
    const int N = A.N, m = A.m;
    const int nlast = (KIND == FK_TOPHAT_K) ? m + 1 : N;
    const int il = (int)(e % nlast);
    long long q = e / nlast;
    const int ij = (int)(q % N);
    const int ii = (A.axes == 3) ? (int)(q / N) : 0;
    const int l1 = il > m ? il - N : il, j1 = ij > m ? ij - N : ij, i1 = ii > m ? ii - N : ii;
    const int d2 = i1 * i1 + j1 * j1 + l1 * l1;
    if (KIND == FK_TOPHAT) {
        A.real[e] = ((float)d2 <= A.R2) ? 1.0f : 0.0f;
    } else if (KIND == FK_GAUSSIAN) {
        A.real[e] = (float)exp(-(double)d2 / (2.0 * (double)A.R2));
    } else {
        const float k = (float)((double)A.kF * sqrt((double)d2));
        float2 v;
        v.x = (e == 0 || (k >= A.kmin && k < A.kmax)) ? 1.0f : 0.0f;       // the DC mode is always kept
        v.y = 0.0f;
        A.cplx[e] = v;
    }
}

}  // namespace pyl
