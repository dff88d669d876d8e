// TEST INFRASTRUCTURE ONLY: runs the per-thread bodies of pylians3_b200/csrc/shell_body.cuh serially on the CPU.
//
// The authoring container has no GPU.  The CUDA kernels of pk_shell.cu are thin wrappers around
// `shell_thread<KIND>(args, t, seg, sink)` and `mode_element<OP>(args, e)`; this file calls the very same
// functions for every (t, seg) / e of the launch geometry the device launcher would use, with a sink that adds
// into plain memory.  tests/test_shell_bodies_host.py compares the result with the oracle: thread/segment
// decomposition, duplicate-mode rule, index arithmetic and per-mode functionals are thereby checked without a
// GPU.  Nothing in the product links or calls this.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "shell_body.cuh"

using namespace pyl;

struct HostSink {
    unsigned long long *base;
    void add(long long w, double v) {
        double x;
        memcpy(&x, base + w, 8);
        x += v;
        memcpy(base + w, &x, 8);
    }
    void count(long long w, unsigned long long c) { base[w] += c; }
};

template <int KIND>
static void run_kind(const ShellArgs &A, HostSink &sink) {
    for (int seg = 0; seg < A.nseg; seg++)
        for (long long t = 0; t < A.T; t++) shell_thread<KIND>(A, t, seg, sink);
}

extern "C" {

int harness_shell_bins(int kind, int dims) {
    const double m = (double)(dims / 2);
    const bool plane = (kind == SK_PLANE || kind == SK_XPLANE);
    return (int)sqrt(plane ? 2.0 * m * m : 3.0 * m * m) + 1;
}

int harness_shell(int kind, const void *const *fields, const int *mas_index, int dims, int axis, float scale,
                  const float *tab_k, const float *tab_P, int tab_n, float kF, double log10_kmin, double deltak,
                  int sms, unsigned long long *out) {
    const int m = dims / 2, m1 = m + 1;
    std::vector<double> tab(2 * m1);
    const int p0 = mas_index ? mas_index[0] : 0;
    const int p1 = (kind == SK_XPLANE) ? mas_index[1] : p0;
    for (int i = 0; i < m1; i++) { tab[i] = shell_window(i, dims, p0); tab[m1 + i] = shell_window(i, dims, p1); }
    ShellArgs A;
    memset(&A, 0, sizeof(A));
    for (int f = 0; f < shell_nfields(kind); f++) A.f[f] = fields[f];
    A.win[0] = tab.data(); A.win[1] = tab.data() + m1;
    A.N = dims; A.m = m; A.even = (dims % 2 == 0);
    const bool plane = (kind == SK_PLANE || kind == SK_XPLANE);
    A.nx = plane ? 1 : dims;
    A.nzs = (kind == SK_XI) ? dims : m1;
    A.hermitian = (kind == SK_XI) ? 0 : 1;
    A.axis = axis;
    A.n3 = harness_shell_bins(kind, dims);
    A.scale = scale;
    A.tab_k = tab_k; A.tab_P = tab_P; A.tab_n = tab_n; A.kF = kF; A.log10_kmin = log10_kmin; A.deltak = deltak;
    shell_geometry(A, sms);
    memset(out, 0, (size_t)(2 + shell_nvals(kind)) * A.n3 * 8);
    HostSink sink{out};
    switch (kind) {
        case SK_THETA: run_kind<SK_THETA>(A, sink); break;
        case SK_DV: run_kind<SK_DV>(A, sink); break;
        case SK_VV: run_kind<SK_VV>(A, sink); break;
        case SK_EXPECTED: run_kind<SK_EXPECTED>(A, sink); break;
        case SK_PLANE: run_kind<SK_PLANE>(A, sink); break;
        case SK_XPLANE: run_kind<SK_XPLANE>(A, sink); break;
        case SK_XI: run_kind<SK_XI>(A, sink); break;
        default: return -1;
    }
    return A.nseg;
}

int harness_modes(int op, float *a, const float *b, int dims, int mas_a, int mas_b) {
    const int m1 = dims / 2 + 1;
    std::vector<double> tab(2 * m1);
    for (int i = 0; i < m1; i++) { tab[i] = shell_window(i, dims, mas_a); tab[m1 + i] = shell_window(i, dims, mas_b); }
    ModeArgs A;
    A.a = reinterpret_cast<float2 *>(a);
    A.b = reinterpret_cast<const float2 *>(b);
    A.win[0] = tab.data(); A.win[1] = tab.data() + m1;
    A.N = dims; A.m = dims / 2; A.even = (dims % 2 == 0);
    A.total = (long long)dims * dims * m1;
    for (long long e = 0; e < A.total; e++) {
        if (op == MO_DECONVOLVE) mode_element<MO_DECONVOLVE>(A, e);
        else mode_element<MO_POWER>(A, e);
    }
    return 0;
}

int harness_filter(int kind, void *out, int dims, int axes, float R2, float kF, float kmin, float kmax) {
    FilterArgs A;
    A.real = reinterpret_cast<float *>(out);
    A.cplx = reinterpret_cast<float2 *>(out);
    A.N = dims; A.m = dims / 2; A.axes = axes;
    A.R2 = R2; A.kF = kF; A.kmin = kmin; A.kmax = kmax;
    const long long last = (kind == FK_TOPHAT_K) ? dims / 2 + 1 : dims;
    A.total = (axes == 3 ? (long long)dims * dims : (long long)dims) * last;
    for (long long e = 0; e < A.total; e++) {
        if (kind == FK_TOPHAT) filter_element<FK_TOPHAT>(A, e);
        else if (kind == FK_GAUSSIAN) filter_element<FK_GAUSSIAN>(A, e);
        else filter_element<FK_TOPHAT_K>(A, e);
    }
    return 0;
}

}  // extern "C"
