/* CPU restatement of the reference's |delta_k|^2 deconvolution + binning loops.
 *
 * TEST INFRASTRUCTURE ONLY (see the header of ma_oracle.c for who may use it).
 *
 * Parity status: PINNED against tests/golden/*.npz (outputs of the compiled, unmodified
 * reference, oracle/_ref) by tests/test_oracle_vs_golden.py, given the same delta_k.
 * The FFT itself (reference: FFTW3 through pyfftw, library/Pk_library/Pk_library.pyx:117-130;
 * un-vendored third-party code, unpinned in setup.py:137) is NOT restated here: the Python
 * side (oracle/pk_oracle.py) feeds this loop with scipy.fft.rfftn of the float32 field, the
 * same transform the pyfftw shim gives the compiled reference.
 *
 * Reference followed (paths relative to /root/reference):
 *   Pk  loop : library/Pk_library/Pk_library.pyx:311-378
 *   XPk loop : library/Pk_library/Pk_library.pyx:623-732
 *   window   : library/Pk_library/Pk_library.pyx:72-84 (MAS_function / MAS_correction)
 *
 * One routine covers both: F fields laid out field-major (F, dims, dims, dims/2+1)
 * complex64; X = F(F-1)/2 pairs in lexicographic order i<j.  `phase` (Pk only) may be NULL.
 * All outputs are RAW sums; units/averaging (Pk_library.pyx:384-418, 735-791) happen in
 * oracle/pk_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline double mas_correction(double x, int mas_index)
{
    /* Pk_library.pyx:83-84 */
    return (x == 0.0) ? 1.0 : pow(x / sin(x), (double)mas_index);
}

/* dk       : (F, dims, dims, dims/2+1) interleaved float (re,im)
 * mas_index: per field, 0..4 (None,NGP,CIC,TSC,PCS)
 * outputs (all double, zero-initialised by the caller, accumulated here):
 *   k3D[kmax+1], Nm3D[kmax+1], Pk3D[(kmax+1)*3*F] (bin,ell,field), PkX3D[(kmax+1)*3*X],
 *   phase[kmax+1] (field 0 only; may be NULL),
 *   k1D[kpar_max+1], Nm1D[..], Pk1D[(kmax_par+1)*F], PkX1D[(kmax_par+1)*X],
 *   Nm2D[n2d], Pk2D[n2d*F], PkX2D[n2d*X],  n2d = (kmax_par+1)*(kmax_per+1)
 */
void oracle_pk_bin(const float *dk, int dims, int F, const int *mas_index, int axis,
                   int kmax_par, int kmax_per, int kmax,
                   double *k3D, double *Nm3D, double *Pk3D, double *PkX3D, double *phase,
                   double *k1D, double *Nm1D, double *Pk1D, double *PkX1D,
                   double *Nm2D, double *Pk2D, double *PkX2D)
{
    const int middle = dims / 2;
    const int nz = middle + 1;
    const int X = F * (F - 1) / 2;
    const double prefact = M_PI / dims;
    const int even = (dims % 2 == 0);
    double *re = (double *)malloc(sizeof(double) * F);
    double *im = (double *)malloc(sizeof(double) * F);
    double *cx = (double *)malloc(sizeof(double) * F);
    double *cy = (double *)malloc(sizeof(double) * F);
    (void)kmax; (void)kmax_per;

    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        for (int f = 0; f < F; f++) cx[f] = mas_correction(prefact * kx, mas_index[f]);

        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            for (int f = 0; f < F; f++) cy[f] = mas_correction(prefact * ky, mas_index[f]);

            for (int kzz = 0; kzz < nz; kzz++) {
                const int kz = kzz; /* kzz <= middle always */

                /* Hermitian duplicates on the kz=0 and kz=middle planes (:324-327) */
                if (kz == 0 || (kz == middle && even)) {
                    if (kx < 0) continue;
                    if (kx == 0 || (kx == middle && even)) {
                        if (ky < 0) continue;
                    }
                }

                const double k = sqrt((double)(kx * kx + ky * ky + kz * kz));
                const int k_index = (int)k;

                int k_par, k_per;
                if (axis == 0) { k_par = kx; k_per = (int)sqrt((double)(ky * ky + kz * kz)); }
                else if (axis == 1) { k_par = ky; k_per = (int)sqrt((double)(kx * kx + kz * kz)); }
                else { k_par = kz; k_per = (int)sqrt((double)(kx * kx + ky * ky)); }

                const double mu = (k == 0.0) ? 0.0 : (double)k_par / k;
                const double mu2 = mu * mu;
                const double val1 = (3.0 * mu2 - 1.0) / 2.0;
                const double val2 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
                if (k_par < 0) k_par = -k_par;

                const int in1d = (k <= (double)middle);
                const int64_t i2 = (int64_t)(kmax_par + 1) * k_per + k_par;

                if (in1d) { k1D[k_par] += (double)k_par; Nm1D[k_par] += 1.0; }
                Nm2D[i2] += 1.0;
                k3D[k_index] += k;
                Nm3D[k_index] += 1.0;

                for (int f = 0; f < F; f++) {
                    const double cz = mas_correction(prefact * kz, mas_index[f]);
                    /* product in double, rounded to float32 (`cdef float MAS_factor`, :272,351) */
                    const float fac = (float)(cx[f] * cy[f] * cz);
                    const int64_t off =
                        ((((int64_t)f * dims + kxx) * dims + kyy) * nz + kzz) * 2;
                    /* complex64 * float32 in single precision (:352), then promoted (:355-357) */
                    volatile float r32 = dk[off] * fac;
                    volatile float i32 = dk[off + 1] * fac;
                    re[f] = (double)r32;
                    im[f] = (double)i32;
                    const double d2 = re[f] * re[f] + im[f] * im[f];
                    if (in1d) Pk1D[(int64_t)k_par * F + f] += d2;
                    Pk2D[i2 * F + f] += d2;
                    Pk3D[((int64_t)k_index * 3 + 0) * F + f] += d2;
                    Pk3D[((int64_t)k_index * 3 + 1) * F + f] += d2 * val1;
                    Pk3D[((int64_t)k_index * 3 + 2) * F + f] += d2 * val2;
                    if (f == 0 && phase) {
                        const double ph = atan2(re[0], sqrt(d2)); /* :358 */
                        phase[k_index] += ph * ph;
                    }
                }

                int ix = 0;
                for (int i = 0; i < F; i++)
                    for (int j = i + 1; j < F; j++) {
                        const double dx = re[i] * re[j] + im[i] * im[j]; /* :716-717 */
                        if (in1d) PkX1D[(int64_t)k_par * X + ix] += dx;
                        PkX2D[i2 * X + ix] += dx;
                        PkX3D[((int64_t)k_index * 3 + 0) * X + ix] += dx;
                        PkX3D[((int64_t)k_index * 3 + 1) * X + ix] += dx * val1;
                        PkX3D[((int64_t)k_index * 3 + 2) * X + ix] += dx * val2;
                        ix++;
                    }
            }
        }
    }
    free(re); free(im); free(cx); free(cy);
}

/* =====================================================================================
 * Sibling estimators of Pk_library.pyx.  Same rules as
 * above: RAW sums only, every rounding the reference performs is spelled out, the FFTs are
 * done by the Python side (oracle/cpu.py) with pocketfft.
 * ===================================================================================== */

/* XPk_imag loop, Pk_library.pyx:905-1016: identical to the XPk loop except for the cross term
 * (:1000-1001)  imag_i*real_j - real_i*imag_j.  Implemented by re-running the XPk loop above
 * and replacing the three cross accumulators. */
void oracle_pk_bin_imag(const float *dk, int dims, int F, const int *mas_index, int axis,
                        int kmax_par, int kmax_per, int kmax,
                        double *k3D, double *Nm3D, double *Pk3D, double *PkX3D,
                        double *k1D, double *Nm1D, double *Pk1D, double *PkX1D,
                        double *Nm2D, double *Pk2D, double *PkX2D)
{
    const int middle = dims / 2, nz = middle + 1, X = F * (F - 1) / 2;
    const int even = (dims % 2 == 0);
    const double prefact = M_PI / dims;
    const int64_t n2 = (int64_t)(kmax_par + 1) * (kmax_per + 1);
    oracle_pk_bin(dk, dims, F, mas_index, axis, kmax_par, kmax_per, kmax, k3D, Nm3D, Pk3D, PkX3D,
                  NULL, k1D, Nm1D, Pk1D, PkX1D, Nm2D, Pk2D, PkX2D);
    for (int64_t i = 0; i < (int64_t)(kmax + 1) * 3 * X; i++) PkX3D[i] = 0.0;
    for (int64_t i = 0; i < (int64_t)(kmax_par + 1) * X; i++) PkX1D[i] = 0.0;
    for (int64_t i = 0; i < n2 * X; i++) PkX2D[i] = 0.0;
    double *re = (double *)malloc(sizeof(double) * F);
    double *im = (double *)malloc(sizeof(double) * F);
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            for (int kz = 0; kz < nz; kz++) {
                if (kz == 0 || (kz == middle && even)) {
                    if (kx < 0) continue;
                    if ((kx == 0 || (kx == middle && even)) && ky < 0) continue;
                }
                const double k = sqrt((double)(kx * kx + ky * ky + kz * kz));
                const int k_index = (int)k;
                int k_par, k_per;
                if (axis == 0) { k_par = kx; k_per = (int)sqrt((double)(ky * ky + kz * kz)); }
                else if (axis == 1) { k_par = ky; k_per = (int)sqrt((double)(kx * kx + kz * kz)); }
                else { k_par = kz; k_per = (int)sqrt((double)(kx * kx + ky * ky)); }
                const double mu = (k == 0.0) ? 0.0 : (double)k_par / k;
                const double mu2 = mu * mu;
                const double val1 = (3.0 * mu2 - 1.0) / 2.0;
                const double val2 = (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0;
                if (k_par < 0) k_par = -k_par;
                const int in1d = (k <= (double)middle);
                const int64_t i2 = (int64_t)(kmax_par + 1) * k_per + k_par;
                for (int f = 0; f < F; f++) {
                    const float fac = (float)(mas_correction(prefact * kx, mas_index[f]) *
                                              mas_correction(prefact * ky, mas_index[f]) *
                                              mas_correction(prefact * kz, mas_index[f]));
                    const int64_t off = ((((int64_t)f * dims + kxx) * dims + kyy) * nz + kz) * 2;
                    volatile float r32 = dk[off] * fac;
                    volatile float i32 = dk[off + 1] * fac;
                    re[f] = (double)r32; im[f] = (double)i32;
                }
                int ix = 0;
                for (int i = 0; i < F; i++)
                    for (int j = i + 1; j < F; j++) {
                        const double dx = im[i] * re[j] - re[i] * im[j];   /* :1000-1001 */
                        if (in1d) PkX1D[(int64_t)k_par * X + ix] += dx;
                        PkX2D[i2 * X + ix] += dx;
                        PkX3D[((int64_t)k_index * 3 + 0) * X + ix] += dx;
                        PkX3D[((int64_t)k_index * 3 + 1) * X + ix] += dx * val1;
                        PkX3D[((int64_t)k_index * 3 + 2) * X + ix] += dx * val2;
                        ix++;
                    }
            }
        }
    }
    free(re); free(im);
}

/* Pk_plane loop (Pk_library.pyx:470-499) and XPk_plane loop (:1151-1200) on (grid, grid/2+1)
 * complex64 images.  F = 1 or 2 images, field-major; PkX may be NULL for F = 1.
 * k2D[kmax+1], Nm[kmax+1], Pk2D[(kmax+1)*F], PkX[kmax+1]. */
void oracle_plane_bin(const float *dk, int grid, int F, const int *mas_index,
                      double *k2D, double *Nm, double *Pk2D, double *PkX)
{
    const int middle = grid / 2, ny = middle + 1, even = (grid % 2 == 0);
    const double prefact = M_PI / grid;
    double re[2], im[2];
    for (int kxx = 0; kxx < grid; kxx++) {
        const int kx = (kxx > middle) ? kxx - grid : kxx;
        for (int ky = 0; ky < ny; ky++) {
            if (ky == 0 || (ky == middle && even)) {
                if (kx < 0) continue;                                  /* :480-481 */
            }
            const double k = sqrt((double)(kx * kx + ky * ky));
            const int k_index = (int)k;
            k2D[k_index] += k;
            Nm[k_index] += 1.0;
            for (int f = 0; f < F; f++) {
                const float fac = (float)(mas_correction(prefact * kx, mas_index[f]) *
                                          mas_correction(prefact * ky, mas_index[f]));   /* :488 */
                const int64_t off = (((int64_t)f * grid + kxx) * ny + ky) * 2;
                volatile float r32 = dk[off] * fac;
                volatile float i32 = dk[off + 1] * fac;
                re[f] = (double)r32; im[f] = (double)i32;
                Pk2D[(int64_t)k_index * F + f] += re[f] * re[f] + im[f] * im[f];
            }
            if (F == 2 && PkX) PkX[k_index] += re[0] * re[1] + im[0] * im[1];            /* :1196-1200 */
        }
    }
}

/* Shell sums of the velocity-divergence estimators.
 *   kind 0: Pk_theta (Pk_library.pyx:1273-1316)  fields Vx,Vy,Vz           -> P1 = |theta|^2
 *   kind 1: XPk_dv   (:1386-1432)                fields delta,Vx,Vy,Vz     -> P1,P2,PX
 *   kind 2: XPk_vv   (:1515-1568)                fields d1,Vx1,Vy1,Vz1,d2,Vx2,Vy2,Vz2 -> P1,P2,PX
 * dk: field-major (nf, dims, dims, dims/2+1) complex64.  One MAS for all fields.
 * The reference forms k.V with C `int * float` products, i.e. in float32 (:1303-1309). */
static void kdotv(const float *dk, int64_t stride, int64_t off, int f0, float fac, int kx, int ky, int kz,
                  float *dot_re, float *dot_im)
{
    float re[3], im[3];
    for (int c = 0; c < 3; c++) {
        const float *p = dk + ((int64_t)(f0 + c) * stride + off) * 2;
        re[c] = p[0] * fac; im[c] = p[1] * fac;
    }
    *dot_re = kx * re[0] + ky * re[1] + kz * re[2];
    *dot_im = kx * im[0] + ky * im[1] + kz * im[2];
}

void oracle_vel_bin(int kind, const float *dk, int dims, int mas_index,
                    double *k3D, double *Nm, double *P1, double *P2, double *PX)
{
    const int middle = dims / 2, nz = middle + 1, even = (dims % 2 == 0);
    const double prefact = M_PI / dims;
    const int64_t stride = (int64_t)dims * dims * nz;
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        const double cx = mas_correction(prefact * kx, mas_index);
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            const double cy = mas_correction(prefact * ky, mas_index);
            for (int kz = 0; kz < nz; kz++) {
                if (kz == 0 || (kz == middle && even)) {
                    if (kx < 0) continue;
                    if ((kx == 0 || (kx == middle && even)) && ky < 0) continue;
                }
                const double kmod = sqrt((double)(kx * kx + ky * ky + kz * kz));
                const int k_index = (int)kmod;
                const float fac = (float)(cx * cy * mas_correction(prefact * kz, mas_index));
                const int64_t off = ((int64_t)kxx * dims + kyy) * nz + kz;
                k3D[k_index] += kmod;
                Nm[k_index] += 1.0;
                if (kind == 0) {
                    float dre, dim_;
                    kdotv(dk, stride, off, 0, fac, kx, ky, kz, &dre, &dim_);
                    const double real = -(double)dim_, imag = (double)dre;        /* :1303-1309 */
                    P1[k_index] += real * real + imag * imag;
                } else {
                    double r1, i1, r2, i2;
                    float dre, dim_;
                    if (kind == 1) {
                        const float a = dk[off * 2] * fac, b = dk[off * 2 + 1] * fac; /* :1410,1416-1417 */
                        r1 = (double)a; i1 = (double)b;
                        kdotv(dk, stride, off, 1, fac, kx, ky, kz, &dre, &dim_);
                        r2 = (double)dim_; i2 = -(double)dre;                     /* :1419-1425 */
                    } else {
                        kdotv(dk, stride, off, 1, fac, kx, ky, kz, &dre, &dim_);
                        r1 = (double)dim_; i1 = -(double)dre;                     /* :1549-1554 */
                        kdotv(dk, stride, off, 5, fac, kx, ky, kz, &dre, &dim_);
                        r2 = (double)dim_; i2 = -(double)dre;                     /* :1556-1561 */
                    }
                    P1[k_index] += r1 * r1 + i1 * i1;
                    P2[k_index] += r2 * r2 + i2 * i2;
                    PX[k_index] += r1 * r2 + i1 * i2;
                }
            }
        }
    }
}

/* expected_Pk mode loop, Pk_library.pyx:2004-2037.  tab_k/tab_P: the log-spaced float32 table
 * built at :1983-1993 (`bins` entries).  `k` is a C float in the reference. */
void oracle_expected_pk(int dims, float kF, const float *tab_k, const float *tab_P, float kmin_in,
                        float deltak, double *k3D, double *Pk3D, double *Nm)
{
    const int middle = dims / 2, even = (dims % 2 == 0);
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            for (int kz = 0; kz <= middle; kz++) {
                if (kz == 0 || (kz == middle && even)) {
                    if (kx < 0) continue;
                    if ((kx == 0 || (kx == middle && even)) && ky < 0) continue;
                }
                float k = (float)sqrt((double)(kx * kx + ky * ky + kz * kz));
                const int k_index = (int)k;
                if (k == 0.0f) continue;
                k = k * kF;
                const int i = (int)((log10((double)k) - log10((double)kmin_in)) / (double)deltak);
                const float Pi = (tab_P[i + 1] - tab_P[i]) / (tab_k[i + 1] - tab_k[i]) * (k - tab_k[i]) + tab_P[i];
                k3D[k_index] += (double)k;
                Pk3D[k_index] += (double)Pi;
                Nm[k_index] += 1.0;
            }
        }
    }
}

/* correct_MAS mode loop, Pk_library.pyx:1909-1929: multiply the INDEPENDENT modes of the
 * half-spectrum by the float32 window factor, in place; Hermitian duplicates on the kz = 0 /
 * Nyquist planes are skipped, i.e. left uncorrected, exactly as the reference does. */
void oracle_correct_mas_modes(float *dk, int dims, int mas_index)
{
    const int middle = dims / 2, nz = middle + 1, even = (dims % 2 == 0);
    const double prefact = M_PI / dims;
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        const double cx = mas_correction(prefact * kx, mas_index);
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            const double cy = mas_correction(prefact * ky, mas_index);
            for (int kz = 0; kz < nz; kz++) {
                if (kz == 0 || (kz == middle && even)) {
                    if (kx < 0) continue;
                    if ((kx == 0 || (kx == middle && even)) && ky < 0) continue;
                }
                const float fac = (float)(cx * cy * mas_correction(prefact * kz, mas_index));
                const int64_t off = (((int64_t)kxx * dims + kyy) * nz + kz) * 2;
                dk[off] *= fac; dk[off + 1] *= fac;
            }
        }
    }
}

/* Xi / XXi mode loop, Pk_library.pyx:2198-2218 and :2335-2362: EVERY stored mode (no skip rule)
 * becomes (re1*re2 + im1*im2, 0) of the deconvolved fields, in float32 (`cdef float real, imag`).
 * d2k == NULL: auto-correlation (d2k = d1k, mas2 = mas1).  Result overwrites d1k. */
void oracle_xi_modes(float *d1k, const float *d2k, int dims, int mas1, int mas2)
{
    const int middle = dims / 2, nz = middle + 1;
    const double prefact = M_PI / dims;
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            for (int kz = 0; kz < nz; kz++) {
                const float f1 = (float)(mas_correction(prefact * kx, mas1) * mas_correction(prefact * ky, mas1) *
                                         mas_correction(prefact * kz, mas1));
                const int64_t off = (((int64_t)kxx * dims + kyy) * nz + kz) * 2;
                const float r1 = d1k[off] * f1, i1 = d1k[off + 1] * f1;
                float r2 = r1, i2 = i1;
                if (d2k) {
                    const float f2 = (float)(mas_correction(prefact * kx, mas2) * mas_correction(prefact * ky, mas2) *
                                             mas_correction(prefact * kz, mas2));
                    r2 = d2k[off] * f2; i2 = d2k[off + 1] * f2;
                }
                d1k[off] = r1 * r2 + i1 * i2;
                d1k[off + 1] = 0.0f;
            }
        }
    }
}

/* Xi / XXi real-space binning, Pk_library.pyx:2233-2267 (= :2378-2412): every cell of the
 * (dims,dims,dims) float32 correlation grid, radial bins of one cell, l = 0,2,4 weights. */
void oracle_xi_bin(const float *xi, int dims, int axis, double *r3D, double *xi3D /* [kmax+1][3] */, double *Nm)
{
    const int middle = dims / 2;
    for (int kxx = 0; kxx < dims; kxx++) {
        const int kx = (kxx > middle) ? kxx - dims : kxx;
        for (int kyy = 0; kyy < dims; kyy++) {
            const int ky = (kyy > middle) ? kyy - dims : kyy;
            for (int kzz = 0; kzz < dims; kzz++) {
                const int kz = (kzz > middle) ? kzz - dims : kzz;
                const double k = sqrt((double)(kx * kx + ky * ky + kz * kz));
                const int k_index = (int)k;
                const int k_par = (axis == 0) ? kx : (axis == 1 ? ky : kz);
                const double mu = (k == 0.0) ? 0.0 : (double)k_par / k;
                const double mu2 = mu * mu;
                const double v = (double)xi[((int64_t)kxx * dims + kyy) * dims + kzz];
                r3D[k_index] += k;
                xi3D[k_index * 3 + 0] += v;
                xi3D[k_index * 3 + 1] += (v * (3.0 * mu2 - 1.0) / 2.0);
                xi3D[k_index * 3 + 2] += (v * (35.0 * mu2 * mu2 - 30.0 * mu2 + 3.0) / 8.0);
                Nm[k_index] += 1.0;
            }
        }
    }
}
