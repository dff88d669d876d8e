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
