/* CPU restatement of the reference's particle->grid mass assignment.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pylians3_b200/ may call, link or import this
 * file; it exists to check the CUDA path (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_golden.py checks every function here
 * against tests/golden/*.npz, which were produced by the compiled, unmodified reference
 * (oracle/_ref, see tests/golden/make_golden.py), and tests/test_oracle_vs_ref.py compares
 * it live against oracle/_ref when that is present.
 *
 * Reference followed (paths relative to /root/reference):
 *   NGP / NGPW : library/MAS_library/MAS_library.pyx:273-292 / 305-325
 *   CIC / CICW : library/MAS_library/MAS_library.pyx:123-166 / 179-216
 *   TSC / TSCW : library/MAS_library/MAS_library.pyx:369-404 / 417-452
 *   PCS / PCSW : library/MAS_library/MAS_library.pyx:463-497 / 510-545
 *   2D mode    : library/MAS_library/MAS_library.pyx:84-110 (third axis pinned to cell 0
 *                with unit weight, so each particle lands 1/2/3/4 times in the plane)
 *   CIC_interp : library/MAS_library/MAS_library.pyx:558-599 (grid -> particle gather)
 *   pos_redshift_space : library/redshift_space_library/redshift_space_library.pyx:29-46
 *
 * The serial particle order of the reference is kept, so float32 accumulation order is the
 * same as the reference's.
 */
#include <math.h>
#include <stdint.h>

enum { ORACLE_NGP = 0, ORACLE_CIC = 1, ORACLE_TSC = 2, ORACLE_PCS = 3 };

/* Per-axis stencil of one particle: `n` cells with their weights.
 * dist is the ROUNDED float32 product pos*inv_cell_size (MAS_library.pyx:152,290,392,485). */
static inline int axis_stencil(int mas, float dist, int dims, int idx[4], float w[4])
{
    switch (mas) {
    case ORACLE_NGP: {
        /* :290-291  index = <int>(pos*inv + 0.5) ; (index+dims)%dims  -- 0.5 is a double */
        int i = (int)((double)dist + 0.5);
        idx[0] = (i + dims) % dims;
        w[0] = 1.0f;
        return 1;
    }
    case ORACLE_CIC: {
        /* :152-157 */
        int i = (int)dist;
        float u = dist - (float)i;
        float d = (float)(1.0 - (double)u);
        idx[0] = i % dims;
        idx[1] = (idx[0] + 1) % dims;
        w[0] = d;
        w[1] = u;
        return 2;
    }
    case ORACLE_TSC: {
        /* :392-399 */
        int minimum = (int)floor((double)dist - 1.5);
        for (int j = 0; j < 3; j++) {
            idx[j] = (minimum + j + 1 + dims) % dims;
            float diff = (float)fabs((double)((float)(minimum + j + 1) - dist));
            if (diff < 0.5)
                w[j] = (float)(0.75 - (double)(diff * diff));
            else if (diff < 1.5)
                w[j] = (float)(0.5 * (1.5 - (double)diff) * (1.5 - (double)diff));
            else
                w[j] = 0.0f;
        }
        return 3;
    }
    default: {
        /* PCS :485-492, polynomial evaluated in double, stored as float */
        int minimum = (int)floor((double)dist - 2.0);
        for (int j = 0; j < 4; j++) {
            idx[j] = (minimum + j + 1 + dims) % dims;
            float diff = (float)fabs((double)((float)(minimum + j + 1) - dist));
            double x = (double)diff;
            if (diff < 1.0)
                w[j] = (float)((4.0 - 6.0 * x * x + 3.0 * x * x * x) / 6.0);
            else if (diff < 2.0)
                w[j] = (float)((2.0 - x) * (2.0 - x) * (2.0 - x) / 6.0);
            else
                w[j] = 0.0f;
        }
        return 4;
    }
    }
}

/* number: float32 grid, dims^axes cells, C order, ACCUMULATED in place (never zeroed).
 * W may be NULL.  axes = 2 or 3.  For axes == 2 the result is the reference's
 * un-renormalised plane (each particle counted n times along the pinned third axis);
 * the /2,/3,/4 of MAS_library.pyx:90-107 is applied by the Python caller. */
void oracle_ma(int mas, const float *pos, float *number, const float *W, long particles,
               int dims, int axes, float BoxSize)
{
    const float inv_cell_size = (float)dims / BoxSize; /* :135 -- float32 division */
    int idx[3][4];
    float w[3][4];
    int n = 1;

    /* 2D: third axis pinned to cell 0 with unit weight (:138-139, :384-386, :477-479) */
    for (int a = 0; a < 3; a++)
        for (int j = 0; j < 4; j++) { idx[a][j] = 0; w[a][j] = 1.0f; }

    const int64_t s0 = (axes == 3) ? (int64_t)dims * dims : dims;
    const int64_t s1 = (axes == 3) ? dims : 1;
    const int64_t s2 = (axes == 3) ? 1 : 0;

    for (long p = 0; p < particles; p++) {
        for (int a = 0; a < axes; a++) {
            volatile float dist = pos[(int64_t)p * axes + a] * inv_cell_size; /* rounded f32 */
            n = axis_stencil(mas, dist, dims, idx[a], w[a]);
        }
        if (W == 0) {
            for (int l = 0; l < n; l++)
                for (int m = 0; m < n; m++)
                    for (int q = 0; q < n; q++)
                        number[idx[0][l] * s0 + idx[1][m] * s1 + idx[2][q] * s2] +=
                            w[0][l] * w[1][m] * w[2][q];
        } else {
            const float wp = W[p];
            for (int l = 0; l < n; l++)
                for (int m = 0; m < n; m++)
                    for (int q = 0; q < n; q++)
                        number[idx[0][l] * s0 + idx[1][m] * s1 + idx[2][q] * s2] +=
                            w[0][l] * w[1][m] * w[2][q] * wp;
        }
    }
}

/* Grid -> particle CIC interpolation (gather), library/MAS_library/MAS_library.pyx:558-599:
 * den[i] = sum over the 8 cells around pos[i] of density[cell]*weight, weights and indices as in
 * CIC (:585-590); each term is ((density*wx)*wy)*wz in float32, summed left to right (:592-599).
 * den is OVERWRITTEN. */
void oracle_cic_interp(const float *density, int dims, float BoxSize, const float *pos, long particles,
                       float *den)
{
    const float inv_cell_size = (float)dims / BoxSize;
    for (long p = 0; p < particles; p++) {
        int id[3], iu[3];
        float u[3], d[3];
        for (int a = 0; a < 3; a++) {
            volatile float dist = pos[(int64_t)p * 3 + a] * inv_cell_size;
            u[a] = dist - (float)(int)dist;
            d[a] = (float)(1.0 - (double)u[a]);
            id[a] = ((int)dist) % dims;
            iu[a] = (id[a] + 1) % dims;
        }
#define DEN(i, j, k) density[((int64_t)(i) * dims + (j)) * dims + (k)]
        den[p] = DEN(id[0], id[1], id[2]) * d[0] * d[1] * d[2] + DEN(id[0], id[1], iu[2]) * d[0] * d[1] * u[2] +
                 DEN(id[0], iu[1], id[2]) * d[0] * u[1] * d[2] + DEN(id[0], iu[1], iu[2]) * d[0] * u[1] * u[2] +
                 DEN(iu[0], id[1], id[2]) * u[0] * d[1] * d[2] + DEN(iu[0], id[1], iu[2]) * u[0] * d[1] * u[2] +
                 DEN(iu[0], iu[1], id[2]) * u[0] * u[1] * d[2] + DEN(iu[0], iu[1], iu[2]) * u[0] * u[1] * u[2];
#undef DEN
    }
}

/* Real -> redshift space along one axis, library/redshift_space_library/redshift_space_library.pyx:29-46:
 * factor = (float)((1.0 + z)/H) ; pos = pos + vel*factor (ONE rounding: the reference binary, built with
 * -O3 -ffast-math for an FMA-capable target, contracts it into a fused multiply-add -- checked against the
 * compiled reference) ; while pos < 0: pos += BoxSize ; if pos > BoxSize: pos = fmodf(pos, BoxSize). */
void oracle_pos_redshift_space(float *pos, const float *vel, long particles, float BoxSize, float Hubble,
                               float redshift, int axis)
{
    const float factor = (float)((1.0 + (double)redshift) / (double)Hubble);
    for (long i = 0; i < particles; i++) {
        float p = fmaf(vel[i * 3 + axis], factor, pos[i * 3 + axis]);
        while (p < 0.0f) p += BoxSize;
        if (p > BoxSize) p = fmodf(p, BoxSize);
        pos[i * 3 + axis] = p;
    }
}
