// Grid -> particle CIC interpolation (gather), the transpose of the CIC deposit.
//
// Replaces the reference's serial loop library/MAS_library/MAS_library.pyx:558-599 (CIC_interp), used for
// marked power spectra (docs/source/Pk.rst:205-220).  One thread per particle: indices and weights exactly as
// the CIC stencil (stencil.cuh), each of the 8 terms is ((density*wx)*wy)*wz in float32 and the terms are
// summed left to right like the reference's expression (:592-599).  Bound by the 8 scattered 4-byte reads per
// particle (4 sectors of 32 bytes: the z-neighbours share a sector); positions are read once, den written once.
#include "common.cuh"
#include "stencil.cuh"

namespace pyl {

__global__ void __launch_bounds__(256) cic_interp_kernel(const float *__restrict__ density, int dims,
                                                         float inv_cell_size, const float *__restrict__ pos,
                                                         int64_t particles, float *__restrict__ den) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= particles) return;
    int idx[3][2];
    float w[3][2];
#pragma unroll
    for (int a = 0; a < 3; a++)
        axis_stencil<PYL_MAS_CIC>(cell_coordinate(__ldg(pos + i * 3 + a), inv_cell_size), dims, idx[a], w[a]);
    float sum = 0.0f;
#pragma unroll
    for (int l = 0; l < 2; l++) {
#pragma unroll
        for (int m = 0; m < 2; m++) {
            const float *row = density + ((int64_t)idx[0][l] * dims + idx[1][m]) * dims;
#pragma unroll
            for (int n = 0; n < 2; n++) {
                const float term = __fmul_rn(__fmul_rn(__fmul_rn(__ldg(row + idx[2][n]), w[0][l]), w[1][m]), w[2][n]);
                sum = (l | m | n) ? __fadd_rn(sum, term) : term;
            }
        }
    }
    den[i] = sum;
}

}  // namespace pyl

using namespace pyl;

extern "C" int pyl_cic_interp(const float *density, int dims, float BoxSize, const float *pos, int64_t particles,
                              float *den, pyl_stream_t stream) {
    PYL_REQUIRE(dims > 0 && BoxSize > 0.0f && particles >= 0, "pyl_cic_interp: bad sizes");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(density != nullptr && pos != nullptr && den != nullptr, "pyl_cic_interp: NULL pointer");
    const float inv = (float)dims / BoxSize;      // float32 division, MAS_library.pyx:572
    cic_interp_kernel<<<(unsigned)((particles + 255) / 256), 256, 0, as_stream(stream)>>>(density, dims, inv, pos,
                                                                                        particles, den);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}

// ---- real -> redshift space along one axis ------------------------------------------------------------------
// Replaces library/redshift_space_library/redshift_space_library.pyx:29-46 (the step before MA in
// Pk_library/Pk_snapshot.py:60-64 and MAS_gadget.py): pos[:,axis] += vel[:,axis]*(1+z)/H, wrapped into the box,
// in place.  The reference binary fuses the multiply-add (one rounding), hence fmaf.
namespace pyl {
__global__ void __launch_bounds__(256) redshift_space_kernel(float *__restrict__ pos, const float *__restrict__ vel,
                                                             int64_t particles, float BoxSize, float factor,
                                                             int axis) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= particles) return;
    float p = fmaf(__ldg(vel + i * 3 + axis), factor, pos[i * 3 + axis]);
    if (isfinite(p)) {                       // (the reference would loop forever on -inf; leave non-finite as is)
        while (p < 0.0f) p = __fadd_rn(p, BoxSize);
        if (p > BoxSize) p = fmodf(p, BoxSize);
    }
    pos[i * 3 + axis] = p;
}
}  // namespace pyl

extern "C" int pyl_pos_redshift_space(float *pos, const float *vel, int64_t particles, float BoxSize, float Hubble,
                                      float redshift, int axis, pyl_stream_t stream) {
    PYL_REQUIRE(particles >= 0 && BoxSize > 0.0f && axis >= 0 && axis <= 2, "pyl_pos_redshift_space: bad arguments");
    if (particles == 0) return PYL_OK;
    PYL_REQUIRE(pos != nullptr && vel != nullptr, "pyl_pos_redshift_space: NULL pointer");
    const float factor = (float)((1.0 + (double)redshift) / (double)Hubble);   // :37, double expression -> float
    redshift_space_kernel<<<(unsigned)((particles + 255) / 256), 256, 0, as_stream(stream)>>>(pos, vel, particles,
                                                                                            BoxSize, factor, axis);
    PYL_LAUNCH_CHECK();
    return PYL_OK;
}
