// Per-axis mass-assignment stencils (device), arithmetic restated from the reference:
//   NGP  library/MAS_library/MAS_library.pyx:290-291
//   CIC  library/MAS_library/MAS_library.pyx:152-157
//   TSC  library/MAS_library/MAS_library.pyx:392-399
//   PCS  library/MAS_library/MAS_library.pyx:485-492
// Every operation that fixes the reference's result is spelled with an explicit-rounding
// intrinsic so nvcc cannot contract it into an FMA: the cell coordinate `dist` is the ROUNDED
// float32 product pos*inv_cell_size (an exact product moves CIC weights
// by up to 3e-2 at dims=1024).
#pragma once
#include <cuda_runtime.h>

namespace pyl {

template <int MAS> struct StencilWidth { static constexpr int value = MAS + 1; };  // 1,2,3,4

// periodic wrap into [0, dims).  Inputs inside the documented domain 0 <= pos <= BoxSize only
// ever need one conditional add/subtract; anything else (the reference would index out of
// bounds there) is folded back with a true modulo so we never write outside the grid.
static __device__ __noinline__ int wrap_index_slow(int i, int dims) {
    i %= dims;
    if (i < 0) i += dims;
    return i;
}

__device__ __forceinline__ int wrap_index(int i, int dims) {
    if (i >= dims) i -= dims;
    else if (i < 0) i += dims;
    // kept out of line: an inlined integer modulo was 55% of the scatter kernel's instructions
    if (__builtin_expect((unsigned)i >= (unsigned)dims, 0)) i = wrap_index_slow(i, dims);
    return i;
}

__device__ __forceinline__ float cell_coordinate(float pos, float inv_cell_size) {
    return __fmul_rn(pos, inv_cell_size);
}

// First (unwrapped) cell of the stencil along one axis.  The reference promotes to float64 before it adds the
// (half-)integer offset; a float32 -> float64 conversion is exact, so its results are functions of floor(dist) and of
// the exact fraction dist - floor(dist), and are formed here without the float64 pipe:
//   NGP  (int)(dist + 0.5)          = floor(dist) + [frac >= 0.5]        (MAS_library.pyx:290; dist >= 0)
//   TSC  (int)floor(dist - 1.5) + 1 = floor(dist) - 1 + [frac >= 0.5]    (:392)
//   PCS  (int)floor(dist - 2.0) + 1 = floor(dist) - 1                    (:485)
template <int MAS>
__device__ __forceinline__ int axis_base(float dist) {
    if (MAS == PYL_MAS_CIC) return __float2int_rz(dist);
    if (MAS == PYL_MAS_NGP && dist < 0.0f)      // outside the documented domain: C truncation, as the reference does
        return __double2int_rz(__dadd_rn((double)dist, 0.5));
    const int fl = __float2int_rd(dist);
    if (MAS == PYL_MAS_PCS) return fl - 1;
    const int up = __fsub_rn(dist, (float)fl) >= 0.5f ? 1 : 0;      // dist - floor(dist) is exact in float32
    return MAS == PYL_MAS_NGP ? fl + up : fl - 1 + up;
}

// The S weights of the cells base, base+1, ... in the reference's own arithmetic (float64 where the reference
// promotes, then rounded to float32).  `base` = axis_base<MAS>(dist).
template <int MAS>
__device__ __forceinline__ void axis_weights(float dist, int base, float *w) {
    if (MAS == PYL_MAS_NGP) {
        w[0] = 1.0f;
    } else if (MAS == PYL_MAS_CIC) {
        const float u = __fsub_rn(dist, (float)base);
        w[0] = __fsub_rn(1.0f, u);
        w[1] = u;
    } else if (MAS == PYL_MAS_TSC) {
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const float diff = fabsf(__fsub_rn((float)(base + j), dist));
            float v;
            if (diff < 0.5f) {
                v = __double2float_rn(__dsub_rn(0.75, (double)__fmul_rn(diff, diff)));
            } else if (diff < 1.5f) {
                const double t = __dsub_rn(1.5, (double)diff);
                v = __double2float_rn(__dmul_rn(__dmul_rn(0.5, t), t));
            } else {
                v = 0.0f;
            }
            w[j] = v;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float diff = fabsf(__fsub_rn((float)(base + j), dist));
            const double x = (double)diff;
            float v;
            if (diff < 1.0f) {
                // (4 - 6x^2 + 3x^3)/6 in double, left to right; the reference is compiled with -ffast-math
                // (setup.py:26-31), which turns the division by 6.0 into a multiplication by its reciprocal
                const double a = __dsub_rn(4.0, __dmul_rn(__dmul_rn(6.0, x), x));
                const double b = __dmul_rn(__dmul_rn(__dmul_rn(3.0, x), x), x);
                v = __double2float_rn(__dmul_rn(__dadd_rn(a, b), 1.0 / 6.0));
            } else if (diff < 2.0f) {
                const double t = __dsub_rn(2.0, x);
                v = __double2float_rn(__dmul_rn(__dmul_rn(__dmul_rn(t, t), t), 1.0 / 6.0));
            } else {
                v = 0.0f;
            }
            w[j] = v;
        }
    }
}

// idx[j] : wrapped cell index along this axis, w[j] : its weight, j < StencilWidth<MAS>
// (one definition of the arithmetic for the atomic, tiled and deterministic deposits)
template <int MAS>
__device__ __forceinline__ void axis_stencil(float dist, int dims, int *idx, float *w) {
    constexpr int S = StencilWidth<MAS>::value;
    const int base = axis_base<MAS>(dist);
    axis_weights<MAS>(dist, base, w);
    idx[0] = wrap_index(base, dims);
#pragma unroll
    for (int j = 1; j < S; j++) {
        if (MAS == PYL_MAS_CIC) idx[j] = (idx[j - 1] + 1 == dims) ? 0 : idx[j - 1] + 1;   // (i_d + 1) % dims
        else idx[j] = wrap_index(base + j, dims);
    }
}

}  // namespace pyl
