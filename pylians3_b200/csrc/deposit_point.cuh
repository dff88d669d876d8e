// One particle -> S^axes red.global.add.f32 (shared by the atomic kernel and by the overflow path of
// the tiled kernel).  Arithmetic: see stencil.cuh.
#pragma once
#include "common.cuh"
#include "stencil.cuh"

namespace pyl {

struct SlabWindow {
    int x_origin;   // global plane stored at local plane 0
    int x_planes;   // number of local planes (== dims for the whole grid)
    float plane_mult = 0.0f;   // 2D only: adds per plane cell; 0 -> S (MA's Cython path), 1 -> MAS_c.c
};

// `dist` = cell coordinates fl32(pos*inv_cell_size) of one particle
template <int MAS, int AXES, bool WEIGHTED, bool SLAB>
__device__ __forceinline__ void deposit_dist(const float *dist, float wp, float *__restrict__ number,
                                             int dims, SlabWindow win, unsigned long long &dropped) {
    constexpr int S = StencilWidth<MAS>::value;
    int idx[3][S];
    float w[3][S];
#pragma unroll
    for (int a = 0; a < AXES; a++) axis_stencil<MAS>(dist[a], dims, idx[a], w[a]);

    if (AXES == 3) {
#pragma unroll
        for (int l = 0; l < S; l++) {
            int plane = idx[0][l];
            if (SLAB) {
                plane -= win.x_origin;
                if (plane < 0) plane += dims;
                if (plane >= win.x_planes) { dropped += 1; continue; }
            }
            const int64_t base_x = (int64_t)plane * dims;
#pragma unroll
            for (int m = 0; m < S; m++) {
                const float wxy = __fmul_rn(w[0][l], w[1][m]);
                float *row = number + (base_x + idx[1][m]) * dims;
#pragma unroll
                for (int n = 0; n < S; n++) {
                    float v = __fmul_rn(wxy, w[2][n]);
                    if (WEIGHTED) v = __fmul_rn(v, wp);
                    atomicAdd(row + idx[2][n], v);   // result unused -> RED.E.ADD.F32
                }
            }
        }
    } else {
        // plane: the reference pins the third axis to cell 0 with unit weight and still loops
        // over its S entries (MAS_library.pyx:138-139), so every cell receives S equal adds.
#pragma unroll
        for (int l = 0; l < S; l++) {
            float *row = number + (int64_t)idx[0][l] * dims;
#pragma unroll
            for (int m = 0; m < S; m++) {
                float v = __fmul_rn(w[0][l], w[1][m]);
                if (WEIGHTED) v = __fmul_rn(v, wp);
                atomicAdd(row + idx[1][m], v * (win.plane_mult > 0.0f ? win.plane_mult : (float)S));
            }
        }
    }
}

template <int MAS, int AXES, bool WEIGHTED, bool SLAB>
__device__ __forceinline__ void deposit_one(const float *p, float wp, float *__restrict__ number,
                                            int dims, float inv_cell_size, SlabWindow win,
                                            unsigned long long &dropped) {
    float dist[AXES];
#pragma unroll
    for (int a = 0; a < AXES; a++) dist[a] = cell_coordinate(p[a], inv_cell_size);
    deposit_dist<MAS, AXES, WEIGHTED, SLAB>(dist, wp, number, dims, win, dropped);
}

}  // namespace pyl
