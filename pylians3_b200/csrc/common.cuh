// Shared helpers for the pyl_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pyl_b200.h"

namespace pyl {

void set_last_error(const char *fmt, ...);

#define PYL_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            pyl::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                                cudaGetErrorString(_e));                                  \
            return PYL_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// every kernel launch of this library is followed by PYL_LAUNCH_CHECK(), which also feeds the
// launch counter read by pyl_kernel_launches() (bench.py reports it as "gpu_launches")
void count_launch();
#define PYL_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        pyl::count_launch();                                                              \
        PYL_CUDA_CHECK(cudaGetLastError());                                               \
    } while (0)

#define PYL_REQUIRE(cond, msg)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            pyl::set_last_error("%s:%d: %s", __FILE__, __LINE__, msg);                    \
            return PYL_ERR_ARG;                                                           \
        }                                                                                 \
    } while (0)

inline cudaStream_t as_stream(pyl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached per device)
int sm_count();

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pyl
