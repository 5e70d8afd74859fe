// common.cuh -- shared helpers for libffb6d_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ffb6d_b200.h"

namespace ffb6d {

// thread-local error text behind ffb6d_last_error()
void set_error(const char *fmt, ...);
// bumps the process-wide launch counter behind ffb6d_launch_count()
void count_launch(int n = 1);

#define FFB6D_CHECK_ARG(cond, ...)                         \
    do {                                                   \
        if (!(cond)) {                                     \
            ::ffb6d::set_error(__VA_ARGS__);               \
            return FFB6D_ERR_INVALID;                      \
        }                                                  \
    } while (0)

#define FFB6D_CUDA(call)                                                             \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::ffb6d::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                  \
            return FFB6D_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)

// after a <<<>>> launch
#define FFB6D_LAUNCH_OK(name)                                                          \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            ::ffb6d::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
            return FFB6D_ERR_CUDA;                                                     \
        }                                                                              \
        ::ffb6d::count_launch();                                                       \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// The reference distance: nanoflann L2_Adaptor::evalMetric tail loop for dim 3
// (NN/nanoflann.hpp:343-346): result = 0; result += d0*d0; += d1*d1; += d2*d2 in fp32,
// compiled without FMA.  The _rn intrinsics are never contracted by nvcc.
__device__ __forceinline__ float ref_sqdist(float qx, float qy, float qz, float sx, float sy,
                                            float sz)
{
    const float dx = __fsub_rn(qx, sx);
    const float dy = __fsub_rn(qy, sy);
    const float dz = __fsub_rn(qz, sz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace ffb6d
