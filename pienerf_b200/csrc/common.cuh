// Shared helpers for the pienerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/pienerf_b200.h"

void pn_set_error(const char *fmt, ...);

#define PN_STREAM(s) (reinterpret_cast<cudaStream_t>(s))

// Check the launch that was just enqueued (no sync).
#define PN_LAUNCH_CHECK(what)                                                          \
    do {                                                                               \
        cudaError_t _e = cudaGetLastError();                                           \
        if (_e != cudaSuccess) {                                                       \
            pn_set_error("%s: %s", what, cudaGetErrorString(_e));                      \
            return PN_ECUDA;                                                           \
        }                                                                              \
    } while (0)

#define PN_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) {                                                       \
            pn_set_error("%s: %s", #call, cudaGetErrorString(_e));                     \
            return PN_ECUDA;                                                           \
        }                                                                              \
    } while (0)

#define PN_REQUIRE(cond, msg)                                                          \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            pn_set_error("%s: %s", __func__, msg);                                     \
            return PN_EINVAL;                                                          \
        }                                                                              \
    } while (0)

template <typename T>
__host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

int pn_sm_count_cached();
