// nm_platform.h -- thin layer between the kernels and the CUDA toolchain.
//
// Product build: nvcc, sm_100a; every macro below maps 1:1 onto CUDA.
//
// Test build (tests/emu/, g++ -DNM_EMULATE): the same kernel sources are compiled as host
// code and each CTA is executed by real std::threads with a std::barrier standing in for
// __syncthreads() (tests/emu/nm_emu.h).  That build exists ONLY so that kernel *logic* can
// be checked against the oracle in a container without a GPU; it is never shipped, never
// loaded by the package, and is not a fallback (py_neuromodulation_b200/_lib.py loads the
// CUDA library or raises).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cmath>

#ifdef NM_EMULATE
#include "nm_emu.h"
#else
#include <cuda_runtime.h>

#define NM_GLOBAL __global__
#define NM_DEV __device__ __forceinline__
#define NM_DEV_NOINLINE __device__ __noinline__
#define NM_HD __host__ __device__ __forceinline__
#define NM_RESTRICT __restrict__
#define NM_SHARED_BYTES(name) extern __shared__ __align__(16) unsigned char name[]
#define NM_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define NM_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)

template <typename T>
NM_DEV T nm_ldg(const T* p) { return __ldg(p); }
#endif

// ---- error handling shared by both builds -------------------------------------------------
void nm_set_error(const char* fmt, ...);

#define NM_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            nm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -1;                                                                        \
        }                                                                                     \
    } while (0)

#define NM_CHECK(cond, ...)                   \
    do {                                      \
        if (!(cond)) {                        \
            nm_set_error(__VA_ARGS__);        \
            return -1;                        \
        }                                     \
    } while (0)
