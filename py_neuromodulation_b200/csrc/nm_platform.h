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
#define NM_LAUNCH(kernel, grid, block, smem, stream, ...) nm_launch((kernel), (grid), (block), (smem), (stream), __VA_ARGS__)
#define NM_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)

// ---- launches that a CUDA graph can replay (streaming entry, nm_stream.cuh) -------------------------
// The per-window sequence of the streaming entry is captured ONCE into a graph; for every later window the same host code runs
// in "update" mode: a launch is not issued but compared with the captured kernel node and, if its arguments changed (window
// counters of the stateful families), patched into the executable graph (cudaGraphExecKernelNodeSetParams).  One
// cudaGraphLaunch then replays the whole window.  Outside a session (every batched path) nm_launch is a plain launch.
#include <cstring>
#include <tuple>
#include <utility>
#include <vector>

struct NmGraphNodeRec {
    cudaGraphNode_t node;
    const void* func;
    dim3 grid, block;
    size_t smem;
    std::vector<unsigned char> blob;  // the kernel arguments, byte for byte
};
struct NmGraphSession {
    int mode = 0;  // 0 eager, 1 capturing, 2 updating an instantiated graph
    cudaGraphExec_t exec = nullptr;
    std::vector<NmGraphNodeRec> nodes;
    size_t cursor = 0;
    bool broken = false;
    long long patched = 0;
};
static thread_local NmGraphSession* nm_graph_session = nullptr;
static inline bool nm_gs_updating() { return nm_graph_session && nm_graph_session->mode == 2; }

static inline void nm_graph_blob(std::vector<unsigned char>& blob, void** ptrs, const size_t* sizes, size_t n) {
    size_t tot = 0;
    for (size_t i = 0; i < n; ++i) tot += sizes[i];
    blob.resize(tot);
    size_t at = 0;
    for (size_t i = 0; i < n; ++i) { std::memcpy(blob.data() + at, ptrs[i], sizes[i]); at += sizes[i]; }
}

static inline void nm_graph_record(NmGraphSession* gs, cudaStream_t stream, const void* func, dim3 grid, dim3 block, size_t smem,
                                   void** ptrs, const size_t* sizes, size_t n) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    if (cudaStreamGetCaptureInfo(stream, &st, nullptr, nullptr, &deps, &n_deps) != cudaSuccess || st != cudaStreamCaptureStatusActive ||
        n_deps != 1) {
        gs->broken = true;  // (after a kernel launch the capturing stream depends on exactly that kernel node)
        return;
    }
    NmGraphNodeRec r;
    r.node = deps[0];
    r.func = func;
    r.grid = grid; r.block = block; r.smem = smem;
    nm_graph_blob(r.blob, ptrs, sizes, n);
    gs->nodes.push_back(std::move(r));
}

static inline void nm_graph_update(NmGraphSession* gs, const void* func, dim3 grid, dim3 block, size_t smem, void** ptrs,
                                   const size_t* sizes, size_t n) {
    if (gs->broken || gs->cursor >= gs->nodes.size() || gs->nodes[gs->cursor].func != func) {
        gs->broken = true;
        return;
    }
    NmGraphNodeRec& r = gs->nodes[gs->cursor++];
    std::vector<unsigned char> blob;
    nm_graph_blob(blob, ptrs, sizes, n);
    const bool same = blob == r.blob && grid.x == r.grid.x && grid.y == r.grid.y && grid.z == r.grid.z && block.x == r.block.x &&
                      block.y == r.block.y && block.z == r.block.z && smem == r.smem;
    if (same) return;
    cudaKernelNodeParams kp;
    std::memset(&kp, 0, sizeof(kp));
    kp.func = const_cast<void*>(func);
    kp.gridDim = grid;
    kp.blockDim = block;
    kp.sharedMemBytes = (unsigned)smem;
    kp.kernelParams = ptrs;
    kp.extra = nullptr;
    if (cudaGraphExecKernelNodeSetParams(gs->exec, r.node, &kp) != cudaSuccess) {
        gs->broken = true;
        return;
    }
    r.blob.swap(blob);
    r.grid = grid; r.block = block; r.smem = smem;
    gs->patched++;
}

template <typename... KArgs, size_t... I>
static inline void nm_launch_impl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  std::tuple<KArgs...>& pack, std::index_sequence<I...>) {
    void* ptrs[] = {static_cast<void*>(&std::get<I>(pack))..., nullptr};
    const size_t sizes[] = {sizeof(KArgs)..., 0};
    NmGraphSession* gs = nm_graph_session;
    if (gs && gs->mode == 2) {
        nm_graph_update(gs, (const void*)kernel, grid, block, smem, ptrs, sizes, sizeof...(KArgs));
        return;
    }
    cudaLaunchKernel((const void*)kernel, grid, block, ptrs, smem, stream);
    if (gs && gs->mode == 1) nm_graph_record(gs, stream, (const void*)kernel, grid, block, smem, ptrs, sizes, sizeof...(KArgs));
}

template <typename... KArgs, typename... Args>
static inline void nm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    std::tuple<KArgs...> pack(static_cast<KArgs>(std::forward<Args>(args))...);
    nm_launch_impl(kernel, grid, block, smem, stream, pack, std::index_sequence_for<KArgs...>{});
}

template <typename T>
NM_DEV T nm_ldg(const T* p) { return __ldg(p); }
// single IEEE operations the compiler must not contract into fused multiply-adds (replays of numpy arithmetic)
NM_DEV double nm_mul_rn(double a, double b) { return __dmul_rn(a, b); }
NM_DEV double nm_add_rn(double a, double b) { return __dadd_rn(a, b); }
NM_DEV double nm_sub_rn(double a, double b) { return __dsub_rn(a, b); }
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
