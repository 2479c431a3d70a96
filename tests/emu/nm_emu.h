// nm_emu.h -- TEST INFRASTRUCTURE.  Host-thread emulation of the CUDA execution model.
//
// Lets the *unchanged* kernel sources under py_neuromodulation_b200/csrc/ be compiled by g++
// (-DNM_EMULATE) so that indexing / synchronisation / algorithm bugs show up in the CPU test
// suite of a container that has no GPU.  One std::thread per CUDA thread of a CTA; CTAs of a
// launch run one after the other; __syncthreads() is a std::barrier.  Not shipped, not a
// fallback: the package never loads a library built this way.
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define NM_GLOBAL
#define NM_DEV inline
#define NM_DEV_NOINLINE inline
#define NM_HD inline
#define NM_RESTRICT __restrict__
#define NM_SHARED_BYTES(name) unsigned char* name = nm_emu::tl.smem
#define NM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    nm_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define NM_LAUNCH_BOUNDS(t, b)
#define __forceinline__ inline

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 { double x, y; };
struct float2 { float x, y; };
struct int2 { int x, y; };
struct float4 { float x, y, z, w; };
inline double2 make_double2(double a, double b) { return {a, b}; }
inline float2 make_float2(float a, float b) { return {a, b}; }

namespace nm_emu {
struct WarpSlots {
    std::unique_ptr<std::barrier<>> bar;
    unsigned long long val[32];
};
struct Cta {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<WarpSlots> warps;
    std::vector<unsigned char> smem;
};
struct Tls {
    dim3 tid, bid, bdim, gdim;
    unsigned char* smem = nullptr;
    Cta* cta = nullptr;
};
inline thread_local Tls tl;

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const unsigned nt = block.x * block.y * block.z;
    Cta cta;
    cta.bar = std::make_unique<std::barrier<>>(nt);
    cta.smem.assign(smem_bytes + 64, 0);
    cta.warps.resize((nt + 31) / 32);
    for (unsigned w = 0; w < cta.warps.size(); ++w) {
        unsigned lanes = std::min(32u, nt - w * 32);
        cta.warps[w].bar = std::make_unique<std::barrier<>>(lanes);
        for (auto& v : cta.warps[w].val) v = 0;
    }
    unsigned char* sm = cta.smem.data();
    sm += (64 - (reinterpret_cast<uintptr_t>(sm) & 63)) & 63;
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (unsigned t = 0; t < nt; ++t) {
        pool.emplace_back([&, t]() {
            tl.cta = &cta;
            tl.smem = sm;
            tl.bdim = block;
            tl.gdim = grid;
            tl.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        tl.bid = dim3(bx, by, bz);
                        body();
                        cta.bar->arrive_and_wait();
                    }
        });
    }
    for (auto& th : pool) th.join();
}
}  // namespace nm_emu

#define threadIdx (nm_emu::tl.tid)
#define blockIdx (nm_emu::tl.bid)
#define blockDim (nm_emu::tl.bdim)
#define gridDim (nm_emu::tl.gdim)

inline void __syncthreads() { nm_emu::tl.cta->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) {
    unsigned lin = threadIdx.x + threadIdx.y * blockDim.x;
    nm_emu::tl.cta->warps[lin / 32].bar->arrive_and_wait();
}

template <typename T>
inline T nm_emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned lin = threadIdx.x + threadIdx.y * blockDim.x;
    auto& w = nm_emu::tl.cta->warps[lin / 32];
    unsigned lane = lin % 32;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.val[lane] = raw;
    w.bar->arrive_and_wait();
    unsigned long long got = w.val[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
    w.bar->arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return nm_emu_exchange(v, src & 31); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
    unsigned lane = (threadIdx.x + threadIdx.y * blockDim.x) % 32;
    return nm_emu_exchange(v, int(lane ^ unsigned(m)));
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    unsigned lane = (threadIdx.x + threadIdx.y * blockDim.x) % 32;
    return nm_emu_exchange(v, lane + d < 32 ? int(lane + d) : int(lane));
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    unsigned lane = (threadIdx.x + threadIdx.y * blockDim.x) % 32;
    return nm_emu_exchange(v, lane >= d ? int(lane - d) : int(lane));
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned lin = threadIdx.x + threadIdx.y * blockDim.x;
    auto& w = nm_emu::tl.cta->warps[lin / 32];
    w.val[lin % 32] = pred ? 1ull : 0ull;
    w.bar->arrive_and_wait();
    unsigned bits = 0;
    for (int l = 0; l < 32; ++l)
        if (w.val[l]) bits |= (1u << l);
    w.bar->arrive_and_wait();
    return bits;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(unsigned(v)); }
inline int __ffs(int v) { return __builtin_ffs(v); }

template <typename T>
inline T nm_ldg(const T* p) { return *p; }

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::max;
using std::min;
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }

// ---- minimal CUDA runtime surface ------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
struct nm_emu_event { double t; };
typedef nm_emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount, cudaDevAttrMaxSharedMemoryPerBlockOptin };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum { cudaStreamNonBlocking = 1 };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 2 : 227 * 1024;
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new nm_emu_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); e->t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = float(b->t - a->t); return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
