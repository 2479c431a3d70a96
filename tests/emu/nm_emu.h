// nm_emu.h -- TEST INFRASTRUCTURE.  Host-thread emulation of the CUDA execution model.
//
// Lets the *unchanged* kernel sources under py_neuromodulation_b200/csrc/ be compiled by g++
// (-DNM_EMULATE) so that indexing / synchronisation / algorithm bugs show up in the CPU test
// suite of a container that has no GPU.  One cooperative fiber per CUDA thread of a CTA (switching at
// barriers / shuffles), CTAs of a launch spread over a few OS threads.  Not shipped, not a
// fallback: the package never loads a library built this way.
#pragma once

#include <algorithm>
#include <atomic>
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

// ---- execution model: one FIBER per CUDA thread -----------------------------------------------
// Every CUDA thread of a CTA is a user-space fiber with its own stack; all fibers of a CTA live on one OS
// thread and switch cooperatively at the synchronisation points (__syncthreads, __syncwarp, shuffles,
// ballots), so a barrier costs a handful of register moves instead of a futex round trip.  CTAs of a launch
// are spread over a few OS worker threads (they are independent on the GPU as well).
extern "C" void nm_emu_swap(void** save_sp, void* load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.globl nm_emu_swap
.type nm_emu_swap,@function
nm_emu_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size nm_emu_swap,.-nm_emu_swap
)");
#else
#error "tests/emu/nm_emu.h: the fiber switch is written for x86-64 only"
#endif

namespace nm_emu {
constexpr size_t kStackBytes = 256 * 1024;

struct Cta;
struct Fiber {
    dim3 tid;
    unsigned lin = 0;
    void* sp = nullptr;
    bool done = false;
    Cta* cta = nullptr;
};
struct WarpSlots {
    unsigned lanes = 0, arrived = 0, gen = 0;
    unsigned long long val[32];
};
struct Cta {
    dim3 bid, bdim, gdim;
    unsigned nt = 0, alive = 0, arrived = 0, gen = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpSlots> warps;
    unsigned char* smem = nullptr;
    const std::function<void()>* body = nullptr;
    void* sched_sp = nullptr;
    unsigned cur = 0;
};
struct Tls {
    dim3 tid, bid, bdim, gdim;
    unsigned char* smem = nullptr;
    Cta* cta = nullptr;
    Fiber* fib = nullptr;
};
inline thread_local Tls tl;

inline void enter(Fiber* f) {
    Cta* c = f->cta;
    tl.tid = f->tid;
    tl.bid = c->bid;
    tl.bdim = c->bdim;
    tl.gdim = c->gdim;
    tl.smem = c->smem;
    tl.cta = c;
    tl.fib = f;
}

// hand the OS thread to the next unfinished fiber of the CTA (round robin); returns when this fiber is resumed
inline void yield() {
    Fiber* me = tl.fib;
    Cta* c = me->cta;
    unsigned nxt = me->lin;
    for (unsigned k = 0; k < c->nt; ++k) {
        nxt = (nxt + 1 == c->nt) ? 0 : nxt + 1;
        if (!c->fibers[nxt].done) break;
    }
    if (nxt == me->lin) {
        if (me->done) nm_emu_swap(&me->sp, c->sched_sp);  // last fiber finished: back to the launcher
        return;                                            // nobody else to run
    }
    Fiber* to = &c->fibers[nxt];
    c->cur = nxt;
    nm_emu_swap(&me->sp, to->sp);
    enter(me);
}

inline void cta_release_if_complete(Cta* c) {
    if (c->arrived && c->arrived >= c->alive) {
        c->arrived = 0;
        c->gen++;
    }
}

[[noreturn]] inline void fiber_main() {
    Fiber* me = tl.fib;
    Cta* c = me->cta;
    (*c->body)();
    me->done = true;
    c->alive--;
    cta_release_if_complete(c);  // exited threads do not take part in later barriers
    WarpSlots& w = c->warps[me->lin / 32];
    w.lanes--;
    if (w.arrived && w.arrived >= w.lanes) { w.arrived = 0; w.gen++; }
    for (;;) yield();  // never resumed once every fiber is done (yield() returns to the launcher then)
}
extern "C" inline void nm_emu_trampoline() {
    enter(tl.cta->fibers.data() + tl.cta->cur);
    fiber_main();
}

struct Worker {
    std::unique_ptr<unsigned char[]> stacks;  // not zero-filled: pages are committed only when a fiber touches them
    size_t stack_cap = 0;
    std::vector<unsigned char> smem;
    void run_cta(dim3 grid, dim3 block, dim3 bid, size_t smem_bytes, const std::function<void()>& body) {
        const unsigned nt = block.x * block.y * block.z;
        if (stack_cap < (size_t)nt * kStackBytes) {
            stack_cap = (size_t)nt * kStackBytes;
            stacks.reset(new unsigned char[stack_cap]);
        }
        smem.assign(smem_bytes + 128, 0);
        Cta cta;
        cta.bid = bid; cta.bdim = block; cta.gdim = grid;
        cta.nt = cta.alive = nt;
        cta.body = &body;
        unsigned char* sm = smem.data();
        sm += (64 - (reinterpret_cast<uintptr_t>(sm) & 63)) & 63;
        cta.smem = sm;
        cta.fibers.resize(nt);
        cta.warps.resize((nt + 31) / 32);
        for (unsigned w = 0; w < cta.warps.size(); ++w) {
            cta.warps[w].lanes = std::min(32u, nt - w * 32);
            for (auto& v : cta.warps[w].val) v = 0;
        }
        for (unsigned t = 0; t < nt; ++t) {
            Fiber& f = cta.fibers[t];
            f.cta = &cta;
            f.lin = t;
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            uintptr_t top = reinterpret_cast<uintptr_t>(stacks.get() + (size_t)(t + 1) * kStackBytes);
            top &= ~uintptr_t(15);
            void** sp = reinterpret_cast<void**>(top);
            *--sp = nullptr;                                           // keeps rsp = 8 (mod 16) at function entry
            *--sp = reinterpret_cast<void*>(&nm_emu_trampoline);       // `ret` target
            for (int r = 0; r < 6; ++r) *--sp = nullptr;               // rbp rbx r12 r13 r14 r15
            f.sp = sp;
        }
        tl.cta = &cta;
        cta.cur = 0;
        nm_emu_swap(&cta.sched_sp, cta.fibers[0].sp);  // returns when every fiber has finished
        tl = Tls{};
    }
};

inline unsigned n_workers() {
    static const unsigned n = [] {
        const char* e = std::getenv("NM_EMU_WORKERS");
        unsigned v = e ? (unsigned)std::atoi(e) : std::thread::hardware_concurrency();
        return std::max(1u, std::min(v, 16u));
    }();
    return n;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    const unsigned long long n_cta = (unsigned long long)grid.x * grid.y * grid.z;
    std::atomic<unsigned long long> next{0};
    auto work = [&]() {
        static thread_local Worker wk;
        for (;;) {
            const unsigned long long i = next.fetch_add(1);
            if (i >= n_cta) break;
            dim3 bid((unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((unsigned long long)grid.x * grid.y)));
            wk.run_cta(grid, block, bid, smem_bytes, body);
        }
    };
    const unsigned nw = (unsigned)std::min<unsigned long long>(n_workers(), n_cta);
    if (nw <= 1) { work(); return; }
    std::vector<std::thread> pool;
    pool.reserve(nw);
    for (unsigned t = 0; t < nw; ++t) pool.emplace_back(work);
    for (auto& th : pool) th.join();
}

inline void cta_barrier() {
    Cta* c = tl.cta;
    const unsigned g = c->gen;
    c->arrived++;
    cta_release_if_complete(c);
    while (c->gen == g) yield();
}
inline void warp_barrier() {
    Cta* c = tl.cta;
    WarpSlots& w = c->warps[tl.fib->lin / 32];
    const unsigned g = w.gen;
    w.arrived++;
    if (w.arrived >= w.lanes) { w.arrived = 0; w.gen++; }
    while (w.gen == g) yield();
}
}  // namespace nm_emu

#define threadIdx (nm_emu::tl.tid)
#define blockIdx (nm_emu::tl.bid)
#define blockDim (nm_emu::tl.bdim)
#define gridDim (nm_emu::tl.gdim)

inline void __syncthreads() { nm_emu::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { nm_emu::warp_barrier(); }

template <typename T>
inline T nm_emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    unsigned lin = threadIdx.x + threadIdx.y * blockDim.x;
    auto& w = nm_emu::tl.cta->warps[lin / 32];
    unsigned lane = lin % 32;
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.val[lane] = raw;
    nm_emu::warp_barrier();
    unsigned long long got = w.val[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
    nm_emu::warp_barrier();
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
    nm_emu::warp_barrier();
    unsigned bits = 0;
    for (int l = 0; l < 32; ++l)
        if (w.val[l]) bits |= (1u << l);
    nm_emu::warp_barrier();
    return bits;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(unsigned(v)); }
inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }

template <typename T>
inline T nm_ldg(const T* p) { return *p; }
inline double nm_mul_rn(double a, double b) { volatile double r = a * b; return r; }  // (volatile: no contraction whatever the flags)
inline double nm_add_rn(double a, double b) { volatile double r = a + b; return r; }
inline double nm_sub_rn(double a, double b) { volatile double r = a - b; return r; }

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
inline bool nm_gs_updating() { return false; }  // (no CUDA graphs in the test build: the streaming entry always launches eagerly)
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 4 : 227 * 1024;
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
enum { cudaHostRegisterPortable = 1 };
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; ++r) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new nm_emu_event{0}; return cudaSuccess; }
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
