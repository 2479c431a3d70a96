// nm_host.h -- host-side helpers of the pipeline: device buffers, FFT plan construction,
// filter-spectrum tables.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nm_common.cuh"

#ifdef NM_EMULATE
#define NM_FFT_THREADS 64
#define NM_ROW_THREADS 64
#else
#define NM_FFT_THREADS 256
#define NM_ROW_THREADS 256
#endif

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        release();
        if (bytes == 0) bytes = 16;
        NM_CUDA_CHECK(cudaMalloc(&p, bytes));
        cap = bytes;
        return 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
    template <typename T>
    int upload(const T* src, size_t n, cudaStream_t s) {
        if (ensure(n ? n * sizeof(T) : 16)) return -1;
        if (n) NM_CUDA_CHECK(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, s));
        // pageable sources are staged by the runtime before the call returns, so `src` may be a temporary
        return 0;
    }
    template <typename T>
    int upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
};

// ---- FFT plans ---------------------------------------------------------------------------------
struct FftPlanHost {
    int n = 0;
    std::vector<int> radix, len, pos;
    bool generic = false;
    DevBuf d_tw, d_pos;

    // explicit radix list (power-of-two convolution plans)
    int build_with(int n_, const std::vector<int>& radices, cudaStream_t s) {
        n = n_;
        radix = radices;
        generic = false;
        return finish(s);
    }

    int build(int n_, cudaStream_t s) {
        n = n_;
        radix.clear();
        len.clear();
        generic = false;
        int m = n;
        std::vector<int> r2, r3, r5, rg;
        while (m % 4 == 0) { r2.push_back(4); m /= 4; }
        while (m % 2 == 0) { r2.push_back(2); m /= 2; }
        while (m % 3 == 0) { r3.push_back(3); m /= 3; }
        while (m % 5 == 0) { r5.push_back(5); m /= 5; }
        for (int f = 7; (long long)f * f <= m; f += 2)
            while (m % f == 0) { rg.push_back(f); m /= f; }
        if (m > 1) rg.push_back(m);
        generic = !rg.empty();
        // forward (DIF) order: generic primes, 4s, 2, 3s and the odd radix 5 last -- the last
        // passes have unit butterfly stride, where an odd radix is shared-memory bank-conflict free
        for (int v : rg) radix.push_back(v);
        for (int v : r2) radix.push_back(v);
        for (int v : r3) radix.push_back(v);
        for (int v : r5) radix.push_back(v);
        return finish(s);
    }

    int finish(cudaStream_t s) {
        NM_CHECK((int)radix.size() <= NM_MAX_PASS, "FFT length %d has too many factors", n);
        len.clear();
        int L = n;
        for (int v : radix) { len.push_back(L); L /= v; }
        pos.assign(n, 0);
        for (int f = 0; f < n; ++f) {
            int rem = f, Lc = n, p = 0;
            for (size_t i = 0; i < radix.size(); ++i) {
                const int k = rem % radix[i];
                rem /= radix[i];
                Lc /= radix[i];
                p += k * Lc;
            }
            pos[f] = p;
        }
        std::vector<cx<double>> tw(n);
        for (int k = 0; k < n; ++k) {
            const double ang = -2.0 * M_PI * (double)k / (double)n;
            tw[k] = {std::cos(ang), std::sin(ang)};
        }
        if (d_tw.upload(tw, s)) return -1;
        if (d_pos.upload(pos, s)) return -1;
        return 0;
    }
    NmFft<double> dev() const {
        NmFft<double> f;
        f.n = n;
        f.npass = (int)radix.size();
        for (int i = 0; i < f.npass; ++i) { f.radix[i] = radix[i]; f.len[i] = len[i]; }
        f.tw = d_tw.as<cx<double>>();
        f.pos = d_pos.as<int>();
        return f;
    }
};

static inline int nm_next_smooth(int n) {
    for (int m = n;; ++m) {
        int k = m;
        while (k % 2 == 0) k /= 2;
        while (k % 3 == 0) k /= 3;
        while (k % 5 == 0) k /= 5;
        if (k == 1) return m;
    }
}

// Real spectrum of nF centred symmetric FIRs on a P-point circle, digit-reversed order, scaled 1/P.
static inline int nm_build_hperm(const double* taps, int nF, int L, const FftPlanHost& plan, std::vector<double>& hperm) {
    const int P = plan.n, Lh = (L - 1) / 2;
    NM_CHECK(L % 2 == 1, "FIR length must be odd (zero-phase), got %d", L);
    std::vector<double> ct(P);
    for (int k = 0; k < P; ++k) ct[k] = std::cos(2.0 * M_PI * (double)k / (double)P);
    hperm.assign((size_t)nF * P, 0.0);
    for (int f = 0; f < nF; ++f) {
        const double* h = taps + (size_t)f * L;
        double hmax = 0;
        for (int j = 0; j < L; ++j) hmax = std::fmax(hmax, std::fabs(h[j]));
        for (int j = 0; j < Lh; ++j)
            NM_CHECK(std::fabs(h[j] - h[L - 1 - j]) <= 1e-12 * hmax, "FIR taps must be symmetric (zero-phase design)");
        for (int k = 0; k < P; ++k) {
            double acc = 0.0;
            for (int j = Lh; j >= 1; --j) acc += h[Lh + j] * ct[(int)(((long long)k * j) % P)];
            hperm[(size_t)f * P + plan.pos[k]] = (h[Lh] + 2.0 * acc) / (double)P;
        }
    }
    return 0;
}
