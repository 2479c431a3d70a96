// nm_common.cuh -- shared device helpers and kernel argument structs.
#pragma once

#include "nm_platform.h"
#include "nm_fft.cuh"

#define NM_DBL_MAX 1.7976931348623157e308

// A batch of (window, channel) rows of W float64 samples each ("the preprocessed window").
// Either views straight into the resident re-referenced recording (off[w] = first sample of
// window w, ch_stride = padded recording length) or a chunk buffer written by the notch kernel.
struct NmRows {
    const double* base;
    long long ch_stride;
    const long long* off;  // device array, one entry per window of the batch
    int n_windows;
    int n_ch;
    int W;
};

// Where a kernel writes its features: out[(row0 + w) * F + colmap[c * per_ch + k]].
struct NmOut {
    double* out;
    long long row0;
    int F;
    const int* colmap;  // device array (n_ch * per_ch), -1 = not requested
    int per_ch;
};

NM_DEV void nm_store(const NmOut& o, int w, int c, int k, double v) {
    const int col = nm_ldg(o.colmap + (size_t)c * o.per_ch + k);
    if (col >= 0) o.out[(size_t)(o.row0 + w) * o.F + col] = v;
}

// numpy.nan_to_num with default arguments
NM_DEV double nm_nan_to_num(double v) {
    if (v != v) return 0.0;
    if (v > NM_DBL_MAX) return NM_DBL_MAX;
    if (v < -NM_DBL_MAX) return -NM_DBL_MAX;
    return v;
}

NM_DEV double nm_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
NM_DEV double nm_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = (u > v || u != u) ? u : v;  // propagate NaN like numpy.max
    }
    return v;
}
NM_DEV double nm_warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double u = __shfl_xor_sync(0xffffffffu, v, o);
        v = (u < v || u != u) ? u : v;
    }
    return v;
}
NM_DEV int nm_warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp-wide sums of NV (power of two, <= 8) values with a halving exchange: in every step a lane keeps one half of
// its values and trades the other half with a partner, so NV values cost NV - 1 + (5 - log2 NV) shuffles instead of
// 5 * NV.  On return the total of value i sits in v[0] of every lane whose index has bits (4, 3, ..) == bits of i,
// i.e. value i is held by lanes with  (lane >> (5 - log2 NV)) == i.
template <int NV>
NM_DEV void nm_warp_sum_multi(double* v, int lane) {
    static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8, "NV must be a power of two <= 8");
    int bit = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double keep = upper ? v[i + n / 2] : v[i];
            const double give = upper ? v[i] : v[i + n / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, give, bit);
        }
        bit >>= 1;
    }
#pragma unroll
    for (int o = bit; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}

// Sum NV values over the whole CTA; every thread receives the totals.  `red` needs
// NV * 32 doubles of shared memory.  Contains two barriers.
template <int NV>
NM_DEV void nm_block_sum(double* v, double* red, int tid, int nt) {
    const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = nm_warp_sum(v[i]);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += red[i * 32 + w];
        v[i] = s;
    }
}
