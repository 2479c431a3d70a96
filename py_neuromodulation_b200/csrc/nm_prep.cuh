// nm_prep.cuh -- once-per-recording preprocessing that is independent of the window grid:
//   nan_to_num -> channel pick -> re-reference            (stream/data_processor.py:255,
//                                                           processing/rereference.py:99-100)
// plus the per-32-sample NaN block map that makes the per-window NaN-channel mask
// (stream/data_processor.py:253) cheap.
//
// Re-referencing is linear and acts across channels, the notch FIR is linear and acts along
// time, so they commute: the reference applies notch then re-reference per window; here the
// re-reference is applied ONCE to the whole recording and the (window-boundary dependent)
// notch afterwards -- identical up to float64 rounding, and 10x less work at 90 % overlap.
//
// The (C x C) reference matrix is handed over factored as  group-sum coefficients + a sparse
// remainder (common-average rows collapse to one coefficient on the per-sample group sum).
#pragma once

#include "nm_common.cuh"

#define NM_MAX_GROUPS 8

struct NmPrepArgs {
    const void* raw;       // (C_all, T) float32 or float64, row pitch raw_pitch elements
    int raw_is_f64;
    long long raw_pitch;
    int C_all;
    long long T;
    int C;                 // feature channels
    const int* pick;       // [C] raw row of feature channel j
    int G;                 // number of group sums (<= NM_MAX_GROUPS)
    const int* group_of;   // [C] group id or -1
    const double* gcoef;   // [C * G]
    const int* sp_ptr;     // [C + 1]
    const int* sp_col;     // feature-channel index
    const double* sp_val;
    double* xr;            // (C, xr_pitch) output
    long long xr_pitch;
    unsigned char* nanblk; // (C_all, nanblk_pitch): 1 if any NaN in the 32-sample block
    long long nanblk_pitch;
    // channel-sharded multi-GPU runs: group sums over ALL ranks' channels, (G, gsum_pitch), all-reduced by
    // the host between nm_gsum_kernel and nm_prep_kernel; nullptr = sum the local channels in-kernel
    const double* gsum_ext;
    long long gsum_pitch;
    long long t0, t1;      // samples [t0, t1) are processed by this launch (time slices of a pipelined upload)
    // fused window kernel (nm_fused.cuh): the re-referenced copy is not materialised (write_xr == 0); the launch only leaves
    // the NaN block map and the local group sums (gsum_out, (G, gsum_pitch), nullptr when gsum_ext already holds them)
    double* gsum_out;
    int write_xr;
};

template <int RT>
NM_DEV double nm_raw_at_t(const NmPrepArgs& a, int row, long long t) {
    const bool f64 = RT < 0 ? a.raw_is_f64 != 0 : RT != 0;
    return f64 ? nm_ldg(reinterpret_cast<const double*>(a.raw) + (size_t)row * a.raw_pitch + t)
               : (double)nm_ldg(reinterpret_cast<const float*>(a.raw) + (size_t)row * a.raw_pitch + t);
}

NM_DEV double nm_raw_at(const NmPrepArgs& a, int row, long long t) {
    return a.raw_is_f64 ? nm_ldg(reinterpret_cast<const double*>(a.raw) + (size_t)row * a.raw_pitch + t)
                        : (double)nm_ldg(reinterpret_cast<const float*>(a.raw) + (size_t)row * a.raw_pitch + t);
}

// CTA = 32 consecutive samples (lanes) x NM_PREP_WARPS channel groups (warps): warp g serves the channels / raw rows
// congruent to g, so a time slice of a few ten thousand samples already fills the GPU and every global access is a coalesced
// row segment.  The per-sample group sums are combined across the warps through shared memory (fixed order: deterministic).
#define NM_PREP_WARPS 8
#define NM_PREP_THREADS (32 * NM_PREP_WARPS)
static NM_HD size_t nm_prep_smem_bytes() { return (size_t)NM_PREP_WARPS * NM_MAX_GROUPS * 32 * sizeof(double); }

// group sums of this CTA's 32 samples over the local channels -> S[NM_MAX_GROUPS] in every thread (contains two barriers)
// NG = compile-time bound of the group count (a.G <= NG), RT = raw type (0 float32, 1 float64, -1 read a.raw_is_f64)
template <int NG, int RT>
NM_DEV void nm_prep_group_sums(const NmPrepArgs& a, long long t, bool active, int lane, int g, double* part, double* S) {
    double P[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) P[q] = 0.0;
    if (active) {
        for (int j = g; j < a.C; j += NM_PREP_WARPS) {
            const int grp = nm_ldg(a.group_of + j);
            if (grp >= 0) {
                const double v = nm_nan_to_num(nm_raw_at_t<RT>(a, nm_ldg(a.pick + j), t));
#pragma unroll
                for (int q = 0; q < NG; ++q)
                    if (q == grp) P[q] += v;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NG; ++q)
        if (q < a.G) part[(g * NM_MAX_GROUPS + q) * 32 + lane] = P[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        double s = 0.0;
        if (q < a.G)
            for (int w = 0; w < NM_PREP_WARPS; ++w) s += part[(w * NM_MAX_GROUPS + q) * 32 + lane];
        S[q] = s;
    }
    __syncthreads();
}

// specialisations: <1, 0> / <1, 1> = at most one reference group (common average) on float32 / float64 input; <NM_MAX_GROUPS, -1> =
// everything else.  Same arithmetic in the same order in all of them.
template <int NG, int RT>
NM_GLOBAL void nm_prep_kernel(NmPrepArgs a) {
    NM_SHARED_BYTES(smem);
    double* part = reinterpret_cast<double*>(smem);
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long t = a.t0 + (long long)blockIdx.x * 32 + lane;  // t0 is a multiple of 32 (NaN block map)
    const bool active = t < a.t1;

    for (int r = g; r < a.C_all; r += NM_PREP_WARPS) {
        const double v = active ? nm_raw_at_t<RT>(a, r, t) : 0.0;
        const unsigned bits = __ballot_sync(0xffffffffu, v != v);
        if (lane == 0 && active) a.nanblk[(size_t)r * a.nanblk_pitch + (t >> 5)] = bits ? 1 : 0;
    }

    double S[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) S[q] = 0.0;
    if (a.G > 0 && a.gsum_ext) {
#pragma unroll
        for (int q = 0; q < NG; ++q)
            if (q < a.G && active) S[q] = a.gsum_ext[(size_t)q * a.gsum_pitch + t];
    } else if (a.G > 0) {
        nm_prep_group_sums<NG, RT>(a, t, active, lane, g, part, S);
    }
    if (!active) return;
    if (a.gsum_out) {
#pragma unroll
        for (int q = 0; q < NG; ++q)
            if (q < a.G && (q % NM_PREP_WARPS) == g) a.gsum_out[(size_t)q * a.gsum_pitch + t] = S[q];
    }
    if (!a.write_xr) return;
    for (int i = g; i < a.C; i += NM_PREP_WARPS) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < NG; ++q)
            if (q < a.G) acc += nm_ldg(a.gcoef + (size_t)i * a.G + q) * S[q];
        const int k1 = nm_ldg(a.sp_ptr + i + 1);
        for (int k = nm_ldg(a.sp_ptr + i); k < k1; ++k)
            acc += nm_ldg(a.sp_val + k) * nm_nan_to_num(nm_raw_at_t<RT>(a, nm_ldg(a.pick + nm_ldg(a.sp_col + k)), t));
        a.xr[(size_t)i * a.xr_pitch + t] = acc;
    }
}

// local per-sample group sums of this rank's shard -> gsum (G, gsum_pitch); same CTA geometry as nm_prep_kernel
NM_GLOBAL void nm_gsum_kernel(NmPrepArgs a, double* gsum) {
    NM_SHARED_BYTES(smem);
    double* part = reinterpret_cast<double*>(smem);
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long long t = a.t0 + (long long)blockIdx.x * 32 + lane;
    const bool active = t < a.t1;
    double S[NM_MAX_GROUPS];
    nm_prep_group_sums<NM_MAX_GROUPS, -1>(a, t, active, lane, g, part, S);
    if (!active) return;
    for (int q = g; q < a.G; q += NM_PREP_WARPS) gsum[(size_t)q * a.gsum_pitch + t] = S[q];
}

// flags[w * C_all + r] = 1 iff raw row r has a NaN inside window [start[w], start[w] + W)
struct NmNanArgs {
    NmPrepArgs p;
    const long long* start;
    int n_windows;
    int W;
    unsigned char* flags;
};

// out[(row0 + w) * F + col] = NaN for every column listed for a flagged raw row
struct NmNanFillArgs {
    const unsigned char* flags;
    int n_windows;
    int C_all;
    const int* col_ptr;  // [C_all + 1]
    const int* cols;
    double* out;
    long long row0;
    int F;
};

// flag of every (window, raw row) from the 32-sample NaN block map (exact check at the window edges) and, if set, the NaN
// fill of the columns listed for that raw row -- one thread per (window, raw row)
NM_GLOBAL void nm_nanfix_kernel(NmNanArgs a, NmNanFillArgs f) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.n_windows * a.p.C_all) return;
    const int w = (int)(idx / a.p.C_all), r = (int)(idx - (long long)w * a.p.C_all);
    const long long i0 = a.start[w], i1 = i0 + a.W;
    int flag = 0;
    for (long long b = i0 >> 5; b <= ((i1 - 1) >> 5) && !flag; ++b) {
        if (!a.p.nanblk[(size_t)r * a.p.nanblk_pitch + b]) continue;
        const long long lo = b << 5, hi = lo + 32;
        if (lo >= i0 && hi <= i1) {
            flag = 1;
        } else {
            const long long s0 = lo > i0 ? lo : i0, s1 = (hi < i1 ? hi : i1) < a.p.T ? (hi < i1 ? hi : i1) : a.p.T;
            for (long long t = s0; t < s1; ++t) {
                const double v = nm_raw_at(a.p, r, t);
                if (v != v) { flag = 1; break; }
            }
        }
    }
    a.flags[idx] = (unsigned char)flag;
    if (!flag) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int k = f.col_ptr[r]; k < f.col_ptr[r + 1]; ++k) f.out[(size_t)(f.row0 + w) * f.F + f.cols[k]] = qnan;
}
