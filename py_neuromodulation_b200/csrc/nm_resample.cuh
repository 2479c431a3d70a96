// nm_resample.cuh -- raw_resampling (reference: processing/resample.py:28-60 -> mne.filter.resample(x, up=ratio, down=1)).
//
// FFT resampling of one window row is a LINEAR map of its samples (reflect-limited padding, forward transform, spectrum cut or
// zero-extended with the Nyquist correction, inverse transform of the new length, padding removed).  The host designs that map
// once per window length as the dense operator R (n_out x n_in) -- exactly as it designs FIR taps -- and every (window, channel)
// row becomes  y = R x :  one float64 GEMM per chunk of windows,
//
//     Y[r, j] = sum_k  X[r, k] * Rt[k, j]        r = (window, channel) row, k < n_in, j < n_out,   Rt = R^T (k-major)
//
// placed where the reference runs it: after the notch, before the (hoisted) re-reference's consumers -- the re-reference is a
// per-sample combination of channels and commutes with a per-row linear map that is the same for all channels.  Works for any
// ratio (2 kHz -> 1 kHz, 1111.111 Hz -> 1 kHz, up-sampling).
//
// Fast path (integer down-sampling, the default 2 kHz -> 1 kHz case): with ratio 1/D the new spectrum is the old one cut at the new
// Nyquist bin, so  y[m] = x_lp[D m]  where x_lp is the padded window after an ideal zero-phase low-pass H[k] = 1 for
// |k| <= P/(2D) (both +-Nyquist bins kept: their sum is MNE's doubled real Nyquist bin).  That is the notch kernel's own shape --
// reflect-limited padding to P points, forward transform, * real symmetric spectrum, inverse -- with an epilogue that stores
// every D-th sample of the window region (NmEpiStoreDecim on nm_convx_kernel<P, REFLECT>): two transforms per channel pair
// instead of 2 * n_in * n_out multiply-adds per row.  Conditions (nm_resample_fast_ok): padded length P in {1024, 2048, 4096},
// even window, pads divisible by D.  Everything else (up-sampling, ratio 0.8 / 0.9, odd windows) uses the GEMM below.
//
// GEMM kernel: 128 x 64 output tile per CTA, 256 threads, 8 x 4 register tile per thread, K in steps of 16 through shared memory
// (A tile stored k-major so that both operands are read with 128-bit shared loads).  FP64-pipe bound by construction:
// 32 DFMA per 6 LDS.128 per k.
#pragma once

#include "nm_common.cuh"
#include "nm_convx.cuh"

// ---------------------------------------------------------------- fast path: decimating store on the FFT-convolution kernel
struct NmEpiStoreDecim {
    double* y;        // (n_windows, n_ch, Wp) resampled rows
    long long Wp;
    int D;            // keep window samples t = 0, D, 2D, ...
    static constexpr bool kRegs = true;
    static constexpr bool kRegsOnly = true, kReflectOk = true, kSameOk = false, kConvxOnly = true, kF32Ok = false, kSplitOk = false;
    static constexpr bool kSyncsInside = false;
    static NM_HD size_t smem_bytes(int /*nt*/) { return 0; }
    NM_DEV bool regs_ok() const { return true; }
    struct State {};
    template <class PL, typename T>
    NM_DEV void consume(const cx<T>* v, cx<T>* /*work*/, double* /*red*/, State& /*st*/, int o0, int W, int n_ch, int w, int c0,
                        bool has2, int /*f*/, int tid) const {
        double* r0 = y + ((size_t)w * n_ch + c0) * Wp;
#pragma unroll
        for (int k = 0; k < PL::V0; ++k) {
            const int t = tid + PL::NT * k - o0;
            if (t >= 0 && t < W && t % D == 0) {
                r0[t / D] = (double)v[k].re;
                if (has2) r0[Wp + t / D] = (double)v[k].im;
            }
        }
    }
    template <class PL, typename T>
    NM_DEV void finish(cx<T>* /*work*/, double* /*red*/, State& /*st*/, int /*o0*/, int /*W*/, int /*n_ch*/, int /*w*/, int /*c0*/,
                       bool /*has2*/, int /*f*/, int /*tid*/) const {}
    NM_DEV bool needs_trailing_barrier() const { return true; }
};

#define NM_RS_BM 128
#define NM_RS_BN 64
#define NM_RS_BK 16
#define NM_RS_THREADS 256

struct NmResampleArgs {
    NmRows in;          // rows of n_in samples (in.W == n_in)
    const double* rt;   // (n_in, n_out_pitch) row-major: R transposed, rows padded with zeros to a multiple of 4
    int n_out, n_out_pitch;
    double* out;        // (n_windows, n_ch, out_pitch)
    long long out_pitch;
};

static NM_HD size_t nm_resample_smem_bytes() { return sizeof(double) * NM_RS_BK * (NM_RS_BM + 4 + NM_RS_BN); }

NM_GLOBAL void NM_LAUNCH_BOUNDS(NM_RS_THREADS, 2) nm_resample_kernel(NmResampleArgs a) {
    NM_SHARED_BYTES(smem);
    typedef double ARow[NM_RS_BM + 4];  // k-major, padded: the transposing stores spread over banks
    typedef double BRow[NM_RS_BN];
    ARow* As = reinterpret_cast<ARow*>(smem);
    BRow* Bs = reinterpret_cast<BRow*>(smem + sizeof(ARow) * NM_RS_BK);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // thread tile: rows ty*8.., columns tx*4..
    const long long n_rows = (long long)a.in.n_windows * a.in.n_ch;
    const long long r0 = (long long)blockIdx.x * NM_RS_BM;
    const int j0 = blockIdx.y * NM_RS_BN;
    const int K = a.in.W;

    // loader roles.  A: thread loads rows (tid >> 1) and (tid >> 1) + ... : 128 rows x 16 k = 2048 values, 8 per thread:
    // row = tid >> 1, k = (tid & 1) * 8 .. + 8 (contiguous in memory).  B: 16 x 64 = 1024 values, 4 per thread.
    const int a_row = tid >> 1, a_k0 = (tid & 1) * 8;
    const long long ar = r0 + a_row;
    const double* a_src = nullptr;
    if (ar < n_rows) {
        const int w = (int)(ar / a.in.n_ch), c = (int)(ar - (long long)w * a.in.n_ch);
        a_src = a.in.base + (size_t)c * a.in.ch_stride + nm_ldg(a.in.off + w);
    }
    const int b_k = tid >> 4, b_j = (tid & 15) * 4;

    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int k0 = 0; k0 < K; k0 += NM_RS_BK) {
        double av[8], bv[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + a_k0 + i;
            av[i] = (a_src && k < K) ? a_src[k] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + b_k, jj = j0 + b_j + j;
            bv[j] = (k < K && jj < a.n_out_pitch) ? nm_ldg(a.rt + (size_t)k * a.n_out_pitch + jj) : 0.0;
        }
        __syncthreads();  // previous tile fully consumed
#pragma unroll
        for (int i = 0; i < 8; ++i) As[a_k0 + i][a_row] = av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[b_k][b_j + j] = bv[j];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NM_RS_BK; ++k) {
            double ra[8], rb[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) ra[i] = As[k][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) rb[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(ra[i], rb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long r = r0 + ty * 8 + i;
        if (r >= n_rows) continue;
        double* dst = a.out + (size_t)r * a.out_pitch;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = j0 + tx * 4 + j;
            if (jj < a.n_out) dst[jj] = acc[i][j];
        }
    }
}
