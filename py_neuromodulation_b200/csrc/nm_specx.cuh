// nm_specx.cuh -- register-blocked segment DFT for the window lengths the oscillatory plug-ins actually use.
//
// Same contract as nm_spec_kernel (nm_spec.cuh: FFT / Welch / STFT band features of one (window, channel pair) item per
// CTA iteration) but the N-point transform is a compile-time three-pass plan N = R0 * R1 * R2 with R0 values per thread:
//
//   N = 1000 = 10 * 10 * 10     (FFT / Welch at 1 kHz)          100 active threads
//   N = 2000 = 20 * 10 * 10     (FFT / Welch at 2 kHz)          100 active threads
//   N =  500 = 10 * 10 *  5     (STFT, nperseg = 500)            50 active threads
//
// The radix-10 / radix-20 butterflies are prime-factor (Good-Thomas) 2x5 / 4x5 transforms -- no internal twiddles.  Pass 0
// runs on samples fetched straight from global memory (mean removal and the taper are applied in registers), so the
// shared buffer is crossed twice instead of once per radix-{4,2,5,5,5} pass of the generic kernel.  Slot order of the
// spectrum is the FftPlanHost digit-reversed order for radices (R0, R1, R2), read through `pos` like in the generic kernel.
#pragma once

#include "nm_spec.cuh"

// ---- in-register DFTs, natural order in and out (forward only)
template <int A>
NM_DEV void nm_dft_small(cx<double>* v) {
    if (A == 2) nm_bfly2<double, false>(v);
    if (A == 4) {
        nm_bfly4<double, false>(v);
    }
    if (A == 5) nm_bfly5<double, false>(v);
}

// R = A * 5 with gcd(A, 5) = 1:  n = (5*n1 + A*n2) mod R,  k = (5*k1*(5^-1 mod A) + A*k2*(A^-1 mod 5)) mod R
template <int R>
NM_DEV void nm_dft_pfa(cx<double>* v) {
    constexpr int A = R / 5;
    constexpr int INV5 = 1;                       // 5^-1 mod 2 = 5^-1 mod 4 = 1
    constexpr int INVA = (A == 2) ? 3 : 4;        // 2^-1 mod 5 = 3, 4^-1 mod 5 = 4
    cx<double> x[A][5];
#pragma unroll
    for (int n1 = 0; n1 < A; ++n1)
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) x[n1][n2] = v[(5 * n1 + A * n2) % R];
#pragma unroll
    for (int n1 = 0; n1 < A; ++n1) nm_bfly5<double, false>(x[n1]);   // over n2 -> k2
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) {
        cx<double> c[A];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) c[n1] = x[n1][k2];
        nm_dft_small<A>(c);                                           // over n1 -> k1
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) v[(5 * k1 * INV5 + A * INVA * k2) % R] = c[k1];
    }
}

template <int R>
NM_DEV void nm_dft_reg(cx<double>* v) {
    if (R == 5) nm_bfly5<double, false>(v);
    if (R == 10 || R == 20) nm_dft_pfa<R>(v);
}

// v[k] *= w1^k, k = 1..R-1, powers built with logarithmic dependency depth
template <int R>
NM_DEV void nm_twiddle_pow(cx<double>* v, cx<double> w1) {
    cx<double> w[R];
    w[1] = w1;
#pragma unroll
    for (int k = 2; k < R; ++k) w[k] = cx_mul(w[k >> 1], w[k - (k >> 1)]);
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cx_mul(v[k], w[k]);
}

template <int N_, int R0_, int R1_, int R2_>
struct NmSxPlan {
    static constexpr int N = N_, R0 = R0_, R1 = R1_, R2 = R2_;
    static constexpr int NA = N / R0;                 // active threads (one pass-0 butterfly each)
    static constexpr int NT = (NA + 31) / 32 * 32;    // launched threads (whole warps for the reductions)
    static constexpr int L1 = N / R0, M1 = L1 / R1;
    static constexpr int PADU = (R2 % 2 == 0) ? R2 : 0;  // slot e lives at e + e / PADU: unit-stride pass conflict free
    static constexpr int NBUF = N + (PADU ? N / PADU : 0) + 2;
    static_assert(R0 * R1 * R2 == N && R0 % R1 == 0 && R0 % R2 == 0, "three-pass plan, R1 | R0 and R2 | R0");
    static NM_HD int phys(int e) { return PADU ? e + e / PADU : e; }
};

// inverse DFT in registers through the conjugation identity  IDFT(v) = conj(DFT(conj(v)))  (unnormalised)
template <int R>
NM_DEV void nm_dft_reg_inv(cx<double>* v) {
#pragma unroll
    for (int k = 0; k < R; ++k) v[k].im = -v[k].im;
    nm_dft_reg<R>(v);
#pragma unroll
    for (int k = 0; k < R; ++k) v[k].im = -v[k].im;
}

// v[k] *= conj(w1)^k, k = 1..R-1
template <int R>
NM_DEV void nm_twiddle_pow_conj(cx<double>* v, cx<double> w1) {
    w1.im = -w1.im;
    nm_twiddle_pow<R>(v, w1);
}

// Forward transform of a natural-order row that already sits in shared memory (`src`, N elements) into `buf` (padded slot
// order, PL::NBUF elements); `src` and `buf` must not overlap.  All threads of the CTA call it; ends with a barrier.
template <class PL>
NM_DEV void nm_sx_forward_smem(const cx<double>* src, cx<double>* buf, const cx<double>* NM_RESTRICT tw, int tid) {
    constexpr int R0 = PL::R0, R1 = PL::R1, R2 = PL::R2, NA = PL::NA;
    const bool active = tid < NA;
    if (active) {
        cx<double> v[R0];
#pragma unroll
        for (int t = 0; t < R0; ++t) v[t] = src[tid + NA * t];
        nm_dft_reg<R0>(v);
        nm_twiddle_pow<R0>(v, nm_ldg(tw + tid));
#pragma unroll
        for (int k = 0; k < R0; ++k) buf[PL::phys(tid + NA * k)] = v[k];
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R1; ++i) {
            const int q = tid + NA * i;
            const int blk = q / PL::M1, j = q - blk * PL::M1;
            const int e0 = blk * PL::L1 + j;
            cx<double> u[R1];
#pragma unroll
            for (int t = 0; t < R1; ++t) u[t] = buf[PL::phys(e0 + t * PL::M1)];
            nm_dft_reg<R1>(u);
            nm_twiddle_pow<R1>(u, nm_ldg(tw + j * R0));
#pragma unroll
            for (int t = 0; t < R1; ++t) buf[PL::phys(e0 + t * PL::M1)] = u[t];
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R2; ++i) {
            const int e0 = (tid + NA * i) * R2;
            cx<double> u[R2];
#pragma unroll
            for (int t = 0; t < R2; ++t) u[t] = buf[PL::phys(e0 + t)];
            nm_dft_reg<R2>(u);
#pragma unroll
            for (int t = 0; t < R2; ++t) buf[PL::phys(e0 + t)] = u[t];
        }
    }
    __syncthreads();
}

// frequency index held by slot e of the padded buffer after nm_sx_forward_smem (digit-reversed order of radices R0, R1, R2)
template <class PL>
NM_DEV int nm_sx_freq_of_slot(int e) {
    const int k0 = e / PL::L1, r = e - k0 * PL::L1;
    const int k1 = r / PL::R2, k2 = r - k1 * PL::R2;
    return k0 + PL::R0 * k1 + PL::R0 * PL::R1 * k2;
}

// Unnormalised inverse of nm_sx_forward_smem: slot-ordered spectrum in `buf` -> natural-order samples n = tid + NA * t left in
// the registers v[t] of the active threads (tid < NA).  Starts after a barrier of the caller, contains two barriers.
template <class PL>
NM_DEV void nm_sx_inverse_regs(cx<double>* buf, const cx<double>* NM_RESTRICT tw, int tid, cx<double>* v) {
    constexpr int R0 = PL::R0, R1 = PL::R1, R2 = PL::R2, NA = PL::NA;
    const bool active = tid < NA;
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R2; ++i) {
            const int e0 = (tid + NA * i) * R2;
            cx<double> u[R2];
#pragma unroll
            for (int t = 0; t < R2; ++t) u[t] = buf[PL::phys(e0 + t)];
            nm_dft_reg_inv<R2>(u);
#pragma unroll
            for (int t = 0; t < R2; ++t) buf[PL::phys(e0 + t)] = u[t];
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R1; ++i) {
            const int q = tid + NA * i;
            const int blk = q / PL::M1, j = q - blk * PL::M1;
            const int e0 = blk * PL::L1 + j;
            cx<double> u[R1];
#pragma unroll
            for (int t = 0; t < R1; ++t) u[t] = buf[PL::phys(e0 + t * PL::M1)];
            nm_twiddle_pow_conj<R1>(u, nm_ldg(tw + j * R0));
            nm_dft_reg_inv<R1>(u);
#pragma unroll
            for (int t = 0; t < R1; ++t) buf[PL::phys(e0 + t * PL::M1)] = u[t];
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int k = 0; k < R0; ++k) v[k] = buf[PL::phys(tid + NA * k)];
        nm_twiddle_pow_conj<R0>(v, nm_ldg(tw + tid));
        nm_dft_reg_inv<R0>(v);
    }
}

static NM_HD size_t nm_specx_smem_bytes(int nbuf, int nk, int nsegv) {
    return (size_t)nbuf * sizeof(cx<double>) + (size_t)2 * nk * nsegv * sizeof(double) + 2 * 32 * sizeof(double);
}

template <class PL>
NM_GLOBAL void NM_LAUNCH_BOUNDS(PL::NT, 4) nm_specx_kernel(NmSpecArgs a) {
    constexpr int N = PL::N, R0 = PL::R0, R1 = PL::R1, R2 = PL::R2, NA = PL::NA, NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<double>* buf = reinterpret_cast<cx<double>*>(smem);
    const int nsegv = a.keep_segments ? a.nseg : 1;
    double* vals = reinterpret_cast<double*>(buf + PL::NBUF);  // [2][nk][nsegv]
    double* red = vals + (size_t)2 * a.nk * nsegv;
    const int tid = threadIdx.x;
    const bool active = tid < NA;
    const int W = a.in.W;
    const int npair = (a.in.n_ch + 1) >> 1;
    const cx<double>* NM_RESTRICT tw = a.fft.tw;
    const cx<double> w0 = nm_ldg(tw + (active ? tid : 0));  // exp(-2*pi*i*tid/N): pass-0 twiddle generator

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item / npair;
        const int c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        const double* NM_RESTRICT r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
        const double* NM_RESTRICT r1 = r0 + (has2 ? a.in.ch_stride : 0);

        if (!a.keep_segments)
            for (int i = tid; i < 2 * a.nk; i += NT) vals[i] = 0.0;

        for (int s = 0; s < a.nseg; ++s) {
            const int base = a.start + s * a.hop;
            cx<double> v[R0];
            double sum[2] = {0.0, 0.0};
            if (active) {
                if (base >= 0 && base + N <= W) {  // segment inside the window (FFT, Welch): plain coalesced loads
#pragma unroll
                    for (int t = 0; t < R0; ++t) {
                        const int n = base + tid + NA * t;
                        v[t].re = r0[n];
                        v[t].im = has2 ? r1[n] : 0.0;
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < R0; ++t) {
                        const int n = tid + NA * t;
                        const double va = nm_spec_sample(r0, base + n, W, a.ext_even, a.ext_len);
                        const double vb = has2 ? nm_spec_sample(r1, base + n, W, a.ext_even, a.ext_len) : 0.0;
                        v[t] = {va, vb};
                    }
                }
#pragma unroll
                for (int t = 0; t < R0; ++t) {
                    sum[0] += v[t].re;
                    sum[1] += v[t].im;
                }
            }
            if (a.detrend) {
                nm_block_sum<2>(sum, red, tid, NT);  // (also orders this segment after the previous one's bin reads)
                sum[0] /= N;
                sum[1] /= N;
            } else {
                sum[0] = sum[1] = 0.0;
                __syncthreads();
            }
            if (active) {
                if (a.detrend || a.win) {
#pragma unroll
                    for (int t = 0; t < R0; ++t) {
                        const double wv = a.win ? nm_ldg(a.win + tid + NA * t) : 1.0;
                        v[t] = {(v[t].re - sum[0]) * wv, (v[t].im - sum[1]) * wv};
                    }
                }
                // ---- pass 0: radix R0 over stride NA, twiddle w_N^(tid*k)
                nm_dft_reg<R0>(v);
                nm_twiddle_pow<R0>(v, w0);
#pragma unroll
                for (int k = 0; k < R0; ++k) buf[PL::phys(tid + NA * k)] = v[k];
            }
            __syncthreads();
            if (active) {
                // ---- pass 1: R0/R1 butterflies of radix R1 inside blocks of L1 = N/R0, twiddle w_L1^(j*k)
#pragma unroll
                for (int i = 0; i < R0 / R1; ++i) {
                    const int q = tid + NA * i;
                    const int blk = q / PL::M1, j = q - blk * PL::M1;
                    const int e0 = blk * PL::L1 + j;
                    cx<double> u[R1];
#pragma unroll
                    for (int t = 0; t < R1; ++t) u[t] = buf[PL::phys(e0 + t * PL::M1)];
                    nm_dft_reg<R1>(u);
                    nm_twiddle_pow<R1>(u, nm_ldg(tw + j * R0));
#pragma unroll
                    for (int t = 0; t < R1; ++t) buf[PL::phys(e0 + t * PL::M1)] = u[t];
                }
            }
            __syncthreads();
            if (active) {
                // ---- pass 2: R0/R2 butterflies of radix R2, unit stride, no twiddle
#pragma unroll
                for (int i = 0; i < R0 / R2; ++i) {
                    const int e0 = (tid + NA * i) * R2;
                    cx<double> u[R2];
#pragma unroll
                    for (int t = 0; t < R2; ++t) u[t] = buf[PL::phys(e0 + t)];
                    nm_dft_reg<R2>(u);
#pragma unroll
                    for (int t = 0; t < R2; ++t) buf[PL::phys(e0 + t)] = u[t];
                }
            }
            __syncthreads();
            for (int i = tid; i < a.nk; i += NT) {
                const int k = a.k0 + i;
                const cx<double> U = buf[PL::phys(nm_ldg(a.fft.pos + k))];
                const cx<double> V = buf[PL::phys(nm_ldg(a.fft.pos + (k == 0 ? 0 : N - k)))];
                nm_spec_bin(a, vals, i, k, s, U, V);
            }
            __syncthreads();
        }
        nm_spec_finish(a, vals, nsegv, w, c0, has2, tid, NT);
        __syncthreads();
    }
}

using NmSx1000 = NmSxPlan<1000, 10, 10, 10>;
using NmSx2000 = NmSxPlan<2000, 20, 10, 10>;
using NmSx500 = NmSxPlan<500, 10, 10, 5>;

static inline bool nm_specx_supported(int n) { return n == 1000 || n == 2000 || n == 500; }
