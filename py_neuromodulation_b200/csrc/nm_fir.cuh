// nm_fir.cuh -- channels-batched zero-phase FIR (bank) by FFT convolution in shared memory.
//
// One CTA works on one (window, channel-pair) item at a time (persistent, grid-strided):
// the two real rows travel as the real / imaginary part of ONE complex transform -- the
// filters are real and symmetric about their centre tap, so their spectra are real and
//   inverse( forward(xA + i*xB) * H ) = (h*xA) + i*(h*xB)
// needs no spectrum un-mixing at all.  The forward transform is shared by all filters of a
// bank; each filter costs one real-by-complex multiply and one inverse transform, followed
// by a fused epilogue that consumes the filtered pair while it is still in shared memory
// (store rows / tail-variance band power / envelope / sharp-wave analysis).
//
// Two padding modes reproduce the reference exactly:
//   NM_FIR_SAME    scipy.signal.fftconvolve(mode="same") with zero padding
//                  (filter/mne_filter.py:110-116, features/sharpwaves.py:244-251)
//   NM_FIR_REFLECT mne _overlap_add_filter(pad="reflect_limited", phase="zero"): odd reflection
//                  about both end samples (filter/notch_filter.py:84-93)
#pragma once

#include "nm_common.cuh"

#define NM_FIR_SAME 0
#define NM_FIR_REFLECT 1

struct NmFirArgs {
    NmRows in;
    NmFft<double> fft;     // transform size P >= W + Lh (+ E for reflect)
    const double* hperm;   // [nF][P]: real spectrum of the centred taps, digit-reversed order, scaled by 1/P
    int nF;
    int mode;
    int E;                 // samples of odd-reflected extension on each side (reflect mode)
    int n_items;           // n_windows * ceil(n_ch / 2)
    int f0;                // index of the first filter of this launch inside its bank
};

// ---------------------------------------------------------------- epilogue: store filtered rows
struct NmEpiStore {
    double* y;             // (n_windows, n_ch, nF, Wp)
    long long Wp;
    int nF;
    static constexpr bool kRegs = true;   // nm_conv_kernel hands over the 16 outputs of each thread in registers
    static constexpr bool kRegsOnly = true, kReflectOk = false, kSameOk = true, kConvxOnly = false, kF32Ok = false, kSplitOk = false;  // nm_convx_kernel instantiation traits
    static NM_HD size_t smem_bytes(int /*nt*/) { return 0; }
    NM_DEV bool regs_ok() const { return true; }
    static constexpr bool kSyncsInside = false;
    struct State {};
    template <class PL, typename T>
    NM_DEV void consume(const cx<T>* v, cx<T>* /*work*/, double* /*red*/, State& /*st*/, int o0, int W, int n_ch, int w, int c0,
                        bool has2, int f, int tid) const {
        double* r0 = y + (((size_t)w * n_ch + c0) * nF + f) * Wp;
#pragma unroll
        for (int k = 0; k < PL::V0; ++k) {
            const int t = tid + PL::NT * k - o0;
            if (t >= 0 && t < W) {
                r0[t] = (double)v[k].re;
                if (has2) r0[(size_t)nF * Wp + t] = (double)v[k].im;
            }
        }
    }
    template <class PL, typename T>
    NM_DEV void finish(cx<T>* /*work*/, double* /*red*/, State& /*st*/, int /*o0*/, int /*W*/, int /*n_ch*/, int /*w*/, int /*c0*/,
                       bool /*has2*/, int /*f*/, int /*tid*/) const {}
    NM_DEV bool needs_trailing_barrier() const { return true; }
    // v[k] = filtered sample n = tid + nt*k of the (padded) row; the window occupies n in [o0, o0 + W)
    NM_DEV void run_regs(const cx<double>* v, int o0, int W, int n_ch, int w, int c0, bool has2, int f,
                         unsigned char* /*scratch*/, int tid, int nt) const {
        double* r0 = y + (((size_t)w * n_ch + c0) * nF + f) * Wp;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = tid + nt * k - o0;
            if (t >= 0 && t < W) {
                r0[t] = v[k].re;
                if (has2) r0[(size_t)nF * Wp + t] = v[k].im;
            }
        }
    }
    NM_DEV void run(const cx<double>* buf, int o0, int W, int n_ch, int w, int c0, bool has2, int f,
                    unsigned char* /*scratch*/, int tid, int nt) const {
        double* r0 = y + (((size_t)w * n_ch + c0) * nF + f) * Wp;
        for (int t = tid; t < W; t += nt) {
            const cx<double> v = buf[o0 + t];
            r0[t] = v.re;
            if (has2) r0[(size_t)nF * Wp + t] = v.im;
        }
    }
};

// ---------------------------------------------------------------- epilogue: band power (features/bandpower.py:165-207)
#define NM_BP_MAX_INLINE 16
struct NmEpiBandpower {
    const int* seglen;     // [nF] tail length in samples (device array, shared-memory epilogue)
    int seglen_k[NM_BP_MAX_INLINE];  // the same values in the kernel parameter space (register epilogue: no global load per band)
    int want_act, want_mob, want_comp, log_act;
    NmOut out;             // per_ch = nF * 3  (activity, mobility, complexity)
    static constexpr bool kRegs = true;
    static constexpr bool kRegsOnly = false, kReflectOk = false, kSameOk = true, kConvxOnly = false, kF32Ok = true, kSplitOk = false;
    static NM_HD size_t smem_bytes(int /*nt*/) { return 12 * 32 * sizeof(double); }
    // nm_convx_kernel register epilogue: tail moments from the registers, one barrier, then ONE warp (rotating with the
    // filter index) finishes the two channels on two lanes while the other warps already run the next filter.
    static constexpr bool kSyncsInside = true;
    struct State { double s[4]; int seg; };
    template <class PL, typename T>
    NM_DEV void consume(const cx<T>* v, cx<T>* /*work*/, double* /*red*/, State& st, int o0, int W, int /*n_ch*/, int /*w*/, int /*c0*/,
                        bool /*has2*/, int f, int tid) const {
        constexpr int NT = PL::NT;
        int seg = (f < NM_BP_MAX_INLINE) ? seglen_k[f] : nm_ldg(seglen + f);
        if (seg > W) seg = W;
        st.seg = seg;
        const int lo = o0 + W - seg, hi = o0 + W;
        // band-pass outputs have (near) zero mean, so the one-pass moments lose nothing in float64
        st.s[0] = st.s[1] = st.s[2] = st.s[3] = 0.0;
#pragma unroll
        for (int k = 0; k < PL::V0; ++k) {
            const int n = tid + NT * k;
            if (n >= lo && n < hi) {
                const double re = (double)v[k].re, im = (double)v[k].im;  // moments are accumulated in float64 in either mode
                st.s[0] += re; st.s[1] += re * re;
                st.s[2] += im; st.s[3] += im * im;
            }
        }
    }
    template <class PL, typename T>
    NM_DEV void finish(cx<T>* /*work*/, double* red, State& st, int /*o0*/, int /*W*/, int /*n_ch*/, int w, int c0, bool has2, int f,
                       int tid) const {
        constexpr int NW = (PL::NT + 31) / 32;
        const int lane = tid & 31, wid = tid >> 5;
        nm_warp_sum_multi<4>(st.s, lane);  // value i ends up in the lanes with lane >> 3 == i
        if ((lane & 7) == 0) red[(lane >> 3) * NW + wid] = st.s[0];
        __syncthreads();
        if (wid == (f & (NW - 1)) && lane < (has2 ? 2 : 1)) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                t0 += red[(2 * lane) * NW + q];
                t1 += red[(2 * lane + 1) * NW + q];
            }
            const double n0 = st.seg, mean = t0 / n0;
            double v0 = t1 / n0 - mean * mean;
            if (v0 < 0.0) v0 = 0.0;
            if (want_act) nm_store(out, w, c0 + lane, f * 3 + 0, nm_nan_to_num(log_act ? log10(v0) : v0));
        }
    }
    // the register path covers the default configuration (variance of the tail only); mobility / complexity need
    // neighbouring samples and go through the shared-memory row
    NM_DEV bool regs_ok() const { return !(want_mob || want_comp); }
    NM_DEV void run_regs(const cx<double>* v, int o0, int W, int /*n_ch*/, int w, int c0, bool has2, int f,
                         unsigned char* scratch, int tid, int nt) const {
        double* red = reinterpret_cast<double*>(scratch);
        int seg = nm_ldg(seglen + f);
        if (seg > W) seg = W;
        const int lo = o0 + W - seg, hi = o0 + W;
        // band-pass outputs have (near) zero mean, so the one-pass moments lose nothing in float64
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int n = tid + nt * k;
            if (n >= lo && n < hi) {
                s[0] += v[k].re; s[1] += v[k].re * v[k].re;
                s[2] += v[k].im; s[3] += v[k].im * v[k].im;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = nm_warp_sum(s[i]);
        const int lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) red[i * 32 + wid] = s[i];
        }
        __syncthreads();
        if (tid == 0) {
            double tot[4] = {0.0, 0.0, 0.0, 0.0};
            for (int q = 0; q < nw; ++q)
                for (int i = 0; i < 4; ++i) tot[i] += red[i * 32 + q];
            const double n0 = seg;
            for (int k = 0; k < (has2 ? 2 : 1); ++k) {
                const double mean = tot[2 * k] / n0;
                double v0 = tot[2 * k + 1] / n0 - mean * mean;
                if (v0 < 0.0) v0 = 0.0;
                if (want_act) nm_store(out, w, c0 + k, f * 3 + 0, nm_nan_to_num(log_act ? log10(v0) : v0));
            }
        }
    }
    NM_DEV void run(const cx<double>* buf, int o0, int W, int /*n_ch*/, int w, int c0, bool has2, int f,
                    unsigned char* scratch, int tid, int nt) const {
        double* red = reinterpret_cast<double*>(scratch);
        int seg = nm_ldg(seglen + f);
        if (seg > W) seg = W;
        const cx<double>* x = buf + o0 + (W - seg);
        const bool need_d = want_mob || want_comp;
        double s[6] = {0, 0, 0, 0, 0, 0};  // sum x, d1, d2 for the two channels
        for (int t = tid; t < seg; t += nt) {
            const cx<double> a = x[t];
            s[0] += a.re; s[3] += a.im;
            if (need_d && t + 1 < seg) {
                const cx<double> b = x[t + 1];
                s[1] += b.re - a.re; s[4] += b.im - a.im;
                if (t + 2 < seg) {
                    const cx<double> c = x[t + 2];
                    s[2] += (c.re - b.re) - (b.re - a.re);
                    s[5] += (c.im - b.im) - (b.im - a.im);
                }
            }
        }
        nm_block_sum<6>(s, red, tid, nt);
        const double n0 = seg, n1 = seg - 1, n2 = seg - 2;
        const double m0a = s[0] / n0, m1a = s[1] / n1, m2a = s[2] / n2;
        const double m0b = s[3] / n0, m1b = s[4] / n1, m2b = s[5] / n2;
        double q[6] = {0, 0, 0, 0, 0, 0};
        for (int t = tid; t < seg; t += nt) {
            const cx<double> a = x[t];
            double d = a.re - m0a; q[0] += d * d;
            d = a.im - m0b; q[3] += d * d;
            if (need_d && t + 1 < seg) {
                const cx<double> b = x[t + 1];
                d = (b.re - a.re) - m1a; q[1] += d * d;
                d = (b.im - a.im) - m1b; q[4] += d * d;
                if (t + 2 < seg) {
                    const cx<double> c = x[t + 2];
                    d = ((c.re - b.re) - (b.re - a.re)) - m2a; q[2] += d * d;
                    d = ((c.im - b.im) - (b.im - a.im)) - m2b; q[5] += d * d;
                }
            }
        }
        nm_block_sum<6>(q, red + 6 * 32, tid, nt);
        if (tid == 0) {
            for (int k = 0; k < (has2 ? 2 : 1); ++k) {
                const double v0 = q[3 * k] / n0, v1 = q[3 * k + 1] / n1, v2 = q[3 * k + 2] / n2;
                const int c = c0 + k;
                if (want_act) nm_store(out, w, c, f * 3 + 0, nm_nan_to_num(log_act ? log10(v0) : v0));
                if (want_mob) nm_store(out, w, c, f * 3 + 1, nm_nan_to_num(sqrt(v1 / v0)));
                if (want_comp) nm_store(out, w, c, f * 3 + 2, nm_nan_to_num(sqrt(v2 / v1) / sqrt(v1 / v0)));
            }
        }
    }
};

// ---------------------------------------------------------------- the kernel
template <class Epi>
NM_GLOBAL void nm_fir_kernel(NmFirArgs a, Epi epi) {
    NM_SHARED_BYTES(smem);
    const int P = a.fft.n;
    cx<double>* spec = reinterpret_cast<cx<double>*>(smem);
    cx<double>* work = (a.nF > 1) ? spec + P : spec;
    unsigned char* scratch = reinterpret_cast<unsigned char*>(spec + (a.nF > 1 ? 2 : 1) * (size_t)P);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int W = a.in.W, E = a.E;
    const int npair = (a.in.n_ch + 1) >> 1;

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item / npair;
        const int c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        const double* r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
        const double* r1 = r0 + (has2 ? a.in.ch_stride : 0);

        if (a.mode == NM_FIR_REFLECT) {
            const double a0 = 2.0 * r0[0], b0 = 2.0 * r1[0], a1 = 2.0 * r0[W - 1], b1 = 2.0 * r1[W - 1];
            for (int n = tid; n < P; n += nt) {
                double va = 0.0, vb = 0.0;
                if (n < E) {
                    va = a0 - r0[E - n]; vb = b0 - r1[E - n];
                } else if (n < E + W) {
                    va = r0[n - E]; vb = r1[n - E];
                } else if (n < W + 2 * E) {
                    const int k = n - (E + W) + 1;
                    va = a1 - r0[W - 1 - k]; vb = b1 - r1[W - 1 - k];
                }
                spec[n] = {va, has2 ? vb : 0.0};
            }
        } else {
            for (int n = tid; n < P; n += nt) {
                double va = 0.0, vb = 0.0;
                if (n < W) { va = r0[n]; vb = r1[n]; }
                spec[n] = {va, has2 ? vb : 0.0};
            }
        }
        __syncthreads();
        nm_fft_forward<double>(spec, nullptr, a.fft, tid, nt);

        const int o0 = (a.mode == NM_FIR_REFLECT) ? E : 0;
        for (int f = 0; f < a.nF; ++f) {
            const double* h = a.hperm + (size_t)f * P;
            for (int n = tid; n < P; n += nt) {
                const double hv = nm_ldg(h + n);
                const cx<double> v = spec[n];
                work[n] = {v.re * hv, v.im * hv};
            }
            __syncthreads();
            nm_fft_inverse<double>(work, nullptr, a.fft, tid, nt);
            epi.run(work, o0, W, a.in.n_ch, w, c0, has2, f + a.f0, scratch, tid, nt);
            __syncthreads();
        }
    }
}
