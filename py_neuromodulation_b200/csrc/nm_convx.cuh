// nm_convx.cuh -- compile-time specialised FFT-convolution kernel for the transform sizes the path actually uses.
//
// Same algorithm, data layout and epilogue contract as nm_conv_kernel (nm_conv.cuh): one (window, channel pair) item
// per CTA iteration, the two real rows packed as one complex signal, P = 16 * R1 * R2 points in three register-blocked
// passes (16 values per thread), real symmetric filter spectra in digit-reversed slot order, padded shared memory.
// What changes is that P, the radices, the padding and the filter-bank / padding mode are TEMPLATE parameters:
//
//   * every shared-memory access of a pass is `one per-thread base + immediate offset` (the padded index
//     e + (e >> PAD) is additive over the strides because all strides are multiples of 2^PAD), so the integer
//     instruction stream of nm_conv_kernel (IMAD/LEA/SHF, ~40 % of its issue slots in SASS) disappears;
//   * the pass-1 twiddles depend on tid & (M1-1) only and are shared by the 16/R1 butterflies of a thread;
//   * no runtime radix dispatch, no mode / bank branches inside the item loop -> half the code size (i-cache);
//   * the epilogue may reuse `work` as a natural-order row buffer (a barrier separates it from the last pass), which
//     lets the notch kernel compute the Hjorth / line-length / raw features of the window it has just filtered.
//
// P in {1024, 2048, 4096} covers every FIR of the default 1 kHz / 2 kHz configurations (999/1999-tap notch and
// band-pass banks, 1651/3301-tap sharp-wave filters); other sizes keep using nm_conv_kernel / nm_fir_kernel.
#pragma once

#include "nm_conv.cuh"

template <int P_>
struct NmCxPlan {
    static constexpr int P = P_;
    static constexpr int NT = P / 16;
    static constexpr int R1 = (P == 1024) ? 8 : 16;
    static constexpr int R2 = P / 16 / R1;
    static constexpr int PAD = (R2 == 16) ? 4 : 3;
    static constexpr int M1 = NT / R1;                        // butterfly stride of pass 1
    static constexpr int S0 = NT + (NT >> PAD);               // physical strides (see header comment)
    static constexpr int S1 = M1 + (M1 >> PAD);
    static constexpr int B1 = NT * R1 + ((NT * R1) >> PAD);
    static constexpr int B2 = NT * R2 + ((NT * R2) >> PAD);
    static constexpr int NBUF = P + (P >> PAD) + 2;
    static constexpr int MINB = (NT >= 256) ? 2 : (NT == 128 ? 4 : 8);  // 128 registers per thread
    static_assert(R2 == (1 << PAD), "padding unit must equal the last radix");
    static_assert(M1 % R2 == 0 && NT % R2 == 0, "strides must be multiples of the padding unit");
    static_assert(R1 * R2 * 16 == P, "three-pass plan");
};

#define NM_CX_RED_BYTES 256  // 4 values x 8 warps of doubles, owned by the register epilogues

static inline bool nm_convx_supported(int P) { return P == 1024 || P == 2048 || P == 4096; }

// multiply the R-1 upper values of a butterfly by the powers of w1 (conjugated for the inverse)
template <int R, bool INV>
NM_DEV void nm_twiddle_w1(cx<double>* v, cx<double> w1) {
    if (INV) w1.im = -w1.im;
    v[1] = cx_mul(v[1], w1);
    const cx<double> w2 = cx_mul(w1, w1), w3 = cx_mul(w2, w1);
    v[2] = cx_mul(v[2], w2);
    v[3] = cx_mul(v[3], w3);
    const cx<double> w4 = cx_mul(w2, w2);
    v[4] = cx_mul(v[4], w4);
    v[5] = cx_mul(v[5], cx_mul(w4, w1));
    v[6] = cx_mul(v[6], cx_mul(w4, w2));
    v[7] = cx_mul(v[7], cx_mul(w4, w3));
    if (R > 8) {
        const cx<double> w8 = cx_mul(w4, w4);
        v[8] = cx_mul(v[8], w8);
        v[9] = cx_mul(v[9], cx_mul(w8, w1));
        v[10] = cx_mul(v[10], cx_mul(w8, w2));
        v[11] = cx_mul(v[11], cx_mul(w8, w3));
        const cx<double> w12 = cx_mul(w8, w4);
        v[12] = cx_mul(v[12], w12);
        v[13] = cx_mul(v[13], cx_mul(w12, w1));
        v[14] = cx_mul(v[14], cx_mul(w12, w2));
        v[15] = cx_mul(v[15], cx_mul(w12, w3));
    }
}

// interior pass (pass 1): 16/R1 butterflies per thread, all sharing the same twiddle set
template <class PL, bool INV>
NM_DEV void nm_cx_pass1(cx<double>* sm, const cx<double> w1, int tid) {
    constexpr int R = PL::R1;
    const int j = tid & (PL::M1 - 1);
    const int base = (tid - j) * R + j;
    cx<double>* p = sm + base + (base >> PL::PAD);
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        cx<double> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = p[i * PL::B1 + t * PL::S1];
        if (INV) {
            if (j != 0) nm_twiddle_w1<R, true>(v, w1);
            nm_bflyR<R, true>(v);
        } else {
            nm_bflyR<R, false>(v);
            if (j != 0) nm_twiddle_w1<R, false>(v, w1);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) p[i * PL::B1 + t * PL::S1] = v[t];
    }
}

// last forward pass (unit stride).  MODE 0: forward butterfly only (bank: spectrum stays in `src`).
// MODE 1: forward butterfly, * H, inverse butterfly in place (single filter).
// MODE 2: load the spectrum from `src`, * H, inverse butterfly, store to `dst` (one filter of a bank).
template <class PL, int MODE>
NM_DEV void nm_cx_pass2(cx<double>* dst, const cx<double>* src, const double* NM_RESTRICT h, int tid) {
    constexpr int R = PL::R2;
    const int off = tid * (R + 1);  // tid*R + ((tid*R) >> PAD)
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        cx<double> v[R];
        double hv[R];
        if (MODE != 0) {
#pragma unroll
            for (int t = 0; t < R; ++t) hv[t] = nm_ldg(h + (tid + PL::NT * i) * R + t);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[off + i * PL::B2 + t];
        if (MODE != 2) nm_bflyR<R, false>(v);
        if (MODE != 0) {
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = {v[t].re * hv[t], v[t].im * hv[t]};
            nm_bflyR<R, true>(v);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) dst[off + i * PL::B2 + t] = v[t];
    }
}

// compact CTA-wide sum of NV values; `red` needs NV * (NT/32) doubles; result valid in every thread.  Two barriers.
template <int NV, int NW>
NM_DEV void nm_cx_block_sum(double* v, double* red, int tid) {
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = nm_warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * NW + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[i * NW + w];
        v[i] = s;
    }
}

// ---------------------------------------------------------------- epilogue: store rows (+ fused scan features)
// The notch kernel's epilogue: writes the filtered window to the chunk buffer for the other families and, when the
// scan family is enabled, computes Hjorth activity / mobility / complexity, line length and the last sample
// (features/hjorth_raw.py:24-42,51-57, features/linelength.py:11-21) from the rows while they are still on chip --
// two-pass moments like numpy.var.  Replaces nm_scan_kernel's re-read of the chunk.
struct NmEpiStoreScan {
    double* y;  // (n_windows, n_ch, Wp) or nullptr when no other family consumes the rows
    long long Wp;
    int want_hjorth, want_raw, want_ll, want_scan;
    NmOut out;  // per_ch = 5: activity, mobility, complexity, raw, linelength
    static constexpr bool kRegs = true;
    static constexpr bool kRegsOnly = true, kReflectOk = true, kSameOk = false, kConvxOnly = true;
    static NM_HD size_t smem_bytes(int /*nt*/) { return 0; }  // reductions use the padding tail of `work`
    NM_DEV bool regs_ok() const { return true; }
    static constexpr bool kSyncsInside = false;  // (not on every path) -> the kernel adds the trailing barrier

    template <class PL>
    NM_DEV void run_x(const cx<double>* v, cx<double>* work, double* /*red*/, int o0, int W, int n_ch, int w, int c0, bool has2, int /*f*/,
                      int tid) const {
        constexpr int NT = PL::NT, NW = (PL::NT + 31) / 32;
        static_assert((size_t)14 * NW * sizeof(double) <= (size_t)(PL::NBUF - PL::P) * sizeof(cx<double>), "reduction scratch must fit the tail");
        if (y) {
            double* r0 = y + ((size_t)w * n_ch + c0) * Wp;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int t = tid + NT * k - o0;
                if (t >= 0 && t < W) {
                    r0[t] = v[k].re;
                    if (has2) r0[Wp + t] = v[k].im;
                }
            }
        }
        if (!want_scan) return;
        double* red = reinterpret_cast<double*>(work + PL::P);
        // natural-order copy so that every thread can see the two samples after each of its own
        __syncthreads();  // every thread has read its pass-0 inputs from `work`
#pragma unroll
        for (int k = 0; k < 16; ++k) work[tid + NT * k] = v[k];
        __syncthreads();
        const cx<double>* x = work + o0;
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // per row: sum x, sum d, sum dd, sum |d|
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int t = tid + NT * k - o0;
            if (t >= 0 && t < W) {
                s[0] += v[k].re; s[4] += v[k].im;
                if (t + 1 < W) {
                    const cx<double> b = x[t + 1];
                    const double da = b.re - v[k].re, db = b.im - v[k].im;
                    s[1] += da; s[5] += db;
                    s[3] += fabs(da); s[7] += fabs(db);
                    if (t + 2 < W) {
                        const cx<double> c = x[t + 2];
                        s[2] += (c.re - b.re) - da;
                        s[6] += (c.im - b.im) - db;
                    }
                }
            }
        }
        nm_cx_block_sum<8, NW>(s, red, tid);
        const double n0 = W, n1 = W - 1, n2 = W - 2;
        double q[6] = {0, 0, 0, 0, 0, 0};
        if (want_hjorth) {
            const double m0a = s[0] / n0, m1a = s[1] / n1, m2a = s[2] / n2;
            const double m0b = s[4] / n0, m1b = s[5] / n1, m2b = s[6] / n2;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int t = tid + NT * k - o0;
                if (t >= 0 && t < W) {
                    const cx<double> a = x[t];  // re-read instead of keeping v[] live across the reduction (registers)
                    double e = a.re - m0a; q[0] += e * e;
                    e = a.im - m0b; q[3] += e * e;
                    if (t + 1 < W) {
                        const cx<double> b = x[t + 1];
                        const double da = b.re - a.re, db = b.im - a.im;
                        e = da - m1a; q[1] += e * e;
                        e = db - m1b; q[4] += e * e;
                        if (t + 2 < W) {
                            const cx<double> c = x[t + 2];
                            e = ((c.re - b.re) - da) - m2a; q[2] += e * e;
                            e = ((c.im - b.im) - db) - m2b; q[5] += e * e;
                        }
                    }
                }
            }
            nm_cx_block_sum<6, NW>(q, red + 8 * NW, tid);
        }
        if (tid < (has2 ? 2 : 1)) {
            const bool k = tid != 0;  // selects by predicate: no dynamically indexed local arrays
            const int c = c0 + (k ? 1 : 0);
            if (want_hjorth) {
                const double v0 = (k ? q[3] : q[0]) / n0, v1 = (k ? q[4] : q[1]) / n1, v2 = (k ? q[5] : q[2]) / n2;
                const double mob = nm_nan_to_num(sqrt(v1 / v0));
                nm_store(out, w, c, 0, nm_nan_to_num(v0));
                nm_store(out, w, c, 1, mob);
                nm_store(out, w, c, 2, nm_nan_to_num(sqrt(v2 / v1) / mob));
            }
            const cx<double> last = x[W - 1];
            if (want_raw) nm_store(out, w, c, 3, k ? last.im : last.re);
            // mean(|dx| / (W-1)) over W-1 samples: the reference divides by (W-1) twice
            if (want_ll) nm_store(out, w, c, 4, ((k ? s[7] : s[3]) / n1) / n1);
        }
    }
};

// ---------------------------------------------------------------- the kernel
// MODE_REFLECT: NM_FIR_REFLECT (notch) or NM_FIR_SAME;  BANK: more than one filter shares the forward transform.
template <int P, bool REFLECT, bool BANK, class Epi>
NM_GLOBAL void NM_LAUNCH_BOUNDS(NmCxPlan<P>::NT, NmCxPlan<P>::MINB) nm_convx_kernel(NmConvArgs a, Epi epi) {
    using PL = NmCxPlan<P>;
    constexpr int NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<double>* work = reinterpret_cast<cx<double>*>(smem);
    cx<double>* spec = BANK ? work + PL::NBUF : work;
    double* red = reinterpret_cast<double*>(work + (BANK ? 2 : 1) * PL::NBUF);  // NM_CX_RED_BYTES of reduction scratch
    unsigned char* scratch = a.scratch_in_tail ? reinterpret_cast<unsigned char*>(work + P)
                                               : reinterpret_cast<unsigned char*>(red) + NM_CX_RED_BYTES;
    const int tid = threadIdx.x;
    const int W = a.in.W, E = a.E;
    const int npair = (a.in.n_ch + 1) >> 1;
    const int o0 = REFLECT ? E : 0;
    const cx<double>* NM_RESTRICT tw = a.fft.tw;
    cx<double>* const p0w = work + tid + (tid >> PL::PAD);
    cx<double>* const p0s = spec + tid + (tid >> PL::PAD);
    const cx<double> wA = nm_ldg(tw + tid);                          // exp(-2*pi*i*tid/P): pass-0 twiddle generator
    const cx<double> wB = nm_ldg(tw + (tid & (PL::M1 - 1)) * 16);    // exp(-2*pi*i*j/NT): pass-1 generator (table stride P/NT)

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item / npair;
        const int c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        const double* NM_RESTRICT r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
        const double* NM_RESTRICT r1 = r0 + (has2 ? a.in.ch_stride : 0);

        // ---- pass 0 fused with the load of the (odd-reflected / zero padded) window
        cx<double> v[16];
        if (REFLECT) {
            const double a0 = 2.0 * r0[0], b0 = 2.0 * r1[0], a1 = 2.0 * r0[W - 1], b1 = 2.0 * r1[W - 1];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int n = tid + NT * t;
                double va = 0.0, vb = 0.0;
                if (n < E) {
                    va = a0 - r0[E - n]; vb = b0 - r1[E - n];
                } else if (n < E + W) {
                    va = r0[n - E]; vb = r1[n - E];
                } else if (n < W + 2 * E) {
                    const int k = n - (E + W) + 1;
                    va = a1 - r0[W - 1 - k]; vb = b1 - r1[W - 1 - k];
                }
                v[t] = {va, has2 ? vb : 0.0};
            }
        } else {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int n = tid + NT * t;
                double va = 0.0, vb = 0.0;
                if (n < W) { va = r0[n]; vb = r1[n]; }
                v[t] = {va, has2 ? vb : 0.0};
            }
        }
        nm_bfly16<false>(v);
        if (tid != 0) nm_twiddle_w1<16, false>(v, wA);
#pragma unroll
        for (int t = 0; t < 16; ++t) p0s[t * PL::S0] = v[t];
        __syncthreads();
        nm_cx_pass1<PL, false>(spec, wB, tid);
        __syncthreads();
        if (BANK) {
            nm_cx_pass2<PL, 0>(spec, spec, nullptr, tid);
            __syncthreads();
        }

        for (int fi = 0; fi < (BANK ? a.nF : 1); ++fi) {
            const double* NM_RESTRICT h = a.hperm + (size_t)fi * P;
            if (BANK) nm_cx_pass2<PL, 2>(work, spec, h, tid);
            else nm_cx_pass2<PL, 1>(work, work, h, tid);
            __syncthreads();
            nm_cx_pass1<PL, true>(work, wB, tid);
            __syncthreads();
            // ---- final inverse pass: padded slots -> registers, natural order n = tid + NT * t
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = p0w[t * PL::S0];
            if (tid != 0) nm_twiddle_w1<16, true>(v, wA);
            nm_bfly16<true>(v);
            // `work` may still be read by slower threads: an epilogue (or the next filter's pass) must not write it
            // before a barrier.  Register epilogues either synchronise inside (kSyncsInside) or get a trailing barrier.
            bool in_regs = Epi::kRegsOnly;
            if constexpr (Epi::kRegs && !Epi::kRegsOnly) in_regs = epi.regs_ok();
            if (in_regs) {
                if constexpr (Epi::kRegs) {
                    epi.template run_x<PL>(v, work, red, o0, W, a.in.n_ch, w, c0, has2, fi, tid);
                    if constexpr (!Epi::kSyncsInside) __syncthreads();
                }
            } else {
                if constexpr (!Epi::kRegsOnly) {
                    __syncthreads();
#pragma unroll
                    for (int t = 0; t < 16; ++t) work[tid + NT * t] = v[t];
                    __syncthreads();
                    epi.run(work, o0, W, a.in.n_ch, w, c0, has2, fi, scratch, tid, NT);
                    __syncthreads();
                }
            }
        }
    }
}
