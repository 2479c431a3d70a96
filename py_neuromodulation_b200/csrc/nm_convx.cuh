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

#include <type_traits>

#include "nm_conv.cuh"

//
// Round 2: P = 3 * 2^k.  A 'same'-mode FIR of L taps over W samples needs a circular length P >= W + (L-1)/2 only (1499 for the
// 999-tap band-pass banks at 1 kHz, 2999 at 2 kHz), so 1536 = 12*8*16 and 3072 = 12*16*16 replace 2048 / 4096 there: a quarter
// fewer points through every pass.  Pass 0 becomes a radix-12 prime-factor butterfly (3 x 4, no internal twiddles) on V0 = 12
// values per thread with NT = P/12 threads (the CTA size); passes 1 and 2 keep 16 values per thread and therefore run on the
// first NT1 = P/16 threads (three of every four warps), the others only meet the barriers.
#ifndef NM_CX_MINB1_NT128
#define NM_CX_MINB1_NT128 4  // single-filter kernels of the 128-thread plans: resident CTAs per SM the registers are capped for
#endif
#ifndef NM_CX_MINBK_1536
#define NM_CX_MINBK_1536 3  // resident CTAs per SM the registers of the 1536-point bank kernels are capped for (4 fit the shared memory)
#endif
template <int P_>
struct NmCxPlan {
    static constexpr int P = P_;
    static constexpr bool MIXED = (P % 3) == 0;
    static constexpr int V0 = MIXED ? 12 : 16;                // radix of pass 0 == values a thread owns in pass 0 and in the epilogue
    static constexpr int NT = P / V0;                         // CTA threads
    static constexpr int NT1 = P / 16;                        // threads of passes 1 and 2
    static constexpr int R1 = (P == 1024 || P == 1536) ? 8 : 16;
    static constexpr int R2 = P / V0 / R1;
    static constexpr int PAD = (R2 == 16) ? 4 : 3;
    static constexpr int M1 = NT / R1;                        // butterfly stride of pass 1 (block length NT)
    static constexpr int S0 = NT + (NT >> PAD);               // physical strides (see header comment)
    static constexpr int S1 = M1 + (M1 >> PAD);
    static constexpr int B1 = NT1 * R1 + ((NT1 * R1) >> PAD);
    static constexpr int B2 = NT1 * R2 + ((NT1 * R2) >> PAD);
    static constexpr int NBUF = P + (P >> PAD) + 2;
    // resident CTAs per SM the register allocation is tuned for: single-filter kernels run at 128 registers,
    // bank kernels (shared-memory limited anyway) at ~168 so that the prefetched filter spectrum stays in registers
    static constexpr int MINB1 = (NT >= 256) ? 2 : (NT == 128 ? NM_CX_MINB1_NT128 : 8);
    static constexpr int MINBK = (NT >= 256) ? 1 : (NT == 128 ? (P == 1536 ? NM_CX_MINBK_1536 : 3) : 6);
    static_assert(R2 == (1 << PAD), "padding unit must equal the last radix");
    static_assert(M1 % R2 == 0 && NT % R2 == 0, "strides must be multiples of the padding unit");
    static_assert(NT1 % M1 == 0 && (NT1 * R1) % R2 == 0, "the butterflies of a thread share one twiddle set");
    static_assert(R1 * R2 * V0 == P && NT1 % 32 == 0, "three-pass plan");
};

#define NM_CX_RED_BYTES 512  // 8 values x 8 warps of doubles, owned by the register epilogues

static inline bool nm_convx_supported(int P) { return P == 1024 || P == 2048 || P == 4096 || P == 1536 || P == 3072; }
// CTA size of the compile-time plan
static inline int nm_convx_threads(int P) { return (P % 3 == 0) ? P / 12 : P / 16; }

// multiply the R-1 upper values of a butterfly by the powers of w1 (conjugated for the inverse)
template <int R, bool INV, typename T>
NM_DEV void nm_twiddle_w1(cx<T>* v, cx<T> w1) {
    if (INV) w1.im = -w1.im;
    v[1] = cx_mul(v[1], w1);
    const cx<T> w2 = cx_mul(w1, w1), w3 = cx_mul(w2, w1);
    v[2] = cx_mul(v[2], w2);
    v[3] = cx_mul(v[3], w3);
    const cx<T> w4 = cx_mul(w2, w2);
    v[4] = cx_mul(v[4], w4);
    v[5] = cx_mul(v[5], cx_mul(w4, w1));
    v[6] = cx_mul(v[6], cx_mul(w4, w2));
    v[7] = cx_mul(v[7], cx_mul(w4, w3));
    if (R > 8) {
        const cx<T> w8 = cx_mul(w4, w4);
        v[8] = cx_mul(v[8], w8);
        v[9] = cx_mul(v[9], cx_mul(w8, w1));
        v[10] = cx_mul(v[10], cx_mul(w8, w2));
        v[11] = cx_mul(v[11], cx_mul(w8, w3));
        if (R > 12) {
            const cx<T> w12 = cx_mul(w8, w4);
            v[12] = cx_mul(v[12], w12);
            v[13] = cx_mul(v[13], cx_mul(w12, w1));
            v[14] = cx_mul(v[14], cx_mul(w12, w2));
            v[15] = cx_mul(v[15], cx_mul(w12, w3));
        }
    }
}

// 12-point DFT in registers, natural order in and out: prime-factor (Good-Thomas) 3 x 4, n = (4*n1 + 3*n2) mod 12,
// k = (4*k1 + 9*k2) mod 12, so that w12^(n*k) = w3^(n1*k1) * w4^(n2*k2) -- no twiddles between the two stages
template <bool INV, typename T>
NM_DEV void nm_bfly12(cx<T>* v) {
    cx<T> u[4][3];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        cx<T> a[3] = {v[(3 * n2) % 12], v[(4 + 3 * n2) % 12], v[(8 + 3 * n2) % 12]};
        nm_bfly3<T, INV>(a);
        u[n2][0] = a[0]; u[n2][1] = a[1]; u[n2][2] = a[2];
    }
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) {
        nm_r4<INV>(u[0][k1], u[1][k1], u[2][k1], u[3][k1]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) v[(4 * k1 + 9 * k2) % 12] = u[k2][k1];
    }
}

// pass-0 butterfly of a plan: radix 16 or 12
template <int R, bool INV, typename T>
NM_DEV void nm_bfly0(cx<T>* v) {
    if (R == 16) nm_bfly16<INV>(v);
    else nm_bfly12<INV>(v);
}

// interior pass (pass 1): 16/R1 butterflies per thread, all sharing the same twiddle set
template <class PL, bool INV, typename T>
NM_DEV void nm_cx_pass1(cx<T>* sm, const cx<T> w1, int tid) {
    constexpr int R = PL::R1;
    if (PL::NT1 < PL::NT && tid >= PL::NT1) return;  // (warp-uniform: NT1 is a multiple of 32)
    const int j = tid & (PL::M1 - 1);
    const int base = (tid - j) * R + j;
    cx<T>* p = sm + base + (base >> PL::PAD);
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        cx<T> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = p[i * PL::B1 + t * PL::S1];
        if (INV) {
            nm_twiddle_w1<R, true>(v, w1);  // j == 0 multiplies by exactly 1
            nm_bflyR<R, true>(v);
        } else {
            nm_bflyR<R, false>(v);
            nm_twiddle_w1<R, false>(v, w1);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) p[i * PL::B1 + t * PL::S1] = v[t];
    }
}

// The 16 filter-spectrum values a thread multiplies with in the last pass are slots (tid + NT*i)*R2 + t -- the same for
// every item.  The host stores them thread-interleaved (`hx`, nm_cx_hx_index) so that a warp's 128-bit loads are
// contiguous (4 L1 tag requests per instruction instead of 16 for the slot-ordered table); they are fetched one
// phase ahead of their use.
template <class PL>
static NM_HD int nm_cx_hx_index(int tid, int i, int t) {  // position of slot (tid + NT*i)*R2 + t inside one filter's hx block
    return ((i * (PL::R2 / 2) + (t >> 1)) * PL::NT1 + tid) * 2 + (t & 1);
}

template <class PL, typename T>
NM_DEV void nm_cx_load_h(T* hv, const T* NM_RESTRICT hx, int tid) {
    constexpr int R = PL::R2;
    if (PL::NT1 < PL::NT && tid >= PL::NT1) return;
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
#pragma unroll
        for (int t = 0; t < R / 2; ++t) {
            // one (2 x T)-wide load per pair: consecutive lanes read consecutive pairs
            const cx<T> d = nm_ldg(reinterpret_cast<const cx<T>*>(hx) + (i * (R / 2) + t) * PL::NT1 + tid);
            hv[i * R + 2 * t] = d.re;
            hv[i * R + 2 * t + 1] = d.im;
        }
    }
}

// packed float32 pairs: the float32 table, every value broadcast to both halves
template <class PL>
NM_DEV void nm_cx_load_h(f32x2* hv, const float* NM_RESTRICT hx, int tid) {
    constexpr int R = PL::R2;
    if (PL::NT1 < PL::NT && tid >= PL::NT1) return;
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
#pragma unroll
        for (int t = 0; t < R / 2; ++t) {
            const cx<float> d = nm_ldg(reinterpret_cast<const cx<float>*>(hx) + (i * (R / 2) + t) * PL::NT1 + tid);
            hv[i * R + 2 * t] = f32x2(d.re);
            hv[i * R + 2 * t + 1] = f32x2(d.im);
        }
    }
}

template <class PL, typename T>
NM_DEV void nm_cx_load_hT(T* hv, const T* NM_RESTRICT hx, int tid) { nm_cx_load_h<PL, T>(hv, hx, tid); }
template <class PL, typename T>
NM_DEV void nm_cx_load_hT(f32x2* hv, const float* NM_RESTRICT hx, int tid) { nm_cx_load_h<PL>(hv, hx, tid); }

// host side: slot-ordered spectrum block (P values) -> thread-interleaved block
template <int P, typename T>
static inline void nm_cx_interleave_h(const double* h, T* hx) {
    using PL = NmCxPlan<P>;
    for (int tid = 0; tid < PL::NT1; ++tid)
        for (int i = 0; i < 16 / PL::R2; ++i)
            for (int t = 0; t < PL::R2; ++t) hx[nm_cx_hx_index<PL>(tid, i, t)] = (T)h[(tid + PL::NT1 * i) * PL::R2 + t];
}
template <typename T>
static inline void nm_cx_interleave_h(int P, const double* h, T* hx) {
    if (P == 1024) nm_cx_interleave_h<1024>(h, hx);
    else if (P == 2048) nm_cx_interleave_h<2048>(h, hx);
    else if (P == 1536) nm_cx_interleave_h<1536>(h, hx);
    else if (P == 3072) nm_cx_interleave_h<3072>(h, hx);
    else nm_cx_interleave_h<4096>(h, hx);
}

// last forward pass (unit stride).  MODE 0: forward butterfly only (bank: spectrum stays in `src`).
// MODE 1: forward butterfly, * H, inverse butterfly in place (single filter).
// MODE 2: load the spectrum from `src`, * H, inverse butterfly, store to `dst` (one filter of a bank).
template <class PL, int MODE, typename T>
NM_DEV void nm_cx_pass2(cx<T>* dst, const cx<T>* src, const T* hv, int tid) {
    constexpr int R = PL::R2;
    if (PL::NT1 < PL::NT && tid >= PL::NT1) return;
    const int off = tid * (R + 1);  // tid*R + ((tid*R) >> PAD)
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        cx<T> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[off + i * PL::B2 + t];
        if (MODE != 2) nm_bflyR<R, false>(v);
        if (MODE != 0) {
#pragma unroll
            for (int t = 0; t < R; ++t) v[t] = {v[t].re * hv[i * R + t], v[t].im * hv[i * R + t]};
            nm_bflyR<R, true>(v);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) dst[off + i * PL::B2 + t] = v[t];
    }
}

// compact CTA-wide sum of NV (power of two <= 8) values; `red` needs NV * NW doubles; result valid in every thread.
// Two barriers; the warp stage is the halving exchange of nm_warp_sum_multi.
template <int NV, int NW>
NM_DEV void nm_cx_block_sum(double* v, double* red, int tid) {
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int SH = (NV == 8) ? 2 : (NV == 4 ? 3 : (NV == 2 ? 4 : 5));  // value i lives in lanes with lane >> SH == i
    nm_warp_sum_multi<NV>(v, lane);
    __syncthreads();
    if ((lane & ((1 << SH) - 1)) == 0) red[(lane >> SH) * NW + wid] = v[0];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[i * NW + w];
        v[i] = s;
    }
}

template <typename T>
NM_DEV cx<double> nm_cx_wide(cx<T> a) { return {(double)a.re, (double)a.im}; }

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
    static constexpr bool kRegsOnly = true, kReflectOk = true, kSameOk = false, kConvxOnly = true, kF32Ok = true, kSplitOk = false;
    static constexpr bool kSyncsInside = false;  // (not on every path) -> the kernel adds the trailing barrier
    static NM_HD size_t smem_bytes(int /*nt*/) { return 0; }  // reductions use the padding tail of `work`
    NM_DEV bool regs_ok() const { return true; }
    struct State {};

    // window sample u lives at work[u + (u >> 3)]: threads then walk chunks of CH = ceil(W / NT) consecutive samples
    // (8 for the default sizes), which this padding makes bank-conflict free
    static NM_HD int phys(int u) { return u + (u >> 3); }

    // part 1: needs the registers -- store the rows, lay the window out in natural order
    template <class PL, typename T>
    NM_DEV void consume(const cx<T>* v, cx<T>* work, double* /*red*/, State& /*st*/, int o0, int W, int n_ch, int w, int c0,
                        bool has2, int /*f*/, int tid) const {
        constexpr int NT = PL::NT;
        if (y) {
            double* r0 = y + ((size_t)w * n_ch + c0) * Wp;
#pragma unroll
            for (int k = 0; k < PL::V0; ++k) {
                const int t = tid + NT * k - o0;
                if (t >= 0 && t < W) {
                    r0[t] = v[k].re;
                    if (has2) r0[Wp + t] = v[k].im;
                }
            }
        }
        if (!want_scan) return;
        __syncthreads();  // every thread has read its pass-0 inputs from `work`
#pragma unroll
        for (int k = 0; k < PL::V0; ++k) {
            const int u = tid + NT * k - o0;
            if (u >= 0 && u < W) work[phys(u)] = v[k];
        }
    }

    // one chunk sample of the two moment passes: PASS 1 sums x, d, dd, |d|; PASS 2 the centred squares (m = the six means)
    template <int PASS>
    NM_DEV static void acc(double* s, const double* m, cx<double> xa, cx<double> xb, cx<double> xc, bool has1, bool has2nd) {
        const double da = xb.re - xa.re, db = xb.im - xa.im;
        const double dda = (xc.re - xb.re) - da, ddb = (xc.im - xb.im) - db;
        if (PASS == 1) {
            s[0] += xa.re; s[4] += xa.im;
            if (has1) { s[1] += da; s[5] += db; s[3] += fabs(da); s[7] += fabs(db); }
            if (has2nd) { s[2] += dda; s[6] += ddb; }
        } else {
            double e = xa.re - m[0]; s[0] += e * e;
            e = xa.im - m[3]; s[3] += e * e;
            if (has1) { e = da - m[1]; s[1] += e * e; e = db - m[4]; s[4] += e * e; }
            if (has2nd) { e = dda - m[2]; s[2] += e * e; e = ddb - m[5]; s[5] += e * e; }
        }
    }

    // one moment pass over this thread's chunk of CH consecutive window samples with a sliding three-sample window (a fully
    // unrolled CH == 8 variant with all ten loads up front was slower: 40 more live registers at the 128-register cap)
    template <int PASS, typename T>
    NM_DEV void chunk_pass(const cx<T>* work, int W, int CH, int tid, double* s, const double* m) const {
        const int u0 = tid * CH;
        if (u0 >= W) return;
        const int u1 = min(W, u0 + CH);
        const cx<double> zero = {0.0, 0.0};
        cx<double> xa = nm_cx_wide(work[phys(u0)]), xb = (u0 + 1 < W) ? nm_cx_wide(work[phys(u0 + 1)]) : zero;
        for (int u = u0; u < u1; ++u) {
            const cx<double> xc = (u + 2 < W) ? nm_cx_wide(work[phys(u + 2)]) : zero;
            acc<PASS>(s, m, xa, xb, xc, u + 1 < W, u + 2 < W);
            xa = xb; xb = xc;
        }
    }

    // part 2: two-pass moments (numpy.var semantics) over chunks of consecutive samples, each sample loaded once per pass.
    // Ends with every thread past its last read of `work` and of the reduction scratch, so the kernel needs no trailing barrier:
    // the two threads that finish the formulas and store the results overlap with the next window's first pass.
    template <class PL, typename T>
    NM_DEV void finish(cx<T>* work, double* red, State& /*st*/, int /*o0*/, int W, int /*n_ch*/, int w, int c0, bool has2, int /*f*/,
                       int tid) const {
        constexpr int NT = PL::NT, NW = (PL::NT + 31) / 32;
        static_assert((size_t)8 * NW * sizeof(double) <= NM_CX_RED_BYTES, "reduction scratch");
        if (!want_scan) return;
        __syncthreads();  // natural-order copy complete
        const int CH = (W + NT - 1) / NT;
        const cx<double> last = nm_cx_wide(work[phys(W - 1)]);
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // per row: sum x, sum d, sum dd, sum |d|
        chunk_pass<1, T>(work, W, CH, tid, s, nullptr);
        nm_cx_block_sum<8, NW>(s, red, tid);
        const double n0 = W, n1 = W - 1, n2 = W - 2;
        double q[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // 6 used; 8 for the power-of-two exchange reduction
        if (want_hjorth) {
            const double m[6] = {s[0] / n0, s[1] / n1, s[2] / n2, s[4] / n0, s[5] / n1, s[6] / n2};
            chunk_pass<2, T>(work, W, CH, tid, q, m);
            nm_cx_block_sum<8, NW>(q, red, tid);  // (its leading barrier orders it after every read of the first reduction)
        } else {
            __syncthreads();  // every thread is past its reads of `work`
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
            if (want_raw) nm_store(out, w, c, 3, k ? last.im : last.re);
            // mean(|dx| / (W-1)) over W-1 samples: the reference divides by (W-1) twice
            if (want_ll) nm_store(out, w, c, 4, ((k ? s[7] : s[3]) / n1) / n1);
        }
    }
    // without the scan part there is no barrier inside the epilogue: the kernel must separate the register reads of `work`
    // from the next pass
    NM_DEV bool needs_trailing_barrier() const { return !want_scan; }
};

// ---------------------------------------------------------------- the kernel
// REFLECT: NM_FIR_REFLECT (notch) or NM_FIR_SAME;  BANK: the filters of a bank share the forward transform.
//
// Register epilogues come in two parts: consume() reads the 16 outputs a thread holds, finish() does whatever is left
// (reductions, final formulas).  Between the two the kernel already issues the global loads of the NEXT item into the
// freed registers, so their latency overlaps the reductions / barriers instead of stalling the next pass 0.
template <int P, bool REFLECT, typename T>
NM_DEV void nm_cx_load_item(cx<T>* v, const NmConvArgs& a, int item, int npair, int tid, int& w, int& c0, bool& has2) {
    constexpr int NT = NmCxPlan<P>::NT, V0 = NmCxPlan<P>::V0;
    const int W = a.in.W, E = a.E;
    w = item / npair;
    c0 = (item - w * npair) * 2;
    has2 = c0 + 1 < a.in.n_ch;
    const double* NM_RESTRICT r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
    const double* NM_RESTRICT r1 = r0 + (has2 ? a.in.ch_stride : 0);
    if (REFLECT) {
        // odd reflection about both end samples, branch-free: value = c + s * row[idx] with (c, s, idx) selected per
        // region, so all 32 loads of a thread are issued back to back (c - x == fma(-1, x, c): same rounding)
        const double a0 = 2.0 * r0[0], b0 = 2.0 * r1[0], a1 = 2.0 * r0[W - 1], b1 = 2.0 * r1[W - 1];
#pragma unroll
        for (int t = 0; t < V0; ++t) {
            const int n = tid + NT * t;
            const bool left = n < E, mid = !left && n < E + W, right = !left && !mid && n < W + 2 * E;
            int idx = left ? E - n : (mid ? n - E : 2 * W + E - 2 - n);  // right: W-1-k with k = n-(E+W)+1
            idx = (left || mid || right) ? idx : 0;
            const double sgn = mid ? 1.0 : ((left || right) ? -1.0 : 0.0);
            const double ca = left ? a0 : (right ? a1 : 0.0), cb = left ? b0 : (right ? b1 : 0.0);
            const double va = fma(sgn, r0[idx], ca), vb = fma(sgn, r1[idx], cb);
            v[t] = {(T)va, (T)(has2 ? vb : 0.0)};
        }
    } else {
#pragma unroll
        for (int t = 0; t < V0; ++t) {
            const int n = tid + NT * t;
            const int idx = n < W ? n : 0;
            const double va = r0[idx], vb = r1[idx];
            v[t] = {(T)(n < W ? va : 0.0), (T)((has2 && n < W) ? vb : 0.0)};
        }
    }
}

// float32 mode: one item = (window, channel QUAD); sub-item A = rows (c0, c0 + 1) in the .x halves, B = rows (c0 + 2, c0 + 3) in
// the .y halves.  Same reflection / zero-padding arithmetic in float64 as above, rounded once to float32.
template <int P, bool REFLECT>
NM_DEV void nm_cx_load_item4(cx<f32x2>* v, const NmConvArgs& a, int item, int nquad, int tid, int& w, int& c0) {
    constexpr int NT = NmCxPlan<P>::NT, V0 = NmCxPlan<P>::V0;
    const int W = a.in.W, E = a.E;
    w = item / nquad;
    c0 = (item - w * nquad) * 4;
    const int nv = min(4, a.in.n_ch - c0);  // valid rows of the quad
    const double* NM_RESTRICT r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
    const double* NM_RESTRICT r1 = r0 + (nv > 1 ? a.in.ch_stride : 0);
    const double* NM_RESTRICT r2 = r0 + (nv > 2 ? 2 * a.in.ch_stride : 0);
    const double* NM_RESTRICT r3 = r0 + (nv > 3 ? 3 * a.in.ch_stride : 0);
    const bool h1 = nv > 1, h2 = nv > 2, h3 = nv > 3;
    if (REFLECT) {
        const double e0[4] = {2.0 * r0[0], 2.0 * r1[0], 2.0 * r2[0], 2.0 * r3[0]};
        const double e1[4] = {2.0 * r0[W - 1], 2.0 * r1[W - 1], 2.0 * r2[W - 1], 2.0 * r3[W - 1]};
#pragma unroll
        for (int t = 0; t < V0; ++t) {
            const int n = tid + NT * t;
            const bool left = n < E, mid = !left && n < E + W, right = !left && !mid && n < W + 2 * E;
            int idx = left ? E - n : (mid ? n - E : 2 * W + E - 2 - n);
            idx = (left || mid || right) ? idx : 0;
            const double sgn = mid ? 1.0 : ((left || right) ? -1.0 : 0.0);
            const double x0 = fma(sgn, r0[idx], left ? e0[0] : (right ? e1[0] : 0.0));
            const double x1 = fma(sgn, r1[idx], left ? e0[1] : (right ? e1[1] : 0.0));
            const double x2 = fma(sgn, r2[idx], left ? e0[2] : (right ? e1[2] : 0.0));
            const double x3 = fma(sgn, r3[idx], left ? e0[3] : (right ? e1[3] : 0.0));
            v[t] = {f32x2((float)x0, (float)(h2 ? x2 : 0.0)), f32x2((float)(h1 ? x1 : 0.0), (float)(h3 ? x3 : 0.0))};
        }
    } else {
#pragma unroll
        for (int t = 0; t < V0; ++t) {
            const int n = tid + NT * t;
            const int idx = n < W ? n : 0;
            const bool in = n < W;
            v[t] = {f32x2((float)(in ? r0[idx] : 0.0), (float)((in && h2) ? r2[idx] : 0.0)),
                    f32x2((float)((in && h1) ? r1[idx] : 0.0), (float)((in && h3) ? r3[idx] : 0.0))};
        }
    }
}

template <typename T> struct NmIsPacked { static constexpr bool value = false; };
template <> struct NmIsPacked<f32x2> { static constexpr bool value = true; };

// float64 / float32 views of the twiddle and filter-spectrum tables of a launch
// twiddle generator k of the table, in the arithmetic type of the kernel
NM_DEV cx<double> nm_cx_twid(const NmConvArgs& a, int k, double) { return nm_ldg(a.fft.tw + k); }
NM_DEV cx<f32x2> nm_cx_twid(const NmConvArgs& a, int k, f32x2) {
    const cx<float> w = nm_ldg(a.tw32 + k);
    return {f32x2(w.re), f32x2(w.im)};
}
NM_DEV cx<float> nm_cx_twid(const NmConvArgs& a, int k, float) { return nm_ldg(a.tw32 + k); }
NM_DEV const double* nm_cx_hx(const NmConvArgs& a, double) { return a.hx; }
NM_DEV const float* nm_cx_hx(const NmConvArgs& a, f32x2) { return a.hx32; }
NM_DEV const float* nm_cx_hx(const NmConvArgs& a, float) { return a.hx32; }

#ifndef NM_CXP_MINB1_128
#define NM_CXP_MINB1_128 4  // packed single-filter kernel, 128-thread plan: resident CTAs per SM the registers are capped for
#endif
template <typename T, int P, bool BANK>
struct NmCxOcc {  // (packed float32 pairs use the float64 kernel's shared-memory footprint; scalar float32 has half of it)
    static constexpr int f32 = (BANK ? 6 : 5) * 128 / NmCxPlan<P>::NT;
    static constexpr int value = sizeof(T) == 4 ? (f32 < 1 ? 1 : f32)
                                 : (BANK ? NmCxPlan<P>::MINBK
                                         : ((NmIsPacked<T>::value && NmCxPlan<P>::NT == 128) ? NM_CXP_MINB1_128 : NmCxPlan<P>::MINB1));
};
template <typename T, int P, bool REFLECT, bool BANK, class Epi>
NM_GLOBAL void NM_LAUNCH_BOUNDS(NmCxPlan<P>::NT, (NmCxOcc<T, P, BANK>::value))
nm_convx_kernel(NmConvArgs a, Epi epi) {
    using PL = NmCxPlan<P>;
    constexpr int NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<T>* work = reinterpret_cast<cx<T>*>(smem);
    cx<T>* spec = BANK ? work + PL::NBUF : work;
    double* red = reinterpret_cast<double*>(work + (BANK ? 2 : 1) * PL::NBUF);  // NM_CX_RED_BYTES of reduction scratch
    unsigned char* scratch = a.scratch_in_tail ? reinterpret_cast<unsigned char*>(work + P)
                                               : reinterpret_cast<unsigned char*>(red) + NM_CX_RED_BYTES;
    const int tid = threadIdx.x;
    const int W = a.in.W;
    constexpr bool PACKED = NmIsPacked<T>::value;
    // items per window: channel pairs (float64) or channel quads (packed float32 pairs)
    const int npair = PACKED ? (a.in.n_ch + 3) >> 2 : (a.in.n_ch + 1) >> 1;
    const int o0 = REFLECT ? a.E : 0;
    const int nF = BANK ? a.nF : 1;
    const auto* NM_RESTRICT hx = nm_cx_hx(a, T());
    cx<T>* const p0w = work + tid + (tid >> PL::PAD);
    cx<T>* const p0s = spec + tid + (tid >> PL::PAD);
    // twiddle generators; tid == 0 / j == 0 multiply by exactly 1, so no thread needs a special case
    const cx<T> wA = nm_cx_twid(a, tid, T());                          // exp(-2*pi*i*tid/P): pass 0
    const cx<T> wB = nm_cx_twid(a, (tid & (PL::M1 - 1)) * PL::V0, T());  // exp(-2*pi*i*j/NT): pass 1 (table stride P/NT)

    cx<T> v[PL::V0];
    T hv[16];
    int item = blockIdx.x, w = 0, c0 = 0;
    bool has2 = false;
    if (item >= a.n_items) return;
    if constexpr (PACKED) nm_cx_load_item4<P, REFLECT>(v, a, item, npair, tid, w, c0);
    else nm_cx_load_item<P, REFLECT, T>(v, a, item, npair, tid, w, c0, has2);
    if (BANK) nm_cx_load_hT<PL, T>(hv, hx, tid);

    while (item < a.n_items) {
        const int next = item + gridDim.x;
        int nw = 0, nc0 = 0;
        bool nhas2 = false;
        // ---- pass 0 on the window that is already in registers
        nm_bfly0<PL::V0, false>(v);
        nm_twiddle_w1<PL::V0, false>(v, wA);
#pragma unroll
        for (int t = 0; t < PL::V0; ++t) p0s[t * PL::S0] = v[t];
        __syncthreads();
        if (!BANK) nm_cx_load_hT<PL, T>(hv, hx, tid);  // (L1 resident) lands while pass 1 computes
        nm_cx_pass1<PL, false>(spec, wB, tid);
        __syncthreads();
        if (BANK) {
            nm_cx_pass2<PL, 0, T>(spec, spec, nullptr, tid);
            __syncthreads();
        }

        for (int fi = 0; fi < nF; ++fi) {
            const bool last = fi + 1 == nF;
            if (BANK) {
                nm_cx_pass2<PL, 2>(work, spec, hv, tid);
#ifdef NM_CX_HV_EARLY
                // prefetch the next filter's spectrum (wrapping to filter 0 for the next item) one phase ahead
                nm_cx_load_hT<PL, T>(hv, hx + (size_t)(last ? 0 : fi + 1) * P, tid);
#endif
            } else {
                nm_cx_pass2<PL, 1>(work, work, hv, tid);
            }
            __syncthreads();
            nm_cx_pass1<PL, true>(work, wB, tid);
            __syncthreads();
            // ---- final inverse pass: padded slots -> registers, natural order n = tid + NT * t
#pragma unroll
            for (int t = 0; t < PL::V0; ++t) v[t] = p0w[t * PL::S0];
            nm_twiddle_w1<PL::V0, true>(v, wA);
            nm_bfly0<PL::V0, true>(v);
#ifndef NM_CX_HV_EARLY
            // next filter's spectrum (wrapping to filter 0 for the next item): issued where the register pressure is lowest -- a
            // load that the compiler has to spill right away stalls on its own L2 latency (ncu: STL of the loaded pair right
            // behind the LDG, round 2) -- and it lands during the epilogue's reduction and barrier
            if (BANK) nm_cx_load_hT<PL, T>(hv, hx + (size_t)(last ? 0 : fi + 1) * P, tid);
#endif
            // `work` may still be read by slower threads: an epilogue (or the next filter's pass) must not write it
            // before a barrier.  Register epilogues either synchronise inside (kSyncsInside) or get a trailing barrier.
            bool in_regs = Epi::kRegsOnly;
            if constexpr (Epi::kRegs && !Epi::kRegsOnly) in_regs = epi.regs_ok();
            if (in_regs) {
                if constexpr (Epi::kRegs && PACKED) {
                    // the two sub-items leave the packed registers as float32 pairs and run the float32 form of the epilogue one
                    // after the other (B in its own half of `work` / of the reduction scratch where the epilogue overlaps its tail
                    // with the next phase)
                    cx<float>* const wf = reinterpret_cast<cx<float>*>(work);
                    const int nv = min(4, a.in.n_ch - c0);
                    {
                        cx<float> u[PL::V0];
#pragma unroll
                        for (int k = 0; k < PL::V0; ++k) u[k] = {v[k].re.x, v[k].im.x};
                        typename Epi::State st;
                        epi.template consume<PL, float>(u, wf, red, st, o0, W, a.in.n_ch, w, c0, nv > 1, fi, tid);
                        epi.template finish<PL, float>(wf, red, st, o0, W, a.in.n_ch, w, c0, nv > 1, fi, tid);
                    }
                    if (nv > 2) {
                        cx<float> u[PL::V0];
#pragma unroll
                        for (int k = 0; k < PL::V0; ++k) u[k] = {v[k].re.y, v[k].im.y};
                        double* const redb = Epi::kSyncsInside ? red + NM_CX_RED_BYTES / 16 : red;
                        typename Epi::State st;
                        cx<float>* const wb = wf + PL::NBUF;  // second half of the (16-byte element) transform buffer
                        epi.template consume<PL, float>(u, wb, redb, st, o0, W, a.in.n_ch, w, c0 + 2, nv > 3, fi, tid);
                        epi.template finish<PL, float>(wb, redb, st, o0, W, a.in.n_ch, w, c0 + 2, nv > 3, fi, tid);
                    }
                    if (last && next < a.n_items) nm_cx_load_item4<P, REFLECT>(v, a, next, npair, tid, nw, nc0);
                    if constexpr (!Epi::kSyncsInside) {
                        if (epi.needs_trailing_barrier()) __syncthreads();
                    }
                } else if constexpr (Epi::kRegs) {
                    typename Epi::State st;
                    epi.template consume<PL, T>(v, work, red, st, o0, W, a.in.n_ch, w, c0, has2, fi + a.f0, tid);
                    if (last && next < a.n_items) nm_cx_load_item<P, REFLECT, T>(v, a, next, npair, tid, nw, nc0, nhas2);
                    epi.template finish<PL, T>(work, red, st, o0, W, a.in.n_ch, w, c0, has2, fi + a.f0, tid);
                    if constexpr (!Epi::kSyncsInside) {
                        if (epi.needs_trailing_barrier()) __syncthreads();
                    }
                }
            } else {
                if constexpr (!Epi::kRegsOnly && std::is_same<T, double>::value) {  // (the shared-memory epilogues take float64 rows)
                    __syncthreads();
#pragma unroll
                    for (int t = 0; t < PL::V0; ++t) work[tid + NT * t] = v[t];
                    __syncthreads();
                    epi.run(work, o0, W, a.in.n_ch, w, c0, has2, fi + a.f0, scratch, tid, NT);
                    __syncthreads();
                    // (no early prefetch here: these epilogues are register hungry and long enough to hide nothing; a prefetch.global.L2 of the
                    // next item's lines in front of the epilogue measured +-0: sharp waves 8.03 -> 8.12 ms, envelopes 6.34 -> 6.35 ms per 591 windows)
                    if (last && next < a.n_items) nm_cx_load_item<P, REFLECT, T>(v, a, next, npair, tid, nw, nc0, nhas2);
                }
            }
        }
        item = next; w = nw; c0 = nc0; has2 = nhas2;
    }
}
