// nm_conv.cuh -- power-of-two FFT convolution kernel for the FIR banks (the hot kernel of the path).
//
// Same contract as nm_fir_kernel (nm_fir.cuh) -- one (window, channel pair) item per CTA iteration, the two real rows
// packed as one complex signal, real symmetric filter spectra, fused epilogues -- but organised around registers:
//
//   * P = 2^k, NT = P/16 threads; in EVERY pass each thread owns exactly 16 complex values and runs 16/R radix-R
//     butterflies on them in registers (radix 16 = 4x4, 8 = 4x2), so a 2048-point transform is 3 passes, not 5+
//   * pass 1 reads the (extended / zero-padded) window straight from global memory, the last inverse pass hands the
//     filtered rows to the epilogue -- the shared buffer is touched once per interior pass only
//   * for a single filter the last forward pass, the spectrum multiply and the first inverse pass are ONE register
//     step (load, butterfly, *H, inverse butterfly, store)
//   * shared-memory index e lives at e + (e >> pad) (pad = log2 of the last radix): every pass, including the
//     unit-stride one, is bank-conflict free for 16-byte accesses (tests/.. bank model in DESIGN.md section 4)
//   * twiddles: one table load per butterfly, the other powers by (log-depth) complex products
//
// Digit-reversed order in the frequency domain is never undone: Hperm is stored in slot order.
#pragma once

#include "nm_fir.cuh"

#define NM_CONV_MAX_PASS 4

struct NmConv {
    int P, NT, npass;
    int radix[NM_CONV_MAX_PASS];  // forward order, radix[0] == 16
    int len[NM_CONV_MAX_PASS];    // block length seen by pass p; len[0] == P
    int pad;                      // physical slot = e + (e >> pad)
    const cx<double>* tw;         // exp(-2*pi*i*k/P)
};

struct NmConvArgs {
    NmRows in;
    NmConv fft;
    const double* hperm;  // [nF][P] real spectra in slot order, scaled by 1/P
    const double* hx;     // same values in nm_convx_kernel's per-thread order (nm_cx_load_h), or nullptr
    const float* hx32;    // float32 copies of hx / the twiddle table for the float32 mode of nm_convx_kernel, or nullptr
    const cx<float>* tw32;
    int nF, mode, E, n_items;
    int f0;               // index of the first filter of this launch inside its bank (banks too large for shared memory run filter by filter)
    int scratch_in_tail;  // epilogue scratch aliases the unused padding tail of `work` (linear output only uses [0, P))
};

#define NM_PHYS(e, pad) ((e) + ((e) >> (pad)))

// The register butterflies are templates on the scalar type: float64 is the default arithmetic of the path, float32 serves
// the optional fast mode of the linear families (nm_set_precision).
template <bool INV, typename T>
NM_DEV cx<T> nm_mulw(cx<T> a, T wr, T wi) {  // a * (wr + i*wi), conjugated for the inverse
    return INV ? cx<T>{a.re * wr + a.im * wi, a.im * wr - a.re * wi} : cx<T>{a.re * wr - a.im * wi, a.im * wr + a.re * wi};
}

template <bool INV>
NM_DEV cx<f32x2> nm_mulw(cx<f32x2> a, f32x2 wr, f32x2 wi) {  // packed pairs: explicit FFMA2
    return INV ? cx<f32x2>{nm_fma2(a.re, wr, a.im * wi), nm_fma2(a.im, wr, -(a.re * wi))}
               : cx<f32x2>{nm_fma2(a.re, wr, -(a.im * wi)), nm_fma2(a.im, wr, a.re * wi)};
}

template <bool INV, typename T>
NM_DEV void nm_r4(cx<T>& a0, cx<T>& a1, cx<T>& a2, cx<T>& a3) {
    const cx<T> t0 = cx_add(a0, a2), t1 = cx_sub(a0, a2), t2 = cx_add(a1, a3);
    const cx<T> t3 = cx_rot<T, INV>(cx_sub(a1, a3));
    a0 = cx_add(t0, t2);
    a2 = cx_sub(t0, t2);
    a1 = cx_add(t1, t3);
    a3 = cx_sub(t1, t3);
}

// 16-point DFT in registers, natural order in and out (4 x 4 Cooley-Tukey)
template <bool INV, typename T>
NM_DEV void nm_bfly16(cx<T>* v) {
    const T c = T(0.92387953251128675613), s = T(0.38268343236508977173), h = T(0.70710678118654752440);
#pragma unroll
    for (int b = 0; b < 4; ++b) nm_r4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);  // v[4*k1 + b] = u_b[k1]
    // twiddles w16^(b*k1)
    v[4 + 1] = nm_mulw<INV>(v[4 + 1], c, -s);
    v[8 + 1] = nm_mulw<INV>(v[8 + 1], h, -h);
    v[12 + 1] = nm_mulw<INV>(v[12 + 1], s, -c);
    v[4 + 2] = nm_mulw<INV>(v[4 + 2], h, -h);
    v[8 + 2] = cx_rot<T, INV>(v[8 + 2]);
    v[12 + 2] = nm_mulw<INV>(v[12 + 2], -h, -h);
    v[4 + 3] = nm_mulw<INV>(v[4 + 3], s, -c);
    v[8 + 3] = nm_mulw<INV>(v[8 + 3], -h, -h);
    v[12 + 3] = nm_mulw<INV>(v[12 + 3], -c, s);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) nm_r4<INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // v[4*k1 + k2] = X[k1 + 4*k2]
    // transpose to natural order
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; ++k2) {
            const cx<T> t = v[4 * k1 + k2];
            v[4 * k1 + k2] = v[4 * k2 + k1];
            v[4 * k2 + k1] = t;
        }
}

// 8-point DFT in registers (4 x 2)
template <bool INV, typename T>
NM_DEV void nm_bfly8(cx<T>* v) {
    const T h = T(0.70710678118654752440);
    nm_r4<INV>(v[0], v[2], v[4], v[6]);  // u_0[k1] at v[2*k1]
    nm_r4<INV>(v[1], v[3], v[5], v[7]);  // u_1[k1] at v[2*k1 + 1]
    v[3] = nm_mulw<INV>(v[3], h, -h);
    v[5] = cx_rot<T, INV>(v[5]);
    v[7] = nm_mulw<INV>(v[7], -h, -h);
    cx<T> x[8];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        x[k1] = cx_add(v[2 * k1], v[2 * k1 + 1]);
        x[k1 + 4] = cx_sub(v[2 * k1], v[2 * k1 + 1]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = x[k];
}

template <int R, bool INV, typename T>
NM_DEV void nm_bflyR(cx<T>* v) {
    if (R == 16) nm_bfly16<INV>(v);
    if (R == 8) nm_bfly8<INV>(v);
    if (R == 4) nm_r4<INV>(v[0], v[1], v[2], v[3]);
    if (R == 2) {
        const cx<T> a = v[0], b = v[1];
        v[0] = cx_add(a, b);
        v[1] = cx_sub(a, b);
    }
}

// v[k] *= w1^k (forward) or conj(w1)^k (inverse), k = 1..R-1; powers by products of depth <= log2(R)
template <int R, bool INV>
NM_DEV void nm_twiddle(cx<double>* v, const cx<double>* NM_RESTRICT tw, int idx1) {
#ifdef NM_TW_LOAD
    // variant: every power comes from the table (idx1 * k < P always holds, see nm_fft.cuh)
#pragma unroll
    for (int k = 1; k < R; ++k) {
        const cx<double> w = nm_ldg(tw + idx1 * k);
        v[k] = INV ? cx_mulc(v[k], w) : cx_mul(v[k], w);
    }
    return;
#endif
    cx<double> w1 = nm_ldg(tw + idx1);
    if (INV) w1.im = -w1.im;
    v[1] = cx_mul(v[1], w1);
    if (R > 2) {
        const cx<double> w2 = cx_mul(w1, w1), w3 = cx_mul(w2, w1);
        v[2] = cx_mul(v[2], w2);
        v[3] = cx_mul(v[3], w3);
        if (R > 4) {
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
    }
}

// One interior pass through shared memory: each thread runs 16/R butterflies of radix R on its own slots.
// Forward: butterfly, then twiddle.  Inverse: conjugate twiddle, then inverse butterfly.
template <int R, bool INV>
NM_DEV void nm_conv_pass(cx<double>* sm, const NmConv& f, int L, int tid) {
    const int m = L / R;  // power of two
    const int ts = f.P / L;
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        const int q = tid + f.NT * i;
        const int j = q & (m - 1);
        const int base = (q - j) * R + j;  // (q / m) * L + j
        cx<double> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = sm[NM_PHYS(base + t * m, f.pad)];
        if (INV) {
            if (m > 1 && j != 0) nm_twiddle<R, true>(v, f.tw, j * ts);
            nm_bflyR<R, true>(v);
        } else {
            nm_bflyR<R, false>(v);
            if (m > 1 && j != 0) nm_twiddle<R, false>(v, f.tw, j * ts);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) sm[NM_PHYS(base + t * m, f.pad)] = v[t];
    }
}

template <bool INV>
NM_DEV void nm_conv_pass_dispatch(cx<double>* sm, const NmConv& f, int p, int tid) {
    switch (f.radix[p]) {
        case 16: nm_conv_pass<16, INV>(sm, f, f.len[p], tid); break;
        case 8: nm_conv_pass<8, INV>(sm, f, f.len[p], tid); break;
        case 4: nm_conv_pass<4, INV>(sm, f, f.len[p], tid); break;
        default: nm_conv_pass<2, INV>(sm, f, f.len[p], tid); break;
    }
}

// Last forward pass (unit stride).  MODE 0: forward butterfly only (spectrum kept in `src` for a bank).
// MODE 1: forward butterfly, * H, inverse butterfly, in place (single filter).
// MODE 2: load spectrum from `src`, * H, inverse butterfly, store to `dst` (one filter of a bank).
template <int R, int MODE>
NM_DEV void nm_conv_last(cx<double>* dst, const cx<double>* src, const double* NM_RESTRICT h, const NmConv& f, int tid) {
#pragma unroll
    for (int i = 0; i < 16 / R; ++i) {
        const int base = (tid + f.NT * i) * R;
        cx<double> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[NM_PHYS(base + t, f.pad)];
        if (MODE != 2) nm_bflyR<R, false>(v);
        if (MODE != 0) {
#pragma unroll
            for (int t = 0; t < R; ++t) {
                const double hv = nm_ldg(h + base + t);
                v[t] = {v[t].re * hv, v[t].im * hv};
            }
            nm_bflyR<R, true>(v);
        }
#pragma unroll
        for (int t = 0; t < R; ++t) dst[NM_PHYS(base + t, f.pad)] = v[t];
    }
}

template <int MODE>
NM_DEV void nm_conv_last_dispatch(cx<double>* dst, const cx<double>* src, const double* h, const NmConv& f, int tid) {
    switch (f.radix[f.npass - 1]) {
        case 16: nm_conv_last<16, MODE>(dst, src, h, f, tid); break;
        case 8: nm_conv_last<8, MODE>(dst, src, h, f, tid); break;
        case 4: nm_conv_last<4, MODE>(dst, src, h, f, tid); break;
        default: nm_conv_last<2, MODE>(dst, src, h, f, tid); break;
    }
}

static NM_HD size_t nm_conv_buf_elems(int P, int pad) { return (size_t)P + ((size_t)P >> pad) + 2; }

template <class Epi>
NM_GLOBAL void NM_LAUNCH_BOUNDS(512, 1) nm_conv_kernel(NmConvArgs a, Epi epi) {  // NT = P/16 <= 512 (P <= 8192)
    NM_SHARED_BYTES(smem);
    const NmConv& f = a.fft;
    const int P = f.P, NT = f.NT;
    const size_t nbuf = nm_conv_buf_elems(P, f.pad);
    cx<double>* work = reinterpret_cast<cx<double>*>(smem);
    cx<double>* spec = (a.nF > 1) ? work + nbuf : work;
    unsigned char* scratch = a.scratch_in_tail ? reinterpret_cast<unsigned char*>(work + P)
                                               : reinterpret_cast<unsigned char*>(work + (a.nF > 1 ? 2 : 1) * nbuf);
    const int tid = threadIdx.x;
    const int W = a.in.W, E = a.E;
    const int npair = (a.in.n_ch + 1) >> 1;
    const int o0 = (a.mode == NM_FIR_REFLECT) ? E : 0;

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item / npair;
        const int c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        const double* r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
        const double* r1 = r0 + (has2 ? a.in.ch_stride : 0);

        // ---- pass 1 fused with the load of the (extended / zero padded) window
        cx<double> v[16];
        if (a.mode == NM_FIR_REFLECT) {
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
        if (tid != 0) nm_twiddle<16, false>(v, f.tw, tid);
#pragma unroll
        for (int t = 0; t < 16; ++t) spec[NM_PHYS(tid + NT * t, f.pad)] = v[t];
        __syncthreads();
        for (int p = 1; p < f.npass - 1; ++p) {
            nm_conv_pass_dispatch<false>(spec, f, p, tid);
            __syncthreads();
        }
        if (a.nF > 1) {
            nm_conv_last_dispatch<0>(spec, spec, nullptr, f, tid);
            __syncthreads();
        }

        for (int fi = 0; fi < a.nF; ++fi) {
            const double* h = a.hperm + (size_t)fi * P;
            if (a.nF > 1) nm_conv_last_dispatch<2>(work, spec, h, f, tid);
            else nm_conv_last_dispatch<1>(work, work, h, f, tid);
            __syncthreads();
            for (int p = f.npass - 2; p >= 1; --p) {
                nm_conv_pass_dispatch<true>(work, f, p, tid);
                __syncthreads();
            }
            // ---- final inverse pass: padded slots -> registers -> natural linear order for the epilogue
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = work[NM_PHYS(tid + NT * t, f.pad)];
            if (tid != 0) nm_twiddle<16, true>(v, f.tw, tid);
            nm_bfly16<true>(v);
            bool in_regs = false;
            if constexpr (Epi::kRegs) in_regs = epi.regs_ok();
            if (in_regs) {
                if constexpr (Epi::kRegs) epi.run_regs(v, o0, W, a.in.n_ch, w, c0, has2, fi + a.f0, scratch, tid, NT);
            } else {
                __syncthreads();
#pragma unroll
                for (int t = 0; t < 16; ++t) work[tid + NT * t] = v[t];
                __syncthreads();
                epi.run(work, o0, W, a.in.n_ch, w, c0, has2, fi + a.f0, scratch, tid, NT);
            }
            __syncthreads();
        }
    }
}
