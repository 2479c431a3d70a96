// nm_fft.cuh -- in-place mixed-radix complex FFT on a shared-memory buffer, cooperative over a CTA.
//
// Forward  = decimation in frequency (natural order in  -> digit-reversed order out)
// Inverse  = decimation in time      (digit-reversed in -> natural order out), unnormalised
//
// Running the inverse passes in the opposite order undoes the forward passes exactly, so a
// circular convolution is  inverse( forward(x) * Hperm )  with the filter spectrum stored in
// the same digit-reversed order (Hperm[pos[f]] = H[f] / n) -- no reordering pass is ever run.
// Where natural-order bins are needed (band power of a few DFT bins) they are read through
// the pos[] table.  Radices 2,3,4,5 have register butterflies; any other prime factor goes
// through a generic O(r^2) pass that needs a scratch buffer of n elements.
//
// Every butterfly reads and writes the same r slots, so one __syncthreads() per pass is all
// the synchronisation required and no ping-pong buffer is needed.
#pragma once

#include "nm_platform.h"

#define NM_MAX_PASS 20

template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T re, im;
};

#ifndef NM_EMULATE
// read-only path for twiddles: one 128-bit load per complex double
NM_DEV cx<double> nm_ldg(const cx<double>* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return {v.x, v.y};
}
NM_DEV cx<float> nm_ldg(const cx<float>* p) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return {v.x, v.y};
}
#endif

// ---- packed float32 pair: the arithmetic type of the float32 mode of the FIR kernels (nm_convx.cuh).
// Two INDEPENDENT transforms (channel pairs A and B of one item) travel side by side in the two halves of a 64-bit register pair,
// so every butterfly instruction is a Blackwell packed-f32 instruction (add / sub / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2 in
// SASS): half the issue slots of scalar FADD / FFMA for the same flops, and 16-byte (re, im) x 2 shared-memory elements -- the
// float64 kernel's conflict-free layout -- instead of 8-byte ones.
struct alignas(8) f32x2 {
    float x, y;
    f32x2() = default;
    NM_HD f32x2(float a, float b) : x(a), y(b) {}
    NM_HD explicit f32x2(double v) : x((float)v), y((float)v) {}  // broadcast (butterfly constants)
    NM_HD explicit f32x2(float v) : x(v), y(v) {}
};
#ifndef NM_EMULATE
NM_DEV unsigned long long nm_f2_bits(f32x2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
NM_DEV f32x2 nm_f2_from(unsigned long long u) { return *reinterpret_cast<f32x2*>(&u); }
NM_DEV f32x2 operator+(f32x2 a, f32x2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(nm_f2_bits(a)), "l"(nm_f2_bits(b)));
    return nm_f2_from(d);
}
NM_DEV f32x2 operator-(f32x2 a, f32x2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(nm_f2_bits(a)), "l"(nm_f2_bits(b)));
    return nm_f2_from(d);
}
NM_DEV f32x2 operator*(f32x2 a, f32x2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(nm_f2_bits(a)), "l"(nm_f2_bits(b)));
    return nm_f2_from(d);
}
NM_DEV f32x2 nm_fma2(f32x2 a, f32x2 b, f32x2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(nm_f2_bits(a)), "l"(nm_f2_bits(b)), "l"(nm_f2_bits(c)));
    return nm_f2_from(d);
}
#else
NM_DEV f32x2 operator+(f32x2 a, f32x2 b) { return {a.x + b.x, a.y + b.y}; }
NM_DEV f32x2 operator-(f32x2 a, f32x2 b) { return {a.x - b.x, a.y - b.y}; }
NM_DEV f32x2 operator*(f32x2 a, f32x2 b) { return {a.x * b.x, a.y * b.y}; }
NM_DEV f32x2 nm_fma2(f32x2 a, f32x2 b, f32x2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
#endif
NM_DEV f32x2 operator-(f32x2 a) { return {-a.x, -a.y}; }

template <typename T>
struct NmFft {
    int n;
    int npass;
    int radix[NM_MAX_PASS];
    int len[NM_MAX_PASS];  // block length L seen by pass p (forward order); len[0] == n
    const cx<T>* tw;       // tw[k] = exp(-2*pi*i*k/n), k in [0, n)
    const int* pos;        // pos[f] = slot that holds frequency f after the forward transform
};

template <typename T>
NM_DEV cx<T> cx_mul(cx<T> a, cx<T> b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
template <typename T>
NM_DEV cx<T> cx_mulc(cx<T> a, cx<T> b) { return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im}; }  // a * conj(b)
template <typename T>
NM_DEV cx<T> cx_add(cx<T> a, cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <typename T>
NM_DEV cx<T> cx_sub(cx<T> a, cx<T> b) { return {a.re - b.re, a.im - b.im}; }

// packed pairs: explicit fused multiply-adds (inline PTX is opaque to the compiler's own contraction)
NM_DEV cx<f32x2> cx_mul(cx<f32x2> a, cx<f32x2> b) { return {nm_fma2(a.re, b.re, -(a.im * b.im)), nm_fma2(a.re, b.im, a.im * b.re)}; }
NM_DEV cx<f32x2> cx_mulc(cx<f32x2> a, cx<f32x2> b) { return {nm_fma2(a.re, b.re, a.im * b.im), nm_fma2(a.im, b.re, -(a.re * b.im))}; }

// multiply by -i (forward) or +i (inverse)
template <typename T, bool INV>
NM_DEV cx<T> cx_rot(cx<T> a) { return INV ? cx<T>{-a.im, a.re} : cx<T>{a.im, -a.re}; }

template <typename T, bool INV>
NM_DEV void nm_bfly2(cx<T>* v) {
    cx<T> a = v[0], b = v[1];
    v[0] = cx_add(a, b);
    v[1] = cx_sub(a, b);
}

template <typename T, bool INV>
NM_DEV void nm_bfly3(cx<T>* v) {
    const T s = T(0.86602540378443864676372317075294);
    cx<T> p = cx_add(v[1], v[2]);
    cx<T> d = cx_rot<T, INV>(cx_sub(v[1], v[2]));  // -i*(a1-a2) forward
    cx<T> c = {v[0].re - T(0.5) * p.re, v[0].im - T(0.5) * p.im};
    v[0] = cx_add(v[0], p);
    v[1] = {c.re + s * d.re, c.im + s * d.im};
    v[2] = {c.re - s * d.re, c.im - s * d.im};
}

template <typename T, bool INV>
NM_DEV void nm_bfly4(cx<T>* v) {
    cx<T> t0 = cx_add(v[0], v[2]);
    cx<T> t1 = cx_sub(v[0], v[2]);
    cx<T> t2 = cx_add(v[1], v[3]);
    cx<T> t3 = cx_rot<T, INV>(cx_sub(v[1], v[3]));
    v[0] = cx_add(t0, t2);
    v[2] = cx_sub(t0, t2);
    v[1] = cx_add(t1, t3);
    v[3] = cx_sub(t1, t3);
}

template <typename T, bool INV>
NM_DEV void nm_bfly5(cx<T>* v) {
    const T c1 = T(0.30901699437494742410229341718282);   // cos(2pi/5)
    const T c2 = T(-0.80901699437494742410229341718282);  // cos(4pi/5)
    const T s1 = T(0.95105651629515357211643933337938);   // sin(2pi/5)
    const T s2 = T(0.58778525229247312916870595463907);   // sin(4pi/5)
    cx<T> p14 = cx_add(v[1], v[4]), m14 = cx_sub(v[1], v[4]);
    cx<T> p23 = cx_add(v[2], v[3]), m23 = cx_sub(v[2], v[3]);
    cx<T> a = {v[0].re + c1 * p14.re + c2 * p23.re, v[0].im + c1 * p14.im + c2 * p23.im};
    cx<T> b = {v[0].re + c2 * p14.re + c1 * p23.re, v[0].im + c2 * p14.im + c1 * p23.im};
    cx<T> u = cx_rot<T, INV>(cx<T>{s1 * m14.re + s2 * m23.re, s1 * m14.im + s2 * m23.im});
    cx<T> w = cx_rot<T, INV>(cx<T>{s2 * m14.re - s1 * m23.re, s2 * m14.im - s1 * m23.im});
    v[0] = {v[0].re + p14.re + p23.re, v[0].im + p14.im + p23.im};
    v[1] = cx_add(a, u);
    v[4] = cx_sub(a, u);
    v[2] = cx_add(b, w);
    v[3] = cx_sub(b, w);
}

template <typename T, int R, bool INV>
NM_DEV void nm_bfly(cx<T>* v) {
    if (R == 2) nm_bfly2<T, INV>(v);
    if (R == 3) nm_bfly3<T, INV>(v);
    if (R == 4) nm_bfly4<T, INV>(v);
    if (R == 5) nm_bfly5<T, INV>(v);
}

// One pass with a register butterfly.  Forward: butterfly then twiddle; inverse: conj-twiddle then butterfly.
template <typename T, int R, bool INV>
NM_DEV void nm_fft_pass_r(cx<T>* a, int n, int L, const cx<T>* NM_RESTRICT tw, int tid, int nt) {
    const int m = L / R;
    const int nb = n / R;
    const int ts = n / L;
    for (int q = tid; q < nb; q += nt) {
        const int b = q / m;
        const int j = q - b * m;
        cx<T>* p = a + (size_t)b * L + j;
        cx<T> v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = p[t * m];
        if (INV) {
            if (j != 0) {
                const cx<T> w1 = nm_ldg(tw + j * ts);
                cx<T> w = w1;
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    v[t] = cx_mulc(v[t], w);
                    if (t + 1 < R) w = cx_mul(w, w1);
                }
            }
            nm_bfly<T, R, true>(v);
        } else {
            nm_bfly<T, R, false>(v);
            if (j != 0) {
                const cx<T> w1 = nm_ldg(tw + j * ts);
                cx<T> w = w1;
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    v[t] = cx_mul(v[t], w);
                    if (t + 1 < R) w = cx_mul(w, w1);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < R; ++t) p[t * m] = v[t];
    }
}

// Generic prime radix: out-of-place into scratch, then copied back (two extra barriers).
template <typename T, bool INV>
NM_DEV void nm_fft_pass_generic(cx<T>* a, cx<T>* scratch, int n, int L, int R, const cx<T>* NM_RESTRICT tw, int tid, int nt) {
    const int m = L / R;
    const int nb = n / R;
    const int ts = n / L;
    const int tr = n / R;
    for (int q = tid; q < nb; q += nt) {
        const int b = q / m;
        const int j = q - b * m;
        const cx<T>* p = a + (size_t)b * L + j;
        cx<T>* o = scratch + (size_t)b * L + j;
        for (int k = 0; k < R; ++k) {
            cx<T> acc = {T(0), T(0)};
            for (int t = 0; t < R; ++t) {
                cx<T> x = p[t * m];
                const cx<T> wr = nm_ldg(tw + ((k * t) % R) * tr);
                if (INV) {
                    const cx<T> wl = nm_ldg(tw + j * t * ts);
                    x = cx_mulc(x, wl);
                    x = cx_mulc(x, wr);
                } else {
                    x = cx_mul(x, wr);
                }
                acc = cx_add(acc, x);
            }
            if (!INV) acc = cx_mul(acc, nm_ldg(tw + j * k * ts));
            o[k * m] = acc;
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) a[i] = scratch[i];
}

template <typename T, bool INV>
NM_DEV void nm_fft_pass(cx<T>* a, cx<T>* scratch, const NmFft<T>& f, int p, int tid, int nt) {
    const int R = f.radix[p], L = f.len[p];
    switch (R) {
        case 2: nm_fft_pass_r<T, 2, INV>(a, f.n, L, f.tw, tid, nt); break;
        case 3: nm_fft_pass_r<T, 3, INV>(a, f.n, L, f.tw, tid, nt); break;
        case 4: nm_fft_pass_r<T, 4, INV>(a, f.n, L, f.tw, tid, nt); break;
        case 5: nm_fft_pass_r<T, 5, INV>(a, f.n, L, f.tw, tid, nt); break;
        default: nm_fft_pass_generic<T, INV>(a, scratch, f.n, L, R, f.tw, tid, nt); break;
    }
}

// Caller must have synchronised the CTA on the contents of `a`.  On return the CTA is synchronised.
template <typename T>
NM_DEV void nm_fft_forward(cx<T>* a, cx<T>* scratch, const NmFft<T>& f, int tid, int nt) {
    for (int p = 0; p < f.npass; ++p) {
        nm_fft_pass<T, false>(a, scratch, f, p, tid, nt);
        __syncthreads();
    }
}

template <typename T>
NM_DEV void nm_fft_inverse(cx<T>* a, cx<T>* scratch, const NmFft<T>& f, int tid, int nt) {
    for (int p = f.npass - 1; p >= 0; --p) {
        nm_fft_pass<T, true>(a, scratch, f, p, tid, nt);
        __syncthreads();
    }
}
