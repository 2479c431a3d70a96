// nm_fused.cuh -- the whole per-window chain of the linear families in ONE persistent kernel.
//
//   raw rows (float32 / float64, as uploaded) + per-sample group sums S_g[t]
//        --cp.async.bulk + mbarrier-->  shared-memory stage (next item's rows land while the current item computes)
//        --nan_to_num, pick, re-reference folded into the load:  x_i[t] = d_i * nan_to_num(raw_i[t]) + g_i * S[t]
//        --odd-reflect extension, P-point forward, * H_notch, inverse-->  notched window (registers)
//        --> Hjorth / line length / raw          (features/hjorth_raw.py:24-57, features/linelength.py:11-21)
//        --> N-point segment DFT band features   (FFT / Welch, features/oscillatory.py:58-182)
//        --> zero-padded forward, nF x (* H_band, inverse, tail variance)   (features/bandpower.py:165-207)
//
// One (window, channel pair) item per CTA iteration, the two real rows packed as one complex signal, same register-blocked
// transform plan as nm_convx_kernel.  Compared with the chain  nm_prep_kernel -> nm_convx<notch> -> nm_specx -> nm_convx<bank>
// the re-referenced recording (`xr`, float64 copy of the recording) and the notched rows (`Y`, n x C x W float64 per chunk)
// never exist in HBM: the kernel reads each window's raw samples once (L2 serves the 90 % overlap between windows) and
// writes only feature columns.  `Y` is still written when a family outside this kernel (bursts, sharp waves, STFT, ...)
// consumes the notched rows.
//
// Shared memory: two transform buffers X / Y of NBUF complex values that swap roles every item.
//   X: stage of this item  -> natural-order notched window (+ DFT bin values in its tail) -> inverse-transform work buffer of the bank
//   Y: notch transform     -> segment-DFT buffer -> spectrum of the bank's forward transform -> stage of the NEXT item
// The next item's bulk copies are issued (one elected thread) as soon as Y is dead: after the last band's spectrum multiply
// when there is a bank -- the last inverse transform hides the copy latency -- otherwise at the end of the item.
#pragma once

#include <type_traits>

#include "nm_convx.cuh"
#include "nm_specx.cuh"

// ---------------------------------------------------------------- async bulk copy (TMA engine, 1-D) + mbarrier
#ifndef NM_EMULATE
NM_DEV unsigned nm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
NM_DEV void nm_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nm_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy accesses to shared memory (ordered before this thread by a CTA barrier) -> async-proxy writes
NM_DEV void nm_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NM_DEV void nm_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nm_smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes % 16 == 0, dst and src 16-byte aligned; completion is signalled on `bar` (complete_tx)
NM_DEV void nm_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nm_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(nm_smem_u32(bar))
                 : "memory");
}
NM_DEV void nm_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned addr = nm_smem_u32(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(addr), "r"(parity)
                     : "memory");
    } while (!ok);
}
#else
// test build: the elected fiber copies synchronously; the wait is a CTA barrier (every thread waits at the same point)
NM_DEV void nm_mbar_init(unsigned long long* bar, unsigned) { *bar = 0; }
NM_DEV void nm_fence_proxy_async() {}
NM_DEV void nm_mbar_expect_tx(unsigned long long*, unsigned) {}
NM_DEV void nm_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
NM_DEV void nm_mbar_wait(unsigned long long*, unsigned) { __syncthreads(); }
#endif

#define NM_FZ_MAX_SPEC 2

struct NmFusedArgs {
    // recording as uploaded + folded preprocessing
    const void* raw;         // (C_all, raw_pitch) float32 / float64
    long long raw_pitch;     // elements, multiple of 4
    const int* pick;         // [n_ch] raw row of feature channel j
    const double* dcoef;     // [n_ch] coefficient on the channel's own samples
    const double* gcoef;     // [n_ch] coefficient on the group sum (0 where the channel is not re-referenced)
    const double* gsum;      // [gsum_pitch] per-sample group sum (nan_to_num'ed), or nullptr
    const long long* start;  // [n_windows] first sample of each window of the batch
    int n_windows, n_ch, W, E, n_items;
    // notch (single filter, reflect-limited padding) and the optional band-pass bank ('same' padding), same plan P
    const cx<double>* tw;    // exp(-2*pi*i*k/P)
    const double* hx_notch;  // nm_cx_load_h order
    const double* hx_bank;   // [nF][P], or nullptr
    int nF;
    // epilogues
    NmEpiStoreScan scan;     // y == nullptr: the notched rows are not needed outside this kernel
    NmEpiBandpower bp;
    // in-kernel segment DFTs (nseg == 1, nper == SX::N, segment inside the window): argument blocks in DEVICE memory (a dynamically
    // indexed kernel parameter would be copied to local memory), written once at nm_finalize with out.row0 == 0
    int n_spec;
    const NmSpecArgs* spec;
    long long row0;          // first output row of this batch
};

// stage layout (bytes): [group sums: cs doubles][row 0: cx elements][row 1: cx elements], all sizes multiples of 16 bytes
template <bool RAW64>
struct NmFzStage {
    static constexpr int ESZ = RAW64 ? 8 : 4;
    static constexpr int XA = RAW64 ? 2 : 4;  // elements per 16 bytes
    static NM_HD int cs(int W) { return (W + 1 + 1) & ~1; }
    static NM_HD int cxe(int W) { return (W + XA - 1 + XA - 1) & ~(XA - 1); }
    static NM_HD size_t bytes(int W) { return (size_t)cs(W) * 8 + (size_t)2 * cxe(W) * ESZ; }
};

// elected thread: bulk copies of item `item` into `stage`, completion on `bar`
template <bool RAW64>
NM_DEV void nm_fz_issue(const NmFusedArgs& a, int item, int npair, unsigned char* stage, unsigned long long* bar) {
    using ST = NmFzStage<RAW64>;
    const int W = a.W;
    const int w = item / npair, c0 = (item - w * npair) * 2;
    const bool has2 = c0 + 1 < a.n_ch;
    const long long s = nm_ldg(a.start + w);
    const int lead_s = (int)(s & 1), lead_x = (int)(s & (ST::XA - 1));
    const unsigned bs = a.gsum ? (unsigned)(((lead_s + W + 1) & ~1) * 8) : 0u;
    const unsigned bx = (unsigned)(((lead_x + W + ST::XA - 1) & ~(ST::XA - 1)) * ST::ESZ);
    const char* raw = reinterpret_cast<const char*>(a.raw);
    nm_fence_proxy_async();
    nm_mbar_expect_tx(bar, bs + bx * (has2 ? 2u : 1u));
    if (a.gsum) nm_bulk_g2s(stage, a.gsum + (s - lead_s), bs, bar);
    unsigned char* x0 = stage + (size_t)ST::cs(W) * 8;
    nm_bulk_g2s(x0, raw + ((size_t)nm_ldg(a.pick + c0) * a.raw_pitch + (size_t)(s - lead_x)) * ST::ESZ, bx, bar);
    if (has2)
        nm_bulk_g2s(x0 + (size_t)ST::cxe(W) * ST::ESZ, raw + ((size_t)nm_ldg(a.pick + c0 + 1) * a.raw_pitch + (size_t)(s - lead_x)) * ST::ESZ, bx,
                    bar);
}

// numpy.nan_to_num of a raw sample, widened to float64: float32 recordings are tested in float32 (NaN -> 0; +-inf -> +-DBL_MAX
// like nan_to_num of the up-cast value)
NM_DEV double nm_fz_clean(double x) { return nm_nan_to_num(x); }
NM_DEV double nm_fz_clean(float x) {
    const float ax = fabsf(x);
    if (!(ax <= 3.402823466e38f)) return (x != x) ? 0.0 : (x > 0.0f ? NM_DBL_MAX : -NM_DBL_MAX);
    return (double)x;
}

// stage -> the 16 pass-0 inputs of a thread: nan_to_num, re-reference, odd reflection about both end samples
// (same arithmetic as nm_prep_kernel followed by nm_cx_load_item<REFLECT>)
template <int P, bool RAW64>
NM_DEV void nm_fz_load(cx<double>* v, const NmFusedArgs& a, const unsigned char* stage, long long s, int c0, bool has2, int tid) {
    using ST = NmFzStage<RAW64>;
    using RT = typename std::conditional<RAW64, double, float>::type;
    constexpr int NT = NmCxPlan<P>::NT;
    const int W = a.W, E = a.E;
    const bool has_g = a.gsum != nullptr;
    const double* S = reinterpret_cast<const double*>(stage) + (int)(s & 1);
    const RT* xa = reinterpret_cast<const RT*>(stage + (size_t)ST::cs(W) * 8) + (int)(s & (ST::XA - 1));
    const RT* xb = xa + ST::cxe(W);
    const double d0 = nm_ldg(a.dcoef + c0), g0 = nm_ldg(a.gcoef + c0);
    const double d1 = has2 ? nm_ldg(a.dcoef + c0 + 1) : 0.0, g1 = has2 ? nm_ldg(a.gcoef + c0 + 1) : 0.0;
    auto sample = [&](int idx, double& ra, double& rb) {
        const double sg = has_g ? S[idx] : 0.0;
        ra = fma(d0, nm_fz_clean(xa[idx]), g0 * sg);
        rb = fma(d1, nm_fz_clean(has2 ? xb[idx] : RT(0)), g1 * sg);
    };
    double a0, b0, a1, b1;
    sample(0, a0, b0);
    sample(W - 1, a1, b1);
    a0 *= 2.0; b0 *= 2.0; a1 *= 2.0; b1 *= 2.0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int n = tid + NT * t;
        const bool left = n < E, mid = !left && n < E + W, right = !left && !mid && n < W + 2 * E;
        int idx = left ? E - n : (mid ? n - E : 2 * W + E - 2 - n);
        idx = (left || mid || right) ? idx : 0;
        const double sgn = mid ? 1.0 : ((left || right) ? -1.0 : 0.0);
        const double ca = left ? a0 : (right ? a1 : 0.0), cb = left ? b0 : (right ? b1 : 0.0);
        double ra, rb;
        sample(idx, ra, rb);
        v[t] = {fma(sgn, ra, ca), has2 ? fma(sgn, rb, cb) : 0.0};
    }
}

// one in-kernel segment DFT family (FFT or single-segment Welch) on the natural-order notched window `nat`
// (sample u at nat[u + (u >> 3)]); `buf` = SX::NBUF elements that do not overlap `nat`, `vals` = [2][nk] doubles.
// Ends with a barrier; all threads of the CTA call it.
template <class SX>
NM_DEV void nm_fz_spectral(const NmSpecArgs& a, const cx<double>* nat, cx<double>* buf, double* vals, double* red, int w, int c0,
                           bool has2, int tid, int nt) {
    constexpr int N = SX::N, R0 = SX::R0, R1 = SX::R1, R2 = SX::R2, NA = SX::NA;
    const bool active = tid < NA;
    const cx<double>* NM_RESTRICT tw = a.fft.tw;
    for (int i = tid; i < 2 * a.nk; i += nt) vals[i] = 0.0;
    cx<double> v[R0];
    double sum[2] = {0.0, 0.0};
    if (active) {
#pragma unroll
        for (int t = 0; t < R0; ++t) {
            const int u = a.start + tid + NA * t;
            v[t] = nat[NmEpiStoreScan::phys(u)];
            sum[0] += v[t].re;
            sum[1] += v[t].im;
        }
    }
    if (a.detrend) {
        nm_block_sum<2>(sum, red, tid, nt);
        sum[0] /= N;
        sum[1] /= N;
    } else {
        sum[0] = sum[1] = 0.0;
    }
    if (active) {
        if (a.detrend || a.win) {
#pragma unroll
            for (int t = 0; t < R0; ++t) {
                const double wv = a.win ? nm_ldg(a.win + tid + NA * t) : 1.0;
                v[t] = {(v[t].re - sum[0]) * wv, (v[t].im - sum[1]) * wv};
            }
        }
        nm_dft_reg<R0>(v);
        nm_twiddle_pow<R0>(v, nm_ldg(tw + tid));
#pragma unroll
        for (int k = 0; k < R0; ++k) buf[SX::phys(tid + NA * k)] = v[k];
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R1; ++i) {
            const int q = tid + NA * i;
            const int blk = q / SX::M1, j = q - blk * SX::M1;
            const int e0 = blk * SX::L1 + j;
            cx<double> u[R1];
#pragma unroll
            for (int t = 0; t < R1; ++t) u[t] = buf[SX::phys(e0 + t * SX::M1)];
            nm_dft_reg<R1>(u);
            nm_twiddle_pow<R1>(u, nm_ldg(tw + j * R0));
#pragma unroll
            for (int t = 0; t < R1; ++t) buf[SX::phys(e0 + t * SX::M1)] = u[t];
        }
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < R0 / R2; ++i) {
            const int e0 = (tid + NA * i) * R2;
            cx<double> u[R2];
#pragma unroll
            for (int t = 0; t < R2; ++t) u[t] = buf[SX::phys(e0 + t)];
            nm_dft_reg<R2>(u);
#pragma unroll
            for (int t = 0; t < R2; ++t) buf[SX::phys(e0 + t)] = u[t];
        }
    }
    __syncthreads();
    for (int i = tid; i < a.nk; i += nt) {
        const int k = a.k0 + i;
        const cx<double> U = buf[SX::phys(nm_ldg(a.fft.pos + k))];
        const cx<double> V = buf[SX::phys(nm_ldg(a.fft.pos + (k == 0 ? 0 : N - k)))];
        nm_spec_bin(a, vals, i, k, 0, U, V);
    }
    __syncthreads();
    nm_spec_finish(a, vals, 1, w, c0, has2, tid, nt);
    __syncthreads();
}

struct NmSxNone {  // no in-kernel segment DFT for this window length
    static constexpr int N = 0, NBUF = 0;
};

template <int P, bool BANK>
struct NmFzOcc {
    static constexpr int NT = NmCxPlan<P>::NT;
    static constexpr int value = (NT >= 256) ? 1 : (NT == 128 ? 3 : 6);
};

// shared memory: X | Y | reduction scratch | mbarrier
template <int P>
static NM_HD size_t nm_fused_smem_bytes() {
    return (size_t)2 * NmCxPlan<P>::NBUF * sizeof(cx<double>) + NM_CX_RED_BYTES + 16;
}

template <int P, class SX, bool RAW64>
NM_GLOBAL void NM_LAUNCH_BOUNDS(NmCxPlan<P>::NT, (NmFzOcc<P, true>::value)) nm_fused_kernel(NmFusedArgs a) {
    using PL = NmCxPlan<P>;
    using T = double;
    constexpr int NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<T>* X = reinterpret_cast<cx<T>*>(smem);
    cx<T>* Y = X + PL::NBUF;
    double* red = reinterpret_cast<double*>(X + 2 * PL::NBUF);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(red) + NM_CX_RED_BYTES);
    const int tid = threadIdx.x;
    const int W = a.W, E = a.E;
    const int npair = (a.n_ch + 1) >> 1;
    const bool has_bank = a.hx_bank != nullptr;
    const cx<T>* NM_RESTRICT tw = a.tw;
    const cx<T> wA = nm_ldg(tw + tid);
    const cx<T> wB = nm_ldg(tw + (tid & (PL::M1 - 1)) * 16);
    const int poff = tid + (tid >> PL::PAD);
    // natural-order window in X: nat_elems complex values, the DFT bin values of the segment transforms behind it
    const int nat_elems = (NmEpiStoreScan::phys(W - 1) + 2) & ~1;

    int item = blockIdx.x;
    if (item >= a.n_items) return;
    if (tid == 0) nm_mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) nm_fz_issue<RAW64>(a, item, npair, reinterpret_cast<unsigned char*>(X), bar);
    unsigned parity = 0;

    cx<T> v[16];
    T hv[16];
    while (item < a.n_items) {
        const int next = item + gridDim.x;
        const int w = item / npair, c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.n_ch;
        cx<T>* const p0y = Y + poff;

        // Every transform pass exists ONCE in the instruction stream (the kernel is instruction-cache sensitive): job 0 is the notch
        // (forward of the reflected stage, * H_notch, inverse), job 1 adds the bank's forward transform of the notched window,
        // jobs 1..nF are the bands (* H_band, inverse, tail variance).
        const int n_jobs = 1 + (has_bank ? a.nF : 0);
#pragma unroll 1
        for (int job = 0; job < n_jobs; ++job) {
            const bool notch = job == 0;
            if (job <= 1) {
                // ---- forward transform into Y: pass-0 inputs from the stage (notch) or from the natural-order window (bank)
                if (notch) {
                    nm_mbar_wait(bar, parity);
                    parity ^= 1u;
                    nm_fz_load<P, RAW64>(v, a, reinterpret_cast<const unsigned char*>(X), nm_ldg(a.start + w), c0, has2, tid);
                } else {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int n = tid + NT * t;
                        const cx<T> z = {0.0, 0.0};
                        v[t] = n < W ? X[NmEpiStoreScan::phys(n < W ? n : 0)] : z;
                    }
                }
                nm_bfly16<false>(v);
                nm_twiddle_w1<16, false>(v, wA);
#pragma unroll
                for (int t = 0; t < 16; ++t) p0y[t * PL::S0] = v[t];
                __syncthreads();  // (notch: every thread is done with the stage in X)
                nm_cx_load_h<PL, T>(hv, notch ? a.hx_notch : a.hx_bank, tid);  // lands while pass 1 computes
                nm_cx_pass1<PL, false>(Y, wB, tid);
                __syncthreads();
                if (!notch) {
                    nm_cx_pass2<PL, 0, T>(Y, Y, nullptr, tid);  // the bank's spectrum stays in Y
                    __syncthreads();
                }
            }
            // ---- * H and first inverse pass: in place (notch) or Y -> X (band)
            cx<T>* const B = notch ? Y : X;
            if (notch) {
                nm_cx_pass2<PL, 1, T>(Y, Y, hv, tid);
            } else {
                nm_cx_pass2<PL, 2, T>(X, Y, hv, tid);
            }
            __syncthreads();
            // Y is dead after the last band's multiply: the next item's rows land while the last inverse transform runs
            if (has_bank && job + 1 == n_jobs && tid == 0 && next < a.n_items)
                nm_fz_issue<RAW64>(a, next, npair, reinterpret_cast<unsigned char*>(Y), bar);
            nm_cx_pass1<PL, true>(B, wB, tid);
            __syncthreads();
            {
                cx<T>* const p0b = B + poff;
#pragma unroll
                for (int t = 0; t < 16; ++t) v[t] = p0b[t * PL::S0];
            }
            nm_twiddle_w1<16, true>(v, wA);
            nm_bfly16<true>(v);
            // next band's spectrum: issued where the register pressure is lowest (a load that the compiler has to spill right
            // away stalls on its own latency); it lands during the epilogue's reduction and barrier
            if (!notch && job + 1 < n_jobs) nm_cx_load_h<PL, T>(hv, a.hx_bank + (size_t)job * P, tid);

            if (notch) {
                // ---- notched window: registers -> (HBM rows for the families outside this kernel) + natural order in X
                if (a.scan.y) {
                    double* r0 = a.scan.y + ((size_t)w * a.n_ch + c0) * a.scan.Wp;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int t = tid + NT * k - E;
                        if (t >= 0 && t < W) {
                            r0[t] = v[k].re;
                            if (has2) r0[a.scan.Wp + t] = v[k].im;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int u = tid + NT * k - E;
                    if (u >= 0 && u < W) X[NmEpiStoreScan::phys(u)] = v[k];
                }
                // Hjorth / line length / raw (begins with the barrier that completes the natural-order copy)
                if (a.scan.want_scan) {
                    NmEpiStoreScan::State st;
                    a.scan.template finish<PL, T>(X, red, st, E, W, a.n_ch, w, c0, has2, 0, tid);
                } else {
                    __syncthreads();
                }
                // segment DFTs (FFT / Welch band features): X natural -> Y
                if constexpr (SX::N != 0) {
#pragma unroll 1
                    for (int si = 0; si < a.n_spec; ++si)
                        nm_fz_spectral<SX>(a.spec[si], X, Y, reinterpret_cast<double*>(X + nat_elems), red, w + (int)a.row0, c0, has2, tid, NT);
                }
            } else {
                NmEpiBandpower::State st;
                a.bp.template consume<PL, T>(v, X, red, st, 0, W, a.n_ch, w, c0, has2, job - 1, tid);
                a.bp.template finish<PL, T>(X, red, st, 0, W, a.n_ch, w, c0, has2, job - 1, tid);  // (one barrier inside)
            }
        }
        if (!has_bank && tid == 0 && next < a.n_items) nm_fz_issue<RAW64>(a, next, npair, reinterpret_cast<unsigned char*>(Y), bar);
        cx<T>* const sw = X;
        X = Y;
        Y = sw;
        item = next;
    }
}

// ================================================================ the "front" kernel: everything up to the notched window
// Default organisation of the window chain since round 2.  nm_fused_kernel above keeps the whole chain in one CTA but pays for
// it with two transform buffers and 168 registers (3 CTAs per SM, measured slower than the staged kernels).  This kernel takes
// the part that profits from staying on chip and leaves the band-pass bank to nm_convx_kernel<.., BANK, NmEpiBandpower>:
//
//   raw rows (f32 / f64 as uploaded) + group sums --cp.async.bulk + mbarrier--> STAGE (its own small region, not a transform
//   buffer) --nan_to_num, pick, folded re-reference, odd reflection--> forward, * H_notch, inverse --> registers
//        --> notched rows to HBM only if a later family reads them (band power, bursts, sharp waves, STFT)
//        --> Hjorth / line length / raw and the N-point segment DFT band features (FFT / Welch) from the on-chip window
//
// One transform buffer + stage: 53 KB per CTA at P = 2048 with float32 recordings -> 4 CTAs per SM at <= 128 registers.  The stage
// is dead as soon as every thread has loaded its pass-0 inputs, so the NEXT item's bulk copies are issued behind the first
// barrier of the current item and have the whole item to land (no register prefetch, no exposed global-load latency).
// Replaces nm_prep_kernel's float64 copy of the recording, the notch kernel and the spectral kernels of the staged path.
template <int P, bool RAW64>
static NM_HD size_t nm_front_smem_bytes(int W) {
    return (size_t)NmCxPlan<P>::NBUF * sizeof(cx<double>) + ((NmFzStage<RAW64>::bytes(W) + 15) & ~(size_t)15) + NM_CX_RED_BYTES + 16;
}

#ifndef NM_FRONT_MINB128
#define NM_FRONT_MINB128 4  // resident CTAs per SM the 128-thread plan (P = 2048) with float32 recordings is register-limited to
#endif
template <int P, bool RAW64>
struct NmFrontOcc {
    static constexpr int NT = NmCxPlan<P>::NT;
    static constexpr int value = (NT >= 256) ? (RAW64 ? 1 : 2) : (NT == 128 ? (RAW64 ? 3 : NM_FRONT_MINB128) : 8);
};

template <int P, class SX, bool RAW64>
NM_GLOBAL void NM_LAUNCH_BOUNDS(NmCxPlan<P>::NT, (NmFrontOcc<P, RAW64>::value)) nm_front_kernel(NmFusedArgs a) {
    using PL = NmCxPlan<P>;
    using T = double;
    constexpr int NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<T>* work = reinterpret_cast<cx<T>*>(smem);
    unsigned char* stage = reinterpret_cast<unsigned char*>(work + PL::NBUF);
    const int W = a.W, E = a.E;
    double* red = reinterpret_cast<double*>(stage + ((NmFzStage<RAW64>::bytes(W) + 15) & ~(size_t)15));
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(red) + NM_CX_RED_BYTES);
    const int tid = threadIdx.x;
    const int npair = (a.n_ch + 1) >> 1;
    const cx<T>* NM_RESTRICT tw = a.tw;
    const cx<T> wA = nm_ldg(tw + tid);
    const cx<T> wB = nm_ldg(tw + (tid & (PL::M1 - 1)) * 16);
    cx<T>* const p0w = work + tid + (tid >> PL::PAD);
    const int nat_elems = (NmEpiStoreScan::phys(W - 1) + 2) & ~1;

    int item = blockIdx.x;
    if (item >= a.n_items) return;
    if (tid == 0) nm_mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) nm_fz_issue<RAW64>(a, item, npair, stage, bar);
    unsigned parity = 0;

    cx<T> v[16];
    T hv[16];
    while (item < a.n_items) {
        const int next = item + gridDim.x;
        const int w = item / npair, c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.n_ch;
        // ---- pass 0 straight from the stage
        nm_mbar_wait(bar, parity);
        parity ^= 1u;
        nm_fz_load<P, RAW64>(v, a, stage, nm_ldg(a.start + w), c0, has2, tid);
        nm_bfly16<false>(v);
        nm_twiddle_w1<16, false>(v, wA);
#pragma unroll
        for (int t = 0; t < 16; ++t) p0w[t * PL::S0] = v[t];
        __syncthreads();  // every thread is done with the stage: the next item's rows may land from here on
        if (tid == 0 && next < a.n_items) nm_fz_issue<RAW64>(a, next, npair, stage, bar);
        nm_cx_load_h<PL, T>(hv, a.hx_notch, tid);  // (L1 resident) lands while pass 1 computes
        nm_cx_pass1<PL, false>(work, wB, tid);
        __syncthreads();
        nm_cx_pass2<PL, 1, T>(work, work, hv, tid);
        __syncthreads();
        nm_cx_pass1<PL, true>(work, wB, tid);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = p0w[t * PL::S0];
        nm_twiddle_w1<16, true>(v, wA);
        nm_bfly16<true>(v);
        // ---- notched window: registers -> HBM rows (only for families outside this kernel) + natural order on chip
        if (a.scan.y) {
            double* r0 = a.scan.y + ((size_t)w * a.n_ch + c0) * a.scan.Wp;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int t = tid + NT * k - E;
                if (t >= 0 && t < W) {
                    r0[t] = v[k].re;
                    if (has2) r0[a.scan.Wp + t] = v[k].im;
                }
            }
        }
        const bool on_chip = a.scan.want_scan || (SX::N != 0 && a.n_spec > 0);
        __syncthreads();  // every thread has read its final-pass inputs from `work`
        if (on_chip) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int u = tid + NT * k - E;
                if (u >= 0 && u < W) work[NmEpiStoreScan::phys(u)] = v[k];
            }
            if (a.scan.want_scan) {
                NmEpiStoreScan::State st;
                a.scan.template finish<PL, T>(work, red, st, E, W, a.n_ch, w, c0, has2, 0, tid);  // (begins with a barrier)
            } else {
                __syncthreads();
            }
            if constexpr (SX::N != 0) {
#pragma unroll 1
                for (int si = 0; si < a.n_spec; ++si)
                    nm_fz_spectral<SX>(a.spec[si], work, work + nat_elems, reinterpret_cast<double*>(work + nat_elems + SX::NBUF), red,
                                       w + (int)a.row0, c0, has2, tid, NT);  // (ends with a barrier)
            }
            // (no trailing barrier: the scan epilogue ends with every thread past its reads of `work`, the DFT ends with a barrier)
        }
        item = next;
    }
}

// ================================================================ notch kernel with bulk-copy staged rows
// nm_convx_kernel<double, P, REFLECT, single filter, NmEpiStoreScan> with ONE change: the two float64 rows of the NEXT item are
// brought into a shared-memory stage by the TMA engine (cp.async.bulk + mbarrier, issued by one elected thread behind the first
// barrier of the current item) instead of being prefetched into registers during the epilogue.  Same arithmetic on the same
// values; the rows are whatever the staged path feeds the notch (re-referenced recording, pre-filter chunk rows).  Frees the
// prefetch registers (32 bytes of stack instead of 288 at the 128-register cap) and takes the global-load latency off the
// critical path: 11.72 -> 11.08 ms per C3 step on one B200 (profiles/r2_ab_window_chain.txt).
template <int P>
static NM_HD size_t nm_notchx_stage_bytes(int W) { return (size_t)2 * ((W + 1 + 1) & ~1) * sizeof(double); }
template <int P>
static NM_HD size_t nm_notchx_smem_bytes(int W) {
    return (size_t)NmCxPlan<P>::NBUF * sizeof(cx<double>) + nm_notchx_stage_bytes<P>(W) + NM_CX_RED_BYTES + 16;
}

// elected thread: rows (c0, c0 + 1) of window w -> stage; every copy starts at the 16-byte aligned sample at or below the window
NM_DEV void nm_nx_issue(const NmConvArgs& a, int item, int npair, unsigned char* stage, unsigned long long* bar) {
    const int W = a.in.W;
    const int w = item / npair, c0 = (item - w * npair) * 2;
    const bool has2 = c0 + 1 < a.in.n_ch;
    const long long s = nm_ldg(a.in.off + w);
    const int lead = (int)(s & 1);
    const unsigned bx = (unsigned)(((lead + W + 1) & ~1) * 8);
    const int rowe = (W + 1 + 1) & ~1;
    const double* r0 = a.in.base + (size_t)c0 * a.in.ch_stride + (s - lead);
    nm_fence_proxy_async();
    nm_mbar_expect_tx(bar, bx * (has2 ? 2u : 1u));
    nm_bulk_g2s(stage, r0, bx, bar);
    if (has2) nm_bulk_g2s(stage + (size_t)rowe * 8, r0 + a.in.ch_stride, bx, bar);
}

template <int P>
NM_GLOBAL void NM_LAUNCH_BOUNDS(NmCxPlan<P>::NT, (NmCxPlan<P>::MINB1)) nm_notchx_kernel(NmConvArgs a, NmEpiStoreScan epi) {
    using PL = NmCxPlan<P>;
    using T = double;
    constexpr int NT = PL::NT;
    NM_SHARED_BYTES(smem);
    cx<T>* work = reinterpret_cast<cx<T>*>(smem);
    unsigned char* stage = reinterpret_cast<unsigned char*>(work + PL::NBUF);
    const int W = a.in.W, E = a.E;
    double* red = reinterpret_cast<double*>(stage + nm_notchx_stage_bytes<P>(W));
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(red) + NM_CX_RED_BYTES);
    const int tid = threadIdx.x;
    const int npair = (a.in.n_ch + 1) >> 1;
    const int rowe = (W + 1 + 1) & ~1;
    const cx<T>* NM_RESTRICT tw = a.fft.tw;
    const T* NM_RESTRICT hx = a.hx;
    const cx<T> wA = nm_ldg(tw + tid);
    const cx<T> wB = nm_ldg(tw + (tid & (PL::M1 - 1)) * 16);
    cx<T>* const p0w = work + tid + (tid >> PL::PAD);

    int item = blockIdx.x;
    if (item >= a.n_items) return;
    if (tid == 0) nm_mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) nm_nx_issue(a, item, npair, stage, bar);
    unsigned parity = 0;

    cx<T> v[16];
    T hv[16];
    while (item < a.n_items) {
        const int next = item + gridDim.x;
        const int w = item / npair, c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        nm_mbar_wait(bar, parity);
        parity ^= 1u;
        {
            // odd reflection about both end samples: the arithmetic of nm_cx_load_item<REFLECT>, rows read from the stage
            const int lead = (int)(nm_ldg(a.in.off + w) & 1);
            const double* r0 = reinterpret_cast<const double*>(stage) + lead;
            const double* r1 = r0 + (has2 ? rowe : 0);
            const double a0 = 2.0 * r0[0], b0 = 2.0 * r1[0], a1 = 2.0 * r0[W - 1], b1 = 2.0 * r1[W - 1];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const int n = tid + NT * t;
                const bool left = n < E, mid = !left && n < E + W, right = !left && !mid && n < W + 2 * E;
                int idx = left ? E - n : (mid ? n - E : 2 * W + E - 2 - n);
                idx = (left || mid || right) ? idx : 0;
                const double sgn = mid ? 1.0 : ((left || right) ? -1.0 : 0.0);
                const double ca = left ? a0 : (right ? a1 : 0.0), cb = left ? b0 : (right ? b1 : 0.0);
                const double va = fma(sgn, r0[idx], ca), vb = fma(sgn, r1[idx], cb);
                v[t] = {va, has2 ? vb : 0.0};
            }
        }
        nm_bfly16<false>(v);
        nm_twiddle_w1<16, false>(v, wA);
#pragma unroll
        for (int t = 0; t < 16; ++t) p0w[t * PL::S0] = v[t];
        __syncthreads();  // every thread is done with the stage: the next item's rows may land from here on
        if (tid == 0 && next < a.n_items) nm_nx_issue(a, next, npair, stage, bar);
        nm_cx_load_h<PL, T>(hv, hx, tid);
        nm_cx_pass1<PL, false>(work, wB, tid);
        __syncthreads();
        nm_cx_pass2<PL, 1, T>(work, work, hv, tid);
        __syncthreads();
        nm_cx_pass1<PL, true>(work, wB, tid);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = p0w[t * PL::S0];
        nm_twiddle_w1<16, true>(v, wA);
        nm_bfly16<true>(v);
        NmEpiStoreScan::State st;
        epi.template consume<PL, T>(v, work, red, st, E, W, a.in.n_ch, w, c0, has2, 0, tid);
        epi.template finish<PL, T>(work, red, st, E, W, a.in.n_ch, w, c0, has2, 0, tid);
        if (epi.needs_trailing_barrier()) __syncthreads();
        item = next;
    }
}
