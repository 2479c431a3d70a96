// nm_firbank.h -- host-side description of one FIR bank (taps -> transform size, spectra, kernel args).
#pragma once

#include <algorithm>

#include "nm_convx.cuh"
#include "nm_fir.cuh"
#include "nm_host.h"

struct nm_pipeline;

struct FirBank {
    int nF = 0, L = 0, Lh = 0, P = 0, mode = NM_FIR_SAME, E = 0;
    FftPlanHost fft;
    DevBuf d_hperm, d_hx, d_hx32, d_tw32;  // *32: float32 copies for the optional float32 mode of nm_convx_kernel
    bool pow2 = false;  // register-blocked kernel (nm_convx.cuh / nm_conv.cuh) vs generic mixed-radix kernel (nm_fir.cuh)
    bool mixed = false; // P = 3 * 2^k on nm_convx_kernel's 12 x R1 x 16 plans ('same' mode only; there is no nm_conv_kernel fallback)
    int pad = 3;
    // 'same'-mode banks need P >= W + (L-1)/2 only: 1536 / 3072 instead of 2048 / 4096 (NMB200_MIXED_RADIX=0: powers of two only)
    static bool mixed_enabled() {
        const char* env = getenv("NMB200_MIXED_RADIX");
        return env ? atoi(env) != 0 : true;
    }
    int build(const double* taps, int nF_, int L_, int W, int mode_, cudaStream_t s) {
        nF = nF_; L = L_; Lh = (L - 1) / 2; mode = mode_;
        int need;
        if (mode == NM_FIR_REFLECT) {
            // mne _overlap_add_filter: n_edge = max(min(len(h), len(x)) - 1, 0) reflected samples per side;
            // only the (L-1)/2 nearest ones can reach the W centre outputs
            const int n_edge = std::max(std::min(L, W) - 1, 0);
            E = std::min(n_edge, Lh);
            need = W + E + Lh;
        } else {
            E = 0;
            need = W + Lh;
        }
        int p2 = 1;
        while (p2 < need) p2 <<= 1;
        pow2 = p2 >= 512 && p2 <= 8192;  // P/16 threads per CTA: 32 .. 512 (larger transforms use the generic kernel)
        mixed = pow2 && mode == NM_FIR_SAME && (p2 == 2048 || p2 == 4096) && need <= p2 / 4 * 3 && mixed_enabled();
        if (mixed) {
            P = p2 / 4 * 3;
            pad = 4;
            if (fft.build_with(P, std::vector<int>{12, P == 1536 ? 8 : 16, 16}, s)) return -1;
        } else if (pow2) {
            P = p2;
            std::vector<int> radices{16};
            int rem = P / 16;
            if (rem == 32) { radices.push_back(8); radices.push_back(4); rem = 1; }
            if (rem == 64) { radices.push_back(8); radices.push_back(8); rem = 1; }
            while (rem > 16) { radices.push_back(16); rem /= 16; }
            if (rem > 1) radices.push_back(rem);
            NM_CHECK((int)radices.size() <= NM_CONV_MAX_PASS && radices.size() >= 2, "internal: bad pass plan for P = %d", P);
            int last = radices.back(), lg = 0;
            while ((1 << lg) < last) ++lg;
            pad = std::max(3, lg);
            if (fft.build_with(P, radices, s)) return -1;
        } else {
            P = nm_next_smooth(need);
            if (fft.build(P, s)) return -1;
            NM_CHECK(!fft.generic, "internal: convolution length %d is not 5-smooth", P);
        }
        std::vector<double> hperm;
        if (nm_build_hperm(taps, nF, L, fft, hperm)) return -1;
        if (pow2 && nm_convx_supported(P)) {
            std::vector<double> hx(hperm.size());
            std::vector<float> hx32(hperm.size());
            for (int f = 0; f < nF; ++f) {
                nm_cx_interleave_h(P, hperm.data() + (size_t)f * P, hx.data() + (size_t)f * P);
                nm_cx_interleave_h(P, hperm.data() + (size_t)f * P, hx32.data() + (size_t)f * P);
            }
            std::vector<cx<float>> tw32(P);
            for (int k = 0; k < P; ++k) {
                const double ang = -2.0 * M_PI * (double)k / (double)P;
                tw32[k] = {(float)std::cos(ang), (float)std::sin(ang)};
            }
            if (d_hx.upload(hx, s) || d_hx32.upload(hx32, s) || d_tw32.upload(tw32, s)) return -1;
        }
        return d_hperm.upload(hperm, s);
    }
    // Ideal zero-phase low-pass of FFT down-sampling by D (nm_resample.cuh): the W-sample window is reflect-padded by `pad_each`
    // samples per side to exactly P_ points and multiplied with H[k] = 1 for |k| <= P_/(2D), else 0 (circular: no taps).
    int build_lowpass(int W, int P_, int pad_each, int D, cudaStream_t s) {
        nF = 1; L = 1; Lh = 0; mode = NM_FIR_REFLECT; E = pad_each; P = P_; pow2 = true;
        NM_CHECK(nm_convx_supported(P) && (P & (P - 1)) == 0 && W + 2 * pad_each == P && P % (2 * D) == 0, "internal: bad down-sampling plan");
        std::vector<int> radices{16};
        int rem = P / 16;
        if (rem == 64) { radices.push_back(8); radices.push_back(8); rem = 1; }
        while (rem > 16) { radices.push_back(16); rem /= 16; }
        if (rem > 1) radices.push_back(rem);
        int last = radices.back(), lg = 0;
        while ((1 << lg) < last) ++lg;
        pad = std::max(3, lg);
        if (fft.build_with(P, radices, s)) return -1;
        const int nyq = P / (2 * D);
        std::vector<double> hperm((size_t)P, 0.0), hx((size_t)P);
        for (int k = 0; k < P; ++k) {
            const int f = k <= P / 2 ? k : P - k;
            hperm[fft.pos[k]] = f <= nyq ? 1.0 / (double)P : 0.0;  // (MNE's new_len / orig_len scale cancels irfft's 1 / new_len)
        }
        nm_cx_interleave_h(P, hperm.data(), hx.data());
        if (d_hx.upload(hx, s)) return -1;
        return d_hperm.upload(hperm, s);
    }
    // args / conv_args(in, f0, n): the launch serves filters [f0, f0 + n) of the bank (n < 0: all of them)
    NmFirArgs args(const NmRows& in, int f0 = 0, int n = -1) const {
        NmFirArgs a;
        a.in = in;
        a.fft = fft.dev();
        a.hperm = d_hperm.as<double>() + (size_t)f0 * P;
        a.nF = n < 0 ? nF : n;
        a.f0 = f0;
        a.mode = mode;
        a.E = E;
        a.n_items = in.n_windows * ((in.n_ch + 1) / 2);
        return a;
    }
    NmConvArgs conv_args(const NmRows& in, int f0 = 0, int n = -1) const {
        NmConvArgs a;
        a.in = in;
        a.fft.P = P;
        a.fft.NT = P / 16;
        a.fft.npass = (int)fft.radix.size();
        for (int i = 0; i < a.fft.npass; ++i) { a.fft.radix[i] = fft.radix[i]; a.fft.len[i] = fft.len[i]; }
        a.fft.pad = pad;
        a.fft.tw = fft.d_tw.as<cx<double>>();
        a.hperm = d_hperm.as<double>() + (size_t)f0 * P;
        a.hx = d_hx.as<double>() + (size_t)f0 * P;
        a.hx32 = d_hx32.as<float>() + (size_t)f0 * P;
        a.tw32 = d_tw32.as<cx<float>>();
        a.nF = n < 0 ? nF : n;
        a.f0 = f0;
        a.mode = mode;
        a.E = E;
        a.n_items = in.n_windows * ((in.n_ch + 1) / 2);
        a.scratch_in_tail = 0;
        return a;
    }
    bool epi_fits_tail(size_t epi) const { return pow2 && epi > 0 && epi <= (nm_conv_buf_elems(P, pad) - (size_t)P) * sizeof(cx<double>); }
    int threads() const { return pow2 ? (mixed ? nm_convx_threads(P) : P / 16) : NM_FFT_THREADS; }
    // nm_convx_kernel: reflect mode is single-filter (one buffer), 'same' mode always runs the bank code (two buffers)
    // f32: 0 float64, 1 scalar float32 (8-byte elements), 2 packed float32 pairs (16-byte elements like cx<double>)
    // single: one filter per launch (BANK = false instantiation: the spectrum is multiplied in place, one transform buffer)
    size_t smem_x(size_t epi, int f32 = 0, bool single = false) const {
        if (f32 == 2) f32 = 0;
        return nm_conv_buf_elems(P, pad) * (f32 ? sizeof(cx<float>) : sizeof(cx<double>)) * ((mode == NM_FIR_REFLECT || single) ? 1 : 2) +
               NM_CX_RED_BYTES + (epi_fits_tail(epi) ? 0 : epi);
    }
    size_t smem(size_t epi, int n_filters = -1) const {
        const size_t buf = pow2 ? nm_conv_buf_elems(P, pad) : (size_t)P;
        return buf * sizeof(cx<double>) * ((n_filters < 0 ? nF : n_filters) > 1 ? 2 : 1) + (epi_fits_tail(epi) ? 0 : epi);
    }
};
