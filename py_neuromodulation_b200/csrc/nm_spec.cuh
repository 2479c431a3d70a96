// nm_spec.cuh -- oscillatory band features from segment DFTs: FFT / Welch / STFT
//   features/oscillatory.py:58-119 (FFT), :122-182 (Welch), :185-250 (STFT)
//
// One kernel serves the three plugins; the host describes them as "segments of nper samples":
//   FFT    1 segment  = last nper samples, no window, value |Z|
//   Welch  nseg hops of nper/2, per-segment mean removal, periodic Hann,
//          value |Z|^2 / (fs * sum w^2), doubled for interior bins, averaged over segments
//   STFT   even extension by nper/2, zero padding, periodic Hamming, value |Z| / sum w,
//          estimators taken over the (bin x segment) matrix
// The transform is the exact nper-point DFT (mixed radix 2/3/4/5 + generic primes) -- window
// lengths are 1000 / 2000 / 500 samples, not powers of two.  Two channels share one complex
// transform (real / imaginary part) and are separated with the conjugate-symmetry identity.
#pragma once

#include "nm_common.cuh"

struct NmSpecArgs {
    NmRows in;
    NmFft<double> fft;      // size nper
    int need_scratch;       // generic radix present
    int nseg, hop, start;   // segment s covers window samples [start + s*hop, ... + nper)
    int ext_even, ext_len;  // even reflection available outside [0, W) (STFT boundary="even")
    int detrend;
    const double* win;      // [nper] or nullptr
    int power;              // 0: |Z| * scale, 1: |Z|^2 * scale with one-sided doubling
    double scale;
    int log;
    int keep_segments;      // 0: average over segments (FFT/Welch), 1: keep (bin, segment) matrix (STFT)
    int k0, nk;             // stored bins [k0, k0 + nk)
    int n_bands;
    const int* band_lo;     // absolute bin ranges [lo, hi)
    const int* band_hi;
    int est_mask;           // bit 0 mean, 1 median, 2 std, 3 max
    int want_spectrum;
    NmOut out;              // per_ch = n_bands * 4 + (nper / 2 + 1)
    int n_items;
};

NM_DEV double nm_spec_sample(const double* r, int i, int W, int ext_even, int ext_len) {
    if (i >= 0 && i < W) return r[i];
    if (!ext_even) return 0.0;
    if (i < 0) {
        const int k = -i;
        return (k <= ext_len && k < W) ? r[k] : 0.0;
    }
    const int k = i - (W - 1);
    return (k <= ext_len && k < W) ? r[W - 1 - k] : 0.0;
}

static NM_HD size_t nm_spec_smem_bytes(int nper, int need_scratch, int nk, int nsegv) {
    return ((size_t)nper * (need_scratch ? 2 : 1)) * sizeof(cx<double>) + (size_t)2 * nk * nsegv * sizeof(double) +
           2 * 32 * sizeof(double);
}

// magnitude / power of bin k of both channels from the packed transform (U = Z[k], V = Z[N-k]) -> vals
NM_DEV void nm_spec_bin(const NmSpecArgs& a, double* vals, int i, int k, int s, cx<double> U, cx<double> V) {
    const int N = a.fft.n;
    const double are = 0.5 * (U.re + V.re), aim = 0.5 * (U.im - V.im);
    const double bre = 0.5 * (U.im + V.im), bim = -0.5 * (U.re - V.re);
    double ma, mb;
    if (a.power) {
        const double dbl = (k == 0 || (2 * k == N)) ? 1.0 : 2.0;
        ma = (are * are + aim * aim) * a.scale * dbl;
        mb = (bre * bre + bim * bim) * a.scale * dbl;
    } else {
        ma = sqrt(are * are + aim * aim) * a.scale;
        mb = sqrt(bre * bre + bim * bim) * a.scale;
    }
    if (a.keep_segments) {
        vals[(size_t)i * a.nseg + s] = ma;
        vals[(size_t)(a.nk + i) * a.nseg + s] = mb;
    } else {
        vals[i] += ma / a.nseg;
        vals[a.nk + i] += mb / a.nseg;
    }
}

// log transform, band estimators (one warp per (channel, band)) and the optional spectrum columns for one item;
// `vals` = [2][nk][nsegv] bin values of the channel pair.  Shared by the generic and the register-blocked kernels.
NM_DEV void nm_spec_finish(const NmSpecArgs& a, double* vals, int nsegv, int w, int c0, bool has2, int tid, int nt) {
    const int lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const int nbins = a.fft.n / 2 + 1;
    if (a.log) {
        for (int i = tid; i < 2 * a.nk * nsegv; i += nt) vals[i] = log10(vals[i]);
        __syncthreads();
    }

    // band estimators: one warp per (channel, band)
    for (int task = wid; task < 2 * a.n_bands; task += nwarp) {
        const int ch = task / a.n_bands, b = task - ch * a.n_bands;
        if (ch == 1 && !has2) continue;
        const int lo = nm_ldg(a.band_lo + b), hi = nm_ldg(a.band_hi + b);
        const double* v = vals + ((size_t)ch * a.nk + (lo - a.k0)) * nsegv;
        const int cnt = (hi - lo) * nsegv;
        double sm = 0.0, mx = -INFINITY;
        bool any_nan = false;
        for (int i = lane; i < cnt; i += 32) {
            const double x = v[i];
            sm += x;
            if (x != x) any_nan = true;
            if (x > mx) mx = x;
        }
        sm = nm_warp_sum(sm);
        mx = nm_warp_max(mx);
        const double mean = sm / cnt;
        const int c = c0 + ch;
        if (a.est_mask & 1) { if (lane == 0) nm_store(a.out, w, c, b * 4 + 0, mean); }
        if (a.est_mask & 2) {
            const int r_lo = (cnt - 1) / 2, r_hi = cnt / 2;
            double vlo = 0.0, vhi = 0.0;
            for (int i = lane; i < cnt; i += 32) {
                const double x = v[i];
                int rank = 0;
                for (int j = 0; j < cnt; ++j) {
                    const double y = v[j];
                    rank += (y < x || (y == x && j < i)) ? 1 : 0;
                }
                if (rank == r_lo) vlo = x;
                if (rank == r_hi) vhi = x;
            }
            // exactly one lane holds each order statistic; everything else contributes +0
            vlo = nm_warp_sum(vlo);
            vhi = nm_warp_sum(vhi);
            if (lane == 0) nm_store(a.out, w, c, b * 4 + 1, (r_lo == r_hi) ? vlo : 0.5 * (vlo + vhi));
        }
        if (a.est_mask & 4) {
            double q = 0.0;
            for (int i = lane; i < cnt; i += 32) {
                const double d = v[i] - mean;
                q += d * d;
            }
            q = nm_warp_sum(q);
            if (lane == 0) nm_store(a.out, w, c, b * 4 + 2, sqrt(q / cnt));
        }
        if (a.est_mask & 8) { if (lane == 0) nm_store(a.out, w, c, b * 4 + 3, any_nan ? mx : mx); }
    }
    if (a.want_spectrum) {
        for (int i = tid; i < 2 * a.nk; i += nt) {
            const int ch = i / a.nk, kk = i - ch * a.nk;
            if (ch == 1 && !has2) continue;
            double val;
            if (a.keep_segments) {
                double sm = 0.0;
                for (int s = 0; s < a.nseg; ++s) sm += vals[(size_t)i * a.nseg + s];
                val = sm / a.nseg;
            } else {
                val = vals[i];
            }
            const int k = a.k0 + kk;
            if (k < nbins) nm_store(a.out, w, c0 + ch, a.n_bands * 4 + k, val);
        }
    }
}

NM_GLOBAL void nm_spec_kernel(NmSpecArgs a) {
    NM_SHARED_BYTES(smem);
    const int N = a.fft.n;
    cx<double>* buf = reinterpret_cast<cx<double>*>(smem);
    cx<double>* scratch = a.need_scratch ? buf + N : nullptr;
    const int nsegv = a.keep_segments ? a.nseg : 1;
    double* vals = reinterpret_cast<double*>(buf + (a.need_scratch ? 2 : 1) * (size_t)N);  // [2][nk][nsegv]
    double* red = vals + (size_t)2 * a.nk * nsegv;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int W = a.in.W;
    const int npair = (a.in.n_ch + 1) >> 1;

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int w = item / npair;
        const int c0 = (item - w * npair) * 2;
        const bool has2 = c0 + 1 < a.in.n_ch;
        const double* r0 = a.in.base + (size_t)c0 * a.in.ch_stride + nm_ldg(a.in.off + w);
        const double* r1 = r0 + (has2 ? a.in.ch_stride : 0);

        if (!a.keep_segments)
            for (int i = tid; i < 2 * a.nk; i += nt) vals[i] = 0.0;

        for (int s = 0; s < a.nseg; ++s) {
            const int base = a.start + s * a.hop;
            double sum[2] = {0.0, 0.0};
            for (int n = tid; n < N; n += nt) {
                const double va = nm_spec_sample(r0, base + n, W, a.ext_even, a.ext_len);
                const double vb = has2 ? nm_spec_sample(r1, base + n, W, a.ext_even, a.ext_len) : 0.0;
                buf[n] = {va, vb};
                sum[0] += va;
                sum[1] += vb;
            }
            if (a.detrend) {
                nm_block_sum<2>(sum, red, tid, nt);
                sum[0] /= N;
                sum[1] /= N;
            } else {
                sum[0] = sum[1] = 0.0;
            }
            if (a.detrend || a.win) {
                for (int n = tid; n < N; n += nt) {  // each thread touches only the slots it wrote
                    const double wv = a.win ? nm_ldg(a.win + n) : 1.0;
                    cx<double> v = buf[n];
                    buf[n] = {(v.re - sum[0]) * wv, (v.im - sum[1]) * wv};
                }
            }
            __syncthreads();
            nm_fft_forward<double>(buf, scratch, a.fft, tid, nt);

            for (int i = tid; i < a.nk; i += nt) {
                const int k = a.k0 + i;
                const cx<double> U = buf[nm_ldg(a.fft.pos + k)];
                const cx<double> V = buf[nm_ldg(a.fft.pos + (k == 0 ? 0 : N - k))];
                nm_spec_bin(a, vals, i, k, s, U, V);
            }
            __syncthreads();
        }
        nm_spec_finish(a, vals, nsegv, w, c0, has2, tid, nt);
        __syncthreads();
    }
}
