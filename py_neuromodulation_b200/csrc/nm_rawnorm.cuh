// nm_rawnorm.cuh -- RawNormalizer (processing/normalization.py:30-111, type "raw"), SURVEY.md 8f-3.
//
// Reference semantics per window g (counted since the processor was built), per channel, on the PREPROCESSED rows d_g
// (filter -> notch -> re-reference of that very window):
//   g == 0 : d_0 is returned unchanged and becomes the history
//   g >= 1 : the last `add` = int(sfreq / rate) samples of d_g are appended to the history, d_g is normalised against the
//            WHOLE history (those new samples included), clipped to +-clip, nan_to_num'ed, and the history is trimmed to
//            its last n_keep - 1 samples.
// The history is a stream of blocks -- block 0 = the W samples of window 0, block g = the `add` tail samples of window g --
// kept in a per-channel ring; every block also leaves (n, mean, M2) behind, so that the mean / variance a window needs are
// Chan-combined from <= ~300 block records plus one partially expired block read from the ring, instead of two passes over
// 30 000 samples per (window, channel).  Histories hold un-normalised samples, so the windows of a chunk are independent
// once their blocks are appended: kernel 1 appends (warp per (window, channel)), kernel 2 combines and normalises.
// Methods: 0 mean, 2 zscore; 1 median and 3 zscore-median take the median of the same history from the sliding
// order-statistic kernel of the burst thresholds (nm_burst_thr_kernel on signed keys, q = 0.5) between the two kernels.
// Round 2: 4 minmax and 5 robust -- the scikit-learn MinMaxScaler / RobustScaler the reference wraps
// (processing/normalization.py:58-70,173-190) -- run the same kernel once per order statistic they need (q = 0, 1 and
// q = 0.5, 0.25, 0.75), each track with a bracket state of its own.
#pragma once

#include "nm_bursts.cuh"

struct NmRawNormArgs {
    NmRows in;              // preprocessed rows of the chunk
    double* out;            // (n_windows, n_ch, Wp) normalised rows
    long long Wp;
    double* ring;           // (n_ch, cap) history samples by stream position
    long long cap;
    double* blk;            // (n_ch, blk_cap, 3): n, mean, M2 of block j at [j % blk_cap]
    int blk_cap;
    long long g0;           // global index of the chunk's first window
    int add;                // min(int(sfreq / rate), W)
    const long long* lo;    // [n_windows] first stream position of the statistics range of window k (unused for g == 0)
    int method;
    double clip;
    const double* med;      // (n_windows, n_ch) order statistics of the histories from the sliding quantile kernel: the median
                            // (methods 1, 3, 5) or the minimum (method 4), else nullptr
    const double* q1;       // second track: maximum (method 4) / 25th percentile (method 5)
    const double* q2;       // third track: 75th percentile (method 5)
};

struct NmStat { double n, mean, m2; };

NM_DEV NmStat nm_stat_merge(NmStat a, NmStat b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    const double n = a.n + b.n, d = b.mean - a.mean;
    return {n, a.mean + d * (b.n / n), a.m2 + b.m2 + d * d * (a.n * b.n / n)};
}

NM_DEV NmStat nm_stat_warp(NmStat s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        NmStat t;
        t.n = __shfl_xor_sync(0xffffffffu, s.n, o);
        t.mean = __shfl_xor_sync(0xffffffffu, s.mean, o);
        t.m2 = __shfl_xor_sync(0xffffffffu, s.m2, o);
        s = nm_stat_merge(s, t);
    }
    return s;
}

// stream geometry: block 0 = [0, W), block j >= 1 = [W + (j-1)*add, W + j*add)
NM_DEV long long nm_rn_block_start(long long j, int W, int add) { return j == 0 ? 0 : (long long)W + (j - 1) * add; }

// kernel 1: append the block of every window of the chunk (samples -> ring, two-pass block statistics -> blk)
NM_GLOBAL void nm_rawnorm_append_kernel(NmRawNormArgs a) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long n_rows = (long long)a.in.n_windows * a.in.n_ch;
    const int W = a.in.W;
    for (long long row = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); row < n_rows; row += (long long)gridDim.x * wpc) {
        const int k = (int)(row / a.in.n_ch), c = (int)(row - (long long)k * a.in.n_ch);
        const long long g = a.g0 + k;
        const double* x = a.in.base + (size_t)c * a.in.ch_stride + nm_ldg(a.in.off + k);
        const int len = (g == 0) ? W : a.add;
        const double* src = x + (W - len);
        const long long s0 = nm_rn_block_start(g, W, a.add);
        double* ring = a.ring + (size_t)c * a.cap;
        double sum = 0.0;
        for (int t = lane; t < len; t += 32) {
            const double v = src[t];
            ring[(s0 + t) % a.cap] = v;
            sum += v;
        }
        const double mean = nm_warp_sum(sum) / len;
        double q = 0.0;
        for (int t = lane; t < len; t += 32) {
            const double e = src[t] - mean;
            q += e * e;
        }
        q = nm_warp_sum(q);
        if (lane == 0) {
            double* b = a.blk + ((size_t)c * a.blk_cap + (size_t)(g % a.blk_cap)) * 3;
            b[0] = (double)len;
            b[1] = mean;
            b[2] = q;
        }
    }
}

// kernel 2: statistics of the history range [lo, end of block g) and normalisation of the window
NM_GLOBAL void nm_rawnorm_apply_kernel(NmRawNormArgs a) {
    const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    const long long n_rows = (long long)a.in.n_windows * a.in.n_ch;
    const int W = a.in.W;
    for (long long row = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); row < n_rows; row += (long long)gridDim.x * wpc) {
        const int k = (int)(row / a.in.n_ch), c = (int)(row - (long long)k * a.in.n_ch);
        const long long g = a.g0 + k;
        const double* x = a.in.base + (size_t)c * a.in.ch_stride + nm_ldg(a.in.off + k);
        double* y = a.out + ((size_t)k * a.in.n_ch + c) * a.Wp;
        if (g == 0) {
            for (int t = lane; t < W; t += 32) y[t] = x[t];
            continue;
        }
        const long long lo = nm_ldg(a.lo + k);
        // first block that lies completely inside the range
        long long jf = (lo <= 0) ? 0 : ((lo <= W) ? 1 : (lo - W + a.add - 1) / a.add + 1);
        const long long jf_start = nm_rn_block_start(jf, W, a.add);
        NmStat s = {0.0, 0.0, 0.0};
        // partially expired block jf - 1: samples [lo, jf_start) straight from the ring (Welford per lane)
        const double* ring = a.ring + (size_t)c * a.cap;
        for (long long i = lo + lane; i < jf_start; i += 32) {
            const double v = ring[i % a.cap];
            s.n += 1.0;
            const double d = v - s.mean;
            s.mean += d / s.n;
            s.m2 += d * (v - s.mean);
        }
        const double* blk = a.blk + (size_t)c * a.blk_cap * 3;
        for (long long j = jf + lane; j <= g; j += 32) {
            const double* b = blk + (size_t)(j % a.blk_cap) * 3;
            s = nm_stat_merge(s, NmStat{b[0], b[1], b[2]});
        }
        s = nm_stat_warp(s);
        // centre / scale per method: mean -> (mean, mean); median -> (med, med); zscore -> (mean, std); zscore-median -> (med, std)
        double centre = s.mean;
        if (a.method == 1 || a.method == 3) centre = nm_ldg(a.med + (size_t)k * a.in.n_ch + c);
        double scale = centre;
        if (a.method == 2 || a.method == 3) {
            scale = sqrt(s.m2 / s.n);
            if (scale == 0.0) scale = 1.0;  // same behaviour as the reference (and sklearn)
        }
        double mm_scale = 0.0, mm_min = 0.0;
        if (a.method == 4) {  // MinMaxScaler: X * scale_ + min_ (sklearn _handle_zeros_in_scale: ranges below 10 eps -> 1)
            const double mn = nm_ldg(a.med + (size_t)k * a.in.n_ch + c), mx = nm_ldg(a.q1 + (size_t)k * a.in.n_ch + c);
            double range = mx - mn;
            if (range < 10.0 * 2.220446049250313e-16) range = 1.0;
            mm_scale = 1.0 / range;
            mm_min = 0.0 - mn * mm_scale;
        } else if (a.method == 5) {  // RobustScaler: (X - median) / (q75 - q25)
            centre = nm_ldg(a.med + (size_t)k * a.in.n_ch + c);
            scale = nm_ldg(a.q2 + (size_t)k * a.in.n_ch + c) - nm_ldg(a.q1 + (size_t)k * a.in.n_ch + c);
            if (scale < 10.0 * 2.220446049250313e-16) scale = 1.0;
        }
        for (int t = lane; t < W; t += 32) {
            double v = (a.method == 4) ? x[t] * mm_scale + mm_min : (x[t] - centre) / scale;
            if (a.clip != 0.0) {  // (NaN compares false both ways and survives the clip like numpy.clip)
                if (v > a.clip) v = a.clip;
                if (v < -a.clip) v = -a.clip;
            }
            y[t] = nm_nan_to_num(v);
        }
    }
}

struct nm_pipeline;

struct RawNormFam {
    int method = 2, n_keep = 30000, add = 100, C = 0, W = 0, chunk = 0, blk_cap = 0;
    double clip = 3.0;
    long long cap = 0, Wp = 0, batch = 0, len_prev = 0;  // len_prev: history length after the last processed window
    DevBuf d_ring, d_blk, d_out, d_lo;
    // sliding median (methods 1, 3): per-window bookkeeping + persistent state of nm_burst_thr_kernel, one row per channel
    // (methods 4, 5: up to three order statistics per window -- "tracks" -- each with parameters, results and a bracket state of its own)
    DevBuf d_e_end, d_n, d_klo[3], d_khi[3], d_gamma[3], d_med[3];
    NmBqState qstate[3];
    int n_tracks() const { return (method == 1 || method == 3) ? 1 : (method == 4 ? 2 : (method == 5 ? 3 : 0)); }
    double track_q(int t) const {  // quantile of track t
        if (method == 4) return t == 0 ? 0.0 : 1.0;
        if (method == 5) return t == 0 ? 0.5 : (t == 1 ? 0.25 : 0.75);
        return 0.5;
    }
    bool need_median() const { return n_tracks() > 0; }
    int build(int method_, double clip_, int n_keep_, int add_, int C_, int W_) {
        method = method_; clip = clip_; n_keep = n_keep_; C = C_; W = W_;
        add = std::max(1, std::min(add_, W));
        return 0;
    }
    int alloc_chunk(int chunk_, long long Wp_) {
        chunk = chunk_;
        Wp = Wp_;
        const long long max_hist = std::max<long long>(W, n_keep > 1 ? n_keep - 1 : (1LL << 22)) + add;
        cap = max_hist + (long long)chunk * add + W;
        blk_cap = (int)(max_hist / add + chunk + 8);
        if (d_ring.ensure((size_t)C * cap * sizeof(double))) return -1;
        if (d_blk.ensure((size_t)C * blk_cap * 3 * sizeof(double))) return -1;
        if (d_out.ensure((size_t)chunk * C * Wp * sizeof(double))) return -1;
        for (int t = 0; t < n_tracks(); ++t) {
            if (d_med[t].ensure((size_t)chunk * C * sizeof(double))) return -1;
            if (qstate[t].alloc((size_t)C)) return -1;
        }
        return 0;
    }
    void reset(cudaStream_t s) {  // (stream-ordered, see BurstsFam::reset)
        batch = 0;
        len_prev = 0;
        for (int t = 0; t < 3; ++t) qstate[t].reset(s);
    }
    int run(nm_pipeline* p, NmRows& rows);
};
