// nm_norm.cuh -- rolling feature normalisation (processing/normalization.py:81-111,151-170), SURVEY 8f-1.
//
// Reference semantics for the feature vector v_g of window g (g counted since the DataProcessor was built):
//   g == 0 : returned unchanged (not clipped, no nan_to_num) and stored as history
//   g >= 1 : history h = raw vectors of windows max(0, g-n_keep+1) .. g (current included);
//            out = clip((v_g - centre(h)) / scale(h), +-clip), then nan_to_num
// Histories hold RAW values, so all windows of a batch are independent: one thread per (window, column).
#pragma once

#include "nm_common.cuh"

struct NmNormArgs {
    const double* ext;    // (n_prev + n_windows, n_cols) raw values, previous windows first
    int n_prev;
    int n_windows;
    int n_cols;
    const int* cols;      // column in `out` of compact column j
    long long g0;         // global index of the batch's first window
    int n_keep;
    int method;           // 0 mean, 1 median, 2 zscore, 3 zscore-median
    double clip;
    double* out;
    int F;
};

NM_GLOBAL void nm_norm_gather_kernel(const double* out, int F, const int* cols, int n_cols, int n_windows, double* ext_rows) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_windows * n_cols) return;
    const int k = (int)(idx / n_cols), j = (int)(idx - (long long)k * n_cols);
    ext_rows[idx] = out[(size_t)k * F + cols[j]];
}

NM_DEV double nm_norm_median(const double* col, int stride, int n) {
    // order statistics by rank counting (NaNs ignored like numpy.nanmedian); O(n^2), n <= n_keep
    int m = 0;
    for (int i = 0; i < n; ++i) m += (col[(size_t)i * stride] == col[(size_t)i * stride]) ? 1 : 0;
    if (m == 0) return __longlong_as_double(0x7ff8000000000000LL);
    const int r_lo = (m - 1) / 2, r_hi = m / 2;
    double vlo = 0.0, vhi = 0.0;
    for (int i = 0; i < n; ++i) {
        const double x = col[(size_t)i * stride];
        if (x != x) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double y = col[(size_t)j * stride];
            rank += (y < x || (y == x && j < i)) ? 1 : 0;
        }
        if (rank == r_lo) vlo = x;
        if (rank == r_hi) vhi = x;
    }
    return (r_lo == r_hi) ? vlo : 0.5 * (vlo + vhi);
}

NM_GLOBAL void nm_norm_kernel(NmNormArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.n_windows * a.n_cols) return;
    const int k = (int)(idx / a.n_cols), j = (int)(idx - (long long)k * a.n_cols);
    const long long g = a.g0 + k;
    if (g == 0) return;  // first window passes through untouched
    const int row = a.n_prev + k;
    long long nh = g + 1 < a.n_keep ? g + 1 : a.n_keep;
    if (nh > row + 1) nh = row + 1;
    const double* col = a.ext + (size_t)(row - (nh - 1)) * a.n_cols + j;
    const double v = a.ext[(size_t)row * a.n_cols + j];
    double sum = 0.0;
    int cnt = 0;
    for (int i = 0; i < nh; ++i) {
        const double x = col[(size_t)i * a.n_cols];
        if (x == x) { sum += x; ++cnt; }
    }
    const double mean = sum / cnt;
    double r;
    if (a.method == 0) {
        r = (v - mean) / mean;
    } else if (a.method == 1) {
        const double med = nm_norm_median(col, a.n_cols, (int)nh);
        r = (v - med) / med;
    } else {
        double q = 0.0;
        for (int i = 0; i < nh; ++i) {
            const double x = col[(size_t)i * a.n_cols];
            if (x == x) q += (x - mean) * (x - mean);
        }
        double sd = sqrt(q / cnt);
        if (sd == 0.0) sd = 1.0;
        const double centre = (a.method == 2) ? mean : nm_norm_median(col, a.n_cols, (int)nh);
        r = (v - centre) / sd;
    }
    if (a.clip > 0.0) {
        if (r < -a.clip) r = -a.clip;
        if (r > a.clip) r = a.clip;
    }
    a.out[(size_t)k * a.F + a.cols[j]] = nm_nan_to_num(r);
}

struct NormFam {
    int method = 2, n_keep = 300, n_cols = 0;
    double clip = 3.0;
    DevBuf d_cols, d_ext, d_hist;
    long long batch = 0;  // windows seen since reset
    int n_prev = 0;       // rows currently held in d_hist (<= n_keep - 1)
    int build(int method_, double clip_, int n_keep_, int n_cols_, const int* cols, cudaStream_t s) {
        method = method_; clip = clip_; n_keep = n_keep_; n_cols = n_cols_;
        if (d_cols.upload(cols, (size_t)n_cols, s)) return -1;
        return d_hist.ensure((size_t)std::max(1, n_keep) * std::max(1, n_cols) * sizeof(double));
    }
    void reset() { batch = 0; n_prev = 0; }
    int run(nm_pipeline* p, int n_windows);
};
