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
    int method;           // 0 mean, 1 median, 2 zscore, 3 zscore-median, 4 minmax, 5 robust, 6 quantile (the last three: restated scikit-learn transformers)
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

// ---- scikit-learn transformers the reference wraps (processing/normalization.py:58-70,173-190), restated ------------------------
// fit on nan_to_num(history), transform the RAW current value; sklearn.preprocessing._data (1.9): _handle_zeros_in_scale sets scales
// below 10 * eps to 1.
#define NM_NORM_TINY_SCALE (10.0 * 2.220446049250313e-16)

// numpy 'linear' percentile of the nan_to_num'ed history: virtual index (n - 1) * q, numpy's _lerp
NM_DEV double nm_norm_lerp(double a, double b, double t) {
    const double diff = b - a;
    double r = a + diff * t;
    if (t >= 0.5) r = b - diff * (1.0 - t);
    return (diff == 0.0) ? a : r;
}

// RobustScaler: (v - median) / (q75 - q25); one O(n^2) rank-counting pass collects the six order statistics
NM_DEV double nm_norm_robust(const double* col, int stride, int n, double v) {
    const double vi25 = (double)(n - 1) * 0.25, vi75 = (double)(n - 1) * 0.75;
    const int k25 = (int)floor(vi25), k75 = (int)floor(vi75);
    const int want[6] = {(n - 1) / 2, n / 2, k25, k25 + 1 < n ? k25 + 1 : n - 1, k75, k75 + 1 < n ? k75 + 1 : n - 1};
    double os[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        const double x = nm_nan_to_num(col[(size_t)i * stride]);
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double y = nm_nan_to_num(col[(size_t)j * stride]);
            rank += (y < x || (y == x && j < i)) ? 1 : 0;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q)
            if (rank == want[q]) os[q] = x;
    }
    const double med = (want[0] == want[1]) ? os[0] : 0.5 * (os[0] + os[1]);
    const double q25 = nm_norm_lerp(os[2], os[3], vi25 - k25), q75 = nm_norm_lerp(os[4], os[5], vi75 - k75);
    double scale = q75 - q25;
    if (scale < NM_NORM_TINY_SCALE) scale = 1.0;
    return (v - med) / scale;
}

// MinMaxScaler(feature_range = (0, 1)): X * scale_ + min_
NM_DEV double nm_norm_minmax(const double* col, int stride, int n, double v) {
    double mn = nm_nan_to_num(col[0]), mx = mn;
    for (int i = 1; i < n; ++i) {
        const double x = nm_nan_to_num(col[(size_t)i * stride]);
        mn = x < mn ? x : mn;
        mx = x > mx ? x : mx;
    }
    double range = mx - mn;
    if (range < NM_NORM_TINY_SCALE) range = 1.0;
    const double scale = 1.0 / range;
    const double min_ = 0.0 - mn * scale;
    return v * scale + min_;
}

// QuantileTransformer(n_quantiles = 300, uniform output) for histories of n <= 300 rows: n_quantiles_ = n, references
// linspace(0, 1, n), quantiles = np.nanpercentile(history, references * 100) -- the sorted history up to the last bit: the virtual
// index (n - 1) * ((j * step * 100) / 100) is not always exactly j, and then the quantile is a lerp that lands one ulp-ish next to
// the order statistic.  For distinct values that moves the result by ~1e-16; among TIES it decides which of the equal quantiles is
// "the last one <= v" (ascending np.interp) and "the first one >= v" (descending np.interp), i.e. a whole reference step.  So the
// roundings of numpy / scikit-learn are replayed exactly for the first and the last tie (no fused multiply-adds: nm_*_rn).
// The transform of a member v of the history (the current row always is one) is 0.5 * (R[j1] + R[m1]); values equal to the largest /
// smallest quantile map to 1 / 0 (the lower bound is applied last).
NM_DEV double nm_norm_vindex(int j, int n) {
    const double step = 1.0 / (double)(n - 1);                                   // numpy.linspace: arange(n) * step, last element := 1
    const double ref = (j == n - 1) ? 1.0 : nm_mul_rn((double)j, step);
    const double q = nm_mul_rn(ref, 100.0) / 100.0;                              // references_ * 100, then percentile's q / 100
    return nm_mul_rn((double)(n - 1), q);                                        // 'linear': (n - 1) * quantiles
}
NM_DEV double nm_norm_qlerp(double a, double b, double t) {                      // numpy _lerp
    const double diff = b - a;
    if (t >= 0.5) return nm_sub_rn(b, nm_mul_rn(diff, 1.0 - t));
    return nm_add_rn(a, nm_mul_rn(diff, t));
}

NM_DEV double nm_norm_quantile(const double* col, int stride, int n, double v) {
    if (v != v) return v;
    int lt = 0, le = 0;
    double mn = nm_nan_to_num(col[0]), mx = mn, pred = -NM_DBL_MAX, succ = NM_DBL_MAX;
    for (int i = 0; i < n; ++i) {
        const double x = nm_nan_to_num(col[(size_t)i * stride]);
        lt += (x < v) ? 1 : 0;
        le += (x <= v) ? 1 : 0;
        mn = x < mn ? x : mn;
        mx = x > mx ? x : mx;
        if (x < v && x > pred) pred = x;
        if (x > v && x < succ) succ = x;
    }
    const double step = 1.0 / (double)(n - 1);
    double up, dn;  // R[j1] and R[m1]
    if (le == lt) {
        // v is not a member of the history (only +-inf can get here: the history holds nan_to_num'ed values): np.interp's end values,
        // or a plain interpolation between the neighbouring order statistics
        if (le == 0) up = dn = 0.0;
        else if (lt == n) up = dn = 1.0;
        else {
            const double r0 = (double)(lt - 1) * step, r1 = (lt == n - 1) ? 1.0 : (double)lt * step;
            up = dn = r0 + (r1 - r0) / (succ - pred) * (v - pred);
        }
    } else {
        // order statistics around the ties of v: sorted[k] = pred (k = lt - 1), v (lt <= k < le), succ (k = le)
        const int ties = le - lt;
        int j1 = le - 1, m1 = lt;
        {
            const double vi = nm_norm_vindex(le - 1, n);
            int f = (int)floor(vi);
            if (f > n - 1) f = n - 1;
            const int nx = f + 1 < n ? f + 1 : n - 1;
            const double a = f < lt ? pred : (f < le ? v : succ), b = nx < lt ? pred : (nx < le ? v : succ);
            if (nm_norm_qlerp(a, b, vi - (double)f) > v && ties >= 2) j1 = le - 2;  // the last tie's quantile was nudged above v
        }
        {
            const double vi = nm_norm_vindex(lt, n);
            int f = (int)floor(vi);
            if (f > n - 1) f = n - 1;
            const int nx = f + 1 < n ? f + 1 : n - 1;
            const double a = f < lt ? pred : (f < le ? v : succ), b = nx < lt ? pred : (nx < le ? v : succ);
            if (nm_norm_qlerp(a, b, vi - (double)f) < v && ties >= 2) m1 = lt + 1;  // the first tie's quantile was nudged below v
        }
        up = (j1 == n - 1) ? 1.0 : nm_mul_rn((double)j1, step);
        dn = (m1 == n - 1) ? 1.0 : nm_mul_rn((double)m1, step);
    }
    double r = 0.5 * (up + dn);
    if (v == mx) r = 1.0;
    if (v == mn) r = 0.0;
    return r;
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
    if (a.method == 4) {
        r = nm_norm_minmax(col, a.n_cols, (int)nh, v);
    } else if (a.method == 5) {
        r = nm_norm_robust(col, a.n_cols, (int)nh, v);
    } else if (a.method == 6) {
        r = nm_norm_quantile(col, a.n_cols, (int)nh, v);
    } else if (a.method == 0) {
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

// ---- order-statistic methods over many windows: one thread per COLUMN walks the windows in order ------------------------------
// median / zscore-median / robust need order statistics of a history that slides by one row per window.  Rank counting per
// (window, column) is O(n_keep^2) -- 135 ms (median) and 412 ms (robust) per 591 windows x 7 936 columns, five to fifteen times the
// rest of the default pipeline.  Here a thread keeps the valid history values of its column SORTED in shared memory (column-
// interleaved: bank-conflict free) and updates it by one insertion and one removal per window (binary search + shift), so the order
// statistics are plain indexing: O(n_keep) per window.  Used for batched runs; a single streamed window keeps the kernel above.
struct NmNormOrderArgs {
    const double* ext;
    int n_prev, n_windows, n_cols;
    const int* cols;
    long long g0;
    int n_keep;
    int method;  // 1 median, 3 zscore-median, 5 robust
    double clip;
    double* out;
    int F;
};

NM_GLOBAL void nm_norm_order_kernel(NmNormOrderArgs a) {
    NM_SHARED_BYTES(smem);
    double* S = reinterpret_cast<double*>(smem);
    const int t = threadIdx.x, T = blockDim.x;
    const int j = blockIdx.x * T + t;
    if (j >= a.n_cols) return;  // (no barriers below)
    const bool to_num = a.method == 5;  // RobustScaler is fitted on nan_to_num(history); the numpy medians skip NaNs
    int m = 0;                          // sorted values S[0 .. m)
#define NM_NS(i) S[(size_t)(i) * T + t]
    auto insert = [&](double x) {
        if (to_num) x = nm_nan_to_num(x);
        if (x != x) return;
        int lo = 0, hi = m;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (NM_NS(mid) <= x) lo = mid + 1; else hi = mid;
        }
        for (int i = m; i > lo; --i) NM_NS(i) = NM_NS(i - 1);
        NM_NS(lo) = x;
        ++m;
    };
    auto remove = [&](double x) {
        if (to_num) x = nm_nan_to_num(x);
        if (x != x) return;
        int lo = 0, hi = m;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (NM_NS(mid) < x) lo = mid + 1; else hi = mid;
        }
        for (int i = lo; i + 1 < m; ++i) NM_NS(i) = NM_NS(i + 1);
        --m;
    };
    int first_row = -1;  // the structure holds rows [first_row, row of the previous window]
    for (int k = 0; k < a.n_windows; ++k) {
        const long long g = a.g0 + k;
        const int row = a.n_prev + k;
        long long nh = g + 1 < a.n_keep ? g + 1 : a.n_keep;
        if (nh > row + 1) nh = row + 1;
        const int want_first = row - (int)(nh - 1);
        if (first_row < 0) {
            for (int r = want_first; r <= row; ++r) insert(a.ext[(size_t)r * a.n_cols + j]);
            first_row = want_first;
        } else {
            for (; first_row < want_first; ++first_row) remove(a.ext[(size_t)first_row * a.n_cols + j]);  // (first: m <= n_keep always)
            insert(a.ext[(size_t)row * a.n_cols + j]);
        }
        if (g == 0) continue;  // first window passes through untouched
        const double v = a.ext[(size_t)row * a.n_cols + j];
        double r;
        if (m == 0) {
            r = __longlong_as_double(0x7ff8000000000000LL);
        } else {
            const int r_lo = (m - 1) / 2, r_hi = m / 2;
            const double med = (r_lo == r_hi) ? NM_NS(r_lo) : 0.5 * (NM_NS(r_lo) + NM_NS(r_hi));
            if (a.method == 1) {
                r = (v - med) / med;
            } else if (a.method == 3) {
                double sum = 0.0, q = 0.0;
                for (int i = 0; i < m; ++i) sum += NM_NS(i);
                const double mean = sum / m;
                for (int i = 0; i < m; ++i) { const double d = NM_NS(i) - mean; q += d * d; }
                double sd = sqrt(q / m);
                if (sd == 0.0) sd = 1.0;
                r = (v - med) / sd;
            } else {
                const double vi25 = (double)(m - 1) * 0.25, vi75 = (double)(m - 1) * 0.75;
                const int k25 = (int)floor(vi25), k75 = (int)floor(vi75);
                const double q25 = nm_norm_lerp(NM_NS(k25), NM_NS(k25 + 1 < m ? k25 + 1 : m - 1), vi25 - k25);
                const double q75 = nm_norm_lerp(NM_NS(k75), NM_NS(k75 + 1 < m ? k75 + 1 : m - 1), vi75 - k75);
                double scale = q75 - q25;
                if (scale < NM_NORM_TINY_SCALE) scale = 1.0;
                r = (v - med) / scale;
            }
        }
        if (a.clip > 0.0) {
            if (r < -a.clip) r = -a.clip;
            if (r > a.clip) r = a.clip;
        }
        a.out[(size_t)k * a.F + a.cols[j]] = nm_nan_to_num(r);
    }
#undef NM_NS
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
    // normalise result rows [w0, w0 + n_windows) of the current run (successive calls continue the history)
    int run(nm_pipeline* p, int n_windows, int w0 = 0, int total = -1);
    // O(n_keep) methods can follow every chunk of a batched run (its rows are then final and can be shipped while later chunks
    // compute); the order-statistic methods keep one sliding pass over all windows at the end
    bool per_chunk_ok() const { return method == 0 || method == 2 || method == 4 || method == 6; }
};
