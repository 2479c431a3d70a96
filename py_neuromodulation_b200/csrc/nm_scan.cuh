// nm_scan.cuh -- Hjorth activity / mobility / complexity, line length and last sample.
//   features/hjorth_raw.py:24-42,51-57   features/linelength.py:11-21
//
// One warp per (window, channel) row; the row is streamed twice (second pass hits L1/L2):
// pass 1 accumulates the means of x, dx, ddx and sum|dx|, pass 2 the centred second moments
// (numpy.var is two-pass, ddof = 0).  HBM-bound: 8 bytes in per sample, 5 values out per row.
#pragma once

#include "nm_common.cuh"

struct NmScanArgs {
    NmRows in;
    int want_hjorth, want_raw, want_ll;
    NmOut out;  // per_ch = 5: activity, mobility, complexity, raw, linelength
};

NM_GLOBAL void nm_scan_kernel(NmScanArgs a) {
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const long long n_rows = (long long)a.in.n_windows * a.in.n_ch;
    const int W = a.in.W;
    for (long long row = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); row < n_rows;
         row += (long long)gridDim.x * warps_per_cta) {
        const int w = (int)(row / a.in.n_ch), c = (int)(row - (long long)w * a.in.n_ch);
        const double* x = a.in.base + (size_t)c * a.in.ch_stride + nm_ldg(a.in.off + w);
        double s0 = 0, s1 = 0, s2 = 0, sl = 0;
        for (int t = lane; t < W; t += 32) {
            const double x0 = x[t];
            s0 += x0;
            if (t + 1 < W) {
                const double x1 = x[t + 1];
                const double d = x1 - x0;
                s1 += d;
                sl += fabs(d);
                if (t + 2 < W) s2 += (x[t + 2] - x1) - d;
            }
        }
        s0 = nm_warp_sum(s0); s1 = nm_warp_sum(s1); s2 = nm_warp_sum(s2); sl = nm_warp_sum(sl);
        const double n0 = W, n1 = W - 1, n2 = W - 2;
        if (a.want_hjorth) {
            const double m0 = s0 / n0, m1 = s1 / n1, m2 = s2 / n2;
            double q0 = 0, q1 = 0, q2 = 0;
            for (int t = lane; t < W; t += 32) {
                const double x0 = x[t];
                double e = x0 - m0;
                q0 += e * e;
                if (t + 1 < W) {
                    const double x1 = x[t + 1];
                    const double d = x1 - x0;
                    e = d - m1;
                    q1 += e * e;
                    if (t + 2 < W) {
                        e = ((x[t + 2] - x1) - d) - m2;
                        q2 += e * e;
                    }
                }
            }
            q0 = nm_warp_sum(q0); q1 = nm_warp_sum(q1); q2 = nm_warp_sum(q2);
            if (lane == 0) {
                const double v0 = q0 / n0, v1 = q1 / n1, v2 = q2 / n2;
                const double mob = nm_nan_to_num(sqrt(v1 / v0));
                nm_store(a.out, w, c, 0, nm_nan_to_num(v0));
                nm_store(a.out, w, c, 1, mob);
                nm_store(a.out, w, c, 2, nm_nan_to_num(sqrt(v2 / v1) / mob));
            }
        }
        if (lane == 0) {
            if (a.want_raw) nm_store(a.out, w, c, 3, x[W - 1]);
            // mean(|dx| / (W-1)) over W-1 samples: the reference divides by (W-1) twice
            if (a.want_ll) nm_store(a.out, w, c, 4, (sl / n1) / n1);
        }
    }
}
