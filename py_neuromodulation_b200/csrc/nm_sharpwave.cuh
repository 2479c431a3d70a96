// nm_sharpwave.cuh -- sharp-wave features (features/sharpwaves.py:225-465) as an epilogue of the FIR kernel.
//
// After the 'same' FIR of one (window, channel pair, filter) item the two filtered rows sit in
// shared memory; four "analysis rows" (2 channels x {+x "Peak" pass, -x "Trough" pass}) are then
// analysed by one warp each:
//   1. local maxima of d and of -d with scipy's plateau-midpoint rule      (scipy _local_maxima_1d)
//   2. greedy minimum-distance suppression in priority order               (scipy _select_by_peak_distance)
//      -- evaluated as the lexicographically-first maximal independent set: a peak survives a round iff it
//      out-ranks every still-undecided neighbour closer than D; identical result, but parallel over peaks
//   3. trough <-> nearest left/right peak pairing incl. the reference's slicing quirk
//   4. per-trough features and estimators; the same estimator then joins the Peak and Trough values.
#pragma once

#include "nm_fir.cuh"

#define NM_SW_MAX_COMBO 40

struct NmSwCfg {
    int D_pk, D_tr;      // ceil(distance) in samples
    int off;             // sharpness offset in samples
    double ms;           // 1000 / sfreq
    int n_combo;
    int feat[NM_SW_MAX_COMBO];
    int est[NM_SW_MAX_COMBO];
    int pair_est;        // apply the estimator between the Peak and Trough passes
    int want_num_peaks;
    int maxn;            // capacity of the per-row peak lists (W/2 + 2)
    int tmp_in_tail;     // the median scratch (4 x maxn doubles) lives behind the filtered rows in the transform buffer
    int need_tmp;        // some estimator is a median (only then the scratch is needed at all)
    int rows_in_tail;    // the peak lists of the first k analysis rows live in that tail too (fewer bytes per CTA -> more CTAs per SM)
};

struct NmSwLists {
    unsigned short *rawP, *rawT, *KP, *KT, *Lp, *Rp;
    unsigned char *st, *wn;  // per candidate: state (0 undecided, 1 kept, 2 suppressed) and "wins this round"
    double* tmp;
};

static NM_HD size_t nm_sw_list_bytes(int maxn) {
    const size_t b = (size_t)6 * maxn * sizeof(unsigned short) + (size_t)2 * maxn;
    return (b + 7) & ~(size_t)7;
}
static NM_HD size_t nm_sw_row_bytes(int maxn, int tmp_in_tail) {
    return nm_sw_list_bytes(maxn) + (tmp_in_tail ? 0 : (size_t)maxn * sizeof(double));
}

NM_DEV double nm_sw_v(const cx<double>* x, int comp, double sign, int t) { return sign * (comp ? x[t].im : x[t].re); }

// local maxima of s*d (s = +1 peaks, -1 troughs) -> ascending midpoints in `list`; returns the count
NM_DEV int nm_sw_local_maxima(const cx<double>* x, int comp, double sign, int W, unsigned short* list, int lane) {
    int n = 0;
    // four 32-sample groups per step: their loads and comparisons are independent, only the list positions (ballot prefix
    // counts) are sequential -- the scan is latency bound otherwise (one warp walks the whole row)
    for (int base = 1; base <= W - 2; base += 128) {
        bool pred[4];
        int mid[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            pred[u] = false;
            mid[u] = 0;
            if (i <= W - 2) {
                const double di = nm_sw_v(x, comp, sign, i);
                if (nm_sw_v(x, comp, sign, i - 1) < di) {
                    int j = i + 1;
                    double dj = nm_sw_v(x, comp, sign, j);
                    while (j < W - 1 && dj == di) dj = nm_sw_v(x, comp, sign, ++j);
                    if (dj < di) {
                        pred[u] = true;
                        mid[u] = (i + j - 1) / 2;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned m = __ballot_sync(0xffffffffu, pred[u]);
            if (pred[u]) list[n + __popc(m & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane))))] = (unsigned short)mid[u];
            n += __popc(m);
        }
    }
    __syncwarp();
    return n;
}

// distance suppression; `kept` receives the surviving indices (ascending); returns their count.
// Race-free by construction: in every phase a lane WRITES only the entries of its own candidates and READS other candidates'
// entries only from an array that no lane writes in that phase (st in phase 1, wn in phase 2); __syncwarp separates the phases.
NM_DEV int nm_sw_select(const cx<double>* x, int comp, double sign, const unsigned short* idx, int n, int D, unsigned char* st,
                        unsigned char* wn, unsigned short* kept, int lane) {
    for (int k = lane; k < n; k += 32) st[k] = 0;
    __syncwarp();
    if (D > 1) {
        while (true) {
            // phase 1: an undecided candidate wins the round iff it out-ranks every undecided neighbour closer than D
            for (int k = lane; k < n; k += 32) {
                bool win = false;
                if (st[k] == 0) {
                    const int ik = idx[k];
                    const double hk = nm_sw_v(x, comp, sign, ik);
                    win = true;
                    for (int q = k - 1; q >= 0 && ik - (int)idx[q] < D && win; --q)
                        if (st[q] == 0 && nm_sw_v(x, comp, sign, idx[q]) > hk) win = false;
                    for (int q = k + 1; q < n && (int)idx[q] - ik < D && win; ++q)
                        if (st[q] == 0 && nm_sw_v(x, comp, sign, idx[q]) >= hk) win = false;
                }
                wn[k] = win ? 1 : 0;
            }
            __syncwarp();
            // phase 2: winners are kept; an undecided candidate next to a winner is suppressed (each lane settles its OWN entries)
            bool pending = false;
            for (int k = lane; k < n; k += 32) {
                if (st[k] != 0) continue;
                if (wn[k]) { st[k] = 1; continue; }
                const int ik = idx[k];
                bool lost = false;
                for (int q = k - 1; q >= 0 && ik - (int)idx[q] < D && !lost; --q) lost = wn[q] != 0;
                for (int q = k + 1; q < n && (int)idx[q] - ik < D && !lost; ++q) lost = wn[q] != 0;
                if (lost) st[k] = 2;
                else pending = true;
            }
            __syncwarp();
            if (!__ballot_sync(0xffffffffu, pending)) break;
        }
    } else {
        for (int k = lane; k < n; k += 32) st[k] = 1;
        __syncwarp();
    }
    int m = 0;
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool keep = k < n && st[k] == 1;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (keep) kept[m + __popc(b & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane))))] = idx[k];
        m += __popc(b);
    }
    __syncwarp();
    return m;
}

// value of feature `fid` for retained trough i; returns false if the trough does not contribute
NM_DEV bool nm_sw_feature(const cx<double>* x, int comp, double sign, int W, const NmSwCfg& c, const NmSwLists& l, int fid, int i,
                          int n_ret, int n_pairs, int first, double* out) {
    const bool paired = i < n_pairs;
    const int T = (i < n_ret) ? l.KT[first + i] : 0;
    const int L = paired ? l.Lp[i] : 0, R = paired ? l.Rp[i] : 0;
    switch (fid) {
        case 0: if (!paired) return false; *out = nm_sw_v(x, comp, sign, L); return true;
        case 1: if (!paired) return false; *out = nm_sw_v(x, comp, sign, R); return true;
        case 3: if (i >= n_ret) return false; *out = nm_sw_v(x, comp, sign, T); return true;
        case 4: if (!paired) return false; *out = (double)(R - L); return true;
        case 5:
            if (!paired || i >= n_ret) return false;
            *out = fabs((nm_sw_v(x, comp, sign, R) + nm_sw_v(x, comp, sign, L)) / 2 - nm_sw_v(x, comp, sign, T));
            return true;
        case 6:
            if (i >= (n_ret > 0 ? n_ret : 1)) return false;
            *out = (i == 0) ? 0.0 : (double)(T - (int)l.KT[first + i - 1]) * c.ms;
            return true;
        case 7: if (!paired || i >= n_ret) return false; *out = (double)(L - T) * c.ms; return true;
        case 8: if (!paired || i >= n_ret) return false; *out = (double)(R - T) * c.ms; return true;
        case 9:
            if (i >= n_ret || !(T - c.off > 0 && T + c.off < W)) return false;
            *out = nm_sw_v(x, comp, sign, T) - 0.5 * (nm_sw_v(x, comp, sign, T - c.off) + nm_sw_v(x, comp, sign, T + c.off));
            return true;
        case 10:
        case 11:
        case 12: {
            if (!paired || i >= n_ret) return false;
            double rise = 0.0, decay = 0.0;
            if (fid != 11)
                for (int t = L; t <= T; ++t) {
                    const double s = (t == 0) ? 0.0 : fabs(nm_sw_v(x, comp, sign, t) - nm_sw_v(x, comp, sign, t - 1));
                    rise = s > rise ? s : rise;
                }
            if (fid != 10)
                for (int t = T; t <= R; ++t) {
                    const double s = (t == 0) ? 0.0 : fabs(nm_sw_v(x, comp, sign, t) - nm_sw_v(x, comp, sign, t - 1));
                    decay = s > decay ? s : decay;
                }
            *out = (fid == 10) ? rise : (fid == 11 ? decay : rise - decay);
            return true;
        }
        default: return false;
    }
}

// one warp: full analysis of one row; results[combo] and results[n_combo] = num_peaks
// nP0 / nT0: local maxima of sign*x in l.rawP and of -sign*x in l.rawT (the latter is the OTHER polarity's rawP list:
// the two passes of a channel share their local-maxima scans)
NM_DEV void nm_sw_analyze(const cx<double>* x, int comp, double sign, int W, const NmSwCfg& c, const NmSwLists& l, int nP0, int nT0,
                          double* results, int lane) {
    const int nP = nm_sw_select(x, comp, sign, l.rawP, nP0, c.D_pk, l.st, l.wn, l.KP, lane);
    const int nT = nm_sw_select(x, comp, -sign, l.rawT, nT0, c.D_tr, l.st, l.wn, l.KT, lane);

    // pairing (features/sharpwaves.py:347-374)
    int first_valid = 0, n_valid = 0, last_valid = 0;
    for (int base = 0; base < nT; base += 32) {
        const int i = base + lane;
        bool noleft = false, valid = false;
        if (i < nT) {
            const int t = l.KT[i];
            int lo = 0, hi = nP;  // first peak with index >= t
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int)l.KP[mid] < t) lo = mid + 1; else hi = mid;
            }
            noleft = (lo == 0);
            valid = !noleft && lo < nP;
        }
        const unsigned bl = __ballot_sync(0xffffffffu, noleft);
        const unsigned bv = __ballot_sync(0xffffffffu, valid);
        first_valid += __popc(bl);
        if (bv) last_valid = base + (31 - __clz((int)bv));
        n_valid += __popc(bv);
    }
    const int n_pairs = n_valid;
    int n_ret = (first_valid <= last_valid && nT > 0) ? (last_valid - first_valid + 1) : 0;
    if (n_valid == 0) n_ret = (first_valid == 0 && nT >= 1) ? 1 : 0;  // reference slicing quirk: trough_idx[0:1]
    // left / right peaks of the valid troughs (they are exactly KT[first_valid .. first_valid + n_pairs))
    for (int i = lane; i < n_pairs; i += 32) {
        const int t = l.KT[first_valid + i];
        int lo = 0, hi = nP;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int)l.KP[mid] < t) lo = mid + 1; else hi = mid;
        }
        l.Lp[i] = l.KP[lo - 1];
        l.Rp[i] = l.KP[lo];
    }
    __syncwarp();

    const int n_iter = n_ret > 1 ? n_ret : 1;
    for (int cb = 0; cb < c.n_combo; ++cb) {
        const int fid = c.feat[cb], est = c.est[cb];
        if (fid == 2) { results[cb] = 0.0; continue; }
        double sum = 0.0, mx = -INFINITY, mn = INFINITY;
        int cnt = 0;
        for (int base = 0; base < n_iter; base += 32) {
            const int i = base + lane;
            double v = 0.0;
            const bool ok = (i < n_iter) && nm_sw_feature(x, comp, sign, W, c, l, fid, i, n_ret, n_pairs, first_valid, &v);
            const unsigned b = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                sum += v;
                mx = (v > mx || v != v) ? v : mx;
                mn = (v < mn || v != v) ? v : mn;
                if (est == 1) l.tmp[cnt + __popc(b & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane))))] = v;
            }
            cnt += __popc(b);
        }
        __syncwarp();
        double r = 0.0;
        if (cnt > 0) {
            sum = nm_warp_sum(sum);
            const double mean = sum / cnt;
            if (est == 0) r = mean;
            else if (est == 2) r = nm_warp_max(mx);
            else if (est == 3) r = nm_warp_min(mn);
            else if (est == 4) {
                double q = 0.0;
                for (int i = lane; i < n_iter; i += 32) {
                    double v;
                    if (nm_sw_feature(x, comp, sign, W, c, l, fid, i, n_ret, n_pairs, first_valid, &v)) q += (v - mean) * (v - mean);
                }
                r = nm_warp_sum(q) / cnt;
            } else {
                const int r_lo = (cnt - 1) / 2, r_hi = cnt / 2;
                double vlo = 0.0, vhi = 0.0;
                for (int i = lane; i < cnt; i += 32) {
                    const double xv = l.tmp[i];
                    int rank = 0;
                    for (int j = 0; j < cnt; ++j) {
                        const double y = l.tmp[j];
                        rank += (y < xv || (y == xv && j < i)) ? 1 : 0;
                    }
                    if (rank == r_lo) vlo = xv;
                    if (rank == r_hi) vhi = xv;
                }
                vlo = nm_warp_sum(vlo);
                vhi = nm_warp_sum(vhi);
                r = (r_lo == r_hi) ? vlo : 0.5 * (vlo + vhi);
            }
        }
        __syncwarp();
        if (lane == 0) results[cb] = r;
    }
    if (lane == 0) results[c.n_combo] = (double)n_ret;
}

NM_DEV double nm_sw_join(int est, double a, double b) {
    switch (est) {
        case 0:
        case 1: return (a + b) / 2;
        case 2: return (a > b || a != a) ? a : b;
        case 3: return (a < b || a != a) ? a : b;
        default: {
            const double m = (a + b) / 2;
            return ((a - m) * (a - m) + (b - m) * (b - m)) / 2;
        }
    }
}

struct NmEpiSharpwave {
    static constexpr bool kRegs = false;
    static constexpr bool kRegsOnly = false, kReflectOk = false, kSameOk = true, kConvxOnly = false, kF32Ok = false, kSplitOk = true;  // nm_convx_kernel instantiation traits
    NmSwCfg cfg;
    NmOut out;  // per_ch = nF * (n_combo + 1) * 2 ; slot (f*(n_combo+1) + combo)*2 + polarity
    static NM_HD size_t smem_bytes_for(int maxn, int n_combo, int tmp_in_tail, int rows_in_tail = 0) {
        return (4 - rows_in_tail) * nm_sw_row_bytes(maxn, tmp_in_tail) + (size_t)4 * (n_combo + 1) * sizeof(double) + 4 * sizeof(int) + 16;
    }
    // tail layout: [median scratch: 4 x maxn doubles, if needed and resident there][lists of rows 0 .. rows_in_tail - 1]
    NM_DEV void lists_for(NmSwLists& l, unsigned char* scratch, const cx<double>* tail, int ar) const {
        unsigned char* tail_lists = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(tail)) +
                                    ((cfg.tmp_in_tail && cfg.need_tmp) ? (size_t)4 * cfg.maxn * sizeof(double) : 0);
        unsigned char* base = ar < cfg.rows_in_tail ? tail_lists + (size_t)ar * nm_sw_list_bytes(cfg.maxn)
                                                    : scratch + (size_t)(ar - cfg.rows_in_tail) * nm_sw_row_bytes(cfg.maxn, cfg.tmp_in_tail);
        l.rawP = reinterpret_cast<unsigned short*>(base);
        l.rawT = l.rawP + cfg.maxn;
        l.KP = l.rawT + cfg.maxn;
        l.KT = l.KP + cfg.maxn;
        l.Lp = l.KT + cfg.maxn;
        l.Rp = l.Lp + cfg.maxn;
        l.st = reinterpret_cast<unsigned char*>(l.Rp + cfg.maxn);
        l.wn = l.st + cfg.maxn;
        l.tmp = cfg.tmp_in_tail ? const_cast<double*>(reinterpret_cast<const double*>(tail)) + (size_t)ar * cfg.maxn
                                : reinterpret_cast<double*>(base + nm_sw_list_bytes(cfg.maxn));
    }
    NM_DEV void run(const cx<double>* buf, int o0, int W, int /*n_ch*/, int w, int c0, bool has2, int f, unsigned char* scratch,
                    int tid, int nt) const {
        const int lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
        const size_t rb = nm_sw_row_bytes(cfg.maxn, cfg.tmp_in_tail);
        double* results = reinterpret_cast<double*>(scratch + (4 - cfg.rows_in_tail) * rb);
        int* nraw = reinterpret_cast<int*>(results + (size_t)4 * (cfg.n_combo + 1));  // local-maxima counts of the 4 rows
        const cx<double>* x = buf + o0;
        const cx<double>* tail = x + W;  // free part of the transform buffer (host checked the capacity: tmp_in_tail)
        // phase 1: analysis row (comp, pol) scans the local maxima of (pol ? -x : x) into its rawP list
        for (int ar = wid; ar < 4; ar += nwarp) {
            const int comp = ar >> 1, pol = ar & 1;
            if (comp == 1 && !has2) continue;
            NmSwLists l;
            lists_for(l, scratch, tail, ar);
            const int n = nm_sw_local_maxima(x, comp, pol ? -1.0 : 1.0, W, l.rawP, lane);
            if (lane == 0) nraw[ar] = n;
        }
        __syncthreads();
        // phase 2: the other polarity's list serves as this row's trough candidates
        for (int ar = wid; ar < 4; ar += nwarp) {
            const int comp = ar >> 1, pol = ar & 1;
            if (comp == 1 && !has2) continue;
            NmSwLists l, other;
            lists_for(l, scratch, tail, ar);
            lists_for(other, scratch, tail, ar ^ 1);
            l.rawT = other.rawP;
            nm_sw_analyze(x, comp, pol ? -1.0 : 1.0, W, cfg, l, nraw[ar], nraw[ar ^ 1], results + (size_t)ar * (cfg.n_combo + 1), lane);
        }
        __syncthreads();
        const int per_f = (cfg.n_combo + 1) * 2;
        for (int i = tid; i < 2 * (cfg.n_combo + 1); i += nt) {
            const int comp = i / (cfg.n_combo + 1), cb = i - comp * (cfg.n_combo + 1);
            if (comp == 1 && !has2) continue;
            const double a = results[(size_t)(comp * 2 + 0) * (cfg.n_combo + 1) + cb];
            const double b = results[(size_t)(comp * 2 + 1) * (cfg.n_combo + 1) + cb];
            const int c = c0 + comp;
            if (cb == cfg.n_combo) {
                if (cfg.want_num_peaks) {
                    if (cfg.pair_est) nm_store(out, w, c, f * per_f + cb * 2, (a + b) / 2);
                    else { nm_store(out, w, c, f * per_f + cb * 2, a); nm_store(out, w, c, f * per_f + cb * 2 + 1, b); }
                }
            } else if (cfg.feat[cb] != 2) {
                if (cfg.pair_est) nm_store(out, w, c, f * per_f + cb * 2, nm_sw_join(cfg.est[cb], a, b));
                else { nm_store(out, w, c, f * per_f + cb * 2, a); nm_store(out, w, c, f * per_f + cb * 2 + 1, b); }
            }
        }
    }
};

struct SharpwaveFam {
    FirBank bank;
    NmSwCfg cfg;
    DevBuf d_colmap;
    int C = 0, per_ch = 0;
    int build(const double* taps, int nF, int L, int C_, int W, int dist_peaks, int dist_troughs, int sharp_offset, double ms, int n_combo,
              const int* feat_ids, const int* est_ids, int pair_est, int want_num_peaks, const int* colmap, cudaStream_t s) {
        NM_CHECK(n_combo <= NM_SW_MAX_COMBO, "at most %d (feature, estimator) pairs are supported", NM_SW_MAX_COMBO);
        NM_CHECK(W <= 65535, "sharp-wave analysis supports windows of at most 65535 samples");
        NM_CHECK(dist_peaks >= 1 && dist_troughs >= 1, "peak distances must be >= 1 sample");
        C = C_;
        if (bank.build(taps, nF, L, W, NM_FIR_SAME, s)) return -1;
        cfg.D_pk = dist_peaks; cfg.D_tr = dist_troughs; cfg.off = sharp_offset; cfg.ms = ms;
        cfg.n_combo = n_combo;
        for (int i = 0; i < n_combo; ++i) {
            NM_CHECK(feat_ids[i] >= 0 && feat_ids[i] <= 12 && est_ids[i] >= 0 && est_ids[i] <= 4, "bad sharp-wave feature/estimator id");
            cfg.feat[i] = feat_ids[i];
            cfg.est[i] = est_ids[i];
        }
        cfg.pair_est = pair_est;
        cfg.want_num_peaks = want_num_peaks;
        cfg.maxn = W / 2 + 2;
        // the transform buffer holds P >= W + (L-1)/2 elements; whatever follows the W filtered samples is free
        const size_t tail_bytes = (size_t)(bank.P - W) * sizeof(cx<double>), tmp_bytes = (size_t)4 * cfg.maxn * sizeof(double);
        cfg.need_tmp = 0;
        for (int i = 0; i < n_combo; ++i) cfg.need_tmp |= (est_ids[i] == 1) ? 1 : 0;
        cfg.tmp_in_tail = (tail_bytes >= tmp_bytes) ? 1 : 0;
        // whatever the median scratch leaves of the tail takes the peak lists of the first rows (no median configured -- the
        // default -- frees all of it): 2 rows at P = 2048 / W = 1000 (4 instead of 3 CTAs per SM), all 4 at P = 4096
        const size_t used = (cfg.tmp_in_tail && cfg.need_tmp) ? tmp_bytes : 0;
        cfg.rows_in_tail = cfg.tmp_in_tail ? (int)std::min<size_t>(4, (tail_bytes - used) / nm_sw_list_bytes(cfg.maxn)) : 0;
        per_ch = nF * (n_combo + 1) * 2;
        return d_colmap.upload(colmap, (size_t)C * per_ch, s);
    }
    size_t epi_smem() const { return NmEpiSharpwave::smem_bytes_for(cfg.maxn, cfg.n_combo, cfg.tmp_in_tail, cfg.rows_in_tail); }
    int allow_smem(const nm_pipeline* p);
    int run(nm_pipeline* p, const NmRows& rows, int w0);
};
