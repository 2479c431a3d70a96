// nm_bursts.cuh -- burst features (features/bursts.py:149-298).
//
//  pass 1  nm_fir_kernel<NmEpiBursts>: band-pass bank ('same' FIR) -> W-point analytic-signal
//          envelope (scipy.signal.hilbert == W-point DFT, one-sided doubling, inverse DFT; the
//          Hilbert multiplier is real-linear, so the channel pair shares one complex transform)
//          -> envelope rows of the chunk + the newest samples appended to the history ring.
//  pass 2  nm_burst_thr_kernel: threshold = numpy 'linear' quantile of the last `ring` envelope
//          samples as of each window -- exact order statistics by radix select on the float64
//          bit patterns (envelopes are >= 0, so the patterns order like the values).
//  pass 3  nm_burst_feat_kernel: env >= thr, run-length statistics with warp ballots.
//
// The history is a TRUE ring of the last time_duration_s (the intended semantics; the reference's
// in-place partition scrambles its buffer once it is full -- SURVEY.md headline facts).  While the
// history is not full (windows 0..290 at defaults) both are identical.
#pragma once

#include "nm_fir.cuh"
#include "nm_specx.cuh"

struct NmEpiBursts {
    static constexpr bool kRegs = false;
    static constexpr bool kRegsOnly = false, kReflectOk = false, kSameOk = true, kConvxOnly = false, kF32Ok = false, kSplitOk = true;  // nm_convx_kernel instantiation traits
    NmFft<double> hfft;     // W-point transform
    int need_scratch;
    double* env;            // chunk envelopes (n_windows, n_ch, nB, Wp)
    long long Wp;
    int nB;
    double* ring;           // (n_ch, nB, cap)
    long long cap;
    long long win0;         // global index (since reset) of the chunk's first window
    int S;                  // samples appended by every window but the first
    int fast_n;             // 1000 / 2000 / 500: W-point transforms through the register-blocked plans of nm_specx.cuh, else 0
    static NM_HD size_t smem_bytes_for(int W, int need_scratch, int fast_n = 0) {
        if (fast_n == 1000) return (size_t)NmSx1000::NBUF * sizeof(cx<double>);
        if (fast_n == 2000) return (size_t)NmSx2000::NBUF * sizeof(cx<double>);
        if (fast_n == 500) return (size_t)NmSx500::NBUF * sizeof(cx<double>);
        return (size_t)W * sizeof(cx<double>) * (need_scratch ? 2 : 1);
    }

    // envelope of one sample pair -> chunk rows + history ring
    NM_DEV void emit(int t, double xa, double xb, double ha, double hb_, int W, int n_ch, int w, int c0, bool has2, int f) const {
        const long long gw = win0 + w;
        const int take = (gw == 0) ? W : S;
        const long long e_prev = (gw == 0) ? 0 : (long long)W + (gw - 1) * S;
        const int j = t - (W - take);
        for (int k = 0; k < (has2 ? 2 : 1); ++k) {
            const int c = c0 + k;
            const double re = k ? xb : xa, im = k ? hb_ : ha;
            const double e = sqrt(re * re + im * im);
            env[(((size_t)w * n_ch + c) * nB + f) * Wp + t] = e;
            if (j >= 0) ring[((size_t)c * nB + f) * cap + (e_prev + j) % cap] = e;
        }
    }

    // W-point analytic signal with a compile-time plan: forward, one-sided multiplier in slot order, inverse into registers
    template <class PL>
    NM_DEV void run_fast(const cx<double>* x, int W, int n_ch, int w, int c0, bool has2, int f, cx<double>* hb, int tid, int nt) const {
        nm_sx_forward_smem<PL>(x, hb, hfft.tw, tid);
        const double inv = 1.0 / W;
        for (int e = tid; e < W; e += nt) {
            const int k = nm_sx_freq_of_slot<PL>(e);
            const cx<double> v = hb[PL::phys(e)];
            cx<double> r = {0.0, 0.0};
            if (k != 0 && 2 * k != W) {
                if (2 * k < W) r = {v.im * inv, -v.re * inv};   // * (-i)
                else r = {-v.im * inv, v.re * inv};             // * (+i)
            }
            hb[PL::phys(e)] = r;
        }
        __syncthreads();
        cx<double> v[PL::R0];
        nm_sx_inverse_regs<PL>(hb, hfft.tw, tid, v);
        if (tid < PL::NA) {
#pragma unroll
            for (int k = 0; k < PL::R0; ++k) {
                const int t = tid + PL::NA * k;
                emit(t, x[t].re, x[t].im, v[k].re, v[k].im, W, n_ch, w, c0, has2, f);
            }
        }
    }

    NM_DEV void run(const cx<double>* buf, int o0, int W, int n_ch, int w, int c0, bool has2, int f,
                    unsigned char* scratch_raw, int tid, int nt) const {
        cx<double>* hb = reinterpret_cast<cx<double>*>(scratch_raw);
        const cx<double>* x = buf + o0;
        if (fast_n == 1000) { run_fast<NmSx1000>(x, W, n_ch, w, c0, has2, f, hb, tid, nt); return; }
        if (fast_n == 2000) { run_fast<NmSx2000>(x, W, n_ch, w, c0, has2, f, hb, tid, nt); return; }
        if (fast_n == 500) { run_fast<NmSx500>(x, W, n_ch, w, c0, has2, f, hb, tid, nt); return; }
        cx<double>* sc = need_scratch ? hb + W : nullptr;
        for (int t = tid; t < W; t += nt) hb[t] = x[t];
        __syncthreads();
        nm_fft_forward<double>(hb, sc, hfft, tid, nt);
        const double inv = 1.0 / W;
        for (int k = tid; k < W; k += nt) {
            const int slot = nm_ldg(hfft.pos + k);
            const cx<double> v = hb[slot];
            cx<double> r = {0.0, 0.0};
            if (k != 0 && 2 * k != W) {
                if (2 * k < W) r = {v.im * inv, -v.re * inv};   // * (-i)
                else r = {-v.im * inv, v.re * inv};             // * (+i)
            }
            hb[slot] = r;
        }
        __syncthreads();
        nm_fft_inverse<double>(hb, sc, hfft, tid, nt);
        for (int t = tid; t < W; t += nt) emit(t, x[t].re, x[t].im, hb[t].re, hb[t].im, W, n_ch, w, c0, has2, f);
    }
};

// ------------------------------------------------------------------ pass 2: thresholds
// threshold = numpy.quantile(history, q) with the 'linear' rule: order statistics k_lo, k_hi of the last n envelope
// samples and a lerp (host supplies k_lo, k_hi, gamma per window).  Envelopes are >= 0, so the float64 bit patterns
// ("keys") order like the values and all selection is done on unsigned 64-bit keys -- exact, no tolerance.
//
// The history slides by S samples per window (S = 100 of 30 000 at defaults), so the threshold is maintained
// INCREMENTALLY per (channel, band) row instead of re-selecting from the whole ring for every window:
//
//   bracket [lo_key, hi_key)   a key interval around the current quantile holding <= NM_BQ_CAP history samples
//   cnt_below                  number of history samples with key < lo_key
//   queue                      the history samples inside the bracket, in TIME order (FIFO: they enter and expire
//                              in time order), with their logical sample index
//
// A window step classifies only the S entering and S expiring samples (cnt_below +-, queue append / pop) and selects
// rank k_lo - cnt_below among the queue entries (<= 2048 keys in shared memory).  When the target rank leaves the
// bracket, the queue overflows or there is no valid state (first window, reset), the bracket is rebuilt from the ring
// with a two-level radix histogram (the former per-window algorithm, which also remains the fallback when a single
// 2^-12-binade sub-bin holds more samples than the queue, e.g. an all-zero signal).
// One CTA owns one row and walks the windows of the chunk in order; the state persists in global memory between
// launches (chunks, streamed windows).
struct NmBurstQRow {
    unsigned long long lo_key, hi_key;
    long long e_prev, first_prev;  // the state describes the history [first_prev, e_prev)
    int cnt_below, count, valid;
    int rebuilds, directs;  // statistics since the last reset: bracket rebuilds, windows served by the direct selection
    int pad;                // (statistics) brackets re-centred without a histogram rebuild
};

struct NmBurstThrArgs {
    const double* ring;
    long long cap;
    int n_ch, nB, n_windows;
    const long long* e_end;  // [n_windows] logical end (exclusive) of the history as of window w
    const int* n_hist;       // [n_windows] number of samples in the history
    const int* k_lo;         // [n_windows] order statistic below the virtual index
    const int* k_hi;
    const double* gamma;     // [n_windows]
    double* thr;             // (n_windows, n_ch, nB)
    NmBurstQRow* qrow;       // [n_ch * nB]
    unsigned long long* qkey;  // [n_ch * nB][NM_BQ_CAP]
    unsigned* qidx;            // [n_ch * nB][NM_BQ_CAP] low 32 bits of the logical sample index
    int incremental;         // 0: re-select from the ring for every window (reference implementation of this kernel)
    // Few rows (channel-sharded runs: 32 channels x 2 bands = 64 rows on 148 SMs): the windows of a launch are split into n_split
    // consecutive ranges per row, one CTA each.  Range 0 continues from the persisted state, the others start with a bracket
    // rebuild from the ring (all envelope samples of the launch are already there), the last one persists its state.  Every
    // threshold is the exact order statistic either way.
    int n_split;
    // where the last range leaves the state for the next launch: the buffers above when n_split == 1 (in place), the other half
    // of a double buffer (zeroed by the host before the launch) otherwise -- range 0 of a row may still be reading the old state
    // while the last range of the same row finishes
    NmBurstQRow* qrow_out;
    unsigned long long* qkey_out;
    unsigned* qidx_out;
};

static inline int nm_burst_thr_split(int n_rows, int n_windows, int n_sm) {
    static const int per_sm = [] { const char* e = getenv("NMB200_BURST_SPLIT_OCC"); const int v = e ? atoi(e) : 4; return v < 0 ? 0 : (v > 4 ? 4 : v); }();
    int s = (per_sm * n_sm) / (n_rows > 0 ? n_rows : 1);
    if (s > 8) s = 8;
    while (s > 1 && n_windows / s < 8) --s;  // a rebuild costs about three incremental steps: keep the ranges long enough
    return s < 1 ? 1 : s;
}

#define NM_SEL_BINS 4096
#define NM_SEL_CAND 1024
#define NM_BQ_CAP 2048
#define NM_BQ_MASK (NM_BQ_CAP - 1)
#define NM_BQ_HALF 448
#define NM_BQ_SLACK 128
#define NM_BQ_THREADS 256

struct NmBqSmem {
    int* hist;                  // NM_SEL_BINS
    unsigned long long* qk;     // NM_BQ_CAP (also the candidate list of the direct selection)
    unsigned* qi;               // NM_BQ_CAP
    int* ctl;                   // 16
    unsigned long long* res;    // 4
    int* wcnt;                  // 32
};

static NM_HD size_t nm_bq_smem_bytes() {
    return NM_SEL_BINS * sizeof(int) + NM_BQ_CAP * 8 + NM_BQ_CAP * 4 + 16 * sizeof(int) + 4 * 8 + 32 * sizeof(int) + 64;
}

// order-preserving map double -> unsigned 64-bit key (all finite values, either sign) and back: the sliding order
// statistics serve the burst envelopes (>= 0) and the raw-normaliser medians (signed samples) with the same code
NM_DEV unsigned long long nm_key_of(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
NM_DEV double nm_val_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
NM_DEV unsigned long long nm_bq_key(const double* rrow, long long cap, long long i) { return nm_key_of(rrow[i % cap]); }

// Visit the history samples i = i0 + lane_or_tid + k * step (i < i1) of the ring row: f(i, key).  `pos0` is the ring
// position of sample 0 of the history (already reduced mod cap); positions wrap at most once.  Four independent loads
// are in flight per thread -- these passes are latency bound otherwise (one CTA streams 240 kB out of L2).
template <class F>
NM_DEV void nm_bq_for_each(const double* NM_RESTRICT rrow, long long cap, long long pos0, int i0, int i1, int first_i, int step, F f) {
    int i = i0 + first_i;
    for (; i + 3 * step < i1; i += 4 * step) {
        long long p0 = pos0 + i, p1 = p0 + step, p2 = p1 + step, p3 = p2 + step;
        if (p0 >= cap) p0 -= cap;
        if (p1 >= cap) p1 -= cap;
        if (p2 >= cap) p2 -= cap;
        if (p3 >= cap) p3 -= cap;
        const double v0 = rrow[p0], v1 = rrow[p1], v2 = rrow[p2], v3 = rrow[p3];
        f(i, nm_key_of(v0));
        f(i + step, nm_key_of(v1));
        f(i + 2 * step, nm_key_of(v2));
        f(i + 3 * step, nm_key_of(v3));
    }
    for (; i < i1; i += step) {
        long long p = pos0 + i;
        if (p >= cap) p -= cap;
        f(i, nm_key_of(rrow[p]));
    }
}

// locate the bin that holds 0-based rank `rank` in hist[0, nbins): ctl[0] = bin, ctl[1] = count below it, ctl[2] = its count.
// Executed by warp 0; nbins is a multiple of 32.  Callers place barriers around it.
NM_DEV void nm_bq_find_bin(const int* hist, int nbins, int rank, int* ctl, int tid) {
    if (tid < 32) {
        const int lane = tid, per = nbins / 32;
        int s = 0;
        for (int i = 0; i < per; ++i) s += hist[lane * per + i];
        int incl = s;
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const int excl = incl - s;
        if (rank >= excl && rank < incl) {
            int run = excl;
            for (int i = 0; i < per; ++i) {
                const int h = hist[lane * per + i];
                if (rank < run + h) {
                    ctl[0] = lane * per + i;
                    ctl[1] = run;
                    ctl[2] = h;
                    break;
                }
                run += h;
            }
        }
    }
}

// Exact order statistics straight from the ring (radix select, 12-bit digits, candidate gather): returns the
// interpolated threshold.  All threads of the CTA call it; result valid in every thread.
NM_DEV double nm_bq_select_direct(const double* rrow, long long cap, long long first, int n, int k_lo, int k_hi, double gamma,
                                  const NmBqSmem& sm, int tid, int nt) {
    const int lane = tid & 31;
    int rank = k_lo;
    unsigned long long prefix = 0ull, mask = 0ull;
    int shift = 52, width = 12;
    bool resolved = false;
    const long long first_mod = first % cap;
    while (true) {
        for (int i = tid; i < NM_SEL_BINS; i += nt) sm.hist[i] = 0;
        __syncthreads();
        const unsigned long long dm = (1ull << width) - 1ull;
        nm_bq_for_each(rrow, cap, first_mod, 0, n, tid, nt, [&](int, unsigned long long key) {
            if ((key & mask) == prefix) atomicAdd(&sm.hist[(int)((key >> shift) & dm)], 1);
        });
        __syncthreads();
        nm_bq_find_bin(sm.hist, NM_SEL_BINS, rank, sm.ctl, tid);
        __syncthreads();
        const int digit = sm.ctl[0], below = sm.ctl[1], cnt = sm.ctl[2];
        rank -= below;
        prefix |= ((unsigned long long)digit) << shift;
        mask |= dm << shift;
        __syncthreads();
        if (shift == 0) { resolved = true; break; }
        if (cnt <= NM_SEL_CAND) break;
        if (shift >= 16) { shift -= 12; width = 12; }
        else { width = shift; shift = 0; }   // final 4 bits
    }
    unsigned long long a_key = prefix;
    if (!resolved) {
        if (tid == 0) sm.ctl[3] = 0;
        __syncthreads();
        nm_bq_for_each(rrow, cap, first_mod, 0, n, tid, nt, [&](int, unsigned long long key) {
            if ((key & mask) == prefix) {
                const int slot = atomicAdd(&sm.ctl[3], 1);
                if (slot < NM_SEL_CAND) sm.qk[slot] = key;
            }
        });
        __syncthreads();
        const int m = sm.ctl[3] < NM_SEL_CAND ? sm.ctl[3] : NM_SEL_CAND;
        for (int i = tid; i < m; i += nt) {
            const unsigned long long x = sm.qk[i];
            int r = 0;
            for (int j = 0; j < m; ++j) {
                const unsigned long long y = sm.qk[j];
                r += (y < x || (y == x && j < i)) ? 1 : 0;
            }
            if (r == rank) sm.res[0] = x;
        }
        __syncthreads();
        a_key = sm.res[0];
    }
    // second order statistic: a again if duplicated far enough, else the smallest larger value
    const double av = nm_val_of(a_key);
    double thr = av;
    if (k_hi != k_lo) {
        if (tid == 0) { sm.ctl[4] = 0; sm.ctl[5] = 0; }
        __syncthreads();
        int less = 0, eq = 0;
        unsigned long long mg = ~0ull;
        nm_bq_for_each(rrow, cap, first_mod, 0, n, tid, nt, [&](int, unsigned long long key) {
            less += key < a_key;
            eq += key == a_key;
            if (key > a_key && key < mg) mg = key;
        });
        less = nm_warp_sum_i(less);
        eq = nm_warp_sum_i(eq);
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xffffffffu, mg, o);
            mg = u < mg ? u : mg;
        }
        if (lane == 0) {
            atomicAdd(&sm.ctl[4], less);
            atomicAdd(&sm.ctl[5], eq);
            sm.qk[tid >> 5] = mg;  // per-warp minima (the candidate list is no longer needed)
        }
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int q = 0; q < ((nt + 31) >> 5); ++q) best = sm.qk[q] < best ? sm.qk[q] : best;
        const double bv = (k_hi < sm.ctl[4] + sm.ctl[5]) ? av : nm_val_of(best);
        const double diff = bv - av;
        double lerp = av + diff * gamma;
        if (gamma >= 0.5) lerp = bv - diff * (1.0 - gamma);
        thr = (diff == 0.0) ? av : lerp;
        __syncthreads();
    }
    return thr;
}

// Fill the queue, in time order, with the history samples of [first, first + n) whose key lies in [lo_key, hi_key) and
// count the samples below lo_key.  Two passes without atomics: warp `wid` owns a contiguous time segment, counts its
// matches, and after a prefix over the warps writes them at its offset.  Returns false (uniformly) if they do not fit.
NM_DEV bool nm_bq_gather(const double* rrow, long long cap, long long first, int n, unsigned long long lo_key, unsigned long long hi_key,
                         const NmBqSmem& sm, int& cnt_below, int& head, int& count, int tid, int nt) {
    const int lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const long long first_mod = first % cap;
    const int seg = (n + nwarp - 1) / nwarp;
    const int s0 = min(n, wid * seg), s1 = min(n, s0 + seg);
    int mine = 0, below = 0;
    nm_bq_for_each(rrow, cap, first_mod, s0, s1, lane, 32, [&](int, unsigned long long key) {
        mine += (key >= lo_key && key < hi_key) ? 1 : 0;
        below += (key < lo_key) ? 1 : 0;
    });
    mine = nm_warp_sum_i(mine);
    below = nm_warp_sum_i(below);
    __syncthreads();
    if (lane == 0) { sm.wcnt[wid] = mine; sm.hist[wid] = below; }
    __syncthreads();
    int off = 0, total = 0, nb = 0;
    for (int q = 0; q < nwarp; ++q) {
        if (q < wid) off += sm.wcnt[q];
        total += sm.wcnt[q];
        nb += sm.hist[q];
    }
    __syncthreads();
    if (total > NM_BQ_CAP - NM_BQ_SLACK) return false;
    for (int i0 = s0; i0 < s1; i0 += 32) {
        const int i = i0 + lane;
        unsigned long long key = 0ull;
        bool in = false;
        if (i < s1) {
            long long p = first_mod + i;
            if (p >= cap) p -= cap;
            key = nm_key_of(rrow[p]);
            in = key >= lo_key && key < hi_key;
        }
        const unsigned bm = __ballot_sync(0xffffffffu, in);
        if (in) {
            const int pos = off + __popc(bm & ((1u << lane) - 1u));
            sm.qk[pos] = key;
            sm.qi[pos] = (unsigned)(first + i);
        }
        off += __popc(bm);
    }
    head = 0;
    count = total;
    cnt_below = nb;
    __syncthreads();
    return true;
}

// (Re)build the bracket around rank k_lo of the history [first, first + n) and fill the queue in time order.
// Returns false (uniformly) when no bracket fits the queue; the caller then falls back to the direct selection.
NM_DEV bool nm_bq_rebuild(const double* rrow, long long cap, long long first, int n, int k_lo, const NmBqSmem& sm, unsigned long long& lo_key,
                          unsigned long long& hi_key, int& cnt_below, int& head, int& count, int tid, int nt) {
    // level 1: sign + exponent
    for (int i = tid; i < NM_SEL_BINS; i += nt) sm.hist[i] = 0;
    __syncthreads();
    const long long first_mod = first % cap;
    nm_bq_for_each(rrow, cap, first_mod, 0, n, tid, nt, [&](int, unsigned long long key) { atomicAdd(&sm.hist[(int)(key >> 52)], 1); });
    __syncthreads();
    nm_bq_find_bin(sm.hist, NM_SEL_BINS, k_lo, sm.ctl, tid);
    __syncthreads();
    const unsigned long long d1 = (unsigned long long)sm.ctl[0];
    const int below1 = sm.ctl[1];
    __syncthreads();
    // level 2: the top 12 mantissa bits inside that binade
    for (int i = tid; i < NM_SEL_BINS; i += nt) sm.hist[i] = 0;
    __syncthreads();
    nm_bq_for_each(rrow, cap, first_mod, 0, n, tid, nt, [&](int, unsigned long long key) {
        if ((key >> 52) == d1) atomicAdd(&sm.hist[(int)((key >> 40) & 0xfffull)], 1);
    });
    __syncthreads();
    nm_bq_find_bin(sm.hist, NM_SEL_BINS, k_lo - below1, sm.ctl, tid);
    __syncthreads();
    if (tid == 0) {
        const int d2 = sm.ctl[0], below2 = sm.ctl[1];
        const int budget = NM_BQ_CAP - NM_BQ_SLACK;
        int total = sm.ctl[2], lo_sub = d2, hi_sub = d2, acc_lo = 0, acc_hi = 0;
        while (lo_sub > 0 && acc_lo < NM_BQ_HALF && total + sm.hist[lo_sub - 1] <= budget) {
            --lo_sub;
            acc_lo += sm.hist[lo_sub];
            total += sm.hist[lo_sub];
        }
        while (hi_sub < NM_SEL_BINS - 1 && acc_hi < NM_BQ_HALF && total + sm.hist[hi_sub + 1] <= budget) {
            ++hi_sub;
            acc_hi += sm.hist[hi_sub];
            total += sm.hist[hi_sub];
        }
        sm.ctl[6] = lo_sub;
        sm.ctl[7] = hi_sub;
        sm.ctl[8] = total;
        sm.ctl[9] = below1 + below2 - acc_lo;
    }
    __syncthreads();
    if (sm.ctl[8] > NM_BQ_CAP - NM_BQ_SLACK) {  // a single sub-bin does not fit (massive ties)
        __syncthreads();
        return false;
    }
    lo_key = (d1 << 52) | ((unsigned long long)sm.ctl[6] << 40);
    hi_key = (d1 << 52) + (((unsigned long long)sm.ctl[7] + 1ull) << 40);
    __syncthreads();
    return nm_bq_gather(rrow, cap, first, n, lo_key, hi_key, sm, cnt_below, head, count, tid, nt);
}

// rank-`r` key (0-based) among the queue entries, plus (when wanted) the next order statistic: res[0] = a, res[1] = b.
// Requires 0 <= r and r + (want_next ? 1 : 0) < count.
NM_DEV void nm_bq_select_queue(const NmBqSmem& sm, int head, int count, unsigned long long lo_key, unsigned long long hi_key, int r,
                               bool want_next, int tid, int nt) {
    const int lane = tid & 31;
    unsigned long long lo = lo_key, range = hi_key - lo_key;
    int rr = r;
    unsigned long long a_key = 0ull;
    while (true) {
        int sh = 56 - __clzll((long long)((range - 1ull) | 1ull));  // smallest shift with (range - 1) >> sh <= 255
        if (sh < 0) sh = 0;
        for (int i = tid; i < 256; i += nt) sm.hist[i] = 0;
        __syncthreads();
        for (int j = tid; j < count; j += nt) {
            const unsigned long long key = sm.qk[(head + j) & NM_BQ_MASK];
            if (key >= lo && key - lo < range) atomicAdd(&sm.hist[(int)((key - lo) >> sh)], 1);
        }
        __syncthreads();
        nm_bq_find_bin(sm.hist, 256, rr, sm.ctl, tid);
        __syncthreads();
        const int tb = sm.ctl[0], below = sm.ctl[1], cnt = sm.ctl[2];
        __syncthreads();
        lo += (unsigned long long)tb << sh;
        range = 1ull << sh;
        rr -= below;
        if (sh == 0) { a_key = lo; break; }  // every key of the bin equals lo
        if (cnt <= 256) {
            // gather the bin's members (order irrelevant) and rank them by counting
            if (tid == 0) sm.ctl[3] = 0;
            __syncthreads();
            unsigned long long* mem = reinterpret_cast<unsigned long long*>(sm.hist + 256);
            for (int j = tid; j < count; j += nt) {
                const unsigned long long key = sm.qk[(head + j) & NM_BQ_MASK];
                if (key >= lo && key - lo < range) mem[atomicAdd(&sm.ctl[3], 1)] = key;
            }
            __syncthreads();
            for (int i = tid; i < cnt; i += nt) {
                const unsigned long long x = mem[i];
                int q = 0;
                for (int j = 0; j < cnt; ++j) {
                    const unsigned long long y = mem[j];
                    q += (y < x || (y == x && j < i)) ? 1 : 0;
                }
                if (q == rr) sm.res[0] = x;
            }
            __syncthreads();
            a_key = sm.res[0];
            break;
        }
    }
    __syncthreads();  // every thread has read res[0] (the counting path publishes the rank-r key there)
    if (tid == 0) { sm.res[0] = a_key; sm.ctl[4] = 0; sm.ctl[5] = 0; }
    __syncthreads();
    if (want_next) {
        int less = 0, eq = 0;
        unsigned long long mg = ~0ull;
        for (int j = tid; j < count; j += nt) {
            const unsigned long long key = sm.qk[(head + j) & NM_BQ_MASK];
            less += key < a_key;
            eq += key == a_key;
            if (key > a_key && key < mg) mg = key;
        }
        less = nm_warp_sum_i(less);
        eq = nm_warp_sum_i(eq);
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xffffffffu, mg, o);
            mg = u < mg ? u : mg;
        }
        if (lane == 0) {
            atomicAdd(&sm.ctl[4], less);
            atomicAdd(&sm.ctl[5], eq);
            reinterpret_cast<unsigned long long*>(sm.hist)[tid >> 5] = mg;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long best = ~0ull;
            for (int q = 0; q < ((nt + 31) >> 5); ++q) {
                const unsigned long long u = reinterpret_cast<unsigned long long*>(sm.hist)[q];
                best = u < best ? u : best;
            }
            sm.res[1] = (r + 1 < sm.ctl[4] + sm.ctl[5]) ? a_key : best;
        }
        __syncthreads();
    }
}

// (64 registers: 4 resident CTAs per SM, so that the 512 rows of 256 channels x 2 bands are ONE wave on 148 SMs)
NM_GLOBAL void NM_LAUNCH_BOUNDS(NM_BQ_THREADS, 4) nm_burst_thr_kernel(NmBurstThrArgs a) {
    NM_SHARED_BYTES(smem);
    NmBqSmem sm;
    sm.hist = reinterpret_cast<int*>(smem);
    sm.qk = reinterpret_cast<unsigned long long*>(sm.hist + NM_SEL_BINS);
    sm.qi = reinterpret_cast<unsigned*>(sm.qk + NM_BQ_CAP);
    sm.ctl = reinterpret_cast<int*>(sm.qi + NM_BQ_CAP);
    sm.res = reinterpret_cast<unsigned long long*>(sm.ctl + 16);
    sm.wcnt = reinterpret_cast<int*>(sm.res + 4);
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const int n_rows = a.n_ch * a.nB;
    const int n_split = a.n_split > 1 ? a.n_split : 1;

    for (int vb = blockIdx.x; vb < n_rows * n_split; vb += gridDim.x) {
        const int row = vb / n_split, part = vb - row * n_split;
        const int w_begin = (int)((long long)a.n_windows * part / n_split), w_end = (int)((long long)a.n_windows * (part + 1) / n_split);
        const double* rrow = a.ring + (size_t)row * a.cap;
        NmBurstQRow st = a.qrow[row];
        unsigned long long lo_key = st.lo_key, hi_key = st.hi_key;
        long long e_prev = st.e_prev, first_prev = st.first_prev;
        int cnt_below = st.cnt_below, count = st.count, head = 0;
        int n_rebuild = 0, n_direct = 0, n_recentre = 0;  // (deltas of this launch: added to the row's counters at the end)
        bool valid = a.incremental && st.valid != 0 && part == 0;
        if (valid) {
            for (int j = tid; j < count; j += nt) {
                sm.qk[j] = a.qkey[(size_t)row * NM_BQ_CAP + j];
                sm.qi[j] = a.qidx[(size_t)row * NM_BQ_CAP + j];
            }
        }
        __syncthreads();

        for (int w = w_begin; w < w_end; ++w) {
            const long long e = a.e_end[w];
            const int n = a.n_hist[w];
            const long long first = e - n;
            const int k_lo = a.k_lo[w], k_hi = a.k_hi[w];
            const bool want_next = k_hi != k_lo;
            if (valid && (e < e_prev || first < first_prev || e - e_prev > a.cap / 2)) valid = false;
            if (valid) {
                // ---- slide: expiring samples [first_prev, first), entering samples [e_prev, e)
                if (tid == 0) { sm.ctl[10] = 0; sm.ctl[11] = 0; }
                __syncthreads();
                int delta = 0;
                nm_bq_for_each(rrow, a.cap, first_prev % a.cap, 0, (int)(first - first_prev), tid, nt,
                               [&](int, unsigned long long key) { delta -= (key < lo_key) ? 1 : 0; });
                int nexp = 0;
                for (int j = tid; j < count; j += nt) nexp += ((int)(sm.qi[(head + j) & NM_BQ_MASK] - (unsigned)first) < 0) ? 1 : 0;
                nexp = nm_warp_sum_i(nexp);
                if (lane == 0 && nexp) atomicAdd(&sm.ctl[11], nexp);
                __syncthreads();
                nexp = sm.ctl[11];
                head = (head + nexp) & NM_BQ_MASK;
                count -= nexp;
                const int n_enter = (int)(e - e_prev);
                const long long enter_mod = e_prev % a.cap;
                for (int base = 0; base < n_enter && valid; base += nt) {
                    const int i = base + tid;
                    unsigned long long key = 0ull;
                    bool in = false;
                    if (i < n_enter) {
                        long long p = enter_mod + i;
                        if (p >= a.cap) p -= a.cap;
                        key = nm_key_of(rrow[p]);
                        if (key < lo_key) delta += 1;
                        else in = key < hi_key;
                    }
                    const unsigned bm = __ballot_sync(0xffffffffu, in);
                    if (lane == 0) sm.wcnt[wid] = __popc(bm);
                    __syncthreads();
                    int off = 0, tot = 0;
                    for (int q = 0; q < nwarp; ++q) {
                        const int c = sm.wcnt[q];
                        if (q < wid) off += c;
                        tot += c;
                    }
                    if (count + tot > NM_BQ_CAP) {
                        valid = false;  // uniform: queue overflow -> rebuild below
                    } else {
                        if (in) {
                            const int pos = (head + count + off + __popc(bm & ((1u << lane) - 1u))) & NM_BQ_MASK;
                            sm.qk[pos] = key;
                            sm.qi[pos] = (unsigned)(e_prev + i);
                        }
                        count += tot;
                    }
                    __syncthreads();
                }
                delta = nm_warp_sum_i(delta);
                if (lane == 0 && delta) atomicAdd(&sm.ctl[10], delta);
                __syncthreads();
                cnt_below += sm.ctl[10];
                const int r = k_lo - cnt_below;
                __syncthreads();
                if (valid && (r < 0 || r + (want_next ? 1 : 0) >= count)) {
                    // the quantile has just drifted past an edge of the bracket: re-centre a bracket of the same key
                    // width on that edge (two atomic-free passes) before falling back to the histogram rebuild
                    const unsigned long long hw = (hi_key - lo_key) >> 1;
                    const unsigned long long c = (r < 0) ? lo_key : hi_key;
                    const unsigned long long nlo = c > hw ? c - hw : 0ull, nhi = c + hw;
                    ++n_recentre;
                    valid = nhi > nlo && nm_bq_gather(rrow, a.cap, first, n, nlo, nhi, sm, cnt_below, head, count, tid, nt);
                    if (valid) {
                        lo_key = nlo;
                        hi_key = nhi;
                        const int r2 = k_lo - cnt_below;
                        if (r2 < 0 || r2 + (want_next ? 1 : 0) >= count) valid = false;
                    }
                }
            }
            double thr;
            bool have = false;
            if (!valid && a.incremental) {
                ++n_rebuild;
                valid = nm_bq_rebuild(rrow, a.cap, first, n, k_lo, sm, lo_key, hi_key, cnt_below, head, count, tid, nt);
                if (valid) {
                    const int r = k_lo - cnt_below;
                    if (r < 0 || r + (want_next ? 1 : 0) >= count) valid = false;  // bracket clipped at a binade edge
                }
            }
            if (valid) {
                nm_bq_select_queue(sm, head, count, lo_key, hi_key, k_lo - cnt_below, want_next, tid, nt);
                const double av = nm_val_of(sm.res[0]);
                thr = av;
                if (want_next) {
                    const double bv = nm_val_of(sm.res[1]);
                    const double g = a.gamma[w];
                    const double diff = bv - av;
                    double lerp = av + diff * g;
                    if (g >= 0.5) lerp = bv - diff * (1.0 - g);
                    thr = (diff == 0.0) ? av : lerp;
                }
                have = true;
                __syncthreads();
            }
            if (!have) ++n_direct;
            if (!have) thr = nm_bq_select_direct(rrow, a.cap, first, n, k_lo, k_hi, a.gamma[w], sm, tid, nt);
            if (tid == 0) a.thr[(size_t)w * n_rows + row] = thr;
            e_prev = e;
            first_prev = first;
            __syncthreads();
        }
        // ---- persist the state of this row
        if (a.incremental) {
            const bool keeper = part == n_split - 1;  // the range that ends the launch hands its state to the next one
            if (valid && keeper) {
                for (int j = tid; j < count; j += nt) {
                    a.qkey_out[(size_t)row * NM_BQ_CAP + j] = sm.qk[(head + j) & NM_BQ_MASK];
                    a.qidx_out[(size_t)row * NM_BQ_CAP + j] = sm.qi[(head + j) & NM_BQ_MASK];
                }
            }
            if (tid == 0) {
                NmBurstQRow* o = a.qrow_out + row;
                if (part == 0 && a.qrow_out != a.qrow) {  // (double-buffered: carry the counters over)
                    n_rebuild += st.rebuilds; n_direct += st.directs; n_recentre += st.pad;
                }
                if (keeper) {  // (field by field: the counters below are updated atomically by every range of the row)
                    o->lo_key = lo_key; o->hi_key = hi_key;
                    o->e_prev = e_prev; o->first_prev = first_prev;
                    o->cnt_below = cnt_below; o->count = count; o->valid = valid ? 1 : 0;
                }
                if (n_rebuild) atomicAdd(&o->rebuilds, n_rebuild);
                if (n_direct) atomicAdd(&o->directs, n_direct);
                if (n_recentre) atomicAdd(&o->pad, n_recentre);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ pass 3: run-length features
struct NmBurstFeatArgs {
    const double* env;   // (n_windows, n_ch, nB, Wp)
    long long Wp;
    const double* thr;   // (n_windows, n_ch, nB)
    int n_windows, n_ch, nB, W;
    double sfreq, seg_s;
    NmOut out;           // per_ch = nB * 6
};

NM_DEV double nm_warp_incl_scan(double v, int lane) {
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

NM_GLOBAL void nm_burst_feat_kernel(NmBurstFeatArgs a) {
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const long long n_rows = (long long)a.n_windows * a.n_ch * a.nB;
    const int W = a.W;
    for (long long row = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); row < n_rows; row += (long long)gridDim.x * wpc) {
        const double* e = a.env + (size_t)row * a.Wp;
        const double thr = a.thr[row];
        int total = 0, trans = 0, nv = 0, maxlen = 0;
        double means = 0.0, amax = 0.0;
        int open_len = 0;
        double open_sum = 0.0;
        unsigned prev_last = 0u;
        const int nsteps = (W + 31) >> 5;
        for (int j = 0; j < nsteps; ++j) {
            const int t = (j << 5) + lane;
            const bool valid = t < W;
            const double ev = valid ? e[t] : 0.0;
            const bool b = valid && (ev >= thr);
            const unsigned mask = __ballot_sync(0xffffffffu, b);
            const unsigned vmask = (j == nsteps - 1 && (W & 31)) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
            total += __popc(mask);
            trans += __popc((mask ^ ((mask << 1) | prev_last)) & vmask);
            if (b && ev > amax) amax = ev;
            const double P = nm_warp_incl_scan(b ? ev : 0.0, lane);
            // a run carried over from the previous step that stops exactly at the step boundary
            if (open_len > 0 && !(mask & 1u)) {
                if (lane == 0) {
                    nv += 1;
                    maxlen = open_len > maxlen ? open_len : maxlen;
                    means += open_sum / open_len;
                }
                open_len = 0;
                open_sum = 0.0;
            }
            // run bookkeeping for lanes inside a run
            const unsigned below = ~mask & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane)));
            const int seg_start = below ? (32 - __clz((int)below)) : 0;
            const bool from_carry = (below == 0u);
            const double Pbefore = __shfl_sync(0xffffffffu, P, seg_start > 0 ? seg_start - 1 : 0);
            const double runsum = P - (seg_start > 0 ? Pbefore : 0.0);
            const bool is_end = b && lane < 31 && !((mask >> (lane + 1)) & 1u);
            if (is_end) {
                const int len = (lane - seg_start + 1) + (from_carry ? open_len : 0);
                const double sum = runsum + (from_carry ? open_sum : 0.0);
                if (t < W - 1) {
                    nv += 1;
                    maxlen = len > maxlen ? len : maxlen;
                    means += sum / len;
                }
            }
            // carry for the next step (warp-uniform: taken from lane 31)
            const int l31_start = __shfl_sync(0xffffffffu, seg_start, 31);
            const int l31_carry = __shfl_sync(0xffffffffu, (int)from_carry, 31);
            const double l31_sum = __shfl_sync(0xffffffffu, runsum, 31);
            if (mask >> 31) {
                if (l31_carry) { open_len += 32; open_sum += l31_sum; }
                else { open_len = 32 - l31_start; open_sum = l31_sum; }
            } else {
                open_len = 0;
                open_sum = 0.0;
            }
            prev_last = mask >> 31;
        }
        nv = nm_warp_sum_i(nv);
        means = nm_warp_sum(means);
        amax = nm_warp_max(amax);
        for (int o = 16; o > 0; o >>= 1) {
            const int u = __shfl_xor_sync(0xffffffffu, maxlen, o);
            maxlen = u > maxlen ? u : maxlen;
        }
        if (lane == 0) {
            const int w = (int)(row / ((long long)a.n_ch * a.nB));
            const int cb = (int)(row - (long long)w * a.n_ch * a.nB);
            const int c = cb / a.nB, bnd = cb - c * a.nB;
            const int num = trans / 2;
            const double dmean = num ? ((double)total / (double)num) / a.sfreq : 0.0;
            nm_store(a.out, w, c, bnd * 6 + 0, dmean);
            nm_store(a.out, w, c, bnd * 6 + 1, (double)maxlen / a.sfreq);
            nm_store(a.out, w, c, bnd * 6 + 2, nv ? means / nv : 0.0);
            nm_store(a.out, w, c, bnd * 6 + 3, amax);
            nm_store(a.out, w, c, bnd * 6 + 4, dmean / a.seg_s);
            nm_store(a.out, w, c, bnd * 6 + 5, (double)((thr <= e[W - 1]) ? 1 : 0));
        }
    }
}

// ------------------------------------------------------------------ host side
// Persistent per-row state of nm_burst_thr_kernel.  Double-buffered: a launch that splits the windows of a row over several CTAs
// reads the state from one half and leaves the new state in the other (no CTA ever writes what another one may still read).
struct NmBqState {
    DevBuf qrow[2], qkey[2], qidx[2];
    int cur = 0;
    size_t rows = 0;
    int alloc(size_t rows_) {
        rows = rows_;
        for (int h = 0; h < 2; ++h) {
            if (qrow[h].ensure(rows * sizeof(NmBurstQRow)) || qkey[h].ensure(rows * NM_BQ_CAP * 8) || qidx[h].ensure(rows * NM_BQ_CAP * 4)) return -1;
            NM_CUDA_CHECK(cudaMemset(qrow[h].p, 0, rows * sizeof(NmBurstQRow)));
        }
        NM_CUDA_CHECK(cudaDeviceSynchronize());  // (legacy-stream memset: finish before any non-blocking stream uses the rows)
        return 0;
    }
    bool allocated() const { return qrow[0].p != nullptr; }
    // stream-ordered: kernels of a run that has not been synchronised yet may still use the queue state (the pipeline's streams
    // are non-blocking, so a cudaMemset on the legacy stream would NOT wait for them)
    void reset(cudaStream_t s) {
        if (allocated()) cudaMemsetAsync(qrow[cur].p, 0, rows * sizeof(NmBurstQRow), s);  // valid = 0 for every row
    }
    const NmBurstQRow* current() const { return qrow[cur].as<NmBurstQRow>(); }
    // fills the state pointers of a launch (ta.n_split already chosen) and makes the written half the current one
    void bind(NmBurstThrArgs& ta, cudaStream_t s) {
        const int o = ta.n_split > 1 ? cur ^ 1 : cur;
        ta.qrow = qrow[cur].as<NmBurstQRow>();
        ta.qkey = qkey[cur].as<unsigned long long>();
        ta.qidx = qidx[cur].as<unsigned>();
        ta.qrow_out = qrow[o].as<NmBurstQRow>();
        ta.qkey_out = qkey[o].as<unsigned long long>();
        ta.qidx_out = qidx[o].as<unsigned>();
        if (o != cur) cudaMemsetAsync(qrow[o].p, 0, rows * sizeof(NmBurstQRow), s);
        cur = o;
    }
};

struct BurstsFam {
    FirBank bank;
    FftPlanHost hfft;
    DevBuf d_env, d_ring, d_thr, d_colmap, d_e_end, d_n, d_lo, d_hi, d_gamma;
    NmBqState qstate;
    int incremental = 1;  // nm_set_burst_threshold_mode(0) selects the per-window re-selection (reference of the kernel)
    int nB = 0, C = 0, W = 0, S = 0, ring_n = 0, chunk = 0;
    long long cap = 0, Wp = 0, batch = 0;
    double q = 0.75, sfreq = 1000, seg_s = 1;

    int build(const double* taps, int nB_, int L, int C_, int W_, int S_, int ring_n_, double q_, double sfreq_, double seg_s_,
              const int* colmap, cudaStream_t s) {
        nB = nB_; C = C_; W = W_; S = S_; ring_n = ring_n_; q = q_; sfreq = sfreq_; seg_s = seg_s_;
        if (bank.build(taps, nB, L, W, NM_FIR_SAME, s)) return -1;
        if (hfft.build(W, s)) return -1;
        return d_colmap.upload(colmap, (size_t)C * nB * 6, s);
    }
    int alloc_chunk(int chunk_, long long Wp_) {
        chunk = chunk_;
        Wp = Wp_;
        cap = (long long)ring_n + (long long)chunk * (S > 0 ? S : 1) + W;
        if (d_env.ensure((size_t)chunk * C * nB * Wp * sizeof(double))) return -1;
        if (d_ring.ensure((size_t)C * nB * cap * sizeof(double))) return -1;
        if (d_thr.ensure((size_t)chunk * C * nB * sizeof(double))) return -1;
        return qstate.alloc((size_t)C * nB);
    }
    void reset(cudaStream_t s) {
        batch = 0;
        qstate.reset(s);
    }
    int fast_n() const { return (nm_specx_supported(W) && bank.threads() >= (W == 500 ? NmSx500::NA : NmSx1000::NA)) ? W : 0; }
    size_t epi_smem() const { return NmEpiBursts::smem_bytes_for(W, hfft.generic, fast_n()); }
    static size_t thr_smem() { return nm_bq_smem_bytes(); }
    int allow_smem(const nm_pipeline* p);
    long long run_base = 0;  // `batch` at the time prepare() ran
    // device arrays prepare() filled for the current run: the DevBufs above (batched runs) or the streaming parameter block
    const long long* a_e_end = nullptr;
    const int *a_n = nullptr, *a_lo = nullptr, *a_hi = nullptr;
    const double* a_gamma = nullptr;
    int prepare(nm_pipeline* p, int n_windows);
    int run(nm_pipeline* p, const NmRows& rows, int w0);
};
